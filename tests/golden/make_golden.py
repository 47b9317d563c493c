"""Collects the reference's committed known-answer artefacts for the hot path into tests/golden/ (run in the
build container, where /root/reference exists; the GPU box only sees the committed copies).

  copper_ss_curve.txt     <- calibration/data/csv/calibration_case1/stress_strain_curve_copper.txt
                             (consumer: calibration/calibration_case1_singleCrystalCopper_GB.py:120-121,166-167)
  tantalum_ss_curve.txt   <- calibration/data/csv/calibration_case2/ss_curve_e-2.txt
                             (consumer: calibration/calibration_case2_singleCrystalTa_GB.py:107-108,146)
  steel304_uq_zz_curve.txt<- calibration/data/csv/calibration_case4/UQ/stress_zz_curve_scenario0.txt
  quat_304.txt            <- polycrystal_304steel/data/csv/polycrystal_304steel/quat.txt
"""
import os
import shutil

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = {
    'copper_ss_curve.txt': 'calibration/data/csv/calibration_case1/stress_strain_curve_copper.txt',
    'tantalum_ss_curve.txt': 'calibration/data/csv/calibration_case2/ss_curve_e-2.txt',
    'steel304_uq_zz_curve.txt': 'calibration/data/csv/calibration_case4/UQ/stress_zz_curve_scenario0.txt',
    'quat_304.txt': 'polycrystal_304steel/data/csv/polycrystal_304steel/quat.txt',
}
if __name__ == '__main__':
    for dst, src in FILES.items():
        shutil.copyfile(os.path.join(REF, src), os.path.join(HERE, dst))
        print('copied', src, '->', dst)
