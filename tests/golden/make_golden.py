"""Collects the reference's committed known-answer artefacts for the hot path into tests/golden/ (run in the
build container, where /root/reference exists; the GPU box only sees the committed copies).

  copper_ss_curve.txt     <- calibration/data/csv/calibration_case1/stress_strain_curve_copper.txt
                             (consumer: calibration/calibration_case1_singleCrystalCopper_GB.py:120-121,166-167)
  tantalum_ss_curve.txt   <- calibration/data/csv/calibration_case2/ss_curve_e-2.txt
                             (consumer: calibration/calibration_case2_singleCrystalTa_GB.py:107-108,146)
  steel304_uq_zz_curve.txt<- calibration/data/csv/calibration_case4/UQ/stress_zz_curve_scenario0.txt
  quat_304.txt            <- polycrystal_304steel/data/csv/polycrystal_304steel/quat.txt
  mesh2.msh               <- singlecrystal_copper/data/neper/singlecrystal_copper/mesh2.msh        (Neper, 2^3 cells, 1 grain)
  domain0_mesh5.msh       <- polycrystal_304steel/data/neper/polycrystal_304steel/domain0_mesh5.msh (Neper, 5^3 cells, 8 grains)
  box.msh                 <- calibration/data/msh/box.msh                                           (Gmsh, 1 hex + lower-dim. elements)
                             (reader fixtures for cpfem_b200.generate_mesh.read_gmsh22_hex)
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
    'mesh2.msh': 'singlecrystal_copper/data/neper/singlecrystal_copper/mesh2.msh',
    'domain0_mesh5.msh': 'polycrystal_304steel/data/neper/polycrystal_304steel/domain0_mesh5.msh',
    'box.msh': 'calibration/data/msh/box.msh',
}
if __name__ == '__main__':
    for dst, src in FILES.items():
        shutil.copyfile(os.path.join(REF, src), os.path.join(HERE, dst))
        print('copied', src, '->', dst)


# ---- VTU series the reference commits (outputs of its own drivers): decoded into small .npz fixtures --------------
#   steel304_vtu.npz  <- polycrystal_304steel/data/vtk/polycrystal_304steel/u_000..u_007.vtu
#                        (driver polycrystal_304steel/polycrystal_304steel.py:83-233, mesh domain0_mesh16.msh)
#   dpsteel_vtu.npz   <- polycrystal_DPsteel/data/vtk/polycrystal_DPsteel/u_inhomo_000..007.vtu
#                        (driver polycrystal_DPsteel/polycrystal_DPsteel_inhomo.py:78-260, mesh n1-id1-mesh10.msh); the
#                        cell fields phase_inds / cell_ori_inds / C11.. make the driver's unseeded random draw recoverable
def read_vtu(path):
    """meshio's binary VTU: base64(uint32 header [nblocks, blocksize, lastsize, csizes...]) + base64(zlib blocks)."""
    import base64, re, struct, zlib
    import numpy as np
    txt = open(path).read()
    out = {}
    types = {'Float32': np.float32, 'Float64': np.float64, 'Int32': np.int32, 'Int64': np.int64, 'UInt8': np.uint8}
    for m in re.finditer(r'<DataArray([^>]*)>(.*?)</DataArray>', txt, re.S):
        attrs = dict(re.findall(r'(\w+)="([^"]*)"', m.group(1)))
        if attrs.get('format') != 'binary':
            continue
        raw = m.group(2).strip().encode()
        nblocks = struct.unpack('<I', base64.b64decode(raw[:24])[:4])[0]
        hbytes = 4 * (3 + nblocks)
        hlen = ((hbytes + 2) // 3) * 4
        hdr = np.frombuffer(base64.b64decode(raw[:hlen])[:hbytes], dtype=np.uint32)
        data = base64.b64decode(raw[hlen:])
        buf, off = b'', 0
        for c in hdr[3:]:
            buf += zlib.decompress(data[off:off + c])
            off += c
        arr = np.frombuffer(buf, dtype=types[attrs['type']])
        nc = int(attrs.get('NumberOfComponents', 1))
        out[attrs.get('Name', 'noname')] = arr.reshape(-1, nc) if nc > 1 else arr
    return out


def vtu_fixtures(nsteps=8):
    import numpy as np
    d304 = os.path.join(REF, 'polycrystal_304steel/data/vtk/polycrystal_304steel')
    v = [read_vtu(os.path.join(d304, f'u_{k:03d}.vtu')) for k in range(nsteps)]
    np.savez_compressed(os.path.join(HERE, 'steel304_vtu.npz'),
                        points=v[0]['Points'], cells=v[0]['connectivity'].reshape(-1, 8).astype(np.int32),
                        cell_ori_inds=v[0]['cell_ori_inds'].astype(np.int16),
                        sigma_zz=np.stack([x['sigma_zz'] for x in v]), sigma_xx=np.stack([x['sigma_xx'] for x in v]),
                        sigma_yy=np.stack([x['sigma_yy'] for x in v]))
    ddp = os.path.join(REF, 'polycrystal_DPsteel/data/vtk/polycrystal_DPsteel')
    v = [read_vtu(os.path.join(ddp, f'u_inhomo_{k:03d}.vtu')) for k in range(nsteps)]
    np.savez_compressed(os.path.join(HERE, 'dpsteel_vtu.npz'),
                        points=v[0]['Points'], cells=v[0]['connectivity'].reshape(-1, 8).astype(np.int32),
                        cell_ori_inds=v[0]['cell_ori_inds'].astype(np.int16), phase_inds=v[0]['phase_inds'].astype(np.int8),
                        C11=v[0]['C11'].astype(np.float64), C12=v[0]['C12'].astype(np.float64), C44=v[0]['C44'].astype(np.float64),
                        sol=np.stack([x['sol'] for x in v]), sigma_zz=np.stack([x['sigma_zz'] for x in v]),
                        sigma_xx=np.stack([x['sigma_xx'] for x in v]), sigma_yy=np.stack([x['sigma_yy'] for x in v]))
    # tantalum_vtu.npz <- singlecrystal_tantalum/data/vtk/singlecrystal_tantalum/u_000..007.vtu (10^3 cells, BCC12, quat = identity,
    #                     driver singlecrystal_tantalum/singlecrystal_tantalum.py:65-251)
    dta = os.path.join(REF, 'singlecrystal_tantalum/data/vtk/singlecrystal_tantalum')
    v = [read_vtu(os.path.join(dta, f'u_{k:03d}.vtu')) for k in range(nsteps)]
    np.savez_compressed(os.path.join(HERE, 'tantalum_vtu.npz'),
                        points=v[0]['Points'], cells=v[0]['connectivity'].reshape(-1, 8).astype(np.int32),
                        sol=np.stack([x['sol'] for x in v]), sigma_zz=np.stack([x['sigma_zz'] for x in v]),
                        sigma_xx=np.stack([x['sigma_xx'] for x in v]))
    shutil.copyfile(os.path.join(REF, 'polycrystal_DPsteel/data/csv/polycrystal_DPsteel/quat.txt'), os.path.join(HERE, 'quat_dp.txt'))
    print('wrote steel304_vtu.npz, dpsteel_vtu.npz, quat_dp.txt')


def case4_fixture():
    """calibration_case4 (304 steel, 20^3 cells / 50 grains; calibration_case4_UQ_polyCrystalSteel_1D_GB.py:100-230): the
    Neper mesh n50-id0.msh is exactly the structured box (checked here), so the fixture is the grain id of every cell +
    the 50 quaternions; the known answer is steel304_uq_zz_curve.txt (copied above)."""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(HERE, '..', '..', 'jax-cpfem_b200'))
    from cpfem_b200.generate_mesh import read_gmsh22_hex, box_mesh
    m = read_gmsh22_hex(os.path.join(REF, 'calibration/data/neper/calibration_case4/UQ/n50-id0.msh'))
    L = m.points.max(0)
    b = box_mesh(20, 20, 20, *L)
    assert np.allclose(b.points, m.points, atol=1e-12) and np.array_equal(b.cells_dict['hexahedron'], m.cells_dict['hexahedron'])
    quat = np.loadtxt(os.path.join(REF, 'calibration/data/csv/calibration_case4/quat.txt'))[:50, 1:]
    np.savez_compressed(os.path.join(HERE, 'steel304_case4.npz'), cell_grain_inds=(m.cell_data['gmsh:physical'][0] - 1).astype(np.int8),
                        quat=quat, L=L)
    print('wrote steel304_case4.npz')


if __name__ == '__main__':
    vtu_fixtures()
    case4_fixture()
