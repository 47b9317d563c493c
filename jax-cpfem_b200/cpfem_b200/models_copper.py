"""FCC copper, 12 slip systems - parameter set of singlecrystal_copper/models_copper.py:54-56,94-96,141-149,231."""
from .problem import CrystalPlasticityBase, get_rot_mat, get_rot_mat_vmap  # noqa: F401
from . import slip_systems


class CrystalPlasticity(CrystalPlasticityBase):
    slip_file = slip_systems.FCC12
    gss_initial = 60.8
    C11, C12, C44 = 1.684e5, 1.214e5, 0.754e5
    h, t_sat, gss_a, xm = 541.5, 109.8, 2.5, 0.1
    max_sub_step = 5
