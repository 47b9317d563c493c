"""BCC tantalum, 12 slip systems - singlecrystal_tantalum/models_tantalum.py:56,59,96-98,143-151,234."""
from .problem import CrystalPlasticityBase, get_rot_mat, get_rot_mat_vmap  # noqa: F401
from . import slip_systems


class CrystalPlasticity(CrystalPlasticityBase):
    slip_file = slip_systems.BCC12
    gss_initial = 67.4641
    C11, C12, C44 = 2.670e5, 1.610e5, 0.825e5
    h, t_sat, gss_a, xm = 1959.1320, 7295.1754, 200.0, 1.0 / 45.2726
    max_sub_step = 5
