"""FCC 304 steel, 12 slip systems - polycrystal_304steel/models_304steel.py:56,95-97,143-151,232."""
from .problem import CrystalPlasticityBase, get_rot_mat, get_rot_mat_vmap  # noqa: F401
from . import slip_systems


class CrystalPlasticity(CrystalPlasticityBase):
    slip_file = slip_systems.FCC12
    gss_initial = 90.0
    C11, C12, C44 = 2.622e5, 1.120e5, 0.746e5
    h, t_sat, gss_a, xm = 392.9772, 7295.1754, 8.0, 1.0 / 120.0
    max_sub_step = 8
