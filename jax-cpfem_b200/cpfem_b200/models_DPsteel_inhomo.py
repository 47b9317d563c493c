"""Two-phase (ferrite / martensite) BCC steel, 24 slip systems, per-quadrature-point material parameters and
elastic tensor - polycrystal_DPsteel/models_DPsteel_inhomo.py:55-229 (setup), :240-361 (per-point maps)."""
import numpy as onp

from . import api, slip_systems
from .problem import CrystalPlasticityBase, get_rot_mat, get_rot_mat_vmap  # noqa: F401


def cubic_C(C11, C12, C44):
    """models_DPsteel_inhomo.py:120-180."""
    C = onp.zeros((3, 3, 3, 3))
    for i in range(3):
        C[i, i, i, i] = C11
        for j in range(3):
            if i != j:
                C[i, i, j, j] = C12
                C[i, j, i, j] = C44
                C[i, j, j, i] = C44
    return C


class CrystalPlasticity(CrystalPlasticityBase):
    slip_file = slip_systems.BCC24
    max_sub_step = 5
    phase1_volume = 0.4
    # phase 0 ferrite / phase 1 martensite (models_DPsteel_inhomo.py:73-102)
    phase = dict(r0=(1., 1.), gss_initial=(170.0, 435.0), h0=(400.0, 950.0), gss_a0=(4.0, 4.0), t_sat0=(2500.0, 5300.0),
                 xm0=(0.05, 0.05), C11=(2.314e5, 4.174e5), C12=(1.347e5, 2.424e5), C44=(1.164e5, 2.111e5))
    phase1_array = None       # set before construction (or pass through additional_info[2]) for a fixed draw

    def custom_init(self, quat, cell_ori_inds, phase1_array=None):
        nc, nq = self.fes[0].num_cells, self.fes[0].num_quads
        if phase1_array is None:
            phase1_array = self.phase1_array
        if phase1_array is None:          # reference: unseeded shuffle (models_DPsteel_inhomo.py:58-65)
            phase1_array = onp.zeros(nc, dtype=int)
            phase1_array[:int(nc * self.phase1_volume)] = 1
            onp.random.shuffle(phase1_array)
        self.phase1_array = onp.asarray(phase1_array, dtype=int)
        ph = self.phase1_array
        pick = lambda k: onp.array(self.phase[k])[ph]
        self.slip_table = onp.asarray(self.slip_file, dtype=onp.float64)
        ns = len(self.slip_table)
        self.num_slip_sys = ns
        quat = onp.asarray(quat, dtype=onp.float64)
        ori = onp.clip(onp.asarray(cell_ori_inds, dtype=onp.int64), 0, len(quat) - 1)    # JAX gather clamps (:212)
        rot_mats_gp = onp.repeat(get_rot_mat(quat)[ori][:, None], nq, axis=1)
        Fp_inv_gp = onp.tile(onp.eye(3)[None, None], (nc, nq, 1, 1))
        slip_resistance_gp = onp.repeat(onp.repeat(pick('gss_initial')[:, None], nq, axis=1)[:, :, None], ns, axis=2)
        slip_gp = onp.zeros_like(slip_resistance_gp)
        rep = lambda v: onp.repeat(v[:, None], nq, axis=1)
        C_phase = onp.array([cubic_C(self.phase['C11'][k], self.phase['C12'][k], self.phase['C44'][k]) for k in (0, 1)])[ph]
        C_gp = onp.repeat(C_phase[:, None], nq, axis=1)
        # uniform fields of the material struct are unused when the per-point arrays are present
        self.material = api.make_material(self.phase['C11'][0], self.phase['C12'][0], self.phase['C44'][0],
                                          self.phase['h0'][0], self.phase['t_sat0'][0], self.phase['gss_a0'][0],
                                          self.phase['xm0'][0], 1.0, 0.001, self.tol, self.max_sub_step)
        self.internal_vars = [Fp_inv_gp, slip_resistance_gp, slip_gp, rot_mats_gp, rep(pick('gss_a0')), rep(pick('h0')),
                              rep(pick('t_sat0')), rep(pick('xm0')), rep(pick('r0')), C_gp]
