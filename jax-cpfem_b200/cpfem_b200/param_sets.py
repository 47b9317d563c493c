"""The reference's material parameter sets as plain data (no torch, no JAX): what the `models_*.py` files of JAX-CPFEM
hard-code inside `custom_init` / `get_maps`.  Used by the JAX-side shim (jax_problem.py); the torch mirror keeps the same
numbers as class attributes of its `models_*.CrystalPlasticity` classes (tests/test_jax_ffi.py checks the two agree)."""
from . import slip_systems

# field order of cpfem_material (include/cpfem.h)
MATERIAL_FIELDS = ('C11', 'C12', 'C44', 'h', 't_sat', 'gss_a', 'ao', 'xm', 'r', 'tol', 'max_sub_step', 'max_iter')


def _mat(C11, C12, C44, h, t_sat, gss_a, xm, max_sub_step, gss_initial, slip, r=1.0, ao=0.001, tol=1e-8, max_iter=200):
    return dict(material=dict(C11=C11, C12=C12, C44=C44, h=h, t_sat=t_sat, gss_a=gss_a, ao=ao, xm=xm, r=r, tol=tol,
                              max_sub_step=max_sub_step, max_iter=max_iter), gss_initial=gss_initial, slip=slip)


PRESETS = {
    # singlecrystal_copper/models_copper.py:54-56,94-96,141-149,231
    'copper': _mat(1.684e5, 1.214e5, 0.754e5, 541.5, 109.8, 2.5, 0.1, 5, 60.8, slip_systems.FCC12),
    # singlecrystal_tantalum/models_tantalum.py:56,59,96-98,143-151,234
    'tantalum': _mat(2.670e5, 1.610e5, 0.825e5, 1959.1320, 7295.1754, 200.0, 1.0 / 45.2726, 5, 67.4641, slip_systems.BCC12),
    # polycrystal_304steel/models_304steel.py:56,95-97,143-151,232
    '304steel': _mat(2.622e5, 1.120e5, 0.746e5, 392.9772, 7295.1754, 8.0, 1.0 / 120.0, 8, 90.0, slip_systems.FCC12),
    # polycrystal_DPsteel/models_DPsteel_inhomo.py:73-86 (phase 0; the per-point arrays of internal_vars override these)
    'dpsteel': _mat(2.314e5, 1.347e5, 1.164e5, 400.0, 2500.0, 4.0, 0.05, 5, 170.0, slip_systems.BCC24),
}
