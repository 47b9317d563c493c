"""cpfem_b200: B200-native (sm_100a) hot path of JAX-CPFEM behind the reference's Problem API.

Host-side mirror of the reference interface:
  generate_mesh.Mesh / box_mesh            (jax_fem.generate_mesh names used by the drivers)
  problem.Problem, problem.CrystalPlasticityBase
  models_copper / models_tantalum / models_304steel / models_DPsteel_inhomo .CrystalPlasticity
All numerical work happens in libcpfem_b200.so (csrc/), called through the C ABI of include/cpfem.h.
"""
from ._lib import build, lib, LIB_PATH, CpfemError  # noqa: F401
try:
    from .api import Plan, make_material, LAYOUT_AOS, LAYOUT_SOA  # noqa: F401
except ImportError as _e:          # a JAX-only box without torch: the XLA-FFI path (jax_ffi, jax_problem) still works
    if getattr(_e, 'name', '') != 'torch':
        raise
