"""JAX side of the XLA-FFI adapter (csrc/cpfem_ffi.cc): builds libcpfem_ffi.so against the headers of
`jax.ffi.include_dir()`, registers the handlers as CUDA custom-call targets, and creates plans from JAX device arrays.

Needs JAX with the FFI API (jax >= 0.4.35: `jax.ffi`, or `jax.extend.ffi` on 0.4.3x).  JAX cannot be installed in the
container this repo is built in (no index), so nothing here runs there; tests/test_jax_ffi.py skips without JAX and
tests/test_abi.py type-checks the C++ side against a stand-in header.  torch is NOT needed by this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as onp

from . import _lib
from .param_sets import MATERIAL_FIELDS

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(os.path.dirname(_HERE), 'csrc')
_INCLUDE = os.path.join(os.path.dirname(os.path.dirname(_HERE)), 'include')
FFI_LIB_PATH = os.path.join(_HERE, 'libcpfem_ffi.so')
TARGETS = ('cpfem_update_state_ffi', 'cpfem_avg_stress_ffi', 'cpfem_update_avg_ffi', 'cpfem_residual_ffi',
           'cpfem_newton_update_ffi', 'cpfem_point_eval_ffi', 'cpfem_dirichlet_ffi', 'cpfem_bicgstab_ffi',
           'cpfem_point_jac_x_ffi', 'cpfem_vjp_params_ffi', 'cpfem_csr_transpose_ffi')

_registered = False


def jax_ffi_module():
    """`jax.ffi` (or `jax.extend.ffi` on older releases); raises ImportError when JAX or its FFI API is missing."""
    import jax
    mod = getattr(jax, 'ffi', None)
    if mod is None or not hasattr(mod, 'ffi_call'):
        from jax.extend import ffi as mod            # jax 0.4.31 ... 0.4.34
    return mod


def available():
    try:
        jax_ffi_module()
        return True
    except Exception:
        return False


def build(force=False, cuda_include='/usr/local/cuda/include'):
    """g++ -shared csrc/cpfem_ffi.cc against the XLA FFI headers JAX ships, linked to the in-tree libcpfem_b200.so."""
    ffi = jax_ffi_module()
    src = os.path.join(_CSRC, 'cpfem_ffi.cc')
    if not force and os.path.exists(FFI_LIB_PATH) and os.path.getmtime(FFI_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(_lib.LIB_PATH)):
        return FFI_LIB_PATH
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    cmd = ['g++', '-std=c++17', '-O2', '-shared', '-fPIC', '-I', ffi.include_dir(), '-I', cuda_include, '-I', _INCLUDE, src,
           '-L', _HERE, '-lcpfem_b200', '-Wl,-rpath,$ORIGIN', '-o', FFI_LIB_PATH]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('building libcpfem_ffi.so failed:\n' + r.stdout + r.stderr)
    return FFI_LIB_PATH


def register():
    """Registers every handler of libcpfem_ffi.so with XLA (platform CUDA).  Idempotent."""
    global _registered
    if _registered:
        return
    ffi = jax_ffi_module()
    lib = ctypes.CDLL(build())
    for name in TARGETS:
        ffi.register_ffi_target(name, ffi.pycapsule(getattr(lib, name)), platform='CUDA')
    _registered = True


def material_attr(mat: dict):
    """cpfem_material as the dictionary attribute the handlers decode (field types must match the struct)."""
    out = {}
    for k in MATERIAL_FIELDS:
        out[k] = onp.int32(mat[k]) if k in ('max_sub_step', 'max_iter') else onp.float64(mat[k])
    return out


class JaxPlan:
    """cpfem_plan built from JAX device arrays (cells int32 (nc, 8), points float64 (nnodes, 3)) through the C ABI.
    The handle is passed to the FFI handlers as an int64 attribute; the CSR pattern is copied to the host once."""

    def __init__(self, cells, points, slip):
        import jax
        import jax.numpy as jnp
        L = _lib.lib()
        self.cells = jnp.asarray(onp.ascontiguousarray(cells, dtype=onp.int32))
        self.points = jnp.asarray(onp.ascontiguousarray(points, dtype=onp.float64))
        jax.block_until_ready((self.cells, self.points))
        slip = onp.ascontiguousarray(slip, dtype=onp.float64)
        self.nc, self.nn, self.ns = int(self.cells.shape[0]), int(self.points.shape[0]), int(slip.shape[0])
        h = ctypes.c_void_p()
        _lib.check(L.cpfem_plan_create(ctypes.c_void_p(self.cells.unsafe_buffer_pointer()), self.nc,
                                       ctypes.c_void_p(self.points.unsafe_buffer_pointer()), self.nn,
                                       slip.ctypes.data_as(ctypes.c_void_p), self.ns, None, ctypes.byref(h)), 'cpfem_plan_create')
        self._h = h
        info = (ctypes.c_int64 * 6)()
        _lib.check(L.cpfem_plan_info(self._h, info), 'cpfem_plan_info')
        self.nnz = int(info[3])
        self.ndof = 3 * self.nn
        self.indptr = onp.empty(self.ndof + 1, dtype=onp.int64)
        self.indices = onp.empty(self.nnz, dtype=onp.int32)
        _lib.check(L.cpfem_plan_csr_copy(self._h, self.indptr.ctypes.data_as(ctypes.c_void_p),
                                         self.indices.ctypes.data_as(ctypes.c_void_p), None), 'cpfem_plan_csr_copy')
        import ctypes.util                                        # the copies above ran on the legacy stream: wait for them
        ctypes.CDLL(ctypes.util.find_library('cudart') or 'libcudart.so').cudaDeviceSynchronize()

    @property
    def handle(self):
        return onp.int64(self._h.value)

    def __del__(self):
        try:
            if getattr(self, '_h', None):
                _lib.lib().cpfem_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass
