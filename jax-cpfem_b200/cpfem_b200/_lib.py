"""ctypes binding of libcpfem_b200.so (the C ABI declared in include/cpfem.h).

The library is built in-tree by `build()` (nvcc, sm_100a).  There is no CPU fallback: if the shared
library is missing or cannot be loaded, every entry point raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(os.path.dirname(_HERE), 'csrc')
_INCLUDE = os.path.join(os.path.dirname(os.path.dirname(_HERE)), 'include')
# CPFEM_B200_LIB: load another build of the same library (kernel-tuning experiments); still CUDA-only
LIB_PATH = os.environ.get('CPFEM_B200_LIB') or os.path.join(_HERE, 'libcpfem_b200.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC']

c_i64, c_i32, c_dbl, c_vp = ctypes.c_int64, ctypes.c_int32, ctypes.c_double, ctypes.c_void_p


class Material(ctypes.Structure):
    """cpfem_material (include/cpfem.h)."""
    _fields_ = [('C11', c_dbl), ('C12', c_dbl), ('C44', c_dbl), ('h', c_dbl), ('t_sat', c_dbl), ('gss_a', c_dbl),
                ('ao', c_dbl), ('xm', c_dbl), ('r', c_dbl), ('tol', c_dbl), ('max_sub_step', c_i32), ('max_iter', c_i32)]


class State(ctypes.Structure):
    """cpfem_state."""
    _fields_ = [(n, c_vp) for n in ('Fp_inv', 'g', 'slip', 'rot', 'gss_a', 'h', 't_sat', 'xm', 'r', 'C')] + [('layout', c_i32)]


class StateGrad(ctypes.Structure):
    """cpfem_state_grad."""
    _fields_ = [(n, c_vp) for n in ('Fp_inv', 'g', 'slip', 'rot', 'gss_a', 'h', 't_sat', 'xm', 'r', 'C')]


class StateOut(ctypes.Structure):
    """cpfem_state_out."""
    _fields_ = [('Fp_inv', c_vp), ('g', c_vp), ('slip', c_vp), ('layout', c_i32)]


# name -> (restype, argtypes); kept in one table so tests can check every symbol of include/cpfem.h is exported
SIGNATURES = {
    'cpfem_plan_create': (ctypes.c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_i32, c_vp, ctypes.POINTER(c_vp)]),
    'cpfem_plan_destroy': (ctypes.c_int, [c_vp]),
    'cpfem_plan_csr': (ctypes.c_int, [c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), ctypes.POINTER(c_i64)]),
    'cpfem_plan_csr_copy': (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp]),
    'cpfem_plan_set_active_cells': (ctypes.c_int, [c_vp, c_i64]),
    'cpfem_plan_set_progress_event': (ctypes.c_int, [c_vp, c_i64, c_vp]),
    'cpfem_plan_info': (ctypes.c_int, [c_vp, ctypes.POINTER(c_i64)]),
    'cpfem_update_state': (ctypes.c_int, [c_vp, ctypes.POINTER(Material), c_vp, ctypes.POINTER(State),
                                          ctypes.POINTER(StateOut), c_dbl, c_vp, c_vp]),
    'cpfem_update_state_cells': (ctypes.c_int, [c_vp, ctypes.POINTER(Material), c_vp, ctypes.POINTER(State),
                                                ctypes.POINTER(StateOut), c_dbl, c_i64, c_i64, c_vp, c_vp]),
    'cpfem_update_state_avg_stress': (ctypes.c_int, [c_vp, ctypes.POINTER(Material), c_vp, ctypes.POINTER(State),
                                                     ctypes.POINTER(StateOut), c_dbl, c_vp, c_vp, c_vp]),
    'cpfem_residual': (ctypes.c_int, [c_vp, ctypes.POINTER(Material), c_vp, ctypes.POINTER(State), c_dbl, c_vp, c_vp, c_vp]),
    'cpfem_newton_update': (ctypes.c_int, [c_vp, ctypes.POINTER(Material), c_vp, ctypes.POINTER(State), c_dbl,
                                           c_vp, c_vp, c_vp, c_vp, c_vp]),
    'cpfem_avg_stress': (ctypes.c_int, [c_vp, ctypes.POINTER(Material), c_vp, ctypes.POINTER(State), c_dbl, c_vp, c_vp, c_vp]),
    'cpfem_point_stress_tangent': (ctypes.c_int, [c_vp, ctypes.POINTER(Material), c_vp, c_i64, ctypes.POINTER(State),
                                                  c_dbl, c_vp, c_vp, c_vp, c_vp]),
    'cpfem_point_update_state': (ctypes.c_int, [c_vp, ctypes.POINTER(Material), c_vp, c_i64, ctypes.POINTER(State),
                                                ctypes.POINTER(StateOut), c_dbl, c_vp, c_vp]),
    'cpfem_point_eval': (ctypes.c_int, [c_vp, ctypes.POINTER(Material), c_vp, c_i64, ctypes.POINTER(State), c_dbl, c_vp, c_vp,
                                        ctypes.POINTER(StateOut), c_vp, c_vp, c_vp]),
    'cpfem_point_jac_x': (ctypes.c_int, [c_vp, ctypes.POINTER(Material), c_vp, c_i64, ctypes.POINTER(State), c_dbl, c_i32, c_vp, c_vp,
                                         c_vp, c_vp, c_vp]),
    'cpfem_point_vjp': (ctypes.c_int, [c_vp, ctypes.POINTER(Material), c_vp, c_i64, ctypes.POINTER(State), c_dbl, c_i32, c_vp, c_vp,
                                       c_vp, c_vp]),
    'cpfem_vjp_params': (ctypes.c_int, [c_vp, ctypes.POINTER(Material), c_vp, ctypes.POINTER(State), c_dbl, c_vp,
                                        ctypes.POINTER(StateGrad), c_vp, c_vp]),
    'cpfem_check_cubic': (ctypes.c_int, [c_vp, c_i64, c_dbl, c_vp, c_vp]),
    'cpfem_apply_dirichlet': (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    'cpfem_spmv': (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    'cpfem_csr_diagonal': (ctypes.c_int, [c_vp, c_vp, c_vp, c_i32, c_vp]),
    'cpfem_csr_transpose': (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp]),
    'cpfem_bicgstab': (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i32, c_dbl, c_dbl, c_i64, ctypes.POINTER(c_i64),
                                      ctypes.POINTER(c_dbl), c_vp]),
    'cpfem_bicgstab_enqueue': (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i32, c_dbl, c_dbl, c_i64, c_i64, c_vp, c_vp, c_vp]),
    'cpfem_scatter_add': (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    'cpfem_gather': (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    'cpfem_sumsq': (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp]),
    'cpfem_peer_alloc': (ctypes.c_int, [c_i64, ctypes.POINTER(c_vp), c_vp]),
    'cpfem_peer_open': (ctypes.c_int, [c_vp, ctypes.POINTER(c_vp)]),
    'cpfem_peer_close': (ctypes.c_int, [c_vp]),
    'cpfem_peer_free': (ctypes.c_int, [c_vp]),
    'cpfem_peer_put': (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, ctypes.c_uint64, c_vp]),
    'cpfem_peer_signal': (ctypes.c_int, [c_vp, ctypes.c_uint64, c_vp]),
    'cpfem_peer_wait': (ctypes.c_int, [c_vp, ctypes.c_uint64, c_dbl, c_vp, c_vp]),
    'cpfem_aos_to_soa': (ctypes.c_int, [c_vp, c_i64, c_i32, c_vp, c_vp]),
    'cpfem_soa_to_aos': (ctypes.c_int, [c_vp, c_i64, c_i32, c_vp, c_vp]),
    'cpfem_dfma_peak_kernel': (ctypes.c_int, [c_i64, c_vp, ctypes.POINTER(c_dbl), c_vp]),
    'cpfem_launch_count': (c_i64, []),
    'cpfem_last_error': (ctypes.c_char_p, []),
    'cpfem_version': (ctypes.c_int, []),
}

_lib = None
_lock = threading.Lock()


def sources():
    return [os.path.join(_CSRC, 'cpfem_kernels.cu'), os.path.join(_CSRC, 'cpfem_solver.cu'), os.path.join(_CSRC, 'cpfem_peer.cu')]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(_CSRC, 'cp_point.cuh'), os.path.join(_CSRC, 'cp_adjoint.cuh'), os.path.join(_CSRC, 'cpfem_internal.h'), os.path.join(_INCLUDE, 'cpfem.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = out or LIB_PATH
    if not force and out == LIB_PATH and not needs_build():
        return LIB_PATH
    cmd = ['nvcc'] + NVCC_FLAGS + list(extra_flags) + (['-Xptxas', '-v'] if verbose else []) + ['-I', _INCLUDE, '-o', out] + sources()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return out


def lib():
    """Load the shared library (once).  Raises if it has not been built - there is no fallback."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(f'{LIB_PATH} not found: build it with __graft_entry__.build() / cpfem_b200.build(); '
                                   'there is no CPU fallback for the CUDA path')
            L = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(L, name)
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


class CpfemError(RuntimeError):
    pass


def check(rc, who=''):
    if rc != 0:
        msg = lib().cpfem_last_error()
        raise CpfemError(f'{who} failed ({rc}): {msg.decode() if msg else ""}')
