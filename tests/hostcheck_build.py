"""Builds/loads tests/hostcheck/libhostcheck.so: the product's per-point header compiled for the host (g++).
Test infrastructure only - lets the CPU suite compare the hand-derived algebra with the oracle without a GPU."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'hostcheck')
SO = os.path.join(HERE, 'libhostcheck.so')
SRC = os.path.join(HERE, 'hostcheck.cpp')
HDR = os.path.join(os.path.dirname(HERE), '..', 'jax-cpfem_b200', 'csrc', 'cp_point.cuh')


class CMat(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in ['C11', 'C12', 'C44', 'h', 't_sat', 'gss_a', 'ao', 'xm', 'r', 'tol']] + \
               [('max_sub_step', ctypes.c_int32), ('max_iter', ctypes.c_int32)]


def load():
    if (not os.path.exists(SO)) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(['g++', '-O2', '-fopenmp', '-std=c++17', '-shared', '-fPIC', '-x', 'c++', SRC, '-o', SO])
    return ctypes.CDLL(SO)


def set_threads(lib, n):
    """OpenMP threads of the host build (0 = leave as is); returns the number in use."""
    lib.hostcheck_threads.restype = ctypes.c_int
    return int(lib.hostcheck_threads(ctypes.c_int(int(n))))


def evaluate(lib, mat, dt, H, A, g, sl, R, pp=None, tangent=True, pown=0):
    """mat: oracle Material. Returns P, T, A_new, g_new, slip_new, info(iters, evals, status)."""
    n = len(H)
    ns = g.shape[1]
    H, A, g, sl, R = [np.ascontiguousarray(x, dtype=np.float64).reshape(n, -1) for x in (H, A, g, sl, R)]
    P = np.zeros((n, 9)); T = np.zeros((n, 81)); An = np.zeros((n, 9)); gn = np.zeros((n, ns)); sn = np.zeros((n, ns))
    it = np.zeros((n, 3), np.int32)
    slip = np.ascontiguousarray(mat.slip, dtype=np.float64)
    m = CMat(mat.C11, mat.C12, mat.C44, mat.h, mat.t_sat, mat.gss_a, mat.ao, mat.xm, mat.r, mat.tol, mat.max_sub_step, 200)
    p_ = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    if pp is not None:
        pp = np.ascontiguousarray(pp, dtype=np.float64)
    rc = lib.hostcheck_points(ctypes.c_int(ns), ctypes.c_int(pown), p_(slip), ctypes.byref(m), ctypes.c_double(dt), ctypes.c_int64(n), p_(H), p_(A),
                              p_(g), p_(sl), p_(R), p_(pp), p_(P), p_(T) if tangent else None, p_(An), p_(gn), p_(sn), p_(it))
    assert rc == 0
    return P.reshape(n, 3, 3), T.reshape(n, 3, 3, 3, 3), An.reshape(n, 3, 3), gn, sn, it
