"""The experimental CP_PRUNE / CP_PEEL code shapes of the local Newton solve (csrc/cp_point.cuh: line-search trials whose rejection is
certain from tau/g alone are not evaluated) must give bitwise the results of the default path - stresses, tangents, new
state, iteration AND evaluation counts (models_copper.py:204-249 is followed literally either way) - and must actually
skip evaluations where the rate exponent makes rejected trials overshoot (304 steel: n = 120, tantalum: n = 45.3).
CPU only: the header compiled for the host, single-threaded."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import cases
import cpfem_oracle as O
import hostcheck_build as hb

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'hostcheck')


@pytest.fixture(scope='module')
def libs(tmp_path_factory):
    d = tmp_path_factory.mktemp('prune')
    out = []
    for k, flag in enumerate(('-DCP_PRUNE=1', '-DCP_PRUNE=0', '-DCP_PEEL=2')):
        so = str(d / ('h%d.so' % k))
        subprocess.check_call(['g++', '-O2', '-std=c++17', flag, '-shared', '-fPIC', os.path.join(HERE, 'prune_hook.cpp'), '-o', so])
        L = ctypes.CDLL(so)
        L.hostcheck_pruned_count.restype = ctypes.c_longlong
        out.append(L)
    return out


@pytest.mark.parametrize('name,pown,expect', [('304steel', 119, True), ('tantalum', 0, True), ('copper', 9, False), ('dp_ferrite', 19, False)])
def test_pruned_solve_is_bitwise_identical(libs, name, pown, expect):
    on, off, peel = libs
    fac, deps, dt = cases.MATERIALS[name]
    mat = fac()
    rng = np.random.default_rng(0)
    n, ns = 600, len(mat.slip)
    R = O.get_rot_mat(cases.rand_quat(rng, n))
    A = np.tile(np.eye(3), (n, 1, 1))
    g = mat.gss_initial * np.ones((n, ns))
    sl = np.zeros((n, ns))
    pruned = peeled = evals = 0
    for step in range(1, 13):
        eps = deps * step
        H = np.zeros((n, 3, 3))
        H[:, 2, 2] = eps
        H[:, 0, 0] = H[:, 1, 1] = -0.3 * eps
        H += rng.uniform(-1, 1, size=H.shape) * abs(deps) * 0.1
        a = hb.evaluate(on, mat, dt, H, A, g, sl, R, tangent=True, pown=pown)
        b = hb.evaluate(off, mat, dt, H, A, g, sl, R, tangent=True, pown=pown)
        c = hb.evaluate(peel, mat, dt, H, A, g, sl, R, tangent=True, pown=pown)      # CP_PEEL=2: first evaluation peeled,
        for x, y, z in zip(a, b, c):                                                  # first line search decided up front
            assert np.array_equal(x, y) and np.array_equal(z, y), (name, step)
        peeled += peel.hostcheck_pruned_count()
        pruned += on.hostcheck_pruned_count()
        assert off.hostcheck_pruned_count() == 0
        evals += int(a[5][:, 1].sum())
        A, g, sl = a[2], a[3], a[4]
    print(f'{name}: {pruned} of {evals} residual evaluations skipped (CP_PRUNE), {peeled} (CP_PEEL=2)')
    assert (pruned > 0.05 * evals) if expect else (pruned >= 0)
    assert (peeled > 0.02 * evals) if expect else (peeled >= 0)
