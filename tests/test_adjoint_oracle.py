"""Adjoint row on the CPU: the product's dual-number header (csrc/cp_adjoint.cuh, compiled for the host by
tests/hostcheck) against the oracle's autodiff of the restated implicit_residual / first_PK_stress.

Reference: f_jvp's jac_x / jac_y (singlecrystal_copper/models_copper.py:251-259; DP form polycrystal_DPsteel/
models_DPsteel_inhomo.py:245,352-361) and their reverse mode inside implicit_vjp (crystal_plasticity_OR_design/solver.py:801-853).
Tolerance: 1e-10 of each block's largest entry (north_star's fp64 bar)."""
import ctypes

import numpy as np
import pytest
import torch

import cases
import cpfem_oracle as O
import hostcheck_build


@pytest.fixture(scope='module')
def hostcheck():
    return hostcheck_build.load()


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def host_jac(lib, mat, dt, H, A, g, R, S, pp, with_params, with_C):
    n, ns = len(H), g.shape[1]
    nx = 27 + 2 * ns + (5 if with_params else 0) + (81 if with_C else 0)
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64).reshape(n, -1)
    H, A, g, R, S, pp = map(c, (H, A, g, R, S, pp))
    jx, jy, dPx, dPs = np.zeros((n, 9, nx)), np.zeros((n, 9, 9)), np.zeros((n, 9, nx)), np.zeros((n, 9, 9))
    slip = np.ascontiguousarray(mat.slip, dtype=np.float64)
    rc = lib.hostcheck_jac(ctypes.c_int(ns), _p(slip), ctypes.c_double(mat.ao * dt), ctypes.c_int64(n), _p(H), _p(A), _p(g), _p(R), _p(pp),
                           _p(S), ctypes.c_int(with_params), ctypes.c_int(with_C), _p(jx), _p(jy), _p(dPx), _p(dPs))
    assert rc == 0
    return jx, jy, dPx, dPs


def host_vjp(lib, mat, dt, H, A, g, R, S, pp, W, with_params):
    n, ns = len(H), g.shape[1]
    nd = 27 + 2 * ns + (5 if with_params else 0)
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64).reshape(n, -1)
    H, A, g, R, S, pp, W = map(c, (H, A, g, R, S, pp, W))
    grad = np.zeros((n, nd))
    slip = np.ascontiguousarray(mat.slip, dtype=np.float64)
    rc = lib.hostcheck_vjp(ctypes.c_int(ns), _p(slip), ctypes.c_double(mat.ao * dt), ctypes.c_int64(n), _p(H), _p(A), _p(g), _p(R), _p(pp),
                           _p(S), _p(W), ctypes.c_int(with_params), _p(grad))
    assert rc == 0
    return grad


def blocks(ns, nextra):
    """(name, slice) of x's blocks in the reference's ravel order."""
    b = [('u_grad', slice(0, 9)), ('Fp_inv', slice(9, 18)), ('g', slice(18, 18 + ns)), ('slip', slice(18 + ns, 18 + 2 * ns)),
         ('rot', slice(18 + 2 * ns, 27 + 2 * ns))]
    o = 27 + 2 * ns
    if nextra >= 5:
        b += [(k, slice(o + i, o + i + 1)) for i, k in enumerate(('gss_a', 'h', 't_sat', 'xm', 'r'))]
        o += 5
    if nextra >= 6:
        b.append(('C', slice(o, o + 81)))
    return b


def assert_blocks(got, want, ns, nextra, tol=1e-10, what=''):
    """Every block of columns to `tol` of the block's largest entry (blocks differ by orders of magnitude: d/dg ~ n/g,
    d/dC ~ strain); identically-zero blocks must be exactly zero."""
    for name, sl in blocks(ns, nextra):
        a, b = got[..., sl], want[..., sl]
        scale = np.abs(b).max()
        if scale == 0.0:
            assert np.abs(a).max() == 0.0, (what, name)
        else:
            assert np.abs(a - b).max() <= tol * scale, (what, name, np.abs(a - b).max() / scale)


@pytest.mark.parametrize('name', ['copper', '304steel', 'tantalum', 'dp_ferrite'])
def test_jac_x_jac_y_vs_oracle_autodiff(name, hostcheck):
    """All 51 (75 for BCC24) columns of jac_x, jac_y, and the explicit dP/dx, dP/dy at the converged S."""
    for step, mat, dt, H, A, g, sl, R in cases.point_history(name, n=12, steps=6, seed=3):
        if step not in (1, 4, 6):
            continue
        pb = O.PointBatch(A, g, sl, R, mat)
        y = pb.newton_solver(H, dt)
        ns = g.shape[1]
        pp = np.tile([mat.C11, mat.C12, mat.C44, mat.xm], (len(H), 1))
        jx, jy, dPx, dPs = host_jac(hostcheck, mat, dt, H, A, g, R, y.numpy(), pp, 0, 0)
        Hh = torch.as_tensor(H)
        assert_blocks(jx, pb.jac_x(Hh, y, dt).numpy(), ns, 0, what=f'{name} step {step} jac_x')
        jy_o = pb.jac_y(Hh, y, dt).numpy()
        assert np.abs(jy - jy_o).max() <= 1e-10 * np.abs(jy_o).max()
        # total derivative: explicit part + dP/dS . dS/dx
        tot = dPx + dPs @ np.linalg.solve(jy, -jx)
        assert_blocks(tot, pb.dP_dx(Hh, dt, y).numpy(), ns, 0, tol=1e-9, what=f'{name} step {step} dP/dx')


def test_jac_x_dp_form_all_161_columns(hostcheck):
    """DP-steel form: x carries the five per-point parameters and the 81 entries of C (models_DPsteel_inhomo.py:245): 161
    columns for BCC24.  Only xm and C enter the residual; gss_a, h, t_sat, r give exact zeros."""
    n = 10
    params, ph, quat, ori = cases.dp_params(n, seed=5)
    rng = np.random.default_rng(7)
    sel = lambda a: a[:, 0]
    A, g, sl, R, ga, h, ts, xm, r, C = [sel(a) for a in params]
    f = O.dp_ferrite()
    pb = O.PointBatch(A, g, sl, R, gss_a=ga, h=h, t_sat=ts, xm=xm, r=r, C=C, slip_table=O.SLIP_BCC24, ao=f.ao, max_sub_step=f.max_sub_step,
                      tol=f.tol)
    dt = 0.2
    H = np.zeros((n, 3, 3))
    H[:, 2, 2] = 4e-3
    H[:, 0, 0] = H[:, 1, 1] = -1.2e-3
    H += rng.uniform(-1, 1, size=H.shape) * 2e-4
    # advance two steps so that Fp_inv != I and g != g0
    An, gn, sn = pb.update_int_vars(H, dt)
    pb = O.PointBatch(An.numpy(), gn.numpy(), sn.numpy(), R, gss_a=ga, h=h, t_sat=ts, xm=xm, r=r, C=C, slip_table=O.SLIP_BCC24, ao=f.ao,
                      max_sub_step=f.max_sub_step, tol=f.tol)
    H = 1.5 * H
    y = pb.newton_solver(H, dt)
    pp = np.stack([C[:, 0, 0, 0, 0], C[:, 0, 0, 1, 1], C[:, 0, 1, 0, 1], xm], axis=1)
    jx, jy, dPx, dPs = host_jac(hostcheck, f, dt, H, An.numpy(), gn.numpy(), R, y.numpy(), pp, 1, 1)
    assert jx.shape[2] == 161
    Hh = torch.as_tensor(H)
    assert_blocks(jx, pb.jac_x(Hh, y, dt, nextra=6).numpy(), 24, 6, what='dp jac_x')
    tot = dPx + dPs @ np.linalg.solve(jy, -jx)
    assert_blocks(tot, pb.dP_dx(Hh, dt, y, nextra=6).numpy(), 24, 6, tol=1e-9, what='dp dP/dx')


@pytest.mark.parametrize('name', ['copper', '304steel'])
def test_point_vjp_vs_oracle(name, hostcheck):
    """cp_point_vjp (one transposed 9x9 solve per point instead of nx forward solves) = W : dP/dx of the oracle."""
    rng = np.random.default_rng(11)
    for step, mat, dt, H, A, g, sl, R in cases.point_history(name, n=10, steps=5, seed=4):
        if step != 5:
            continue
        pb = O.PointBatch(A, g, sl, R, mat)
        y = pb.newton_solver(H, dt)
        W = rng.normal(size=(len(H), 9))
        pp = np.tile([mat.C11, mat.C12, mat.C44, mat.xm], (len(H), 1))
        got = host_vjp(hostcheck, mat, dt, H, A, g, R, y.numpy(), pp, W, 1)
        want = np.einsum('pi,pic->pc', W, pb.dP_dx(torch.as_tensor(H), dt, y, nextra=5).numpy())
        assert_blocks(got, want, g.shape[1], 5, tol=1e-9, what=f'{name} vjp')


def test_oracle_dP_dx_vs_central_differences():
    """The oracle's own total derivative against central differences of its forward map (independent of autodiff)."""
    for step, mat, dt, H, A, g, sl, R in cases.point_history('copper', n=4, steps=4, seed=9):
        if step != 4:
            continue
        pb = O.PointBatch(A, g, sl, R, mat)
        D = pb.dP_dx(torch.as_tensor(H), dt).numpy()
        P = lambda A_, g_, R_: O.PointBatch(A_, g_, sl, R_, dataclass_tol(mat)).first_PK_stress(H, dt).numpy().reshape(-1, 9)
        for blk, arr, off in (('A', A, 9), ('g', g, 18), ('R', R, 18 + 2 * g.shape[1])):
            flat = arr.reshape(len(H), -1)
            for c in (0, flat.shape[1] // 2, flat.shape[1] - 1):
                e = 1e-6 * max(abs(flat[:, c]).max(), 1.0)
                up, dn = flat.copy(), flat.copy()
                up[:, c] += e
                dn[:, c] -= e
                args = lambda v: (v.reshape(arr.shape) if blk == 'A' else A, v.reshape(arr.shape) if blk == 'g' else g,
                                  v.reshape(arr.shape) if blk == 'R' else R)
                fd = (P(*args(up)) - P(*args(dn))) / (2 * e)
                assert np.abs(fd - D[:, :, off + c]).max() <= 2e-4 * max(np.abs(D[:, :, off + c]).max(), 1e-30), (blk, c)


def dataclass_tol(mat):
    """Same material with a tighter local tolerance, so that finite differences see a smooth map."""
    import dataclasses
    return dataclasses.replace(mat, tol=5e-10)
