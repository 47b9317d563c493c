"""CPU oracle for the JAX-CPFEM hot path.  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (torch fp64 + torch.func forward-mode autodiff, which plays the
role jax.jacfwd plays in the reference) of the per-quadrature-point Kalidindi crystal-plasticity
update and of the hex8 residual / tangent assembly that feeds on it.  Nothing in the product path
(`jax-cpfem_b200/`) imports it: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may.

Every function cites the reference lines it follows (paths relative to the JAX-CPFEM tree):

  * per-point physics:  singlecrystal_copper/models_copper.py:21-48, 135-271
                        polycrystal_DPsteel/models_DPsteel_inhomo.py:240-361 (per-point parameters)
  * field operations:   singlecrystal_copper/models_copper.py:273-282, 297-319
  * Jacobian-to-CSR:    crystal_plasticity_OR_design/solver.py:279-293
  * Dirichlet rows:     crystal_plasticity_OR_design/solver.py:119-133, 213-237, 387-414
  * FE layer (cell kernel, jacfwd over the 24 cell dofs, I/J rule, scatter-add, hex8 basis and
    2x2x2 Gauss rule): third-party `jax_fem` (deepmodeling/jax-fem, unpinned, ~v0.0.5-0.0.8), which
    the reference imports (`models_copper.py:9`) but does not vendor.  Its published algorithm is
    restated here; parity is anchored on the reference's call sites (`models_copper.py:84,277,315`,
    `solver.py:244,281,392`) and on the golden curves / VTU files the reference commits.

Pinning status (see tests/test_oracle_golden.py, tests/golden/):
  * stress / state / residual path: pinned against the reference's committed golden stress-strain
    curves (`calibration/data/csv/calibration_case{1,2}/...`) and committed VTU outputs.
  * tangent (V, CSR data): "parity unpinned" by any reference artefact - the reference commits no
    Jacobian values.  Here the tangent is produced by forward-mode autodiff of the restated
    residual (the same construction the reference uses), and cross-checked by central differences.

Differences from the reference that are deliberate and do not change results:
  * `abs(x)**n * sign(x)` is guarded at x == 0 so that torch's batched forward-mode rule returns
    the same 0 that JAX returns there (torch would produce NaN from 0*inf).
  * Forward assembly only ever pushes tangents through `u_grad`; the 42 (152) other columns of
    `jac_x` (`models_copper.py:256`) multiply zero tangents and are not formed there.  The full `jac_x`
    (all 51 / 56 / 161 columns) and what reverse mode makes of f_jvp inside `implicit_vjp`
    (`crystal_plasticity_OR_design/solver.py:801-853`) are restated by `PointBatch.jac_x`, `PointBatch.dP_dx`
    and `FEOracle.vjp_params` for the adjoint row.
"""
from __future__ import annotations

import dataclasses
import os
from typing import Optional, Sequence

import numpy as onp
import scipy.sparse
import torch
from torch.func import jacfwd, vmap

torch.set_default_dtype(torch.float64)

DIM = 3

# --------------------------------------------------------------------------------------------
# Slip-system tables.  Rows are "normal(3) direction(3)", un-normalised, exactly as in
# */data/csv/input_slip_sys*.txt (e.g. singlecrystal_copper/data/csv/input_slip_sys.txt).
# --------------------------------------------------------------------------------------------
SLIP_FCC12 = onp.array([
    [1, 1, -1, 0, 1, 1], [1, 1, -1, 1, 0, 1], [1, 1, -1, 1, -1, 0],
    [1, -1, -1, 0, 1, -1], [1, -1, -1, 1, 0, 1], [1, -1, -1, 1, 1, 0],
    [1, -1, 1, 0, 1, 1], [1, -1, 1, 1, 0, -1], [1, -1, 1, 1, 1, 0],
    [1, 1, 1, 0, 1, -1], [1, 1, 1, 1, 0, -1], [1, 1, 1, 1, -1, 0]], dtype=onp.float64)

SLIP_BCC12 = onp.array([
    [1, 1, 0, -1, 1, 1], [1, 1, 0, 1, -1, 1], [1, -1, 0, 1, 1, 1], [1, -1, 0, 1, 1, -1],
    [1, 0, 1, 1, 1, -1], [1, 0, 1, -1, 1, 1], [1, 0, -1, 1, 1, 1], [1, 0, -1, 1, -1, 1],
    [0, 1, 1, 1, 1, -1], [0, 1, 1, 1, -1, 1], [0, 1, -1, 1, 1, 1], [0, 1, -1, -1, 1, 1]],
    dtype=onp.float64)

SLIP_BCC24 = onp.concatenate([SLIP_BCC12, onp.array([
    [1, 1, 2, 1, 1, -1], [-1, 1, 2, 1, -1, 1], [1, -1, 2, -1, 1, 1], [1, 1, -2, 1, 1, 1],
    [1, 2, 1, 1, -1, 1], [-1, 2, 1, 1, 1, -1], [1, -2, 1, 1, 1, 1], [1, 2, -1, -1, 1, 1],
    [2, 1, 1, -1, 1, 1], [-2, 1, 1, 1, 1, 1], [2, -1, 1, 1, 1, -1], [2, 1, -1, 1, -1, 1]],
    dtype=onp.float64)])


@dataclasses.dataclass
class Material:
    """Parameter set of one `models_*.py` file."""
    name: str
    slip: onp.ndarray          # (ns, 6) normal | direction
    gss_initial: float
    h: float
    t_sat: float
    gss_a: float
    xm: float
    C11: float
    C12: float
    C44: float
    r: float = 1.0
    ao: float = 0.001
    max_sub_step: int = 5
    tol: float = 1e-8


def copper():      # singlecrystal_copper/models_copper.py:54-56,94-96,141-149,231
    return Material('copper', SLIP_FCC12, 60.8, 541.5, 109.8, 2.5, 0.1, 1.684e5, 1.214e5, 0.754e5)


def tantalum():    # singlecrystal_tantalum/models_tantalum.py:56,59,96-98,143-151,234
    return Material('tantalum', SLIP_BCC12, 67.4641, 1959.1320, 7295.1754, 200.0, 1.0 / 45.2726,
                    2.670e5, 1.610e5, 0.825e5)


def steel304():    # polycrystal_304steel/models_304steel.py:56,95-97,143-151,232
    return Material('304steel', SLIP_FCC12, 90.0, 392.9772, 7295.1754, 8.0, 1.0 / 120.0,
                    2.622e5, 1.120e5, 0.746e5, max_sub_step=8)


def dp_ferrite():  # polycrystal_DPsteel/models_DPsteel_inhomo.py:73-86 (phase 0), :321
    return Material('dp_ferrite', SLIP_BCC24, 170.0, 400.0, 2500.0, 4.0, 0.05, 2.314e5, 1.347e5, 1.164e5)


def dp_martensite():  # polycrystal_DPsteel/models_DPsteel_inhomo.py:89-102 (phase 1)
    return Material('dp_martensite', SLIP_BCC24, 435.0, 950.0, 5300.0, 4.0, 0.05, 4.174e5, 2.424e5, 2.111e5)


# --------------------------------------------------------------------------------------------
# models_copper.py:21-48
# --------------------------------------------------------------------------------------------
def rotate_tensor_rank_4(R, T):
    """models_copper.py:21-26: out_ijkl = R_ia R_jb R_kc R_ld T_abcd."""
    return torch.einsum('ia,jb,kc,ld,abcd->ijkl', R, R, R, R, T)


def rotate_tensor_rank_2(R, T):
    """models_copper.py:29-32: out_ij = R_ia R_jb T_ab."""
    return torch.einsum('ia,jb,ab->ij', R, R, T)


def get_rot_mat(q):
    """models_copper.py:37-45, quaternion (w,x,y,z) -> rotation matrix.  numpy, batched on axis 0."""
    q = onp.asarray(q, dtype=onp.float64)
    q0, q1, q2, q3 = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    return onp.stack([
        onp.stack([q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, 2 * q1 * q2 - 2 * q0 * q3, 2 * q1 * q3 + 2 * q0 * q2], -1),
        onp.stack([2 * q1 * q2 + 2 * q0 * q3, q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, 2 * q2 * q3 - 2 * q0 * q1], -1),
        onp.stack([2 * q1 * q3 - 2 * q0 * q2, 2 * q2 * q3 + 2 * q0 * q1, q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3], -1)], -2)


def schmid_tensors(slip):
    """models_copper.py:62-69: normalise, then outer(direction, normal) -> (ns,3,3)."""
    slip = onp.asarray(slip, dtype=onp.float64)
    d = slip[:, DIM:]
    d = d / onp.linalg.norm(d, axis=1)[:, None]
    n = slip[:, :DIM]
    n = n / onp.linalg.norm(n, axis=1)[:, None]
    return onp.einsum('ai,aj->aij', d, n)


def cubic_C(C11, C12, C44):
    """models_copper.py:92-130."""
    C = onp.zeros((3, 3, 3, 3))
    for i in range(3):
        C[i, i, i, i] = C11
    for i in range(3):
        for j in range(3):
            if i != j:
                C[i, i, j, j] = C12
                C[i, j, i, j] = C44
                C[i, j, j, i] = C44
    return C


def latent_q(ns, r):
    """models_copper.py:71-76 / models_DPsteel_inhomo.py:261-266: r everywhere, 1 on 'coplanar' triples."""
    q = r * torch.ones((ns, ns))
    mask = torch.zeros((ns, ns), dtype=torch.bool)
    for i in range(ns):
        for j in range(3):
            mask[i, i // 3 * 3 + j] = True
    return torch.where(mask, torch.ones(()), q)


# --------------------------------------------------------------------------------------------
# Per-point physics (one quadrature point; batched with vmap)
# --------------------------------------------------------------------------------------------
def _signed_pow(x, n):
    """abs(x)**n * sign(x) with JAX's value/derivative (0) at x == 0 (see module docstring)."""
    zero = x == 0
    ax = torch.where(zero, torch.ones_like(x), torch.abs(x))
    return torch.where(zero, torch.zeros_like(x), ax ** n * torch.sign(x))


def helper(u_grad, Fp_inv_old, g_old, slip_old, rot_mat, S, gss_a, h, t_sat, xm, r, schmid, dt, ao):
    """models_copper.py:172-192 (DP form: models_DPsteel_inhomo.py:260-284)."""
    ns = schmid.shape[0]
    q = latent_q(ns, r)
    M = torch.einsum('ia,jb,sab->sij', rot_mat, rot_mat, schmid)            # rotate_tensor_rank_2_vmap
    tau = torch.sum(S[None, :, :] * M, dim=(1, 2))
    gamma_inc = ao * dt * _signed_pow(tau / g_old, 1. / xm)
    tmp = h * torch.abs(gamma_inc) * _signed_pow(1 - g_old / t_sat, gss_a)
    # reference: abs(1-g/t)**a * sign(1-g/t)  == _signed_pow(1-g/t, a)
    g_inc = (q @ tmp[:, None]).reshape(-1)
    g_new = g_old + g_inc
    slip_new = slip_old + gamma_inc
    F = u_grad + torch.eye(DIM)
    L_plastic_inc = torch.sum(gamma_inc[:, None, None] * M, dim=0)
    Fp_inv_new = Fp_inv_old @ (torch.eye(DIM) - L_plastic_inc)
    Fe = F @ Fp_inv_new
    return Fp_inv_new, g_new, slip_new, Fe, F


def implicit_residual(u_grad, Fp_inv_old, g_old, slip_old, rot_mat, y, gss_a, h, t_sat, xm, r, C, schmid, dt, ao):
    """models_copper.py:195-201: res = ravel(S - rot4(R,C) : 1/2 (Fe^T Fe - I))."""
    S = y.reshape(DIM, DIM)
    _, _, _, Fe, _ = helper(u_grad, Fp_inv_old, g_old, slip_old, rot_mat, S, gss_a, h, t_sat, xm, r, schmid, dt, ao)
    E = 0.5 * (Fe.T @ Fe - torch.eye(DIM))
    S_ = torch.sum(rotate_tensor_rank_4(rot_mat, C) * E[None, None, :, :], dim=(2, 3))
    return (S - S_).reshape(-1)


def det3(M):
    """Explicit 3x3 determinant (torch.linalg.det's forward-mode rule goes through an LU and returns NaN
    under vmap for some exactly-diagonal inputs; the cofactor expansion has no such problem)."""
    return (M[0, 0] * (M[1, 1] * M[2, 2] - M[1, 2] * M[2, 1])
            - M[0, 1] * (M[1, 0] * M[2, 2] - M[1, 2] * M[2, 0])
            + M[0, 2] * (M[1, 0] * M[2, 1] - M[1, 1] * M[2, 0]))


def inv3(M):
    """Explicit 3x3 inverse (adjugate / determinant)."""
    c = lambda i, j: (M[(i + 1) % 3, (j + 1) % 3] * M[(i + 2) % 3, (j + 2) % 3]
                      - M[(i + 1) % 3, (j + 2) % 3] * M[(i + 2) % 3, (j + 1) % 3])
    adj = torch.stack([torch.stack([c(j, i) for j in range(3)]) for i in range(3)])
    return adj / det3(M)


def _pk1_from_S(u_grad, Fp_inv_old, g_old, slip_old, rot_mat, y, gss_a, h, t_sat, xm, r, schmid, dt, ao):
    """models_copper.py:158-161: sigma = Fe S Fe^T / det Fe ; P = det F sigma F^-T."""
    S = y.reshape(DIM, DIM)
    _, _, _, Fe, F = helper(u_grad, Fp_inv_old, g_old, slip_old, rot_mat, S, gss_a, h, t_sat, xm, r, schmid, dt, ao)
    sigma = 1. / det3(Fe) * Fe @ S @ Fe.T
    P = det3(F) * sigma @ inv3(F).T
    return P


class PointBatch:
    """A batch of quadrature points with their state and (per-point or uniform) material data.

    Arrays follow the reference's `internal_vars` order (`models_copper.py:133`,
    `models_DPsteel_inhomo.py:229`), flattened over (cell, quad) to one leading axis.
    """

    def __init__(self, Fp_inv, g, slip, rot, mat: Optional[Material] = None, *, gss_a=None, h=None, t_sat=None,
                 xm=None, r=None, C=None, slip_table=None, ao=0.001, max_sub_step=5, tol=1e-8):
        t = lambda a: torch.as_tensor(onp.asarray(a), dtype=torch.float64)
        self.Fp_inv, self.g, self.slip, self.rot = t(Fp_inv), t(g), t(slip), t(rot)
        n = self.Fp_inv.shape[0]
        if mat is not None:
            gss_a = mat.gss_a if gss_a is None else gss_a
            h = mat.h if h is None else h
            t_sat = mat.t_sat if t_sat is None else t_sat
            xm = mat.xm if xm is None else xm
            r = mat.r if r is None else r
            C = cubic_C(mat.C11, mat.C12, mat.C44) if C is None else C
            slip_table = mat.slip if slip_table is None else slip_table
            ao, max_sub_step, tol = mat.ao, mat.max_sub_step, mat.tol
        per_pt = lambda a: (t(a) * torch.ones(n)) if onp.ndim(a) == 0 else t(a).reshape(n)
        self.gss_a, self.h, self.t_sat, self.xm, self.r = map(per_pt, (gss_a, h, t_sat, xm, r))
        self.C = t(C)                                  # (3,3,3,3) or (n,3,3,3,3)
        self.schmid = t(schmid_tensors(slip_table))
        self.ao, self.max_sub_step, self.tol = ao, max_sub_step, tol
        self.n = n

    def _args(self, idx):
        C = self.C if self.C.dim() == 4 else self.C[idx]
        return (self.Fp_inv[idx], self.g[idx], self.slip[idx], self.rot[idx]), \
               (self.gss_a[idx], self.h[idx], self.t_sat[idx], self.xm[idx], self.r[idx]), C

    # ---- batched wrappers -------------------------------------------------------------------
    def _res_fn(self, dt):
        cdim = None if self.C.dim() == 4 else 0

        def f(u_grad, A, g, sl, R, y, a, h, ts, xm, r, C):
            return implicit_residual(u_grad, A, g, sl, R, y, a, h, ts, xm, r, C, self.schmid, dt, self.ao)
        return f, (0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, cdim)

    def residual(self, u_grad, y, dt, idx=slice(None)):
        f, dims = self._res_fn(dt)
        st, pm, C = self._args(idx)
        return vmap(f, in_dims=dims)(u_grad, *st, y, *pm, C)

    def jac_y(self, u_grad, y, dt, idx=slice(None)):
        f, dims = self._res_fn(dt)
        st, pm, C = self._args(idx)
        return vmap(jacfwd(f, argnums=5), in_dims=dims)(u_grad, *st, y, *pm, C)

    def jac_ugrad(self, u_grad, y, dt, idx=slice(None)):
        f, dims = self._res_fn(dt)
        st, pm, C = self._args(idx)
        return vmap(jacfwd(f, argnums=0), in_dims=dims)(u_grad, *st, y, *pm, C)   # (n, 9, 3, 3)

    # ---- models_copper.py:204-249 -----------------------------------------------------------
    def newton_solver(self, u_grad, dt, return_iters=False):
        """Local Newton on S with the 'cut-half' line search, per point, literal control flow:

            y = 0; r = res(y)
            while ||r|| > tol:                                   (:212-216)
                inc = solve(jac_y(y), -r)                         (:226-227)
                relax, crt, sub = 1, r, 0
                while ||crt|| >= ||r|| and sub < max_sub_step:    (:235)
                    crt = res(y + relax*inc); relax *= .5; sub += 1
                y = y + 2*relax*inc ; r = crt                     (:245)

        Points are independent: converged points are frozen by removing them from the active index set,
        which is what the scalar reference does for each point on its own (not what `vmap` of a
        `while_loop` costs, but the same values).
        """
        u_grad = torch.as_tensor(u_grad, dtype=torch.float64)
        n = self.n
        y = torch.zeros(n, 9)
        res = self.residual(u_grad, y, dt)
        iters = torch.zeros(n, dtype=torch.int64)
        evals = torch.ones(n, dtype=torch.int64)
        active = torch.nonzero(torch.linalg.norm(res, dim=1) > self.tol).reshape(-1)
        guard = 0
        while active.numel() > 0:
            guard += 1
            if guard > 200:
                raise RuntimeError('oracle: local Newton did not converge in 200 iterations')
            ya, ra, ua = y[active], res[active], u_grad[active]
            J = self.jac_y(ua, ya, dt, active)
            inc = torch.linalg.solve(J, -ra)
            relax = torch.ones(active.numel())
            crt = ra.clone()
            sub = torch.zeros(active.numel(), dtype=torch.int64)
            rn = torch.linalg.norm(ra, dim=1)
            while True:
                m = (torch.linalg.norm(crt, dim=1) >= rn) & (sub < self.max_sub_step)
                # NaN: comparisons are False, the point leaves the loop like in the reference.
                k = torch.nonzero(m).reshape(-1)
                if k.numel() == 0:
                    break
                ak = active[k]
                crt[k] = self.residual(ua[k], ya[k] + relax[k, None] * inc[k], dt, ak)
                relax[k] = 0.5 * relax[k]
                sub[k] += 1
                evals[ak] += 1
            y[active] = ya + 2. * relax[:, None] * inc
            res[active] = crt
            iters[active] += 1
            active = active[torch.linalg.norm(crt, dim=1) > self.tol]
        if return_iters:
            return y, iters, evals
        return y

    # ---- models_copper.py:155-162 -----------------------------------------------------------
    def first_PK_stress(self, u_grad, dt, y=None):
        u_grad = torch.as_tensor(u_grad, dtype=torch.float64)
        if y is None:
            y = self.newton_solver(u_grad, dt)

        def f(u_grad, A, g, sl, R, y, a, h, ts, xm, r):
            return _pk1_from_S(u_grad, A, g, sl, R, y, a, h, ts, xm, r, self.schmid, dt, self.ao)
        st, pm, _ = self._args(slice(None))
        return vmap(f)(u_grad, *st, y, *pm)

    # ---- models_copper.py:251-259 + forward-mode through first_PK_stress ---------------------
    def tangent(self, u_grad, dt, y=None):
        """dP_ij/dH_kl (n,3,3,3,3): what `jax.jacfwd` of `tensor_map` w.r.t. `u_grad` yields.

        f_jvp (:251-259): dy = solve(jac_y, -(jac_x @ v)); only the u_grad block of v is non-zero in
        forward assembly.  Then first_PK_stress (:155-162) pushes (du_grad, dy) through `helper`.
        """
        u_grad = torch.as_tensor(u_grad, dtype=torch.float64)
        if y is None:
            y = self.newton_solver(u_grad, dt)
        n = self.n
        jy = self.jac_y(u_grad, y, dt)                                  # (n,9,9)
        jx = self.jac_ugrad(u_grad, y, dt).reshape(n, 9, 9)             # (n,9,9)  d res / d u_grad
        dy = torch.linalg.solve(jy, -jx)                                # (n,9,9)  dy/du_grad

        def f(u_grad, A, g, sl, R, y, a, h, ts, xm, r):
            return _pk1_from_S(u_grad, A, g, sl, R, y, a, h, ts, xm, r, self.schmid, dt, self.ao)
        st, pm, _ = self._args(slice(None))
        dP_du = vmap(jacfwd(f, argnums=0))(u_grad, *st, y, *pm).reshape(n, 9, 9)
        dP_dy = vmap(jacfwd(f, argnums=5))(u_grad, *st, y, *pm).reshape(n, 9, 9)
        A = dP_du + dP_dy @ dy
        return A.reshape(n, 3, 3, 3, 3)

    # ---- adjoint row F5: full jac_x and the total derivative of tensor_map ---------------------
    def x_size(self, nextra=0):
        """Entries of x = ravel([u_grad, Fp_inv_old, slip_resistance_old, slip_old, rot_mat]) (models_copper.py:156), then
        [gss_a, h, t_sat, xm, r] when nextra >= 5 (calibration form) and C (81) when nextra == 6 (DP form,
        models_DPsteel_inhomo.py:245)."""
        return 27 + 2 * self.g.shape[1] + (5 if nextra >= 5 else 0) + (81 if nextra >= 6 else 0)

    def _x_argnums(self, nextra):
        return (0, 1, 2, 3, 4) + ((6, 7, 8, 9, 10) if nextra >= 5 else ()) + ((11,) if nextra >= 6 else ())

    def jac_x(self, u_grad, y, dt, nextra=0):
        """jax.jacfwd(implicit_residual, argnums=0)(x, y) of f_jvp (models_copper.py:256): (n, 9, nx), columns in the
        reference's ravel order.  rot_mat's nine entries are independent variables, like jacfwd sees them."""
        u_grad = torch.as_tensor(u_grad, dtype=torch.float64)
        n = self.n
        st, pm, C = self._args(slice(None))
        if C.dim() == 4:
            C = C[None].expand(n, 3, 3, 3, 3)

        def f(u_grad, A, g, sl, R, y, a, h, ts, xm, r, C):
            return implicit_residual(u_grad, A, g, sl, R, y, a, h, ts, xm, r, C, self.schmid, dt, self.ao)
        J = vmap(jacfwd(f, argnums=self._x_argnums(nextra)))(u_grad, *st, y, *pm, C)
        return torch.cat([j.reshape(n, 9, -1) for j in J], dim=2)

    def dP_dx(self, u_grad, dt, y=None, nextra=0):
        """Total derivative of tensor_map with respect to x through the local solve (n, 9, nx): what jacfwd / vjp of
        first_PK_stress (models_copper.py:155-162) yields with newton_solver's custom f_jvp (:251-259):
        dP/dx = dP/dx|_y + dP/dy . dy/dx,  dy/dx = solve(jac_y, -jac_x)."""
        u_grad = torch.as_tensor(u_grad, dtype=torch.float64)
        if y is None:
            y = self.newton_solver(u_grad, dt)
        n = self.n
        jy = self.jac_y(u_grad, y, dt)
        jx = self.jac_x(u_grad, y, dt, nextra)
        dy = torch.linalg.solve(jy, -jx)                                # (n, 9, nx)

        def f(u_grad, A, g, sl, R, y, a, h, ts, xm, r):
            return _pk1_from_S(u_grad, A, g, sl, R, y, a, h, ts, xm, r, self.schmid, dt, self.ao)
        st, pm, _ = self._args(slice(None))
        argn = tuple(a for a in self._x_argnums(nextra) if a != 11)
        Jp = vmap(jacfwd(f, argnums=argn))(u_grad, *st, y, *pm)
        dP_x = torch.cat([j.reshape(n, 9, -1) for j in Jp], dim=2)
        if nextra >= 6:
            dP_x = torch.cat([dP_x, torch.zeros(n, 9, 81)], dim=2)      # P does not depend on C at fixed y
        dP_dy = vmap(jacfwd(f, argnums=5))(u_grad, *st, y, *pm).reshape(n, 9, 9)
        return dP_x + dP_dy @ dy

    # ---- models_copper.py:164-169 -----------------------------------------------------------
    def update_int_vars(self, u_grad, dt, y=None):
        u_grad = torch.as_tensor(u_grad, dtype=torch.float64)
        if y is None:
            y = self.newton_solver(u_grad, dt)

        def f(u_grad, A, g, sl, R, y, a, h, ts, xm, r):
            S = y.reshape(3, 3)
            Fp_inv_new, g_new, slip_new, _, _ = helper(u_grad, A, g, sl, R, S, a, h, ts, xm, r, self.schmid, dt, self.ao)
            return Fp_inv_new, g_new, slip_new
        st, pm, _ = self._args(slice(None))
        return vmap(f)(u_grad, *st, y, *pm)


# --------------------------------------------------------------------------------------------
# FE layer [jax_fem restated] - SURVEY Appendix D
# --------------------------------------------------------------------------------------------
# hex8 reference nodes in meshio/Gmsh order on [0,1]^3 (matches e.g.
# singlecrystal_copper/data/neper/singlecrystal_copper/mesh2.msh:39 "1 2 5 4 10 11 14 13")
HEX8_NODES = onp.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0],
                        [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], dtype=onp.float64)
_G = onp.array([(1 - 1 / onp.sqrt(3)) / 2, (1 + 1 / onp.sqrt(3)) / 2])
# 2x2x2 Gauss points on [0,1]^3, x slowest / z fastest, weights 1/8
HEX8_QUAD_PTS = onp.array([[_G[i], _G[j], _G[k]] for i in range(2) for j in range(2) for k in range(2)])
HEX8_QUAD_W = onp.full(8, 1.0 / 8.0)


def hex8_shape_grads_ref():
    """d N_a / d xi at each quad point, shape (8 quads, 8 nodes, 3)."""
    out = onp.zeros((8, 8, 3))
    for q, xi in enumerate(HEX8_QUAD_PTS):
        for a, na in enumerate(HEX8_NODES):
            f = onp.where(na == 1, xi, 1 - xi)
            df = onp.where(na == 1, 1.0, -1.0)
            for d in range(3):
                v = df[d]
                for e in range(3):
                    if e != d:
                        v = v * f[e]
                out[q, a, d] = v
    return out


def shape_grads_JxW(points, cells):
    """jax_fem FiniteElement.get_shape_grads: physical shape-function gradients (nc,8,8,3), JxW (nc,8)."""
    points = onp.asarray(points, dtype=onp.float64)
    cells = onp.asarray(cells)
    dN = hex8_shape_grads_ref()                                  # (q, a, 3)
    X = points[cells]                                            # (nc, a, 3)
    jac = onp.einsum('cai,qaj->cqij', X, dN)                     # dX_i / dxi_j
    jinv = onp.linalg.inv(jac)
    sg = onp.einsum('qaj,cqji->cqai', dN, jinv)                  # dN_a/dX_i = dN_a/dxi_j dxi_j/dX_i
    JxW = onp.linalg.det(jac) * HEX8_QUAD_W[None, :]
    return sg, JxW


def compute_u_grads(sol, cells, shape_grads):
    """models_copper.py:277-278: u_grads[c,q,i,j] = sum_a sol[cells[c,a], i] * shape_grads[c,q,a,j]."""
    return onp.einsum('cai,cqaj->cqij', onp.asarray(sol)[onp.asarray(cells)], shape_grads)


def coo_indices(cells, vec=3):
    """jax_fem Problem I/J rule (consumed at solver.py:281): inds[c,3a+i] = 3*cells[c,a]+i;
    I = repeat over columns, J = repeat over rows, so V[c,p,q] pairs with (inds[c,p], inds[c,q])."""
    cells = onp.asarray(cells, dtype=onp.int64)
    inds = (vec * cells[:, :, None] + onp.arange(vec)[None, None, :]).reshape(len(cells), -1)
    nd = inds.shape[1]
    I = onp.repeat(inds[:, :, None], nd, axis=2).reshape(-1)
    J = onp.repeat(inds[:, None, :], nd, axis=1).reshape(-1)
    return I, J


def csr_from_coo(V, I, J, ndof):
    """solver.py:281: scipy canonical CSR (duplicates summed, columns sorted, explicit zeros kept)."""
    return scipy.sparse.csr_array((onp.asarray(V), (I, J)), shape=(ndof, ndof))


class FEOracle:
    """Field-level oracle: one hex8 mesh + one PointBatch factory (reference: Problem subclass)."""

    def __init__(self, points, cells, make_batch):
        self.points = onp.asarray(points, dtype=onp.float64)
        self.cells = onp.asarray(cells, dtype=onp.int64)
        self.nc = len(self.cells)
        self.nn = len(self.points)
        self.shape_grads, self.JxW = shape_grads_JxW(self.points, self.cells)
        self.make_batch = make_batch          # params(list of (nc,8,...) arrays) -> PointBatch
        self.I, self.J = coo_indices(self.cells)

    def u_grads(self, sol):
        return compute_u_grads(sol, self.cells, self.shape_grads).reshape(-1, 3, 3)

    def update_int_vars_gp(self, sol, params, dt):
        """models_copper.py:273-282."""
        pb = self.make_batch(params)
        A, g, sl = pb.update_int_vars(self.u_grads(sol), dt)
        nc = self.nc
        out = list(params)
        out[0] = A.numpy().reshape(nc, 8, 3, 3)
        out[1] = g.numpy().reshape(nc, 8, -1)
        out[2] = sl.numpy().reshape(nc, 8, -1)
        return out

    def point_stress(self, sol, params, dt):
        pb = self.make_batch(params)
        return pb.first_PK_stress(self.u_grads(sol), dt).numpy().reshape(self.nc, 8, 3, 3)

    def cell_residual(self, sol, params, dt, P=None):
        """jax_fem laplace kernel: val[c,a,i] = sum_q sum_j P[c,q,i,j] shape_grads[c,q,a,j] JxW[c,q]."""
        if P is None:
            P = self.point_stress(sol, params, dt)
        return onp.einsum('cqij,cqaj,cq->cai', P, self.shape_grads, self.JxW)

    def compute_residual(self, sol, params, dt):
        """jax_fem compute_residual_vars_helper: scatter-add of the cell residuals (A10)."""
        wf = self.cell_residual(sol, params, dt)
        res = onp.zeros((self.nn, 3))
        onp.add.at(res, self.cells.reshape(-1), wf.reshape(-1, 3))
        return res

    def newton_update(self, sol, params, dt):
        """jax_fem newton_update: residual (nn,3) and V (nc*576,) with V[c, 3a+i, 3b+k] = d val[a,i]/d u[b,k]."""
        pb = self.make_batch(params)
        ug = self.u_grads(sol)
        y = pb.newton_solver(ug, dt)
        P = pb.first_PK_stress(ug, dt, y).numpy().reshape(self.nc, 8, 3, 3)
        A = pb.tangent(ug, dt, y).numpy().reshape(self.nc, 8, 3, 3, 3, 3)
        wf = onp.einsum('cqij,cqaj,cq->cai', P, self.shape_grads, self.JxW)
        res = onp.zeros((self.nn, 3))
        onp.add.at(res, self.cells.reshape(-1), wf.reshape(-1, 3))
        # d u_grad_kl / d u[b,k'] = delta_kk' shape_grads[b,l]
        Ke = onp.einsum('cqaj,cqijkl,cqbl,cq->caibk', self.shape_grads, A, self.shape_grads, self.JxW)
        V = Ke.reshape(-1)
        return res, V

    def vjp_params(self, sol, params, dt, adjoint):
        """vjp_linear_fn of implicit_vjp (crystal_plasticity_OR_design/solver.py:832-838): adjoint (nn, 3) contracted with
        d(compute_residual)/d(internal_vars), returned as a list shaped like `params`.  The residual is linear in P with the
        cell weights shape_grads x JxW, so the cotangent of P at a point is W_ij = sum_a adjoint[node_a, i] dN_a/dX_j JxW
        and the rest is the per-point total derivative `PointBatch.dP_dx`."""
        pb = self.make_batch(params)
        nextra = {4: 0, 9: 5, 10: 6}[len(params)]
        lam = onp.asarray(adjoint, dtype=onp.float64).reshape(self.nn, 3)[self.cells]          # (nc, 8 nodes, 3)
        W = onp.einsum('cai,cqaj,cq->cqij', lam, self.shape_grads, self.JxW).reshape(-1, 9)
        D = pb.dP_dx(self.u_grads(sol), dt, nextra=nextra).numpy()                              # (np, 9, nx)
        G = onp.einsum('pi,pic->pc', W, D)[:, 9:]                                               # drop the u_grad columns
        out, o = [], 0
        for a in params:
            k = int(onp.prod(a.shape[2:])) if a.ndim > 2 else 1
            out.append(G[:, o:o + k].reshape(a.shape))
            o += k
        return out

    def compute_avg_stress(self, sol, params, dt):
        """models_copper.py:297-319."""
        P = self.point_stress(sol, params, dt)
        F = self.u_grads(sol).reshape(self.nc, 8, 3, 3) + onp.eye(3)
        sigma = onp.einsum('cqij,cqkj->cqik', P, F) / onp.linalg.det(F)[:, :, None, None]
        return onp.sum(sigma * self.JxW[:, :, None, None], 1) / onp.sum(self.JxW, axis=1)[:, None, None]


def make_uniform_batch_factory(mat: Material):
    """internal_vars = [Fp_inv_gp, slip_resistance_gp, slip_gp, rot_mats_gp]  (models_copper.py:133)."""
    def f(params):
        Fp, g, sl, R = params[:4]
        ns = g.shape[-1]
        return PointBatch(Fp.reshape(-1, 3, 3), g.reshape(-1, ns), sl.reshape(-1, ns), R.reshape(-1, 3, 3), mat)
    return f


def make_dp_batch_factory(max_sub_step=5, slip_table=SLIP_BCC24):
    """internal_vars = [Fp_inv, g, slip, rot, gss_a, h, t_sat, xm, r, C]  (models_DPsteel_inhomo.py:229)."""
    def f(params):
        Fp, g, sl, R, a, h, ts, xm, r, C = params
        ns = g.shape[-1]
        return PointBatch(Fp.reshape(-1, 3, 3), g.reshape(-1, ns), sl.reshape(-1, ns), R.reshape(-1, 3, 3),
                          gss_a=a.reshape(-1), h=h.reshape(-1), t_sat=ts.reshape(-1), xm=xm.reshape(-1),
                          r=r.reshape(-1), C=C.reshape(-1, 3, 3, 3, 3), slip_table=slip_table,
                          max_sub_step=max_sub_step)
    return f


def initial_internal_vars(nc, mat: Material, rot_mats):
    """models_copper.py:79-91,133."""
    ns = len(mat.slip)
    Fp = onp.tile(onp.eye(3)[None, None], (nc, 8, 1, 1))
    g = mat.gss_initial * onp.ones((nc, 8, ns))
    sl = onp.zeros_like(g)
    R = onp.repeat(onp.asarray(rot_mats)[:, None, :, :], 8, axis=1)
    return [Fp, g, sl, R]


# --------------------------------------------------------------------------------------------
# Mesh + Dirichlet + outer Newton (only to reproduce the reference's golden end-to-end artefacts)
# --------------------------------------------------------------------------------------------
def box_mesh(Nx, Ny, Nz, Lx=1., Ly=1., Lz=1.):
    """Structured hex8 mesh: node id = ix + (Nx+1) iy + (Nx+1)(Ny+1) iz, cell id x-fastest, Gmsh node order
    (same numbering as the Neper files, e.g. mesh2.msh:39)."""
    xs, ys, zs = onp.linspace(0, Lx, Nx + 1), onp.linspace(0, Ly, Ny + 1), onp.linspace(0, Lz, Nz + 1)
    Z, Y, X = onp.meshgrid(zs, ys, xs, indexing='ij')
    points = onp.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    nid = lambda i, j, k: i + (Nx + 1) * j + (Nx + 1) * (Ny + 1) * k
    k, j, i = onp.meshgrid(onp.arange(Nz), onp.arange(Ny), onp.arange(Nx), indexing='ij')
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    cells = onp.stack([nid(i, j, k), nid(i + 1, j, k), nid(i + 1, j + 1, k), nid(i, j + 1, k),
                       nid(i, j, k + 1), nid(i + 1, j, k + 1), nid(i + 1, j + 1, k + 1), nid(i, j + 1, k + 1)], axis=1)
    return points, cells.astype(onp.int32)


def solve_load_step(fe: FEOracle, sol, params, dt, bc_nodes, bc_comps, bc_vals, tol=1e-6, rel_tol=1e-8,
                    dense_lstsq=False, max_outer=50):
    """solver.py:310-437 ('row elimination' Newton) with a direct linear solve instead of BiCGStab.

    bc_*: flat arrays of Dirichlet (node, component, value).  dense_lstsq: minimum-norm dense solve for the
    1-element calibration cases whose boundary conditions leave a rigid rotation free.
    """
    ndof = fe.nn * 3
    rows = onp.asarray(bc_nodes) * 3 + onp.asarray(bc_comps)
    sol = onp.array(sol, dtype=onp.float64)

    def assemble(sol):
        res, V = fe.newton_update(sol, params, dt)
        res = res.reshape(-1).copy()
        res[rows] = sol.reshape(-1)[rows] - bc_vals                       # apply_bc_vec, solver.py:119-133
        A = csr_from_coo(V, fe.I, fe.J, ndof).tolil()
        for r_ in rows:                                                   # zeroRows (diag=1), solver.py:290-293
            A.rows[r_] = [int(r_)]
            A.data[r_] = [1.0]
        return res, A.tocsr()

    res, A = assemble(sol)
    r0 = onp.linalg.norm(res)
    rn = r0
    it = 0
    while (rn / r0 > rel_tol) and (rn > tol):
        if dense_lstsq:
            inc = onp.linalg.lstsq(A.toarray(), -res, rcond=1e-12)[0]
        else:
            inc = scipy.sparse.linalg.spsolve(A.tocsc(), -res)
        sol = sol + inc.reshape(-1, 3)
        res, A = assemble(sol)
        rn = onp.linalg.norm(res)
        it += 1
        if it > max_outer:
            raise RuntimeError('oracle: outer Newton did not converge')
    return sol, it


# --------------------------------------------------------------------------------------------
# Linear solver of the reference: jax_solve (crystal_plasticity_OR_design/solver.py:19-48).
# The Krylov method itself is third-party: jax.scipy.sparse.linalg.bicgstab (JAX, jax/_src/scipy/sparse/linalg.py,
# function _bicgstab_solve; JAX is not vendored and not installable here, the model files name version 0.4.13 at
# models_copper.py:206).  Its published algorithm is restated below statement by statement; parity is anchored on the
# reference's call site (solver.py:34-40: x0, M = Jacobi, tol = atol = 1e-10, maxiter = 10000) and on its acceptance
# test ||A x - b|| < 0.1 (solver.py:43-45).
# --------------------------------------------------------------------------------------------
def bicgstab_ref(A, b, x0=None, M=None, tol=1e-5, atol=0.0, maxiter=None):
    """numpy restatement of jax.scipy.sparse.linalg.bicgstab.  Returns (x, k): k = iterations taken, or JAX's
    breakdown codes -10 (rho == 0) / -11 (omega == 0 or alpha == 0)."""
    b = onp.asarray(b, dtype=onp.float64)
    x = onp.zeros_like(b) if x0 is None else onp.array(x0, dtype=onp.float64)
    if maxiter is None:
        maxiter = 10 * len(b)
    Mf = (lambda v: v) if M is None else M
    bs = float(b @ b)
    atol2 = max(tol ** 2 * bs, atol ** 2)
    r = b - A @ x
    rhat, p, q = r.copy(), r.copy(), r.copy()
    alpha = omega = rho = 1.0
    k = 0
    while (float(r @ r) > atol2) and (k < maxiter) and (k >= 0):
        rho_ = float(rhat @ r)
        beta = rho_ / rho * alpha / omega
        p = r + beta * (p - omega * q)
        phat = Mf(p)
        q = A @ phat
        alpha_ = rho_ / float(rhat @ q)
        s = r - alpha_ * q
        exit_early = float(s @ s) < atol2
        shat = Mf(s)
        t = A @ shat
        omega_ = float(t @ s) / float(t @ t)
        if exit_early:
            x = x + alpha_ * phat
            r = s
        else:
            x = x + (alpha_ * phat + omega_ * shat)
            r = s - omega_ * t
        k_ = -11 if (omega_ == 0 or alpha_ == 0) else k + 1
        if rho_ == 0:
            k_ = -10
        alpha, omega, rho, k = alpha_, omega_, rho_, k_
    return x, k


def jax_solve_ref(A, b, x0, precond=True):
    """solver.py:19-48 on a scipy CSR matrix: Jacobi-preconditioned BiCGStab + the residual acceptance test."""
    jacobi = A.diagonal()
    M = (lambda v: v * (1. / jacobi)) if precond else None
    x, k = bicgstab_ref(A, b, x0=x0, M=M, tol=1e-10, atol=1e-10, maxiter=10000)
    err = onp.linalg.norm(A @ x - b)
    assert err < 0.1, f'linear solver failed to converge with err = {err}'
    return x, k, err


def solve_load_step_bicgstab(fe: 'FEOracle', sol, params, dt, bc_nodes, bc_comps, bc_vals, tol=1e-6, rel_tol=1e-8,
                             max_outer=50):
    """solver.py:310-437 with the reference's own linear solver path (jax_solve) and its initial guess
    (linear_incremental_solver, solver.py:213-237: x0 = assign_bc(0) - copy_bc(dofs))."""
    ndof = fe.nn * 3
    rows = onp.asarray(bc_nodes) * 3 + onp.asarray(bc_comps)
    bc_vals = onp.asarray(bc_vals, dtype=onp.float64)
    sol = onp.array(sol, dtype=onp.float64)

    def assemble(sol):
        res, V = fe.newton_update(sol, params, dt)
        res = res.reshape(-1).copy()
        res[rows] = sol.reshape(-1)[rows] - bc_vals
        A = csr_from_coo(V, fe.I, fe.J, ndof).tolil()
        for r_ in rows:
            A.rows[r_] = [int(r_)]
            A.data[r_] = [1.0]
        return res, A.tocsr()

    res, A = assemble(sol)
    r0 = onp.linalg.norm(res)
    rn = r0
    it, lin_its = 0, []
    while (rn / r0 > rel_tol) and (rn > tol):
        x0 = onp.zeros(ndof)
        x0[rows] = bc_vals - sol.reshape(-1)[rows]
        inc, k, err = jax_solve_ref(A, -res, x0, True)
        lin_its.append(k)
        sol = sol + inc.reshape(-1, 3)
        res, A = assemble(sol)
        rn = onp.linalg.norm(res)
        it += 1
        if it > max_outer:
            raise RuntimeError('oracle: outer Newton did not converge')
    return sol, it, lin_its
