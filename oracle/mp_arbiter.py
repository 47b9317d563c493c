"""High-precision arbiter for the per-point constitutive solve.  TEST INFRASTRUCTURE ONLY (like cpfem_oracle.py).

The reference stops its local Newton iteration at ||r||_2 <= 1e-8 (models_copper.py:212-216) and decides its line search
on `||crt|| >= ||r||` (:235).  Two correct fp64 implementations can sit on different sides of such a comparison when it
is decided in the last bit; they then stop one iteration apart, both below the reference's tolerance, and differ from
each other by ~1e-10 of the stress.  This module restates the residual (models_copper.py:172-201, DP form
models_DPsteel_inhomo.py:260-295) in mpmath at 50 digits, solves r(S) = 0 to 1e-35 with a plain Newton iteration, and
returns the exact root S* and the exact first Piola-Kirchhoff stress P* (:155-162), so that a disputed point can be
judged against the truth instead of against another fp64 answer.

Pure Python loops over 3x3 / ns-sized objects: meant for a handful of points (about a second each).
"""
from __future__ import annotations

import mpmath as mp
import numpy as onp

mp.mp.dps = 50


def _m(a):
    return mp.matrix(onp.asarray(a, dtype=onp.float64).tolist())


def _eye():
    return mp.eye(3)


def _cubic_C(C11, C12, C44):
    """models_copper.py:92-130: C_ijkl of a cubic crystal in its own frame."""
    C = [[[[mp.mpf(0) for _ in range(3)] for _ in range(3)] for _ in range(3)] for _ in range(3)]
    for i in range(3):
        for j in range(3):
            C[i][i][j][j] = mp.mpf(C11) if i == j else mp.mpf(C12)
    for i in range(3):
        for j in range(3):
            if i != j:
                C[i][j][i][j] = mp.mpf(C44)
                C[i][j][j][i] = mp.mpf(C44)
    return C


def _rot4(R, C):
    """rotate_tensor_rank_4 (models_copper.py:21-26): R_ia R_jb R_kc R_ld C_abcd."""
    out = [[[[mp.mpf(0) for _ in range(3)] for _ in range(3)] for _ in range(3)] for _ in range(3)]
    nz = [(a, b, c, d) for a in range(3) for b in range(3) for c in range(3) for d in range(3) if C[a][b][c][d] != 0]
    for i in range(3):
        for j in range(3):
            for k in range(3):
                for l in range(3):
                    s = mp.mpf(0)
                    for a, b, c, d in nz:
                        s += R[i, a] * R[j, b] * R[k, c] * R[l, d] * C[a][b][c][d]
                    out[i][j][k][l] = s
    return out


def _spow(x, n):
    """abs(x)**n * sign(x)."""
    if x == 0:
        return mp.mpf(0)
    return mp.sign(x) * mp.power(abs(x), n)


class PointProblem:
    """One quadrature point: u_grad H, Fp_inv_old A, slip resistances g, rotation R, slip table (ns, 6) rows
    normal(3) direction(3), material scalars."""

    def __init__(self, H, A, g, R, slip_table, C11, C12, C44, xm, ao, dt):
        self.F = _m(H) + _eye()
        self.A = _m(A)
        self.g = [mp.mpf(float(v)) for v in onp.asarray(g).reshape(-1)]
        self.R = _m(R)
        self.n_exp = mp.mpf(1) / mp.mpf(float(xm))
        self.cdt = mp.mpf(float(ao)) * mp.mpf(float(dt))
        self.M = []
        for row in onp.asarray(slip_table, dtype=onp.float64):
            n = _m(row[:3].reshape(3, 1)); d = _m(row[3:].reshape(3, 1))
            n = n / mp.sqrt((n.T * n)[0]); d = d / mp.sqrt((d.T * d)[0])          # models_copper.py:62-66
            M0 = d * n.T                                                            # Schmid tensor d (x) n (:69)
            self.M.append(self.R * M0 * self.R.T)                                   # rotate_tensor_rank_2 (:29-34,185)
        self.Crot = _rot4(self.R, _cubic_C(C11, C12, C44))

    def parts(self, S):
        dgam = []
        for a, Ma in enumerate(self.M):
            tau = sum(S[i, j] * Ma[i, j] for i in range(3) for j in range(3))       # :173
            dgam.append(self.cdt * _spow(tau / self.g[a], self.n_exp))              # :174
        Lp = mp.zeros(3, 3)
        for a, Ma in enumerate(self.M):
            Lp += dgam[a] * Ma
        A_new = self.A * (_eye() - Lp)                                              # :188
        Fe = self.F * A_new                                                         # :190
        return dgam, A_new, Fe

    def residual(self, S):
        """models_copper.py:195-201."""
        _, _, Fe = self.parts(S)
        E = (Fe.T * Fe - _eye()) / 2
        out = mp.zeros(3, 3)
        for i in range(3):
            for j in range(3):
                out[i, j] = S[i, j] - sum(self.Crot[i][j][k][l] * E[k, l] for k in range(3) for l in range(3))
        return out

    def solve(self, S0=None, tol=mp.mpf('1e-35'), max_iter=60):
        """Newton on the 9 unknowns with a central-difference Jacobian (step 1e-20 at 50 digits: error ~1e-40 relative).
        Starts from S0 (e.g. an fp64 answer).  Returns S* as an mp 3x3 matrix."""
        S = _m(S0) if S0 is not None else mp.zeros(3, 3)
        h = mp.mpf('1e-20')
        for _ in range(max_iter):
            r = self.residual(S)
            rn = mp.sqrt(sum(r[i, j] ** 2 for i in range(3) for j in range(3)))
            if rn < tol:
                return S
            J = mp.zeros(9, 9)
            for c in range(9):
                dS = mp.zeros(3, 3)
                dS[c // 3, c % 3] = h
                rp, rm = self.residual(S + dS), self.residual(S - dS)
                for k in range(9):
                    J[k, c] = (rp[k // 3, k % 3] - rm[k // 3, k % 3]) / (2 * h)
            rhs = mp.matrix([-r[k // 3, k % 3] for k in range(9)])
            inc = mp.lu_solve(J, rhs)
            # damped update: never accept a step that increases the residual (the fp64 start is inside the basin anyway)
            lam = mp.mpf(1)
            for _ in range(40):
                St = mp.matrix([[S[i, j] + lam * inc[3 * i + j] for j in range(3)] for i in range(3)])
                rt = self.residual(St)
                if mp.sqrt(sum(rt[i, j] ** 2 for i in range(3) for j in range(3))) < rn:
                    break
                lam /= 2
            S = St
        raise RuntimeError('mp_arbiter: Newton did not reach %s' % tol)

    def first_PK(self, S):
        """models_copper.py:158-161: sigma = Fe S Fe^T / det Fe, P = det F sigma F^-T."""
        _, _, Fe = self.parts(S)
        sigma = Fe * S * Fe.T / mp.det(Fe)
        return mp.det(self.F) * sigma * (self.F ** -1).T

    def residual_norm(self, S):
        r = self.residual(_m(S))
        return mp.sqrt(sum(r[i, j] ** 2 for i in range(3) for j in range(3)))


def exact_point(H, A, g, R, slip_table, C11, C12, C44, xm, ao, dt, S_start=None):
    """Exact root S* (3x3, float64-rounded) and exact P* for one point; S_start = an fp64 solution to start from."""
    pb = PointProblem(H, A, g, R, slip_table, C11, C12, C44, xm, ao, dt)
    S = pb.solve(S_start)
    P = pb.first_PK(S)
    to_np = lambda Mx: onp.array([[float(Mx[i, j]) for j in range(3)] for i in range(3)])
    return to_np(S), to_np(P), pb
