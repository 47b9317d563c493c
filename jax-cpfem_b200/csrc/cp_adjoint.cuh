// cp_adjoint.cuh - the columns of jac_x = d(implicit_residual)/dx and the per-point vector-Jacobian product of the
// stress map, by hand-coded forward-mode dual numbers.
//
// Replaces (reference JAX-CPFEM):
//   f_jvp's  jac_x = jax.jacfwd(implicit_residual, argnums=0)(x, y)    singlecrystal_copper/models_copper.py:251-259
//            jac_y = jax.jacfwd(implicit_residual, argnums=1)(x, y)
//   with x = ravel([u_grad, Fp_inv_old, slip_resistance_old, slip_old, rot_mat])   (51 entries, FCC/BCC12; :156)
//   or   x = ravel([..., gss_a, h, t_sat, xm, r, C])                               (56 / 161 entries, calibration / DP form;
//                                                              polycrystal_DPsteel/models_DPsteel_inhomo.py:245)
//   and what reverse mode makes of them inside implicit_vjp (crystal_plasticity_OR_design/solver.py:801-853):
//   w : dP/dx = w : dP/dx|_S  -  (J_y^-T (dP/dS)^T w) . jac_x
//
// Unlike the forward kernels (cp_point.cuh: crystal frame, symmetric 6-vector, orthogonality of R used freely), the
// functions here follow the reference's formulation LITERALLY - lab frame, 9 unknowns, Schmid tensors R M R^T, elastic
// tensor rotate_tensor_rank_4(R, C) - because jacfwd differentiates with respect to the nine entries of R as independent
// numbers: off the rotation group the two formulations are different functions, and parity is with the reference's.
// Two simplifications that are exact for ANY 3x3 R:  R (d n^T) R^T = (R d)(R n)^T,  and for a cubic C
//   rot4(R,C) : E = C12 B (B : E) + 2 C44 B E B^T + (C11 - C12 - 2 C44) sum_m r_m r_m^T (r_m . E r_m),  B = R R^T, r_m = column m of R
// (E symmetric), which replaces the 3^8-term contraction.
//
// The code is generic in the scalar type T (double or CpDual) and __host__ __device__, so tests/hostcheck compares it
// with the oracle's autodiff on a machine without a GPU.
#pragma once
#include "cp_point.cuh"

struct CpDual {
    double v, d;
    CP_HD CpDual() : v(0.0), d(0.0) {}
    CP_HD CpDual(double v_) : v(v_), d(0.0) {}
    CP_HD CpDual(double v_, double d_) : v(v_), d(d_) {}
};
CP_HD CpDual operator+(CpDual a, CpDual b) { return CpDual(a.v + b.v, a.d + b.d); }
CP_HD CpDual operator-(CpDual a, CpDual b) { return CpDual(a.v - b.v, a.d - b.d); }
CP_HD CpDual operator-(CpDual a) { return CpDual(-a.v, -a.d); }
CP_HD CpDual operator*(CpDual a, CpDual b) { return CpDual(a.v * b.v, a.v * b.d + a.d * b.v); }
CP_HD CpDual operator/(CpDual a, CpDual b) {
    const double q = a.v / b.v;
    return CpDual(q, (a.d - q * b.d) / b.v);
}
CP_HD CpDual& operator+=(CpDual& a, CpDual b) { a.v += b.v; a.d += b.d; return a; }
CP_HD double cp_value(double a) { return a; }
CP_HD double cp_value(CpDual a) { return a.v; }
CP_HD double cp_deriv(double) { return 0.0; }
CP_HD double cp_deriv(CpDual a) { return a.d; }

// |x|^n sign(x) with the value and derivative 0 at x == 0 (what JAX returns there for n > 1; the oracle guards the same way)
CP_HD double cp_signed_pow(double x, double n) {
    if (x == 0.0) return 0.0;
    const double p = pow(fabs(x), n);
    return x > 0.0 ? p : -p;
}
CP_HD CpDual cp_signed_pow(CpDual x, CpDual n) {
    if (x.v == 0.0) return CpDual(0.0, 0.0);
    const double ax = fabs(x.v);
    const double p = pow(ax, n.v);
    const double sp = x.v > 0.0 ? p : -p;
    // d/dx |x|^n sign(x) = n |x|^(n-1);  d/dn = |x|^n ln|x| sign(x)
    return CpDual(sp, n.v * (p / ax) * x.d + sp * log(ax) * n.d);
}

template <class T>
CP_HD void ad_mul(const T* A, const T* B, T* C) {          // C = A B
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
template <class T>
CP_HD void ad_mul_nt(const T* A, const T* B, T* C) {       // C = A B^T
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
template <class T>
CP_HD void ad_mul_tn(const T* A, const T* B, T* C) {       // C = A^T B
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
template <class T>
CP_HD T ad_det(const T* M) {
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}
template <class T>
CP_HD void ad_inv(const T* M, T* Mi) {
    const T c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
    const T det = M[0] * c00 + M[1] * c01 + M[2] * c02;
    const T id = T(1.0) / det;
    Mi[0] = c00 * id; Mi[1] = (M[2] * M[7] - M[1] * M[8]) * id; Mi[2] = (M[1] * M[5] - M[2] * M[4]) * id;
    Mi[3] = c01 * id; Mi[4] = (M[0] * M[8] - M[2] * M[6]) * id; Mi[5] = (M[2] * M[3] - M[0] * M[5]) * id;
    Mi[6] = c02 * id; Mi[7] = (M[1] * M[6] - M[0] * M[7]) * id; Mi[8] = (M[0] * M[4] - M[1] * M[3]) * id;
}

// implicit_residual (models_copper.py:195-201 via helper :172-192) and first_PK_stress (:155-162) at given (x, y), literal
// formulation.  H = u_grad, A = Fp_inv_old, g = slip resistances, R = rot_mat (9 independent entries), xm, C11/C12/C44 the
// cubic constants of C, S = y reshaped (9 independent entries).  r (9) = ravel(S - rot4(R,C) : E);  P (9) if not null.
// E_out (9, optional) returns the Green strain (the closed-form dC columns need it).
template <int NS, class T>
CP_HD void cp_ref_residual(const CpSlip& sl, double cdt, const T* H, const T* A, const T* g, const T* R, T xm, T C11, T C12, T C44,
                           const T* S, T* r, T* P, T* E_out) {
    T F[9];
    for (int i = 0; i < 9; ++i) F[i] = H[i];
    F[0] = F[0] + T(1.0); F[4] = F[4] + T(1.0); F[8] = F[8] + T(1.0);
    const T n_exp = T(1.0) / xm;
    T Lp[9];
    for (int i = 0; i < 9; ++i) Lp[i] = T(0.0);
    for (int a = 0; a < NS; ++a) {
        const CpSlipSys& y = sl.sys[a];
        T Rd[3], Rn[3], SRn[3];
        for (int i = 0; i < 3; ++i) {
            Rd[i] = R[3 * i] * T(y.d[0]) + R[3 * i + 1] * T(y.d[1]) + R[3 * i + 2] * T(y.d[2]);
            Rn[i] = R[3 * i] * T(y.n[0]) + R[3 * i + 1] * T(y.n[1]) + R[3 * i + 2] * T(y.n[2]);
        }
        for (int i = 0; i < 3; ++i) SRn[i] = S[3 * i] * Rn[0] + S[3 * i + 1] * Rn[1] + S[3 * i + 2] * Rn[2];
        const T tau = Rd[0] * SRn[0] + Rd[1] * SRn[1] + Rd[2] * SRn[2];                 // S : (R d)(R n)^T   (:173)
        const T dg = T(cdt) * cp_signed_pow(tau / g[a], n_exp);                         // (:174)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Lp[3 * i + j] += dg * Rd[i] * Rn[j];
    }
    T ImL[9], An[9], Fe[9], E[9];
    for (int i = 0; i < 9; ++i) ImL[i] = -Lp[i];
    ImL[0] = ImL[0] + T(1.0); ImL[4] = ImL[4] + T(1.0); ImL[8] = ImL[8] + T(1.0);
    ad_mul(A, ImL, An);                                                                   // Fp_inv_new (:188)
    ad_mul(F, An, Fe);                                                                    // Fe (:190)
    ad_mul_tn(Fe, Fe, E);
    for (int i = 0; i < 9; ++i) E[i] = T(0.5) * E[i];
    E[0] = E[0] - T(0.5); E[4] = E[4] - T(0.5); E[8] = E[8] - T(0.5);
    if (E_out)
        for (int i = 0; i < 9; ++i) E_out[i] = E[i];
    // S_ = rot4(R, C) : E for cubic C and any R
    T B[9], BE[9], BEB[9];
    ad_mul_nt(R, R, B);
    ad_mul(B, E, BE);
    ad_mul_nt(BE, B, BEB);
    T trBE = T(0.0);
    for (int i = 0; i < 9; ++i) trBE += B[i] * E[i];
    const T Cp = C11 - C12 - T(2.0) * C44;
    T Sx[9];
    for (int i = 0; i < 9; ++i) Sx[i] = C12 * B[i] * trBE + T(2.0) * C44 * BEB[i];
    for (int m = 0; m < 3; ++m) {
        const T r0 = R[m], r1 = R[3 + m], r2 = R[6 + m];
        const T e0 = E[0] * r0 + E[1] * r1 + E[2] * r2, e1 = E[3] * r0 + E[4] * r1 + E[5] * r2, e2 = E[6] * r0 + E[7] * r1 + E[8] * r2;
        const T q = Cp * (r0 * e0 + r1 * e1 + r2 * e2);
        const T rm[3] = {r0, r1, r2};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Sx[3 * i + j] += q * rm[i] * rm[j];
    }
    for (int i = 0; i < 9; ++i) r[i] = S[i] - Sx[i];
    if (P) {
        // sigma = Fe S Fe^T / det Fe ; P = det F sigma F^-T   (:158-161)
        T FS[9], sg[9], Fi[9];
        ad_mul(Fe, S, FS);
        ad_mul_nt(FS, Fe, sg);
        const T s = ad_det(F) / ad_det(Fe);
        ad_inv(F, Fi);
        T t[9];
        ad_mul_nt(sg, Fi, t);                     // sigma F^-T
        for (int i = 0; i < 9; ++i) P[i] = s * t[i];
    }
}

// Number of entries of x for a state with `nextra` trailing parameter arrays: 0 (uniform material: 4 arrays), 5
// (calibration form: gss_a, h, t_sat, xm, r) or 6 (DP form: ... + C).
CP_HD int cp_nx(int ns, int nextra) { return 27 + 2 * ns + (nextra >= 5 ? 5 : 0) + (nextra >= 6 ? 81 : 0); }

// One column c of x (0 <= c < 27 + 2 ns + 5; the C block is handled in closed form by the callers): dr/dx_c (9) and, if
// dP != nullptr, dP/dx_c at fixed S (9).  Columns that do not enter the residual (slip_old; gss_a, h, t_sat, r) give zeros
// without evaluating anything.
template <int NS>
CP_HD void cp_jac_x_column(const CpSlip& sl, double cdt, const double* H, const double* A, const double* g, const double* R,
                           double xm, double C11, double C12, double C44, const double* S, int c, double* dr, double* dP) {
    const int o_A = 9, o_g = 18, o_sl = 18 + NS, o_R = 18 + 2 * NS, o_p = 27 + 2 * NS;
    const bool zero = (c >= o_sl && c < o_R) || (c >= o_p && c != o_p + 3);
    if (zero) {
        for (int i = 0; i < 9; ++i) { dr[i] = 0.0; if (dP) dP[i] = 0.0; }
        return;
    }
    CpDual Hd[9], Ad[9], gd[NS], Rd[9], Sd[9];
    for (int i = 0; i < 9; ++i) { Hd[i] = CpDual(H[i]); Ad[i] = CpDual(A[i]); Rd[i] = CpDual(R[i]); Sd[i] = CpDual(S[i]); }
    for (int a = 0; a < NS; ++a) gd[a] = CpDual(g[a]);
    CpDual xmd(xm);
    if (c < o_A) Hd[c].d = 1.0;
    else if (c < o_g) Ad[c - o_A].d = 1.0;
    else if (c < o_sl) gd[c - o_g].d = 1.0;
    else if (c < o_p) Rd[c - o_R].d = 1.0;
    else xmd.d = 1.0;
    CpDual r[9], P[9];
    cp_ref_residual<NS, CpDual>(sl, cdt, Hd, Ad, gd, Rd, xmd, CpDual(C11), CpDual(C12), CpDual(C44), Sd, r, dP ? P : nullptr, nullptr);
    for (int i = 0; i < 9; ++i) { dr[i] = r[i].d; if (dP) dP[i] = P[i].d; }
}

// Column m of y (the nine entries of S): dr/dS_m (9) and dP/dS_m (9)
template <int NS>
CP_HD void cp_jac_y_column(const CpSlip& sl, double cdt, const double* H, const double* A, const double* g, const double* R,
                           double xm, double C11, double C12, double C44, const double* S, int m, double* dr, double* dP) {
    CpDual Hd[9], Ad[9], gd[NS], Rd[9], Sd[9];
    for (int i = 0; i < 9; ++i) { Hd[i] = CpDual(H[i]); Ad[i] = CpDual(A[i]); Rd[i] = CpDual(R[i]); Sd[i] = CpDual(S[i]); }
    for (int a = 0; a < NS; ++a) gd[a] = CpDual(g[a]);
    Sd[m].d = 1.0;
    CpDual r[9], P[9];
    cp_ref_residual<NS, CpDual>(sl, cdt, Hd, Ad, gd, Rd, CpDual(xm), CpDual(C11), CpDual(C12), CpDual(C44), Sd, r, dP ? P : nullptr, nullptr);
    for (int i = 0; i < 9; ++i) { dr[i] = r[i].d; if (dP) dP[i] = P[i].d; }
}

// dr/dC_abcd (the 81 trailing columns of the DP form): r = S - R_ia R_jb R_kc R_ld C_abcd E_kl  =>
// dr_ij/dC_abcd = -R_ia R_jb Ehat_cd,  Ehat = R^T E R.  P does not depend on C at fixed S.
CP_HD void cp_jac_C_prepare(const double* R, const double* E, double* Ehat) {
    double t[9];
    m3_mul_tn(R, E, t);
    m3_mul(t, R, Ehat);
}
CP_HD double cp_jac_C_entry(const double* R, const double* Ehat, int ij, int abcd) {
    const int i = ij / 3, j = ij % 3, a = abcd / 27, b = (abcd / 9) % 3, cd = abcd % 9;
    return -R[3 * i + a] * R[3 * j + b] * Ehat[cd];
}

// Solve M^T lam = b for a 9x9 M (row-major, destroyed), partial pivoting.  Returns false on a singular matrix.
CP_HD bool cp_solve9_transposed(double* M, double* b) {
    // work on T = M^T in place: T[i][j] = M[j][i]
    for (int i = 0; i < 9; ++i)
        for (int j = i + 1; j < 9; ++j) { const double t = M[9 * i + j]; M[9 * i + j] = M[9 * j + i]; M[9 * j + i] = t; }
    for (int k = 0; k < 9; ++k) {
        int p = k;
        double best = fabs(M[9 * k + k]);
        for (int i = k + 1; i < 9; ++i)
            if (fabs(M[9 * i + k]) > best) { best = fabs(M[9 * i + k]); p = i; }
        if (!(best > 0.0)) return false;
        if (p != k) {
            for (int j = 0; j < 9; ++j) { const double t = M[9 * k + j]; M[9 * k + j] = M[9 * p + j]; M[9 * p + j] = t; }
            const double t = b[k]; b[k] = b[p]; b[p] = t;
        }
        const double ip = 1.0 / M[9 * k + k];
        for (int i = k + 1; i < 9; ++i) {
            const double l = M[9 * i + k] * ip;
            for (int j = k + 1; j < 9; ++j) M[9 * i + j] -= l * M[9 * k + j];
            b[i] -= l * b[k];
        }
    }
    for (int i = 8; i >= 0; --i) {
        for (int j = i + 1; j < 9; ++j) b[i] -= M[9 * i + j] * b[j];
        b[i] /= M[9 * i + i];
    }
    return true;
}

// Per-point vector-Jacobian product of tensor_map with respect to x at the converged S (lab frame, 9 entries):
//   grad[c] = W : dP/dx_c|_S  -  lam . dr/dx_c,     J_y^T lam = (dP/dS)^T W
// for the columns [c0, c1) of x (the C block, if present, is appended by the caller through cp_jac_C_*).  `lam` (9) is
// returned for that purpose.  Returns false if J_y is singular.
template <int NS>
CP_HD bool cp_point_vjp(const CpSlip& sl, double cdt, const double* H, const double* A, const double* g, const double* R, double xm,
                        double C11, double C12, double C44, const double* S, const double* W, int c0, int c1, double* grad, double* lam) {
    double Jy[81];
    for (int m = 0; m < 9; ++m) {
        double dr[9], dP[9];
        cp_jac_y_column<NS>(sl, cdt, H, A, g, R, xm, C11, C12, C44, S, m, dr, dP);
        double b = 0.0;
        for (int i = 0; i < 9; ++i) { Jy[9 * i + m] = dr[i]; b += W[i] * dP[i]; }
        lam[m] = b;
    }
    if (!cp_solve9_transposed(Jy, lam)) return false;
    for (int c = c0; c < c1; ++c) {
        double dr[9], dP[9];
        cp_jac_x_column<NS>(sl, cdt, H, A, g, R, xm, C11, C12, C44, S, c, dr, dP);
        double v = 0.0;
        for (int i = 0; i < 9; ++i) v += W[i] * dP[i] - lam[i] * dr[i];
        grad[c - c0] = v;
    }
    return true;
}
