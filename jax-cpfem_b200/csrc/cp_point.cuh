// Per-quadrature-point Kalidindi crystal-plasticity update, hand-derived, fp64.
//
// Replaces (reference JAX-CPFEM, paths relative to its tree):
//   helper             singlecrystal_copper/models_copper.py:172-192
//   implicit_residual  singlecrystal_copper/models_copper.py:195-201
//   newton_solver      singlecrystal_copper/models_copper.py:204-249   (literal control flow)
//   f_jvp              singlecrystal_copper/models_copper.py:251-259   (implicit-function tangent)
//   first_PK_stress    singlecrystal_copper/models_copper.py:155-162
//   update_int_vars    singlecrystal_copper/models_copper.py:164-169
// and the per-point-parameter form polycrystal_DPsteel/models_DPsteel_inhomo.py:240-361.
//
// Formulation (see DESIGN.md "Per-point algebra"): everything is done in the CRYSTAL frame, where the
// Schmid tensors are the constant d (x) n of the slip table and the elastic tensor is cubic
// (C11, C12, C44).  With  Fc = R^T F R,  Ac = R^T Fp_inv_old R,  G = Fc Ac  the reference's residual
// r(S) = S - rot4(R,C) : 1/2 (Fe^T Fe - I) becomes  r_c(S_c) = S_c - C : 1/2 (Fe_c^T Fe_c - I)  with
// Fe_c = G (I - sum_a dgamma_a d_a n_a^T); its Frobenius norm, the Newton iterates and the line-search
// decisions are those of the reference up to rounding.  All iterates are symmetric (r and the Newton
// increment are), so the unknown is the 6-vector s = (S00,S11,S22,S12,S02,S01) and the 9x9 solve of the
// reference collapses to a 6x6 one.  The Newton matrix J = I + sum_a w_a (C:e_a) p_a^T is solved in the
// scaled form  N z = -C^-1 r,  N = C^-1 D^-1 + sum_a w_a e_a etilde_a^T,  inc = D^-1 z  (D = diag(1,1,1,2,2,2)),
// which is symmetric positive definite up to O(strain) terms, so the LU needs no pivoting.
//
// Code shape (sm_100a: 64 DFMA/clk/SM, 32 KB L1.5 instruction cache, 64 K registers per SM):
//   * the loops over slip systems are ROLLED (unrolled by 4 / 2 only, for ILP in the power-law chains); the
//     per-system arrays (1/g and w = d dgamma / d tau) live behind an accessor `Arr` that the kernels point at a
//     per-thread column of shared memory, so they cost no registers and the Newton loop body stays inside the
//     instruction cache;
//   * the per-system constants (etilde, d n^T, d, n) are one 192-byte record in the kernel-parameter constant bank;
//   * |x|^(n-1) with a compile-time integer exponent (POWN = 9, 19, 119: copper, DP steel, 304 steel) is a
//     straight-line square-and-multiply chain; POWN = 0 takes the exponent at run time (per point);
//   * the local Newton solve is ONE loop over residual evaluations (a small state machine), so that the residual
//     code exists once and lanes of a warp that are at different line-search trials still share every evaluation.
//
// The functions are __host__ __device__ so that tests can compile this header with g++ and compare the
// algebra with the oracle on a machine without a GPU (tests/hostcheck).  The product only ever calls
// them from the CUDA kernels in cpfem_kernels.cu.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define CP_HD __host__ __device__ __forceinline__
#else
#define CP_HD inline
#endif

#define CP_MAX_NS 24
#ifndef CP_TRACE_X
#define CP_TRACE_X(a, ax)      // test hook (tests only): observe |tau/g| of every residual evaluation
#endif

// Uniform material description.  Per-point overrides (DP steel) come through CpPointParams.
struct CpMaterial {
    double C11, C12, C44;   // cubic elastic constants (models_copper.py:94-96)
    double h;               // hardening modulus            (:141)
    double t_sat;           // saturation slip resistance   (:143)
    double gss_a;           // hardening exponent           (:145)
    double ao;              // reference slip rate          (:147)
    double xm;              // rate sensitivity; exponent is 1/xm (:149,174)
    double r;               // latent hardening ratio       (:54)
    double tol;             // local Newton tolerance       (:212)
    int max_sub_step;       // line-search halvings         (:231)
    int max_iter;           // safety cap (reference has none; hitting it is reported in the status word)
};

// One slip system in the crystal frame, from the normalised normal n and direction d of the slip table
// (models_copper.py:62-69).  24 doubles = 192 bytes.
struct
#ifdef __CUDACC__
    __align__(16)
#endif
    CpSlipSys {
    double Et[6];   // strain-like Voigt of sym(d n^T): d0n0, d1n1, d2n2, (d1n2+d2n1)/2, (d0n2+d2n0)/2, (d0n1+d1n0)/2
    double M[9];    // d n^T, row-major
    double pad0;
    double d[3];    // 16-byte aligned together with n: the six doubles load as three pairs
    double n[3];
    double pad[2];
};
struct CpSlip {
    CpSlipSys sys[CP_MAX_NS];
};
// The device keeps two copies of the table: the kernel parameter (constant bank: free operands for loops whose
// index is uniform and known at compile time) and a shared-memory copy (for the data-dependent indices of the
// active-set loops).  CpSlipRef carries both; on the host they are the same table.
struct CpSlipRef {
    const CpSlip* u;    // uniform / compile-time indices
    const CpSlip* d;    // data-dependent indices
};

// rows of `slip6`: normal(3) direction(3), un-normalised, as in data/csv/input_slip_sys*.txt.  Returns false on a zero vector.
inline bool cp_slip_init(CpSlip* sl, const double* slip6, int ns) {
    for (int a = 0; a < CP_MAX_NS; ++a)
        for (int i = 0; i < 24; ++i) ((double*)&sl->sys[a])[i] = 0.0;
    for (int a = 0; a < ns; ++a) {
        const double* row = slip6 + 6 * a;
        const double nn = sqrt(row[0] * row[0] + row[1] * row[1] + row[2] * row[2]);
        const double dn = sqrt(row[3] * row[3] + row[4] * row[4] + row[5] * row[5]);
        if (!(nn > 0.0) || !(dn > 0.0)) return false;
        CpSlipSys& y = sl->sys[a];
        for (int i = 0; i < 3; ++i) { y.n[i] = row[i] / nn; y.d[i] = row[3 + i] / dn; }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) y.M[3 * i + j] = y.d[i] * y.n[j];
        y.Et[0] = y.M[0]; y.Et[1] = y.M[4]; y.Et[2] = y.M[8];
        y.Et[3] = 0.5 * (y.M[5] + y.M[7]); y.Et[4] = 0.5 * (y.M[2] + y.M[6]); y.Et[5] = 0.5 * (y.M[1] + y.M[3]);
    }
    return true;
}

struct CpPointParams {   // per-point values actually used at one quadrature point
    double C11, C12, C44, n_exp /* = 1/xm */;     // needed inside the local Newton solve
    double S11, S12, S44h;                        // cubic compliance: (C11+C12)/den, -C12/den, 1/(2 C44)
    double x_lo;                                  // slip systems with |tau/g| < x_lo are inactive (see cp_x_lo)
    double h, t_sat, gss_a, r;                    // hardening law: only the state update reads them (set after the solve)
};
// Activity threshold of the power law: a system with |tau/g| < x_lo has |tau/g|^(n-1) < 1e-20, i.e. a slip increment
// below 1e-24 and a Newton-matrix contribution below 1e-19 of the elastic compliance for every parameter set of the
// reference - no effect on any double-precision result, so such systems are skipped (their dgamma and w are set to 0).
// With rate exponent 120 fewer than 5 of the 12 FCC systems are active in an average residual evaluation.
CP_HD double cp_x_lo(double n_exp) {
    return (n_exp > 1.0) ? exp(-46.051701859880914 / (n_exp - 1.0)) : 0.0;     // 1e-20 ^ (1/(n-1))
}
CP_HD void cp_params_elastic(CpPointParams& pm, double C11, double C12, double C44, double xm, double x_lo = -1.0) {
    pm.C11 = C11; pm.C12 = C12; pm.C44 = C44; pm.n_exp = 1.0 / xm;
    pm.x_lo = (x_lo >= 0.0) ? x_lo : cp_x_lo(pm.n_exp);
    const double iden = 1.0 / ((C11 - C12) * (C11 + 2.0 * C12));
    pm.S11 = (C11 + C12) * iden; pm.S12 = -C12 * iden; pm.S44h = 0.5 / C44;
}

// Per-thread array of one double per slip system.  STRIDE = 1 on the host; on the device the kernels point it at
// column threadIdx.x of a [NS][blockDim] shared-memory tile (conflict-free).
template <int STRIDE>
struct CpArr {
    double* p;
    CP_HD double& operator[](int a) const { return p[a * STRIDE]; }
};

// ---------------------------------------------------------------------------------------------------
// small 3x3 helpers (row-major double[9])
// ---------------------------------------------------------------------------------------------------
CP_HD void m3_mul(const double* A, const double* B, double* C) {          // C = A B
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
CP_HD void m3_mul_tn(const double* A, const double* B, double* C) {       // C = A^T B
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
CP_HD void m3_mul_nt(const double* A, const double* B, double* C) {       // C = A B^T
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
CP_HD double m3_det(const double* M) {
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}
CP_HD void m3_inv(const double* M, double* Mi, double* det_out) {
    double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
    double det = M[0] * c00 + M[1] * c01 + M[2] * c02;
    double id = 1.0 / det;
    Mi[0] = c00 * id; Mi[1] = (M[2] * M[7] - M[1] * M[8]) * id; Mi[2] = (M[1] * M[5] - M[2] * M[4]) * id;
    Mi[3] = c01 * id; Mi[4] = (M[0] * M[8] - M[2] * M[6]) * id; Mi[5] = (M[2] * M[3] - M[0] * M[5]) * id;
    Mi[6] = c02 * id; Mi[7] = (M[1] * M[6] - M[0] * M[7]) * id; Mi[8] = (M[0] * M[4] - M[1] * M[3]) * id;
    *det_out = det;
}
// Voigt map used throughout: 0:(0,0) 1:(1,1) 2:(2,2) 3:(1,2) 4:(0,2) 5:(0,1)
CP_HD void sym6_to_m3(const double* s, double* S) {
    S[0] = s[0]; S[4] = s[1]; S[8] = s[2];
    S[5] = S[7] = s[3]; S[2] = S[6] = s[4]; S[1] = S[3] = s[5];
}

// x_u^N for U values at once, N a compile-time integer >= 1.  The chains of the U values advance in lock step
// (statement order = interleaved), so that consecutive DMULs are independent and the FP64 pipe latency is hidden
// by instruction-level parallelism.  Square-and-multiply, except that a factor 7 or 17 of N is peeled off first
// (119 = 7 x 17: 4 + 5 = 9 multiplications instead of 11).
template <int N, int U>
struct CpIpow {
    static CP_HD void run(const double* x, double* out) {
        if (N == 1) {
#pragma unroll
            for (int u = 0; u < U; ++u) out[u] = x[u];
        } else if (N > 17 && N % 17 == 0) {
            double y[U];
            CpIpow<17, U>::run(x, y);
            CpIpow<(N % 17 == 0 ? N / 17 : 1), U>::run(y, out);
        } else if (N > 7 && N % 7 == 0) {
            double y[U];
            CpIpow<7, U>::run(x, y);
            CpIpow<(N % 7 == 0 ? N / 7 : 1), U>::run(y, out);
        } else {
            double h[U];
            CpIpow<(N > 1 ? N / 2 : 1), U>::run(x, h);
#pragma unroll
            for (int u = 0; u < U; ++u) out[u] = h[u] * h[u];
            if (N & 1) {
#pragma unroll
                for (int u = 0; u < U; ++u) out[u] *= x[u];
            }
        }
    }
};

// x^e for x >= 0, run-time e.  Integers take the square-and-multiply loop, half-integers add one sqrt
// (copper's hardening exponent 2.5), everything else (tantalum: 44.2726) goes through pow().
CP_HD double cp_pow_pos(double x, double e) {
    const double e2 = e + e;
    if (e2 == floor(e2) && e >= 0.0 && e < 2048.0) {
        int k = (int)e;
        double r = ((double)k == e) ? 1.0 : sqrt(x);
        while (k) {
            if (k & 1) r *= x;
            x *= x;
            k >>= 1;
        }
        return r;
    }
    return pow(x, e);
}

// |x_u|^(n-1) for U slip systems at once.  POWN > 0: compile-time integer exponent; POWN == 0: run-time n1.
template <int POWN, int U>
CP_HD void cp_rate_pow(const double* ax /*U, >= 0*/, double n1, double* out) {
    if (POWN > 0) {
        CpIpow<(POWN > 0 ? POWN : 1), U>::run(ax, out);
    } else {
        if (n1 == floor(n1) && n1 >= 0.0 && n1 < 2048.0) {
            int k = (int)n1;
            double x[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { out[u] = 1.0; x[u] = ax[u]; }
            while (k) {
                if (k & 1) {
#pragma unroll
                    for (int u = 0; u < U; ++u) out[u] *= x[u];
                }
#pragma unroll
                for (int u = 0; u < U; ++u) x[u] *= x[u];
                k >>= 1;
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) out[u] = pow(ax[u], n1);
        }
    }
}

// union of a per-lane bit mask over the lanes of the warp that are executing together (host: one lane)
CP_HD unsigned cp_warp_or(unsigned m) {
#ifdef __CUDA_ARCH__
    return __reduce_or_sync(__activemask(), m);
#else
    return m;
#endif
}
CP_HD int cp_ffs0(unsigned m) {          // index of the lowest set bit (m != 0)
#ifdef __CUDA_ARCH__
    return __ffs((int)m) - 1;
#else
    return __builtin_ctz(m);
#endif
}
CP_HD int cp_popc(unsigned m) {
#ifdef __CUDA_ARCH__
    return __popc(m);
#else
    return __builtin_popcount(m);
#endif
}

// ---------------------------------------------------------------------------------------------------
// One residual evaluation at s (crystal frame).  Also leaves what the next Newton matrix needs (w, Fe).
//   tau_a  = d_a . S n_a = etilde_a . (D s)            (models_copper.py:173)
//   dg_a   = ao dt |tau/g|^(1/xm) sign(tau)            (:174)
//   w_a    = d dg_a / d tau_a = ao dt n |tau/g|^(n-1) / g
//   Fe     = G (I - sum_a dg_a d_a n_a^T)              (:188-191)
//   r      = s - C : 1/2 (Fe^T Fe - I)                 (:199)
// returns ||r||_F over the 9 entries (:212, np.linalg.norm of the 9-vector).
// `s_is_zero`: the caller knows s == 0 and n > 1, where every tau, dg and w vanishes (first evaluation of every solve).
// ---------------------------------------------------------------------------------------------------
// Two passes over the slip systems: (1) x_a = tau_a / g_a for all of them (parked in w[a]) and the set of ACTIVE
// systems |x_a| >= x_lo, united over the warp so that control flow stays uniform; (2) the power law, w and the Lp
// accumulation for the active ones only, U at a time (U independent multiplication chains in flight).  A lane whose
// own |x_a| is below x_lo gets dgamma_a = w_a = 0 even when the system is processed because another lane of the warp
// needs it, so every point's result depends on its own data only (bitwise, whatever the warp composition).
// `mact` returns the warp's active set (what the Newton matrix and the tangent loop over), `mask` the set that was
// processed (mact padded to a multiple of U); w[a] outside `mask` is meaningless until cp_newton zeroes it at the end.
template <int NS, int POWN, class Arr>
CP_HD double cp_residual(const CpSlipRef& sl, const CpPointParams& pm, double cdt, const double* G, const Arr& ginv,
                         const Arr& w, const double* s, bool s_is_zero, double* r, double* Fe, double* Lp, unsigned& mask,
                         unsigned& mact) {
    constexpr int U = 4;
    static_assert(NS % U == 0, "slip systems are processed four at a time");
#pragma unroll
    for (int i = 0; i < 9; ++i) Lp[i] = 0.0;
    mask = 0u;
    mact = 0u;
    if (!s_is_zero) {
        const double n1 = pm.n_exp - 1.0;
        const double cn = cdt * pm.n_exp;
        const double s3 = s[3] + s[3], s4 = s[4] + s[4], s5 = s[5] + s[5];
        unsigned act = 0u;
#pragma unroll
        for (int a = 0; a < NS; ++a) {
            const CpSlipSys& y = sl.u->sys[a];
            const double tau = y.Et[0] * s[0] + y.Et[1] * s[1] + y.Et[2] * s[2] + y.Et[3] * s3 + y.Et[4] * s4 + y.Et[5] * s5;
            const double x = tau * ginv[a];
            w[a] = x;
            CP_TRACE_X(a, fabs(x));
            act |= (fabs(x) >= pm.x_lo ? 1u : 0u) << a;
        }
        unsigned m = cp_warp_or(act);
        mact = m;
        // pad the set to a multiple of U with inactive systems (they are evaluated honestly: tiny values)
        {
            const int k = (U - (cp_popc(m) & (U - 1))) & (U - 1);
            for (int i = 0; i < k; ++i) {
                const unsigned z = ~m & ((NS == 32) ? 0xffffffffu : ((1u << NS) - 1u));
                m |= z & (0u - z);
            }
        }
        mask = m;
        while (m) {
            int a[U];
            double x[U], ax[U], pw[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                a[u] = cp_ffs0(m);
                m &= m - 1u;
                x[u] = w[a[u]];
                ax[u] = fabs(x[u]);
            }
            cp_rate_pow<POWN, U>(ax, n1, pw);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const CpSlipSys& y = sl.d->sys[a[u]];
                if (!(ax[u] >= pm.x_lo)) pw[u] = 0.0;
                const double dg = (cdt * pw[u]) * x[u];
                w[a[u]] = (cn * pw[u]) * ginv[a[u]];
#pragma unroll
                for (int i = 0; i < 9; ++i) Lp[i] += dg * y.M[i];
            }
        }
    }
    // Fe = G - G Lp
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            Fe[3 * i + j] = G[3 * i + j] - (G[3 * i] * Lp[j] + G[3 * i + 1] * Lp[3 + j] + G[3 * i + 2] * Lp[6 + j]);
    // E = 1/2 (Fe^T Fe - I)
    const double E0 = 0.5 * (Fe[0] * Fe[0] + Fe[3] * Fe[3] + Fe[6] * Fe[6] - 1.0);
    const double E1 = 0.5 * (Fe[1] * Fe[1] + Fe[4] * Fe[4] + Fe[7] * Fe[7] - 1.0);
    const double E2 = 0.5 * (Fe[2] * Fe[2] + Fe[5] * Fe[5] + Fe[8] * Fe[8] - 1.0);
    const double E3 = 0.5 * (Fe[1] * Fe[2] + Fe[4] * Fe[5] + Fe[7] * Fe[8]);
    const double E4 = 0.5 * (Fe[0] * Fe[2] + Fe[3] * Fe[5] + Fe[6] * Fe[8]);
    const double E5 = 0.5 * (Fe[0] * Fe[1] + Fe[3] * Fe[4] + Fe[6] * Fe[7]);
    r[0] = s[0] - (pm.C11 * E0 + pm.C12 * (E1 + E2));
    r[1] = s[1] - (pm.C11 * E1 + pm.C12 * (E0 + E2));
    r[2] = s[2] - (pm.C11 * E2 + pm.C12 * (E0 + E1));
    r[3] = s[3] - 2.0 * pm.C44 * E3;
    r[4] = s[4] - 2.0 * pm.C44 * E4;
    r[5] = s[5] - 2.0 * pm.C44 * E5;
    return sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + 2.0 * (r[3] * r[3] + r[4] * r[4] + r[5] * r[5]));
}

// ---------------------------------------------------------------------------------------------------
// Newton matrix in scaled form, LU-factorised in place (no pivoting, see header comment).
//   N = C^-1 D^-1 + sum_a w_a e_a etilde_a^T,  e_a = voigt(sym(K d_a n_a^T)), K = Fe^T G,
//   etilde_a = voigt(sym(d_a n_a^T))   (strain-like: shear entries carry the 1/2)
// After the call N holds L (unit lower) and U; piv[i] = 1/U_ii.
// ---------------------------------------------------------------------------------------------------
template <int NS, class Arr>
CP_HD void cp_newton_matrix(const CpSlipRef& sl, const CpPointParams& pm, const double* G, const double* Fe,
                            const Arr& w, unsigned mask, double* N /*36*/, double* piv /*6*/) {
    double K[9];
    m3_mul_tn(Fe, G, K);
    const double S11 = pm.S11, S12 = pm.S12, S44q = 0.5 * pm.S44h;
#pragma unroll
    for (int i = 0; i < 36; ++i) N[i] = 0.0;
    N[0] = N[7] = N[14] = S11;
    N[1] = N[2] = N[6] = N[8] = N[12] = N[13] = S12;
    N[21] = N[28] = N[35] = S44q;
    // only the warp's active systems contribute (w = 0 for all others); two per trip for instruction-level parallelism
    for (unsigned m = mask; m;) {
        int a2[2];
        double w2[2];
        a2[0] = cp_ffs0(m);
        m &= m - 1u;
        w2[0] = w[a2[0]];
        a2[1] = a2[0];
        w2[1] = 0.0;
        if (m) {
            a2[1] = cp_ffs0(m);
            m &= m - 1u;
            w2[1] = w[a2[1]];
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const CpSlipSys& y = sl.d->sys[a2[u]];
            const double k0 = K[0] * y.d[0] + K[1] * y.d[1] + K[2] * y.d[2];
            const double k1 = K[3] * y.d[0] + K[4] * y.d[1] + K[5] * y.d[2];
            const double k2 = K[6] * y.d[0] + K[7] * y.d[1] + K[8] * y.d[2];
            const double wa = w2[u], wh = 0.5 * wa;
            double e[6];
            e[0] = wa * (k0 * y.n[0]); e[1] = wa * (k1 * y.n[1]); e[2] = wa * (k2 * y.n[2]);
            e[3] = wh * (k1 * y.n[2] + k2 * y.n[1]); e[4] = wh * (k0 * y.n[2] + k2 * y.n[0]); e[5] = wh * (k0 * y.n[1] + k1 * y.n[0]);
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) N[6 * i + j] += e[i] * y.Et[j];
        }
    }
    // in-place LU (Doolittle), no pivoting
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double ip = 1.0 / N[6 * k + k];
        piv[k] = ip;
#pragma unroll
        for (int i = k + 1; i < 6; ++i) {
            const double l = N[6 * i + k] * ip;
            N[6 * i + k] = l;
#pragma unroll
            for (int j = k + 1; j < 6; ++j) N[6 * i + j] -= l * N[6 * k + j];
        }
    }
}

CP_HD void cp_lu_solve(const double* N, const double* piv, double* b /*6, in: rhs, out: z*/) {
#pragma unroll
    for (int i = 1; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < i; ++j) b[i] -= N[6 * i + j] * b[j];
#pragma unroll
    for (int i = 5; i >= 0; --i) {
#pragma unroll
        for (int j = i + 1; j < 6; ++j) b[i] -= N[6 * i + j] * b[j];
        b[i] *= piv[i];
    }
}

// b = -C^-1 r (strain-like vector)
CP_HD void cp_compliance_neg(const CpPointParams& pm, const double* r, double* b) {
    const double S11 = pm.S11, S12 = pm.S12, S44h = pm.S44h;
    b[0] = -(S11 * r[0] + S12 * (r[1] + r[2]));
    b[1] = -(S11 * r[1] + S12 * (r[0] + r[2]));
    b[2] = -(S11 * r[2] + S12 * (r[0] + r[1]));
    b[3] = -S44h * r[3]; b[4] = -S44h * r[4]; b[5] = -S44h * r[5];
}

struct CpSolveInfo {
    int iters;      // outer Newton iterations taken
    int evals;      // residual evaluations
    int status;     // bit0: hit max_iter ; bit1: non-finite residual
};

// ---------------------------------------------------------------------------------------------------
// Local Newton solve, control flow of models_copper.py:204-249:
//   y = 0 ; r = res(y)
//   while ||r|| > tol:  inc = solve(J(y), -r); relax = 1; crt = r; sub = 0
//        while ||crt|| >= ||r|| and sub < max_sub_step:  crt = res(y + relax inc); relax /= 2; sub++
//        y += 2 relax inc ; r = crt
// written as ONE loop over residual evaluations: `st` is the point being evaluated, (s, rn) the accepted iterate.
// The first Newton step needs no matrix: at y = 0 every w_a is 0 (n > 1), J = I and inc = -r exactly.
// On return s, w, Fe, Lp are consistent with the last residual evaluation, which is the one at the
// returned s (bitwise: y + relax*inc and y + 2*(relax/2)*inc are the same number).
// ---------------------------------------------------------------------------------------------------
template <int NS, int POWN, class Arr>
CP_HD void cp_newton(const CpSlipRef& sl, const CpPointParams& pm, double cdt, double tol, int max_sub, int max_iter,
                     const double* G, const Arr& ginv, const Arr& w, double* s, double* Fe, double* Lp,
                     unsigned& mask, unsigned& mact, CpSolveInfo& info) {
    double r[6], st[6], inc[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { s[i] = 0.0; st[i] = 0.0; inc[i] = 0.0; }
    const bool zero_ok = pm.n_exp > 1.0;
    double rn = 0.0, relax = 1.0;
    int sub = 0;
    bool first = true;
    info.iters = 0; info.evals = 0; info.status = 0;
    for (;;) {
        const double crtn = cp_residual<NS, POWN>(sl, pm, cdt, G, ginv, w, st, first && zero_ok, r, Fe, Lp, mask, mact);
        ++info.evals;
        if (!first) {
            relax *= 0.5;
            ++sub;
            if (crtn >= rn && sub < max_sub) {           // line search: next trial y + relax inc
#pragma unroll
                for (int i = 0; i < 6; ++i) st[i] = s[i] + relax * inc[i];
                continue;
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) s[i] = st[i];    // accept: y + 2 relax inc == the point just evaluated
            ++info.iters;
        }
        rn = crtn;
        if (!(rn > tol)) break;
        if (info.iters >= max_iter) { info.status |= 1; break; }
        if (first && zero_ok) {
#pragma unroll
            for (int i = 0; i < 6; ++i) inc[i] = -r[i];
        } else {
            double N[36], piv[6];
            cp_newton_matrix<NS>(sl, pm, G, Fe, w, mact, N, piv);
            cp_compliance_neg(pm, r, inc);
            cp_lu_solve(N, piv, inc);
            inc[3] *= 0.5; inc[4] *= 0.5; inc[5] *= 0.5;          // inc = D^-1 z
        }
        first = false;
        relax = 1.0;
        sub = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) st[i] = s[i] + inc[i];
    }
    if (!(rn == rn)) info.status |= 2;
    // systems outside the last processed set: w = 0 (the output stages read w of every system)
#pragma unroll 4
    for (int a = 0; a < NS; ++a)
        if (!((mask >> a) & 1u)) w[a] = 0.0;
}

// ---------------------------------------------------------------------------------------------------
// Frame change helpers
// ---------------------------------------------------------------------------------------------------
CP_HD void cp_to_crystal(const double* R, const double* M, double* Mc) {     // Mc = R^T M R
    double T[9];
    m3_mul_tn(R, M, T);
    m3_mul(T, R, Mc);
}
CP_HD void cp_to_lab(const double* R, const double* Mc, double* M) {         // M = R Mc R^T
    double T[9];
    m3_mul(R, Mc, T);
    m3_mul_nt(T, R, M);
}

template <class Arr>
struct CpPointState {       // everything the output stages need, crystal frame
    double G[9], Ac[9];
    double s[6], Fe[9], Lp[9];
    Arr ginv, w;            // per-slip-system arrays (1/g_old, d dgamma / d tau at the solution)
    unsigned mask, mact;    // slip systems processed by / active in the last residual evaluation (w = 0 for all others)
    double cdt;
    CpSolveInfo info;
};

// Set-up + solve for one point.  H = u_grad (lab), A = Fp_inv_old (lab), g = slip resistances (any indexable),
// R = rot_mat.  ps.ginv / ps.w must point at storage for NS doubles each.  ps.Ac is NOT set here: the output stages
// need it, the Newton loop does not, so the callers fill it afterwards with cp_point_frame (the kernels reload A and
// R from memory for that, which keeps 27 doubles out of the loop's registers).
template <int NS, int POWN, class Arr, class GIn>
CP_HD void cp_point_solve(const CpSlipRef& sl, const CpMaterial& mat, const CpPointParams& pm, double dt,
                          const double* H, const double* A, const GIn& g, const double* R, CpPointState<Arr>& ps) {
    {
        double F[9], Fc[9], Ac[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) F[i] = H[i];
        F[0] += 1.0; F[4] += 1.0; F[8] += 1.0;
        cp_to_crystal(R, F, Fc);
        cp_to_crystal(R, A, Ac);
        m3_mul(Fc, Ac, ps.G);
    }
#pragma unroll 4
    for (int a = 0; a < NS; ++a) ps.ginv[a] = 1.0 / g[a];
    ps.cdt = mat.ao * dt;
    cp_newton<NS, POWN>(sl, pm, ps.cdt, mat.tol, mat.max_sub_step, mat.max_iter, ps.G, ps.ginv, ps.w, ps.s, ps.Fe, ps.Lp,
                        ps.mask, ps.mact, ps.info);
}

template <class Arr>
CP_HD void cp_point_frame(const double* A, const double* R, CpPointState<Arr>& ps) {   // Ac = R^T Fp_inv_old R
    cp_to_crystal(R, A, ps.Ac);
}

// New state (models_copper.py:164-169 via helper :172-192): Fp_inv_new (lab), g_new, slip_new.
// g / slip_old are read and g_new / slip_new written through indexable accessors (global memory in the kernels);
// ps.w is overwritten with the hardening terms t_a.
template <int NS, class Arr, class GIn, class GOut>
CP_HD void cp_point_state_update(const CpSlipRef& sl, const CpPointParams& pm, const CpPointState<Arr>& ps,
                                 const GIn& g, const GIn& slip_old, const double* R,
                                 double* A_new_lab, const GOut& g_new, const GOut& slip_new) {
    {
        double ImL[9], Anc[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) ImL[i] = -ps.Lp[i];
        ImL[0] += 1.0; ImL[4] += 1.0; ImL[8] += 1.0;
        m3_mul(ps.Ac, ImL, Anc);
        cp_to_lab(R, Anc, A_new_lab);
    }
    const double s3 = ps.s[3] + ps.s[3], s4 = ps.s[4] + ps.s[4], s5 = ps.s[5] + ps.s[5];
    const double inv_n = 1.0 / pm.n_exp, inv_tsat = 1.0 / pm.t_sat;
    double tsum = 0.0;
#pragma unroll 2
    for (int a = 0; a < NS; ++a) {
        const CpSlipSys& y = sl.u->sys[a];
        // dgamma_a recomputed from w_a: dg = w tau / n
        const double tau = y.Et[0] * ps.s[0] + y.Et[1] * ps.s[1] + y.Et[2] * ps.s[2] + y.Et[3] * s3 + y.Et[4] * s4 + y.Et[5] * s5;
        const double dg = ps.w[a] * tau * inv_n;
        slip_new[a] = slip_old[a] + dg;
        const double yy = 1.0 - g[a] * inv_tsat;
        const double sg = (yy > 0.0) ? 1.0 : ((yy < 0.0) ? -1.0 : 0.0);
        const double t = pm.h * fabs(dg) * cp_pow_pos(fabs(yy), pm.gss_a) * sg;       // :178
        ps.w[a] = t;
        tsum += t;
    }
    // g_inc = q t with q = r everywhere, 1 on the triples {3k,3k+1,3k+2}  (:71-76,179)
#pragma unroll 1
    for (int b0 = 0; b0 < NS; b0 += 3) {
        const double trip = ps.w[b0] + ps.w[b0 + 1] + ps.w[b0 + 2];
        const double ginc = pm.r * (tsum - trip) + trip;
#pragma unroll
        for (int u = 0; u < 3; ++u) g_new[b0 + u] = g[b0 + u] + ginc;
    }
}

// First Piola-Kirchhoff stress (lab): P = Fe S A_new^T / det A_new  (== det F sigma F^-T, :160-161).
// Also returns the crystal-frame pieces the tangent needs.
struct CpStressAux {
    double Anc[9];     // A_new crystal
    double idet;       // 1/det(A_new)
    double Pc[9];      // P crystal
};

template <class Arr>
CP_HD void cp_point_stress(const CpPointState<Arr>& ps, const double* R, double* P_lab, CpStressAux& ax) {
    double ImL[9], S[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) ImL[i] = -ps.Lp[i];
    ImL[0] += 1.0; ImL[4] += 1.0; ImL[8] += 1.0;
    m3_mul(ps.Ac, ImL, ax.Anc);
    ax.idet = 1.0 / m3_det(ax.Anc);
    sym6_to_m3(ps.s, S);
    double T[9], T2[9];
    m3_mul(ps.Fe, S, T);
    m3_mul_nt(T, ax.Anc, T2);
#pragma unroll
    for (int i = 0; i < 9; ++i) ax.Pc[i] = T2[i] * ax.idet;
    cp_to_lab(R, ax.Pc, P_lab);
}

// ---------------------------------------------------------------------------------------------------
// Consistent tangent dP_ij/dH_kl in the lab frame, out[(3i+j)*ld + (3k+l)*ls].
//   ds/dF   = -J^-1 dr/dF   (f_jvp, models_copper.py:251-259),  dr/dF[dF] = -C : sym(Fe^T dF A_new)
//   dP      = dF Z + [ dFe S A_new^T + Fe dS A_new^T + Fe S dA_new^T ] / det - P tr(A_new^-1 dA_new)
//   with dLp = sum_a w_a (p_a . ds) d_a n_a^T,  dA_new = -Ac dLp,  dFe = -G dLp,
//        tr(A_new^-1 dA_new) = -tr((I-Lp)^-1 dLp),  Z = A_new S A_new^T / det.
// The loop runs over LAB directions dF = e_k e_l^T, i.e. crystal dF_c = (R^T e_k)(R^T e_l)^T, and rotates each
// dP_c column back with R, which is cheaper than rotating the rank-4 tensor.
// `scale` multiplies the whole tangent (JxW for the element integration).
// ---------------------------------------------------------------------------------------------------
template <int NS, class Arr, typename OutT>
CP_HD void cp_point_tangent(const CpSlipRef& sl, const CpPointParams& pm, const CpPointState<Arr>& ps,
                            const CpStressAux& ax, const double* R, double scale, OutT out, long ld, long ls) {
    double N[36], piv[6];
    cp_newton_matrix<NS>(sl, pm, ps.G, ps.Fe, ps.w, ps.mact, N, piv);
    // constant pieces
    double S[9], Z[9], T1[9] /* S A_new^T */, T2[9] /* Fe S */, Y[9] /* (I-Lp)^-1 */, tmp[9], dY;
    sym6_to_m3(ps.s, S);
    m3_mul_nt(S, ax.Anc, T1);
    m3_mul(ax.Anc, T1, Z);
#pragma unroll
    for (int i = 0; i < 9; ++i) Z[i] *= ax.idet;
    m3_mul(ps.Fe, S, T2);
#pragma unroll
    for (int i = 0; i < 9; ++i) tmp[i] = -ps.Lp[i];
    tmp[0] += 1.0; tmp[4] += 1.0; tmp[8] += 1.0;
    m3_inv(tmp, Y, &dY);
    // U_k = Fe^T R^T e_k  -> rows of (R Fe) ; V_l = A_new^T-contracted: row l of (R A_new)
    double RFe[9], RAn[9], RZ[9];
    m3_mul(R, ps.Fe, RFe);      // RFe[k][:] = sum_c R[k][c] Fe[c][:]   (= Fe^T r_k as a row)
    m3_mul(R, ax.Anc, RAn);     // RAn[l][:] = sum_c R[l][c] A_new[c][:]
    m3_mul(R, Z, RZ);           // RZ[l][:]  = sum_c R[l][c] Z[c][:]
#pragma unroll 1
    for (int kl = 0; kl < 9; ++kl) {
        const int k = kl / 3, l = kl - 3 * k;
        const double* u = &RFe[3 * k];     // Fe^T r_k
        const double* v = &RAn[3 * l];     // (r_l^T A_new)
        // z = N^-1 voigt(sym(u v^T))
        double z[6];
        z[0] = u[0] * v[0]; z[1] = u[1] * v[1]; z[2] = u[2] * v[2];
        z[3] = 0.5 * (u[1] * v[2] + u[2] * v[1]); z[4] = 0.5 * (u[0] * v[2] + u[2] * v[0]); z[5] = 0.5 * (u[0] * v[1] + u[1] * v[0]);
        cp_lu_solve(N, piv, z);
        // ds = D^-1 z ; dS full
        double dS[9];
        dS[0] = z[0]; dS[4] = z[1]; dS[8] = z[2];
        dS[5] = dS[7] = 0.5 * z[3]; dS[2] = dS[6] = 0.5 * z[4]; dS[1] = dS[3] = 0.5 * z[5];
        // dLp = sum_a w_a (etilde_a . z) d_a n_a^T      (p_a . ds == etilde_a . z)
        double dLp[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) dLp[i] = 0.0;
        for (unsigned m = ps.mact; m; m &= m - 1u) {
            const int a = cp_ffs0(m);
            const CpSlipSys& y = sl.d->sys[a];
            const double dtau = y.Et[0] * z[0] + y.Et[1] * z[1] + y.Et[2] * z[2] + y.Et[3] * z[3] + y.Et[4] * z[4] + y.Et[5] * z[5];
            const double dgm = ps.w[a] * dtau;
#pragma unroll
            for (int i = 0; i < 9; ++i) dLp[i] += dgm * y.M[i];
        }
        // dPc = r_k (r_l^T Z)  +  [ -(G dLp) T1 + Fe dS A_new^T - T2 (Ac dLp)^T ] idet + Pc tr(Y dLp)
        double GdL[9], AdL[9], M1[9], M2[9], M3[9], dPc[9];
        m3_mul(ps.G, dLp, GdL);
        m3_mul(GdL, T1, M1);
        m3_mul(ps.Fe, dS, tmp);
        m3_mul_nt(tmp, ax.Anc, M2);
        m3_mul(ps.Ac, dLp, AdL);
        m3_mul_nt(T2, AdL, M3);
        const double tr = Y[0] * dLp[0] + Y[1] * dLp[3] + Y[2] * dLp[6] + Y[3] * dLp[1] + Y[4] * dLp[4] + Y[5] * dLp[7] +
                          Y[6] * dLp[2] + Y[7] * dLp[5] + Y[8] * dLp[8];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                dPc[3 * i + j] = R[3 * k + i] * RZ[3 * l + j] + (M2[3 * i + j] - M1[3 * i + j] - M3[3 * i + j]) * ax.idet +
                                 ax.Pc[3 * i + j] * tr;
        double dP[9];
        cp_to_lab(R, dPc, dP);
#pragma unroll
        for (int ij = 0; ij < 9; ++ij) out[ij * ld + kl * ls] = dP[ij] * scale;
    }
}
