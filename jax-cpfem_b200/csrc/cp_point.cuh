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
#include <string.h>

#ifdef __CUDACC__
#define CP_HD __host__ __device__ __forceinline__
#else
#define CP_HD inline
#endif

#define CP_MAX_NS 24
// Two tuning switches, both measured on B200 at 200^3 (profiles/r2/a_variants_n200.txt):
#ifndef CP_FAST_RCP
#define CP_FAST_RCP 1     // reciprocals as MUFU.RCP64H + two Newton steps (cp_rcp) instead of the IEEE division sequence:
                          // update 61.07 -> 59.52 ms, assembly 114.19 -> 109.48 ms
#endif
#ifndef CP_TAIL2
#define CP_TAIL2 1        // active slip set padded to a multiple of 2 instead of 4 (groups of 4, then one group of 2):
                          // update 61.07 -> 58.56 ms, assembly 114.19 -> 113.18 ms
#endif
#ifndef CP_NM_ODD1
#define CP_NM_ODD1 0      // experiment: single-system tail in cp_newton_matrix instead of a half-empty pair
#endif
#ifndef CP_PRUNE
#define CP_PRUNE 0        // experiment, off: line-search trials whose rejection is certain from tau/g alone are not evaluated
                          // (cp_prune_setup).  Bitwise-identical results and 5.8 % fewer FP64 instructions at the 304-steel
                          // benchmark state, but every code shape tried costs more than it saves on sm_100a: the loop sits at
                          // the register budget, and the extra exit makes ptxas spill Fe across the line-search decision or
                          // hoist the slip table's constant-bank operands into spilled registers (update 57.7 -> 61.4 ... 71.3 ms
                          // at 200^3, profiles/r2/h_prune_variants.txt).  tests/test_prune_option.py keeps the path honest.
#endif
// X_cap of the pruning in high-word terms: X_rej = 2^k (1 + m) is skipped up to 2^k (1.5 + m), i.e. X_cap / X_rej <= 1.5
#define CP_PRUNE_SPAN 0x80000u
#ifndef CP_LS_SMEM
#define CP_LS_SMEM 0      // experiment, off: accepted iterate and Newton increment of the local solve in shared memory
                          // (24 registers of loop state): 57.7 -> 58.8 ms at 200^3, ptxas fills the 168 registers either way
#endif
#ifndef CP_PEEL
#define CP_PEEL 0         // experiment: the evaluation at y = 0 peeled out of the Newton loop (1), and the leading trials of the
                          // FIRST line search that are certainly rejected skipped before the loop is entered (2); see cp_newton_p
#endif
#ifndef CP_G_SMEM
#define CP_G_SMEM 0       // experiment, off: G = Fc Ac of the local solve read from shared memory inside the loop (18 registers)
#endif
#ifndef CP_BLOCK_THREADS
#define CP_BLOCK_THREADS 64    // threads per block of the kernels that call cp_newton (checked in cpfem_kernels.cu)
#endif
#ifndef CP_TRACE_PRUNE
#define CP_TRACE_PRUNE()       // test hook (tests only): count the skipped evaluations
#endif
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
        // largest |d . n| of the table (0 up to rounding for a slip table): cp_prune_setup relies on tr(d n^T) = 0
        const double dn_ = fabs(y.d[0] * y.n[0] + y.d[1] * y.n[1] + y.d[2] * y.n[2]);
        if (dn_ > sl->sys[0].pad[0]) sl->sys[0].pad[0] = dn_;
    }
    return true;
}

// 1/x for a normal, finite x.  CP_FAST_RCP: hardware seed (about 20 bits) + two Newton steps = within 1 ulp, without the
// range checks and the correction step of the correctly rounded division.
CP_HD double cp_rcp(double x) {
#if CP_FAST_RCP && defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
#else
    return 1.0 / x;
#endif
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
    pm.C11 = C11; pm.C12 = C12; pm.C44 = C44; pm.n_exp = 1.0 / xm;     // exact: 1/(1/120) must be the integer 120
    pm.x_lo = (x_lo >= 0.0) ? x_lo : cp_x_lo(pm.n_exp);
    const double iden = cp_rcp((C11 - C12) * (C11 + 2.0 * C12));
    pm.S11 = (C11 + C12) * iden; pm.S12 = -C12 * iden; pm.S44h = 0.5 * cp_rcp(C44);
}

// Per-thread array of one double per slip system.  STRIDE = 1 on the host; on the device the kernels point it at
// column threadIdx.x of a [NS][blockDim] shared-memory tile (conflict-free).
template <int STRIDE>
struct CpArr {
    double* p;
    CP_HD double& operator[](int a) const { return p[a * STRIDE]; }
};

// read-only view of a small per-thread vector: plain array (stride 1) or a shared-memory column
template <int STRIDE>
struct CpVec {
    const double* p;
    CP_HD double operator[](int i) const { return p[i * STRIDE]; }
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
    double id = cp_rcp(det);
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
            // non-integer exponent (tantalum: 44.2726): exp2(n1 log2 x) - relative error n1 |log2 x| eps ~ 1e-14, a third of
            // the instructions of pow(), whose last-ulp accuracy buys nothing at the 1e-10 bar; x = 0 gives exp2(-inf) = 0
#pragma unroll
            for (int u = 0; u < U; ++u) out[u] = exp2(n1 * log2(ax[u]));
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
// true if `p` holds on every lane of the warp that is executing together (host: one lane)
CP_HD bool cp_warp_all(bool p) {
#ifdef __CUDA_ARCH__
    return __all_sync(__activemask(), p) != 0;
#else
    return p;
#endif
}
// High word of |x|.  Non-negative doubles order like their bit patterns, so hi(|x|) >= hi(t) is x >= t exactly when the
// low word of t is zero - comparisons on the integer pipe instead of the FP64 pipe.
CP_HD int cp_hi_abs(double x) {
#ifdef __CUDA_ARCH__
    return __double2hiint(x) & 0x7fffffff;
#else
    long long b;
    memcpy(&b, &x, 8);
    return (int)((b >> 32) & 0x7fffffff);
#endif
}

// Power law, w and the Lp accumulation for the next U systems of the set `m` (lowest bits first; they are removed from m).
template <int POWN, int U, class Arr>
CP_HD void cp_slip_group(const CpSlipRef& sl, const CpPointParams& pm, double cdt, double cn, double n1, const Arr& ginv,
                         const Arr& w, unsigned& m, double* Lp) {
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
        // pad the set to a multiple of U (CP_TAIL2: of 2) with inactive systems (they are evaluated honestly: tiny values)
        {
            constexpr int PADTO = (CP_TAIL2 == 2) ? 1 : (CP_TAIL2 ? 2 : U);
            const int k = (PADTO - (cp_popc(m) & (PADTO - 1))) & (PADTO - 1);
            for (int i = 0; i < k; ++i) {
                const unsigned z = ~m & ((NS == 32) ? 0xffffffffu : ((1u << NS) - 1u));
                m |= z & (0u - z);
            }
        }
        mask = m;
        while (m) {
#if CP_TAIL2
            const int left = cp_popc(m);
            if (left < 4) {
                if (left >= 2) cp_slip_group<POWN, 2>(sl, pm, cdt, cn, n1, ginv, w, m, Lp);
#if CP_TAIL2 == 2
                if (left & 1) cp_slip_group<POWN, 1>(sl, pm, cdt, cn, n1, ginv, w, m, Lp);
#endif
                break;
            }
#endif
            cp_slip_group<POWN, U>(sl, pm, cdt, cn, n1, ginv, w, m, Lp);
        }
    }
    // Fe = G - G Lp
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            Fe[3 * i + j] = G[3 * i + j] - (G[3 * i] * Lp[j] + G[3 * i + 1] * Lp[3 + j] + G[3 * i + 2] * Lp[6 + j]);
    // E2x = 2 E = Fe^T Fe - I; the 1/2 rides on the elastic constants (a multiplication by 0.5 is exact, so
    // (C/2) (2E) == C E bit for bit) and the 2 C44 (1/2) of the shear rows cancels
    const double E0 = Fe[0] * Fe[0] + Fe[3] * Fe[3] + Fe[6] * Fe[6] - 1.0;
    const double E1 = Fe[1] * Fe[1] + Fe[4] * Fe[4] + Fe[7] * Fe[7] - 1.0;
    const double E2 = Fe[2] * Fe[2] + Fe[5] * Fe[5] + Fe[8] * Fe[8] - 1.0;
    const double E3 = Fe[1] * Fe[2] + Fe[4] * Fe[5] + Fe[7] * Fe[8];
    const double E4 = Fe[0] * Fe[2] + Fe[3] * Fe[5] + Fe[6] * Fe[8];
    const double E5 = Fe[0] * Fe[1] + Fe[3] * Fe[4] + Fe[6] * Fe[7];
    const double C11h = 0.5 * pm.C11, C12h = 0.5 * pm.C12;
    r[0] = s[0] - (C11h * E0 + C12h * (E1 + E2));
    r[1] = s[1] - (C11h * E1 + C12h * (E0 + E2));
    r[2] = s[2] - (C11h * E2 + C12h * (E0 + E1));
    r[3] = s[3] - pm.C44 * E3;
    r[4] = s[4] - pm.C44 * E4;
    r[5] = s[5] - pm.C44 * E5;
    return sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + 2.0 * (r[3] * r[3] + r[4] * r[4] + r[5] * r[5]));
}

#if CP_PRUNE || CP_LS_SMEM || CP_G_SMEM
// ---- experimental code shape (CP_PRUNE / CP_LS_SMEM, both off by default): pass (1) split off, skip loop ----
// Pass (1): x_a = tau_a / g_a for every system, parked in w[a]; returns this lane's set of active systems and, in `hmax`,
// the high word of max_a |x_a| (what the pruning of certainly rejected line-search trials looks at, cp_prune_setup).
template <int NS, class Arr>
CP_HD unsigned cp_slip_ratios(const CpSlipRef& sl, const CpPointParams& pm, const Arr& ginv, const Arr& w, const double* s, int& hmax) {
    const double s3 = s[3] + s[3], s4 = s[4] + s[4], s5 = s[5] + s[5];
    unsigned act = 0u;
    int h = 0;
#pragma unroll
    for (int a = 0; a < NS; ++a) {
        const CpSlipSys& y = sl.u->sys[a];
        const double tau = y.Et[0] * s[0] + y.Et[1] * s[1] + y.Et[2] * s[2] + y.Et[3] * s3 + y.Et[4] * s4 + y.Et[5] * s5;
        const double x = tau * ginv[a];
        w[a] = x;
        CP_TRACE_X(a, fabs(x));
        act |= (fabs(x) >= pm.x_lo ? 1u : 0u) << a;
#if CP_PRUNE
        const int hx = cp_hi_abs(x);
        h = hx > h ? hx : h;
#endif
    }
    hmax = h;
    return act;
}

// Pass (2) and the rest of the evaluation; `act` from cp_slip_ratios (ignored when s_is_zero).
template <int NS, int POWN, class Arr, class GT>
CP_HD double cp_residual_x(const CpSlipRef& sl, const CpPointParams& pm, double cdt, const GT& G, const Arr& ginv,
                         const Arr& w, const double* s, bool s_is_zero, unsigned act, double* r, double* Fe, double* Lp,
                         unsigned& mask, unsigned& mact) {
    constexpr int U = 4;
    static_assert(NS % U == 0, "slip systems are processed four at a time");
#pragma unroll
    for (int i = 0; i < 9; ++i) Lp[i] = 0.0;
    mask = 0u;
    mact = 0u;
    if (!s_is_zero) {
        const double n1 = pm.n_exp - 1.0;
        const double cn = cdt * pm.n_exp;
        unsigned m = cp_warp_or(act);
        mact = m;
        // pad the set to a multiple of U (CP_TAIL2: of 2) with inactive systems (they are evaluated honestly: tiny values)
        {
            constexpr int PADTO = (CP_TAIL2 == 2) ? 1 : (CP_TAIL2 ? 2 : U);
            const int k = (PADTO - (cp_popc(m) & (PADTO - 1))) & (PADTO - 1);
            for (int i = 0; i < k; ++i) {
                const unsigned z = ~m & ((NS == 32) ? 0xffffffffu : ((1u << NS) - 1u));
                m |= z & (0u - z);
            }
        }
        mask = m;
        while (m) {
#if CP_TAIL2
            const int left = cp_popc(m);
            if (left < 4) {
                if (left >= 2) cp_slip_group<POWN, 2>(sl, pm, cdt, cn, n1, ginv, w, m, Lp);
#if CP_TAIL2 == 2
                if (left & 1) cp_slip_group<POWN, 1>(sl, pm, cdt, cn, n1, ginv, w, m, Lp);
#endif
                break;
            }
#endif
            cp_slip_group<POWN, U>(sl, pm, cdt, cn, n1, ginv, w, m, Lp);
        }
    }
    // Fe = G - G Lp
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            Fe[3 * i + j] = G[3 * i + j] - (G[3 * i] * Lp[j] + G[3 * i + 1] * Lp[3 + j] + G[3 * i + 2] * Lp[6 + j]);
    // E2x = 2 E = Fe^T Fe - I; the 1/2 rides on the elastic constants (a multiplication by 0.5 is exact, so
    // (C/2) (2E) == C E bit for bit) and the 2 C44 (1/2) of the shear rows cancels
    const double E0 = Fe[0] * Fe[0] + Fe[3] * Fe[3] + Fe[6] * Fe[6] - 1.0;
    const double E1 = Fe[1] * Fe[1] + Fe[4] * Fe[4] + Fe[7] * Fe[7] - 1.0;
    const double E2 = Fe[2] * Fe[2] + Fe[5] * Fe[5] + Fe[8] * Fe[8] - 1.0;
    const double E3 = Fe[1] * Fe[2] + Fe[4] * Fe[5] + Fe[7] * Fe[8];
    const double E4 = Fe[0] * Fe[2] + Fe[3] * Fe[5] + Fe[6] * Fe[8];
    const double E5 = Fe[0] * Fe[1] + Fe[3] * Fe[4] + Fe[6] * Fe[7];
    const double C11h = 0.5 * pm.C11, C12h = 0.5 * pm.C12;
    r[0] = s[0] - (C11h * E0 + C12h * (E1 + E2));
    r[1] = s[1] - (C11h * E1 + C12h * (E0 + E2));
    r[2] = s[2] - (C11h * E2 + C12h * (E0 + E1));
    r[3] = s[3] - pm.C44 * E3;
    r[4] = s[4] - pm.C44 * E4;
    r[5] = s[5] - pm.C44 * E5;
    return sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + 2.0 * (r[3] * r[3] + r[4] * r[4] + r[5] * r[5]));
}
#endif

// ---------------------------------------------------------------------------------------------------
// Newton matrix in scaled form, LU-factorised in place (no pivoting, see header comment).
//   N = C^-1 D^-1 + sum_a w_a e_a etilde_a^T,  e_a = voigt(sym(K d_a n_a^T)), K = Fe^T G,
//   etilde_a = voigt(sym(d_a n_a^T))   (strain-like: shear entries carry the 1/2)
// After the call N holds L (unit lower) and U; piv[i] = 1/U_ii.
// ---------------------------------------------------------------------------------------------------
template <int NS, class Arr, class GT>
CP_HD void cp_newton_matrix(const CpSlipRef& sl, const CpPointParams& pm, const GT& G, const double* Fe,
                            const Arr& w, unsigned mask, double* N /*36*/, double* piv /*6*/) {
    double K[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) K[3 * i + j] = Fe[i] * G[j] + Fe[3 + i] * G[3 + j] + Fe[6 + i] * G[6 + j];      // K = Fe^T G
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
#if CP_NM_ODD1
        const int nu = (w2[1] == 0.0 && a2[1] == a2[0]) ? 1 : 2;
#pragma unroll 1
        for (int u = 0; u < nu; ++u) {
#else
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#endif
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
        const double ip = cp_rcp(N[6 * k + k]);
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

// ---------------------------------------------------------------------------------------------------
// Certain rejections of the line search.  The reference's 'cut-half' search (models_copper.py:231-243) evaluates the
// residual at y + relax inc for relax = 1, 1/2, ... and rejects a trial whose norm is not below the accepted one.  With a
// rate exponent of 20 ... 120 most rejected trials overshoot the flow stress by more than 10 %, where the power law makes
// the residual astronomically large (median ||crt|| / ||r|| of the rejected trials of the 304-steel benchmark state:
// 4e9).  For such a trial the decision follows from x_a = tau_a / g_a alone (the first 84 of the ~330 FP64 instructions
// of an evaluation), by a chain of inequalities that holds for ANY signs and magnitudes of the other slip increments:
//   S : Lp = sum_a dgamma_a tau_a = sum_a |dgamma_a| |tau_a|          (dgamma_a and tau_a have the same sign)
//          >= cdt X^n . X g_min,                 X = max_a |x_a|, g_min = min_a g_a
//   ||Lp||_F >= (S : Lp) / ||S||_F =: L                                (Cauchy-Schwarz)
//   ||I - Lp||_F^2 = 3 + ||Lp||_F^2                                    (tr(d n^T) = 0)
//   ||Fe||_F >= sigma_min(G) ||I - Lp||_F,  sigma_min(G) >= gamma = 1 - ||G - I||_F
//   ||2E||_F = ||Fe^T Fe - I||_F >= ||Fe||_F^2 / sqrt(3) - sqrt(3)     (Fe^T Fe is positive semi-definite)
//   ||C : E||_F >= c_min ||E||_F,  c_min = min(C11 + 2 C12, C11 - C12, 2 C44)   (eigenvalues of a cubic C on symmetric E)
//   ||crt|| = ||S - C : E|| >= c_min (gamma^2 L^2 - 3 (1 - gamma^2)) / (2 sqrt(3)) - ||S||
// so the trial is rejected (||crt|| >= ||r||) as soon as X^(n+1) >= T = L_req cap / (cdt g_min), where
//   L_req^2 = (2 sqrt(3) . 1.5 . 3 cap / c_min + 3 (1 - gamma^2)) / gamma^2
// and `cap` bounds both ||r|| of the accepted iterate (with a factor 2 to spare for the rounding of the literal
// evaluation) and ||S|| of the trial (<= ||y|| + ||inc||).  cap = 4 ||r(0)|| is fixed per point, so the (n+1)-th root is
// taken once per point, before the Newton loop (its code and registers stay out of the loop); the Newton loop checks the two conditions on `cap` once per iteration and compares the high word of
// X with one integer per trial.  Trials beyond X_cap (about 1.5 X_rej) are evaluated as before: far enough out the literal
// evaluation overflows to inf - inf = NaN, which the reference's comparison ACCEPTS.  A skipped trial changes nothing
// but the time: the decision, the iterates and the evaluation count (`evals` still counts it) are the reference's.
// At the 304-steel benchmark state 3.9 of the 17.1 evaluations per point are skipped (60 % of the rejected ones).
// ---------------------------------------------------------------------------------------------------
struct CpPrune {
    int lo;         // high word of X_rej, rounded up; 0x7fffffff: nothing is skipped
    int capw;       // high word of cap: hi(v) < capw implies v < cap
};
template <int NS, class Arr>
CP_HD void cp_prune_setup(const CpSlipRef& sl, const CpPointParams& pm, double cdt, const double* G, const Arr& ginv, CpPrune& pr) {
    pr.lo = 0x7fffffff; pr.capw = 0;
#if CP_PRUNE
    double gi = 0.0;
#pragma unroll 4
    for (int a = 0; a < NS; ++a) gi = ginv[a] > gi ? ginv[a] : gi;
    const double d0 = G[0] - 1.0, d4 = G[4] - 1.0, d8 = G[8] - 1.0;
    const double gam = 1.0 - sqrt(d0 * d0 + d4 * d4 + d8 * d8 + G[1] * G[1] + G[2] * G[2] + G[3] * G[3] + G[5] * G[5] + G[6] * G[6] + G[7] * G[7]);
    double cmin = pm.C11 + 2.0 * pm.C12;
    cmin = (pm.C11 - pm.C12) < cmin ? (pm.C11 - pm.C12) : cmin;
    cmin = (2.0 * pm.C44) < cmin ? (2.0 * pm.C44) : cmin;
    // ||r(0)|| = ||C : 1/2 (G^T G - I)|| (what the first evaluation of the solve returns; any value of that size will do here)
    double rn0;
    {
        const double E0 = G[0] * G[0] + G[3] * G[3] + G[6] * G[6] - 1.0, E1 = G[1] * G[1] + G[4] * G[4] + G[7] * G[7] - 1.0;
        const double E2 = G[2] * G[2] + G[5] * G[5] + G[8] * G[8] - 1.0;
        const double E3 = G[1] * G[2] + G[4] * G[5] + G[7] * G[8], E4 = G[0] * G[2] + G[3] * G[5] + G[6] * G[8];
        const double E5 = G[0] * G[1] + G[3] * G[4] + G[6] * G[7];
        const double a = 0.5 * pm.C11, b = 0.5 * pm.C12;
        const double r0 = a * E0 + b * (E1 + E2), r1 = a * E1 + b * (E0 + E2), r2 = a * E2 + b * (E0 + E1);
        rn0 = sqrt(r0 * r0 + r1 * r1 + r2 * r2 + 2.0 * (pm.C44 * pm.C44) * (E3 * E3 + E4 * E4 + E5 * E5));
    }
    const double cap = 4.0 * rn0;
    const double g2 = gam * gam;
    const double Lreq2 = (15.588457268119896 * cap / cmin + 3.0 * (1.0 - g2)) / g2;       // 2 sqrt(3) . 1.5 . 3 = 15.588...
    const double T = sqrt(Lreq2) * cap * gi / cdt;
    // every quantity must be an ordinary positive number, the table a slip table (d . n = 0), and cdt (1.5 X_rej)^n far from
    // overflow: cdt X_rej^n < T cdt <= 1e30 and 1.5^n <= 1e70 for n <= 400
    if (!(gam >= 0.5) || !(cmin > 0.0) || !(cap > 0.0) || !(gi > 0.0) || !(T < 1e300) || !(T * cdt < 1e30) || !(pm.n_exp > 1.0) ||
        !(pm.n_exp <= 400.0) || !(sl.u->sys[0].pad[0] < 1e-12))
        return;
    const double X = (T > 1.0) ? exp(log(T) / (pm.n_exp + 1.0)) * (1.0 + 1e-6) : 1.0;
    pr.lo = cp_hi_abs(X) + 1;
    pr.capw = cp_hi_abs(cap);
#endif
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

#if CP_PEEL
// The same solve with its first evaluation peeled off.  At y = 0 (rate exponent > 1) every tau, dgamma and w vanishes, so the
// evaluation is Fe = G, Lp = 0, r = -C : 1/2 (G^T G - I) and the first Newton step is inc = -r: done before the loop, which then
// needs neither the `first` flag nor the zero-stress path of cp_residual.  CP_PEEL == 2 also decides the leading trials of
// the first line search before the loop: they sit at relax inc with relax = 1, 1/2, ..., so tau_a / g_a of trial j is
// 2^-j x that of the full step EXACTLY, and the chain of inequalities of cp_prune_setup (||r|| = ||r(0)|| and
// ||S|| <= ||inc|| = ||r(0)|| hold by construction) tells which of them are certainly rejected: their evaluations are
// counted and skipped.  This is per lane - a lane only enters the loop further along its own path - and bitwise neutral.
template <int NS, int POWN, class Arr>
CP_HD void cp_newton_p(const CpSlipRef& sl, const CpPointParams& pm, double cdt, double tol, int max_sub, int max_iter,
                       const double* G, const Arr& ginv, const Arr& w, double* s, double* Fe, double* Lp,
                       unsigned& mask, unsigned& mact, CpSolveInfo& info) {
    double r[6], st[6], inc[6];
    info.iters = 0; info.evals = 0; info.status = 0;
    double rn, relax = 1.0;
    int sub = 0;
    bool go = true;
    if (pm.n_exp > 1.0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) { Lp[i] = 0.0; Fe[i] = G[i]; }
        mask = 0u; mact = 0u;
        const double E0 = Fe[0] * Fe[0] + Fe[3] * Fe[3] + Fe[6] * Fe[6] - 1.0;
        const double E1 = Fe[1] * Fe[1] + Fe[4] * Fe[4] + Fe[7] * Fe[7] - 1.0;
        const double E2 = Fe[2] * Fe[2] + Fe[5] * Fe[5] + Fe[8] * Fe[8] - 1.0;
        const double E3 = Fe[1] * Fe[2] + Fe[4] * Fe[5] + Fe[7] * Fe[8];
        const double E4 = Fe[0] * Fe[2] + Fe[3] * Fe[5] + Fe[6] * Fe[8];
        const double E5 = Fe[0] * Fe[1] + Fe[3] * Fe[4] + Fe[6] * Fe[7];
        const double C11h = 0.5 * pm.C11, C12h = 0.5 * pm.C12;
        r[0] = 0.0 - (C11h * E0 + C12h * (E1 + E2));
        r[1] = 0.0 - (C11h * E1 + C12h * (E0 + E2));
        r[2] = 0.0 - (C11h * E2 + C12h * (E0 + E1));
        r[3] = 0.0 - pm.C44 * E3;
        r[4] = 0.0 - pm.C44 * E4;
        r[5] = 0.0 - pm.C44 * E5;
        rn = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + 2.0 * (r[3] * r[3] + r[4] * r[4] + r[5] * r[5]));
        info.evals = 1;
#pragma unroll
        for (int i = 0; i < 6; ++i) s[i] = 0.0;
        if (!(rn > tol)) go = false;
        else if (info.iters >= max_iter) { info.status |= 1; go = false; }
        else {
#pragma unroll
            for (int i = 0; i < 6; ++i) inc[i] = -r[i];
#if CP_PEEL >= 2
            // which of the trials relax = 1, 1/2, ... are certainly rejected (see cp_prune_setup for the inequalities)
            {
                double gi = 0.0, X1 = 0.0;
                const double i3 = inc[3] + inc[3], i4 = inc[4] + inc[4], i5 = inc[5] + inc[5];
#pragma unroll
                for (int a = 0; a < NS; ++a) {
                    const CpSlipSys& y = sl.u->sys[a];
                    const double tau = y.Et[0] * inc[0] + y.Et[1] * inc[1] + y.Et[2] * inc[2] + y.Et[3] * i3 + y.Et[4] * i4 + y.Et[5] * i5;
                    const double ga = ginv[a];
                    const double x = fabs(tau * ga);
                    X1 = x > X1 ? x : X1;
                    gi = ga > gi ? ga : gi;
                }
                const double d0 = G[0] - 1.0, d4 = G[4] - 1.0, d8 = G[8] - 1.0;
                const double gam = 1.0 - sqrt(d0 * d0 + d4 * d4 + d8 * d8 + G[1] * G[1] + G[2] * G[2] + G[3] * G[3] + G[5] * G[5] + G[6] * G[6] + G[7] * G[7]);
                double cmin = pm.C11 + 2.0 * pm.C12;
                cmin = (pm.C11 - pm.C12) < cmin ? (pm.C11 - pm.C12) : cmin;
                cmin = (2.0 * pm.C44) < cmin ? (2.0 * pm.C44) : cmin;
                // here ||r|| of the accepted iterate = ||S|| bound = rn (the trial stresses are relax inc, ||inc|| = rn), so the
                // requirement c_min (gam^2 L^2 - 3 (1 - gam^2)) / (2 sqrt 3) >= 2 ||r|| + ||S|| reads, with 1.5 to spare,
                const double g2 = gam * gam;
                const double Lreq2 = (15.588457268119896 * rn / cmin + 3.0 * (1.0 - g2)) / g2;       // 2 sqrt(3) . 1.5 . 3
                // L >= cdt X^(n+1) g_min / ||S||, ||S|| <= relax rn  =>  certain if X^(n+1) >= L_req relax rn / (cdt g_min); the
                // left side of trial j is X1^(n+1) 2^(-j (n+1)), the right side T 2^-j
                double T = sqrt(Lreq2) * rn * gi / cdt * (1.0 + 1e-6);
                double ax[1] = {X1}, pw[1];
                cp_rate_pow<POWN, 1>(ax, pm.n_exp - 1.0, pw);
                double pX = pw[0] * X1 * X1;                                                    // X1^(n+1)
                const double hn = exp2(-(pm.n_exp + 1.0));
                const bool usable = (gam >= 0.5) && (cmin > 0.0) && (gi > 0.0) && (T < 1e300) && (sl.u->sys[0].pad[0] < 1e-12);
                // a trial whose slip increments could overflow the literal evaluation (inf - inf = NaN, which the reference's
                // comparison ACCEPTS) is evaluated: cdt X^n < 1e60 <= X^(n+1) < 1e60 X / cdt, and X >= 1 wherever pX >= T > 1
                while (usable && sub + 1 < max_sub && pX >= T && T > 1.0 && pX * cdt < 1e60) {
                    CP_TRACE_PRUNE();
                    relax *= 0.5;
                    ++sub;
                    ++info.evals;
                    pX *= hn;
                    T *= 0.5;
                }
            }
#endif
#pragma unroll
            for (int i = 0; i < 6; ++i) st[i] = 0.0 + relax * inc[i];      // s + relax inc with s = +0.0, like the loop forms it
        }
    } else {
        // rate exponent <= 1: the evaluation at y = 0 is an ordinary one; a NaN norm makes the loop accept it whatever it gives
#pragma unroll
        for (int i = 0; i < 6; ++i) { s[i] = 0.0; st[i] = 0.0; inc[i] = 0.0; }
        rn = nan("");
        info.iters = -1;
    }
    while (go) {
        const double crtn = cp_residual<NS, POWN>(sl, pm, cdt, G, ginv, w, st, false, r, Fe, Lp, mask, mact);
        ++info.evals;
        relax *= 0.5;
        ++sub;
        if (crtn >= rn && sub < max_sub) {               // line search: next trial y + relax inc
#pragma unroll
            for (int i = 0; i < 6; ++i) st[i] = s[i] + relax * inc[i];
            continue;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) s[i] = st[i];        // accept: y + 2 relax inc == the point just evaluated
        ++info.iters;
        rn = crtn;
        if (!(rn > tol)) break;
        if (info.iters >= max_iter) { info.status |= 1; break; }
        {
            double N[36], piv[6];
            cp_newton_matrix<NS>(sl, pm, G, Fe, w, mact, N, piv);
            cp_compliance_neg(pm, r, inc);
            cp_lu_solve(N, piv, inc);
            inc[3] *= 0.5; inc[4] *= 0.5; inc[5] *= 0.5;          // inc = D^-1 z
        }
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
#endif

#if CP_PRUNE || CP_LS_SMEM || CP_G_SMEM
// The same solve with the experimental switches: certainly rejected trials skipped (CP_PRUNE), loop state in shared memory
// (CP_LS_SMEM).  Results are bitwise those of cp_newton (tests/test_prune_option.py).
template <int NS, int POWN, class Arr>
CP_HD void cp_newton_x(const CpSlipRef& sl, const CpPointParams& pm, double cdt, double tol, int max_sub, int max_iter,
                     double* G, const Arr& ginv, const Arr& w, double* s, double* Fe, double* Lp,
                     unsigned& mask, unsigned& mact, CpSolveInfo& info) {
    // Loop state that is touched once per evaluation - the accepted iterate y, the Newton increment, the three integers of
    // the pruning - lives in per-thread columns of shared memory on the device: the loop sits exactly at the register
    // budget of 3 blocks per SM (168), and every further live value makes the compiler spill a dozen doubles around the
    // Newton matrix (three more integers: 57.7 -> 61.4 ms at 200^3).
#if defined(__CUDA_ARCH__) && CP_LS_SMEM
    __shared__ double cp_ls_s[12 * CP_BLOCK_THREADS];
    double* const lss = cp_ls_s + threadIdx.x;
#define CP_Y(i) lss[(i) * CP_BLOCK_THREADS]
#define CP_INC(i) lss[(6 + (i)) * CP_BLOCK_THREADS]
#else
    double lss[12];
#define CP_Y(i) lss[i]
#define CP_INC(i) lss[6 + (i)]
#endif
#if defined(__CUDA_ARCH__) && CP_PRUNE
    __shared__ int cp_prune_s[3 * CP_BLOCK_THREADS];
    int* const prs = cp_prune_s + threadIdx.x;
#define CP_PRS(i) prs[(i) * CP_BLOCK_THREADS]
#else
    int prs[3];
#define CP_PRS(i) prs[i]
#endif
#if defined(__CUDA_ARCH__) && CP_G_SMEM
    __shared__ double cp_g_s[9 * CP_BLOCK_THREADS];
#pragma unroll
    for (int i = 0; i < 9; ++i) cp_g_s[i * CP_BLOCK_THREADS + threadIdx.x] = G[i];
    const CpVec<CP_BLOCK_THREADS> Gv = {cp_g_s + threadIdx.x};
#else
    const CpVec<1> Gv = {G};
#endif
    double r[6], st[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { CP_Y(i) = 0.0; st[i] = 0.0; CP_INC(i) = 0.0; }
    const bool zero_ok = pm.n_exp > 1.0;
    double rn = 0.0, relax = 1.0;
    int sub = 0;
    bool first = true;
    info.iters = 0; info.evals = 0; info.status = 0;
    {
        CpPrune pr;
        cp_prune_setup<NS>(sl, pm, cdt, G, ginv, pr);
        CP_PRS(0) = pr.lo; CP_PRS(1) = pr.capw; CP_PRS(2) = 0x7fffffff;
    }
    for (;;) {
        unsigned act = 0u;
        if (!(first && zero_ok)) {
            for (;;) {
                int hmax;
                act = cp_slip_ratios<NS>(sl, pm, ginv, w, st, hmax);
#if CP_PRUNE
                // A trial whose rejection is certain (cp_prune_setup) is not evaluated any further - unless it is the last one
                // the search allows, which is accepted whatever its norm.  Decided per warp: the evaluation is shared.
                const int plo = (sub + 1 < max_sub) ? CP_PRS(2) : 0x7fffffff;
                if (!cp_warp_all((unsigned)(hmax - plo) < CP_PRUNE_SPAN)) break;
                CP_TRACE_PRUNE();
                ++info.evals;
                relax *= 0.5;
                ++sub;
#pragma unroll
                for (int i = 0; i < 6; ++i) st[i] = CP_Y(i) + relax * CP_INC(i);
#else
                break;
#endif
            }
        }
        const double crtn = cp_residual_x<NS, POWN>(sl, pm, cdt, Gv, ginv, w, st, first && zero_ok, act, r, Fe, Lp, mask, mact);
        ++info.evals;
        if (!first) {
            relax *= 0.5;
            ++sub;
            if (crtn >= rn && sub < max_sub) {           // line search: next trial y + relax inc
#pragma unroll
                for (int i = 0; i < 6; ++i) st[i] = CP_Y(i) + relax * CP_INC(i);
                continue;
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) CP_Y(i) = st[i];    // accept: y + 2 relax inc == the point just evaluated
            ++info.iters;
        }
        rn = crtn;
        if (!(rn > tol)) break;
        if (info.iters >= max_iter) { info.status |= 1; break; }
        double inc[6];
        if (first && zero_ok) {
#pragma unroll
            for (int i = 0; i < 6; ++i) inc[i] = -r[i];
        } else {
            double N[36], piv[6];
            cp_newton_matrix<NS>(sl, pm, Gv, Fe, w, mact, N, piv);
            cp_compliance_neg(pm, r, inc);
            cp_lu_solve(N, piv, inc);
            inc[3] *= 0.5; inc[4] *= 0.5; inc[5] *= 0.5;          // inc = D^-1 z
        }
        first = false;
        relax = 1.0;
        sub = 0;
        // y is the point just evaluated (st) or, before the first step, 0 (also st)
#if CP_PRUNE
        {
            // ||y + relax inc||_F <= ||y||_F + ||inc||_F <= the weighted 1-norms below (sqrt(2) < 1.5)
            const double sc = (fabs(st[0]) + fabs(st[1]) + fabs(st[2]) + fabs(inc[0]) + fabs(inc[1]) + fabs(inc[2])) +
                              1.5 * (fabs(st[3]) + fabs(st[4]) + fabs(st[5]) + fabs(inc[3]) + fabs(inc[4]) + fabs(inc[5]));
            const int capw = CP_PRS(1);
            CP_PRS(2) = (cp_hi_abs(rn) < capw && cp_hi_abs(sc) < capw) ? CP_PRS(0) : 0x7fffffff;
        }
#endif
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            CP_INC(i) = inc[i];
            st[i] += inc[i];
        }
    }
    if (!(rn == rn)) info.status |= 2;
#pragma unroll
    for (int i = 0; i < 6; ++i) s[i] = st[i];            // the accepted iterate is the point of the last evaluation
#if defined(__CUDA_ARCH__) && CP_G_SMEM
#pragma unroll
    for (int i = 0; i < 9; ++i) G[i] = Gv[i];            // the caller's copy was dead during the loop (no registers held)
#endif
    // systems outside the last processed set: w = 0 (the output stages read w of every system)
#pragma unroll 4
    for (int a = 0; a < NS; ++a)
        if (!((mask >> a) & 1u)) w[a] = 0.0;
#undef CP_PRS
#undef CP_Y
#undef CP_INC
}
#endif

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
    for (int a = 0; a < NS; ++a) ps.ginv[a] = cp_rcp(g[a]);
    ps.cdt = mat.ao * dt;
#if CP_PEEL
    cp_newton_p<NS, POWN>(sl, pm, ps.cdt, mat.tol, mat.max_sub_step, mat.max_iter, ps.G, ps.ginv, ps.w, ps.s, ps.Fe, ps.Lp,
                          ps.mask, ps.mact, ps.info);
#elif CP_PRUNE || CP_LS_SMEM || CP_G_SMEM
    cp_newton_x<NS, POWN>(sl, pm, ps.cdt, mat.tol, mat.max_sub_step, mat.max_iter, ps.G, ps.ginv, ps.w, ps.s, ps.Fe, ps.Lp,
                          ps.mask, ps.mact, ps.info);
#else
    cp_newton<NS, POWN>(sl, pm, ps.cdt, mat.tol, mat.max_sub_step, mat.max_iter, ps.G, ps.ginv, ps.w, ps.s, ps.Fe, ps.Lp,
                        ps.mask, ps.mact, ps.info);
#endif
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
    const double inv_n = cp_rcp(pm.n_exp), inv_tsat = cp_rcp(pm.t_sat);
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

// First Piola-Kirchhoff stress (lab): P = Fe S A_new^T / det A_new  (== det F sigma F^-T, :160-161), computed as
// P = (R Fe) S (R A_new)^T / det A_new so that the two mixed-frame matrices the tangent needs come for free.
struct CpStressAux {
    double RFe[9];     // R Fe     (lab row, crystal column)
    double RAn[9];     // R A_new  (lab row, crystal column)
    double idet;       // 1/det(A_new)
};

template <class Arr>
CP_HD void cp_point_stress(const CpPointState<Arr>& ps, const double* R, double* P_lab, CpStressAux& ax) {
    {
        double ImL[9], Anc[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) ImL[i] = -ps.Lp[i];
        ImL[0] += 1.0; ImL[4] += 1.0; ImL[8] += 1.0;
        m3_mul(ps.Ac, ImL, Anc);
        ax.idet = cp_rcp(m3_det(Anc));
        m3_mul(R, Anc, ax.RAn);
    }
    m3_mul(R, ps.Fe, ax.RFe);
    double S[9], T[9];
    sym6_to_m3(ps.s, S);
    m3_mul(ax.RFe, S, T);
    m3_mul_nt(T, ax.RAn, P_lab);
#pragma unroll
    for (int i = 0; i < 9; ++i) P_lab[i] *= ax.idet;
}

// ---------------------------------------------------------------------------------------------------
// Consistent tangent dP_ij/dH_kl in the lab frame (times `scale`: JxW for the element integration).  Reproduces jacfwd through the local solve (f_jvp, models_copper.py:251-259):
//   ds/dF = -J^-1 dr/dF,  dr/dF[dF] = -C : sym(Fe^T dF A_new)   =>   z = D ds = N^-1 b,  b = voigt(sym(u_k v_l^T)),
//   u_k = row k of R Fe, v_l = row l of R A_new   (direction dF = e_k e_l^T in the lab frame),
//   dP = dF Z + [ dFe S A_new^T + Fe dS A_new^T + Fe S dA_new^T ] / det - P tr(A_new^-1 dA_new),
//   dLp = sum_a w_a (etilde_a . z) d_a n_a^T,  dA_new = -Ac dLp,  dFe = -G dLp,  Z = A_new S A_new^T / det.
// Everything after z is linear in z, so   dP_ij(kl) = delta_ik Zlab_lj + W_ij . z(kl)   with the 9 x 6 matrix
//   W_ij,m = [ (R Fe) E_m (R A_new)^T ]_ij / det + sum_a w_a etilde_a[m] Q_a,ij        (E_m: unit dS of z_m)
//   Q_a    = -[ (RFe Y d_a)(RAn S n_a)^T + (RFe S n_a)(RAn Y d_a)^T ] / det + P (n_a . Y d_a),   Y = (I - Lp)^-1
// (G = Fe Y, Ac = A_new Y; every slip system enters through four 3-vectors because M_a = d_a n_a^T has rank one), and
//   W_ij . N^-1 b(kl) = (N^-T W_ij) . b(kl) = [ (R Fe) X_ij (R A_new)^T ]_kl,   X_ij = symmetric 3x3 of y = N^-T W_ij,
// i.e. nine transposed 6x6 solves and nine pairs of 3x3 products replace the nine solve + push-forward passes of the
// direct form.  The work is done one row i of P at a time (18 accumulators), and the LU factors of N are parked in
// `park` (CP_TANGENT_PARK doubles: shared memory in the kernels) while the slip loop runs, so the routine needs no
// local-memory spills.
// ---------------------------------------------------------------------------------------------------
#define CP_TANGENT_PARK 42
CP_HD double cp_sel3(int i, double a0, double a1, double a2) { return i == 0 ? a0 : (i == 1 ? a1 : a2); }
// Step 1 (call right after the local solve, before the stress: G and the matrix die before R Fe, R A_new, P are born):
// Newton matrix at the converged state, LU-factorised, parked.
template <int NS, class Arr, class Park>
CP_HD void cp_point_tangent_factor(const CpSlipRef& sl, const CpPointParams& pm, const CpPointState<Arr>& ps, const Park& park) {
    double N[36], piv[6];
    cp_newton_matrix<NS>(sl, pm, ps.G, ps.Fe, ps.w, ps.mact, N, piv);
#pragma unroll
    for (int i = 0; i < 36; ++i) park[i] = N[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) park[36 + i] = piv[i];
}
// Step 2 (after cp_point_stress): the tangent itself.
// `store(ij, kl, v)` receives dP_ij/dH_kl * scale (the kernels pass a streaming global store, the host check an array).
template <int NS, class Arr, class Park, class Store>
CP_HD void cp_point_tangent(const CpSlipRef& sl, const CpPointState<Arr>& ps, const CpStressAux& ax, const double* P_lab,
                            double scale, const Park& park, const Store& store) {
    double Y[9], Zl[9];
    {
        double tmp[9], dY;
#pragma unroll
        for (int i = 0; i < 9; ++i) tmp[i] = -ps.Lp[i];
        tmp[0] += 1.0; tmp[4] += 1.0; tmp[8] += 1.0;
        m3_inv(tmp, Y, &dY);
        double S[9], T[9];
        sym6_to_m3(ps.s, S);
        m3_mul_nt(S, ax.RAn, T);
        m3_mul(ax.RAn, T, Zl);              // R Z R^T = (R A_new) S (R A_new)^T / det
        const double zs = ax.idet * scale;
#pragma unroll
        for (int i = 0; i < 9; ++i) Zl[i] *= zs;
    }
    const double* s = ps.s;
    const double idet = ax.idet;
#pragma unroll 1
    for (int i = 0; i < 3; ++i) {
        double W[3][6];
        // row i of R Fe and of P, picked with selects (a dynamic index would push the arrays into local memory)
        const double ui0 = cp_sel3(i, ax.RFe[0], ax.RFe[3], ax.RFe[6]), ui1 = cp_sel3(i, ax.RFe[1], ax.RFe[4], ax.RFe[7]),
                     ui2 = cp_sel3(i, ax.RFe[2], ax.RFe[5], ax.RFe[8]);
        const double Pi[3] = {cp_sel3(i, P_lab[0], P_lab[3], P_lab[6]), cp_sel3(i, P_lab[1], P_lab[4], P_lab[7]),
                              cp_sel3(i, P_lab[2], P_lab[5], P_lab[8])};
        {   // elastic part: [ RFe E_m RAn^T ]_ij / det
            const double f0 = idet * ui0, f1 = idet * ui1, f2 = idet * ui2;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const double v0 = ax.RAn[3 * j], v1 = ax.RAn[3 * j + 1], v2 = ax.RAn[3 * j + 2];
                W[j][0] = f0 * v0; W[j][1] = f1 * v1; W[j][2] = f2 * v2;
                W[j][3] = 0.5 * (f1 * v2 + f2 * v1); W[j][4] = 0.5 * (f0 * v2 + f2 * v0); W[j][5] = 0.5 * (f0 * v1 + f1 * v0);
            }
        }
        for (unsigned m = ps.mact; m; m &= m - 1u) {        // slip part: sum_a w_a etilde_a[m] Q_a,ij
            const int a = cp_ffs0(m);
            const CpSlipSys& y = sl.d->sys[a];
            const double d0 = y.d[0], d1 = y.d[1], d2 = y.d[2], n0 = y.n[0], n1 = y.n[1], n2 = y.n[2];
            const double yd0 = Y[0] * d0 + Y[1] * d1 + Y[2] * d2, yd1 = Y[3] * d0 + Y[4] * d1 + Y[5] * d2,
                         yd2 = Y[6] * d0 + Y[7] * d1 + Y[8] * d2;
            const double sn0 = s[0] * n0 + s[5] * n1 + s[4] * n2, sn1 = s[5] * n0 + s[1] * n1 + s[3] * n2,
                         sn2 = s[4] * n0 + s[3] * n1 + s[2] * n2;
            const double c = n0 * yd0 + n1 * yd1 + n2 * yd2;
            const double gd = -idet * (ui0 * yd0 + ui1 * yd1 + ui2 * yd2);
            const double t2 = -idet * (ui0 * sn0 + ui1 * sn1 + ui2 * sn2);
            double q[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const double v0 = ax.RAn[3 * j], v1 = ax.RAn[3 * j + 1], v2 = ax.RAn[3 * j + 2];
                q[j] = gd * (v0 * sn0 + v1 * sn1 + v2 * sn2) + t2 * (v0 * yd0 + v1 * yd1 + v2 * yd2) + Pi[j] * c;
            }
            const double wa = ps.w[a];
#pragma unroll
            for (int mm = 0; mm < 6; ++mm) {
                const double cf = wa * y.Et[mm];
#pragma unroll
                for (int j = 0; j < 3; ++j) W[j][mm] += cf * q[j];
            }
        }
        // y_j = N^-T W_j for the three j at once (N = L U parked; every entry is read once per row i):
        //   U^T t = w (forward substitution),  L^T y = t (backward, unit diagonal)
#pragma unroll
        for (int r = 0; r < 6; ++r) {
#pragma unroll
            for (int k = 0; k < r; ++k) {
                const double u = park[6 * k + r];
#pragma unroll
                for (int j = 0; j < 3; ++j) W[j][r] -= u * W[j][k];
            }
            const double ip = park[36 + r];
#pragma unroll
            for (int j = 0; j < 3; ++j) W[j][r] *= ip;
        }
#pragma unroll
        for (int r = 4; r >= 0; --r) {
#pragma unroll
            for (int k = r + 1; k < 6; ++k) {
                const double l = park[6 * k + r];
#pragma unroll
                for (int j = 0; j < 3; ++j) W[j][r] -= l * W[j][k];
            }
        }
        // dP_ij(kl) = scale [ RFe X_ij RAn^T ]_kl + delta_ik scale Zlab_lj,  X = sym3(y): shear entries carry the 1/2 of b
        const double hs = 0.5 * scale;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double x00 = W[j][0] * scale, x11 = W[j][1] * scale, x22 = W[j][2] * scale;
            const double x12 = W[j][3] * hs, x02 = W[j][4] * hs, x01 = W[j][5] * hs;
            double T[9];                 // T[p][l] = sum_q X[p][q] RAn[l][q]
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                const double v0 = ax.RAn[3 * l], v1 = ax.RAn[3 * l + 1], v2 = ax.RAn[3 * l + 2];
                T[l] = x00 * v0 + x01 * v1 + x02 * v2;
                T[3 + l] = x01 * v0 + x11 * v1 + x12 * v2;
                T[6 + l] = x02 * v0 + x12 * v1 + x22 * v2;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int l = 0; l < 3; ++l) {
                    double v = ax.RFe[3 * k] * T[l] + ax.RFe[3 * k + 1] * T[3 + l] + ax.RFe[3 * k + 2] * T[6 + l];
                    if (k == i) v += Zl[3 * l + j];
                    store(3 * i + j, 3 * k + l, v);
                }
        }
    }
}
