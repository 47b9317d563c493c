// TEST INFRASTRUCTURE ONLY: compiles the product's per-point header (jax-cpfem_b200/csrc/cp_point.cuh) for the
// host so that the hand-derived algebra can be compared with the autodiff oracle on a machine without a GPU.
// Nothing in the product path loads this library.  bench.py's `cpu_baseline.same_algorithm` leg times it (built with
// -fopenmp) as "the GPU path's own algorithm on the host cores", next to the oracle port of the reference's algorithm.
#include <stdint.h>
#include <string.h>
#include "../../jax-cpfem_b200/csrc/cp_point.cuh"
#include "../../jax-cpfem_b200/csrc/cp_adjoint.cuh"

typedef CpArr<1> HArr;

// POWN: 0 = run-time rate exponent (what the kernels use for per-point / non-integer exponents),
//       9 / 19 / 119 = the compile-time integer chains the kernels instantiate for copper / DP steel / 304 steel.
template <int NS, int POWN>
static void run(const double* slip6, const CpMaterial* mat, double dt, int64_t np, const double* H, const double* A,
                const double* g, const double* slip_old, const double* R, const double* pp /* np x 8 or null */,
                double* P, double* tangent, double* A_new, double* g_new, double* slip_new, int32_t* iters) {
    CpSlip table;
    cp_slip_init(&table, slip6, NS);
    const CpSlipRef sl = {&table, &table};
    // points are independent: the OpenMP build (bench.py's same-algorithm CPU baseline) spreads them over the host cores
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t p = 0; p < np; ++p) {
        CpPointParams pm;
        if (pp) {
            const double* q = pp + 8 * p;   // C11 C12 C44 h t_sat gss_a xm r
            cp_params_elastic(pm, q[0], q[1], q[2], q[6]);
            pm.h = q[3]; pm.t_sat = q[4]; pm.gss_a = q[5]; pm.r = q[7];
        } else {
            cp_params_elastic(pm, mat->C11, mat->C12, mat->C44, mat->xm);
            pm.h = mat->h; pm.t_sat = mat->t_sat; pm.gss_a = mat->gss_a; pm.r = mat->r;
        }
        double ginv[NS], w[NS];
        CpPointState<HArr> ps;
        ps.ginv.p = ginv; ps.w.p = w;
        cp_point_solve<NS, POWN>(sl, *mat, pm, dt, H + 9 * p, A + 9 * p, g + NS * p, R + 9 * p, ps);
        cp_point_frame(A + 9 * p, R + 9 * p, ps);
        if (iters) { iters[3 * p] = ps.info.iters; iters[3 * p + 1] = ps.info.evals; iters[3 * p + 2] = ps.info.status; }
        double parkbuf[CP_TANGENT_PARK];
        HArr park; park.p = parkbuf;
        if (tangent) cp_point_tangent_factor<NS>(sl, pm, ps, park);
        CpStressAux ax;
        cp_point_stress(ps, R + 9 * p, P + 9 * p, ax);
        if (tangent) {
            double* T = tangent + 81 * p;
            cp_point_tangent<NS>(sl, ps, ax, P + 9 * p, 1.0, park, [T](int ij, int kl, double v) { T[9 * ij + kl] = v; });
        }
        // last: the state update overwrites ps.w
        if (A_new) cp_point_state_update<NS>(sl, pm, ps, g + NS * p, slip_old + NS * p, R + 9 * p, A_new + 9 * p, g_new + NS * p, slip_new + NS * p);
    }
}

#ifdef _OPENMP
#include <omp.h>
extern "C" int hostcheck_threads(int n) { if (n > 0) omp_set_num_threads(n); return omp_get_max_threads(); }
#else
extern "C" int hostcheck_threads(int) { return 1; }
#endif

extern "C" int hostcheck_points(int ns, int pown, const double* slip6, const CpMaterial* mat, double dt, int64_t np, const double* H,
                                const double* A, const double* g, const double* slip_old, const double* R, const double* pp,
                                double* P, double* tangent, double* A_new, double* g_new, double* slip_new, int32_t* iters) {
#define RUN(NS, PW) run<NS, PW>(slip6, mat, dt, np, H, A, g, slip_old, R, pp, P, tangent, A_new, g_new, slip_new, iters)
    if (ns == 12 && pown == 0) RUN(12, 0);
    else if (ns == 12 && pown == 9) RUN(12, 9);
    else if (ns == 12 && pown == 119) RUN(12, 119);
    else if (ns == 24 && pown == 0) RUN(24, 0);
    else if (ns == 24 && pown == 19) RUN(24, 19);
    else return -1;
#undef RUN
    return 0;
}

// jac_x / jac_y / explicit dP columns of the adjoint header at given (x, y): per point jac_x (9, nx), jac_y (9, 9),
// dPdx (9, nx), dPdS (9, 9), row-major; nx = 27 + 2 ns (+5 with_params, +81 with_C).  pp: (np, 4) = C11 C12 C44 xm.
template <int NS>
static void run_jac(const double* slip6, double cdt, int64_t np, const double* H, const double* A, const double* g, const double* R,
                    const double* pp, const double* S, int with_params, int with_C, double* jac_x, double* jac_y, double* dPdx,
                    double* dPdS) {
    CpSlip table;
    cp_slip_init(&table, slip6, NS);
    const int nd = 27 + 2 * NS + (with_params ? 5 : 0);
    const int nx = nd + (with_C ? 81 : 0);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t p = 0; p < np; ++p) {
        const double* q = pp + 4 * p;
        for (int c = 0; c < nd; ++c) {
            double dr[9], dP[9];
            cp_jac_x_column<NS>(table, cdt, H + 9 * p, A + 9 * p, g + NS * p, R + 9 * p, q[3], q[0], q[1], q[2], S + 9 * p, c, dr, dP);
            for (int i = 0; i < 9; ++i) { jac_x[(p * 9 + i) * nx + c] = dr[i]; dPdx[(p * 9 + i) * nx + c] = dP[i]; }
        }
        if (with_C) {
            double r[9], E[9], Eh[9];
            cp_ref_residual<NS, double>(table, cdt, H + 9 * p, A + 9 * p, g + NS * p, R + 9 * p, q[3], q[0], q[1], q[2], S + 9 * p, r, nullptr, E);
            cp_jac_C_prepare(R + 9 * p, E, Eh);
            for (int i = 0; i < 9; ++i)
                for (int c = 0; c < 81; ++c) { jac_x[(p * 9 + i) * nx + nd + c] = cp_jac_C_entry(R + 9 * p, Eh, i, c); dPdx[(p * 9 + i) * nx + nd + c] = 0.0; }
        }
        for (int m = 0; m < 9; ++m) {
            double dr[9], dP[9];
            cp_jac_y_column<NS>(table, cdt, H + 9 * p, A + 9 * p, g + NS * p, R + 9 * p, q[3], q[0], q[1], q[2], S + 9 * p, m, dr, dP);
            for (int i = 0; i < 9; ++i) { jac_y[(p * 9 + i) * 9 + m] = dr[i]; dPdS[(p * 9 + i) * 9 + m] = dP[i]; }
        }
    }
}
extern "C" int hostcheck_jac(int ns, const double* slip6, double cdt, int64_t np, const double* H, const double* A, const double* g,
                             const double* R, const double* pp, const double* S, int with_params, int with_C, double* jac_x,
                             double* jac_y, double* dPdx, double* dPdS) {
    if (ns == 12) run_jac<12>(slip6, cdt, np, H, A, g, R, pp, S, with_params, with_C, jac_x, jac_y, dPdx, dPdS);
    else if (ns == 24) run_jac<24>(slip6, cdt, np, H, A, g, R, pp, S, with_params, with_C, jac_x, jac_y, dPdx, dPdS);
    else return -1;
    return 0;
}
// per-point VJP (cp_point_vjp) for the differentiable columns; grad (np, nd)
extern "C" int hostcheck_vjp(int ns, const double* slip6, double cdt, int64_t np, const double* H, const double* A, const double* g,
                             const double* R, const double* pp, const double* S, const double* W, int with_params, double* grad) {
    CpSlip table;
    if (ns != 12 && ns != 24) return -1;
    cp_slip_init(&table, slip6, ns);
    const int nd = 27 + 2 * ns + (with_params ? 5 : 0);
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t p = 0; p < np; ++p) {
        const double* q = pp + 4 * p;
        double lam[9];
        bool ok = (ns == 12) ? cp_point_vjp<12>(table, cdt, H + 9 * p, A + 9 * p, g + ns * p, R + 9 * p, q[3], q[0], q[1], q[2], S + 9 * p, W + 9 * p, 0, nd, grad + nd * p, lam)
                             : cp_point_vjp<24>(table, cdt, H + 9 * p, A + 9 * p, g + ns * p, R + 9 * p, q[3], q[0], q[1], q[2], S + 9 * p, W + 9 * p, 0, nd, grad + nd * p, lam);
        if (!ok) bad = 1;
    }
    return bad ? -2 : 0;
}
