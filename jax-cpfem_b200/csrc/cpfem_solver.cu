// cpfem_solver.cu - device linear algebra on the plan's CSR pattern: the hand-off row F2 of SURVEY.md section 8(f).
//
// Replaces (reference JAX-CPFEM, crystal_plasticity_OR_design/solver.py):
//   jax_solve            solver.py:19-48    scipy CSR -> BCOO, Jacobi preconditioner, jax.scipy.sparse.linalg.bicgstab
//                                            (tol = atol = 1e-10, maxiter = 10000), residual check ||A x - b|| < 0.1
// The reference copies the assembled matrix to the host (get_A, solver.py:279-288), back to the device as BCOO and
// runs BiCGStab there; here the CSR data never leaves the device.  jax.scipy.sparse.linalg.bicgstab lives in JAX
// (jax/_src/scipy/sparse/linalg.py, _bicgstab_solve; the model comments name JAX 0.4.13), not in the reference tree:
// its published algorithm is restated below statement by statement (and once more, in numpy, by the test oracle).
//
// Kernels
//   k_bicg_spmv<MODE>   node-block SpMV y = A x: one warp per node, i.e. per three CSR rows that share one sorted
//                       neighbour list; lane j owns neighbour j (3 x entries, a 3 x 3 block of the matrix); the 9 m(n)
//                       entries of the node are one contiguous run of memory, column indices come from the node
//                       adjacency (4 bytes per 9 entries instead of 4 per entry).  HBM-bound: 8 + 4/9 bytes per stored entry.  The dot products BiCGStab needs of
//                       the result are reduced in the same pass.
//   k_bicg_vec<MODE>    the fused vector updates of one BiCGStab iteration with their reductions
//   k_csr_diag          diagonal of A (Jacobi preconditioner, solver.py:32-33)
// Reductions are deterministic: per-block partial sums in a fixed order, the last block to arrive adds them up and
// runs the scalar recurrences (alpha, omega, beta, breakdown codes, convergence flag) on the device; the host only
// polls the convergence flag every few iterations, converged iterations are no-ops, so the iterates do not depend on
// the polling interval.  Eight iterations (40 kernel launches) are captured once per plan into a CUDA graph and replayed:
// on the meshes of the reference's own drivers the iteration is launch-bound.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <new>

#include "cpfem_internal.h"

struct BicgScal {
    double rho, alpha, omega;      // values of the previous iteration (JAX: rho, alpha, omega)
    double rho_;                   // <rhat, r> of the current iterate
    double alpha_, omega_;         // values of this iteration
    double ss, rs, bs, atol2;
    long long k, maxiter;
    int done, exit_early;
    unsigned int ticket;
    unsigned int pad;
};

// Every pointer a BiCGStab iteration touches.  The iteration kernels read this block from DEVICE memory (one uniform load
// each), so that the captured CUDA graph of the iteration does not depend on the caller's buffers.
struct BicgVecs {
    const double *data, *b, *minv;
    double *x, *r, *rhat, *p, *q, *phat, *s, *shat, *t;
};

#define RED_BLOCK 256
#define MAX_PARTIAL_BLOCKS 2048

// block-wide sums of NV values -> partials[block][NV]; returns true in every thread of the LAST block to publish, with
// the grid totals (fixed summation order for a fixed grid) in tot[].
template <int NV>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], double* partials, unsigned int* ticket, double (&tot)[NV]) {
    __shared__ double sh[NV][RED_BLOCK / 32];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sh[i][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = 0.0;
            for (int w = 0; w < RED_BLOCK / 32; ++w) s += sh[i][w];
            partials[(size_t)blockIdx.x * NV + i] = s;
        }
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
    // the last block adds the partials: thread t takes blocks t, t + RED_BLOCK, ... then a fixed tree
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = 0.0;
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += RED_BLOCK) s += ((volatile double*)partials)[(size_t)b * NV + i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        __syncthreads();
        if (lane == 0) sh[i][warp] = s;
        __syncthreads();
        double tt = 0.0;
        for (int w = 0; w < RED_BLOCK / 32; ++w) tt += sh[i][w];
        tot[i] = tt;
    }
    if (threadIdx.x == 0) *ticket = 0u;
    return true;
}

// MODE 0: r = b - A x ; rhat = p = q = r ; rs = <r,r>, bs = <b,b>      (initial_value of _bicgstab_solve)
// MODE 1: q = A phat ; <rhat, q>  -> alpha_
// MODE 2: t = A shat ; <t,s>, <t,t> -> omega_
// MODE 3: y = A x only (plain SpMV: vin -> vout)
#ifndef SPMV_MIN_BLOCKS
#define SPMV_MIN_BLOCKS 4
#endif
template <int MODE>
__global__ void __launch_bounds__(RED_BLOCK, SPMV_MIN_BLOCKS)
k_bicg_spmv(const int64_t* __restrict__ nbr_ptr, const int32_t* __restrict__ nbr, const double* __restrict__ data_arg,
            int64_t nn, BicgScal* sc, double* partials, const BicgVecs* __restrict__ Vp, const double* __restrict__ vin,
            double* __restrict__ vout, double tol, double atol, long long maxiter) {
    if (MODE == 1 || MODE == 2) {
        if (sc->done) return;
    }
    BicgVecs V = {};
    if (MODE != 3) V = *Vp;
    const double* __restrict__ data = (MODE == 3) ? data_arg : V.data;
    const double* __restrict__ xin = (MODE == 0) ? V.x : (MODE == 1) ? V.phat : (MODE == 2) ? V.shat : vin;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * RED_BLOCK + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * RED_BLOCK) >> 5;
    double acc1 = 0.0;                 // lanes 0..2: first dot product, lanes 3..5: second one
    // The index data of a node (its two nbr_ptr entries, then the neighbour id of every lane) sit in front of the x gather
    // in a chain of three dependent memory latencies.  They are fetched ONE NODE AHEAD: while the matrix loads of node n
    // are in flight the warp already holds (b0, m, col) of node n and requests those of node n + nwarps.
    int64_t b0 = 0;
    int m = 0, col = 0;
    if (warp0 < nn) {
        b0 = nbr_ptr[warp0];
        m = (int)(nbr_ptr[warp0 + 1] - b0);
        col = (lane < m) ? nbr[b0 + lane] : 0;
    }
    for (int64_t n = warp0; n < nn; n += nwarps) {
        const int64_t nxt = n + nwarps;
        int64_t b0n = 0, e0n = 0;
        if (nxt < nn) { b0n = nbr_ptr[nxt]; e0n = nbr_ptr[nxt + 1]; }
        const int m3 = 3 * m;
        const double* __restrict__ base = data + 9 * b0;
        // the vector entry the epilogue of lanes 0..2 needs (b, rhat or s of this node's rows) is requested up front, so
        // that its latency hides under the matrix loads instead of trailing the warp reduction
        // (lanes 3..5 mirror lanes 0..2 and carry the second dot product of modes 0 and 2, so that the loop holds ONE
        // accumulator per lane: at 64 registers a second one spilled)
        const int j3 = (lane < 3) ? lane : lane - 3;
        double aux = 0.0;
        if (MODE != 3 && lane < ((MODE == 1) ? 3 : 6)) {
            const double* __restrict__ av = (MODE == 0) ? V.b : (MODE == 1) ? (const double*)V.rhat : (const double*)V.s;
            aux = av[3 * n + j3];
        }
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        // lane j takes neighbour j: its three x entries once, then the 3 x 3 block of the node's rows.  For a fixed (i, k)
        // the lanes read with a 24-byte stride; the three k-loads of a row hit the same sectors (L1), so every sector
        // of the matrix still crosses L2/HBM exactly once, and there is no index arithmetic beyond one multiply-add.
        if (lane < m) {
            const double* __restrict__ xp = xin + 3 * (int64_t)col;
            const double* __restrict__ ap = base + 3 * lane;
            const double a00 = __ldcs(ap), a01 = __ldcs(ap + 1), a02 = __ldcs(ap + 2);
            const double a10 = __ldcs(ap + m3), a11 = __ldcs(ap + m3 + 1), a12 = __ldcs(ap + m3 + 2);
            const double a20 = __ldcs(ap + 2 * m3), a21 = __ldcs(ap + 2 * m3 + 1), a22 = __ldcs(ap + 2 * m3 + 2);
            const double x0 = xp[0], x1 = xp[1], x2 = xp[2];
            s0 = a00 * x0 + a01 * x1 + a02 * x2;
            s1 = a10 * x0 + a11 * x1 + a12 * x2;
            s2 = a20 * x0 + a21 * x1 + a22 * x2;
        }
        for (int j = lane + 32; j < m; j += 32) {                  // m > 32 (unstructured meshes): the remaining neighbours
            const double* __restrict__ xp = xin + 3 * (int64_t)nbr[b0 + j];
            const double* __restrict__ ap = base + 3 * j;
            const double x0 = xp[0], x1 = xp[1], x2 = xp[2];
            s0 += __ldcs(ap) * x0 + __ldcs(ap + 1) * x1 + __ldcs(ap + 2) * x2;
            s1 += __ldcs(ap + m3) * x0 + __ldcs(ap + m3 + 1) * x1 + __ldcs(ap + m3 + 2) * x2;
            s2 += __ldcs(ap + 2 * m3) * x0 + __ldcs(ap + 2 * m3 + 1) * x1 + __ldcs(ap + 2 * m3 + 2) * x2;
        }
        // next node's neighbour ids: b0n has arrived by now (it was requested before the matrix loads were)
        const int mn = (int)(e0n - b0n);
        const int coln = (lane < mn) ? nbr[b0n + lane] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane < 6) {
            const double y = (j3 == 0) ? s0 : (j3 == 1) ? s1 : s2;
            const int64_t row = 3 * n + j3;
            if (lane < 3) {
                if (MODE == 0) {
                    const double rr = aux - y;
                    V.r[row] = rr; V.rhat[row] = rr; V.p[row] = rr; V.q[row] = rr;
                    acc1 += rr * rr;
                } else if (MODE == 1) {
                    V.q[row] = y;
                    acc1 += aux * y;
                } else if (MODE == 2) {
                    V.t[row] = y;
                    acc1 += y * aux;
                } else {
                    vout[row] = y;
                }
            } else {
                if (MODE == 0) acc1 += aux * aux;          // <b, b>
                else if (MODE == 2) acc1 += y * y;         // <t, t>
            }
        }
        b0 = b0n; m = mn; col = coln;
    }
    if (MODE == 3) return;
    double acc[2] = {(lane < 3) ? acc1 : 0.0, (lane >= 3 && lane < 6) ? acc1 : 0.0};
    double tot2[2];
    if (!grid_reduce<2>(acc, partials, &sc->ticket, tot2)) return;
    if (threadIdx.x == 0) {
        if (MODE == 0) {
            const double t2 = tol * tol * tot2[1], a2 = atol * atol;
            sc->bs = tot2[1];
            sc->atol2 = (t2 > a2) ? t2 : a2;                   // jnp.maximum(square(tol) * bs, square(atol))
            sc->rs = tot2[0];
            sc->rho = 1.0; sc->alpha = 1.0; sc->omega = 1.0;   // rho0 = alpha0 = omega0 = 1
            sc->rho_ = tot2[0];                                // <rhat, r> with rhat = r0
            sc->k = 0; sc->maxiter = maxiter;
            sc->exit_early = 0;
            sc->done = !((tot2[0] > sc->atol2) && (0 < maxiter));
        } else if (MODE == 1) {
            sc->alpha_ = sc->rho_ / tot2[0];                   // alpha_ = rho_ / <rhat, q_>
        } else {
            sc->omega_ = tot2[0] / tot2[1];                    // omega_ = <t,s> / <t,t>
        }
    }
}

// MODE 0: beta = rho_/rho * alpha/omega ; p = r + beta (p - omega q) ; phat = M p
// MODE 1: s = r - alpha_ q ; ss = <s,s> ; shat = M s ; exit_early = ss < atol2
// MODE 2: x += alpha_ phat (+ omega_ shat) ; r = s (- omega_ t) ; rs = <r,r> ; rho_next = <rhat, r> ; k, breakdown, done
template <int MODE>
__global__ void __launch_bounds__(RED_BLOCK)
k_bicg_vec(int64_t n, BicgScal* sc, double* partials, const BicgVecs* __restrict__ Vp) {
    if (sc->done) return;
    const BicgVecs V = *Vp;
    const int64_t i0 = (int64_t)blockIdx.x * RED_BLOCK + threadIdx.x, st = (int64_t)gridDim.x * RED_BLOCK;
    double acc[2] = {0.0, 0.0};
    if (MODE == 0) {
        const double omega = sc->omega;
        const double beta = sc->rho_ / sc->rho * sc->alpha / omega;
        for (int64_t i = i0; i < n; i += st) {
            const double pn = V.r[i] + beta * (V.p[i] - omega * V.q[i]);
            V.p[i] = pn;
            V.phat[i] = V.minv ? pn * V.minv[i] : pn;
        }
        return;
    } else if (MODE == 1) {
        const double al = sc->alpha_;
        for (int64_t i = i0; i < n; i += st) {
            const double sv = V.r[i] - al * V.q[i];
            V.s[i] = sv;
            V.shat[i] = V.minv ? sv * V.minv[i] : sv;
            acc[0] += sv * sv;
        }
    } else {
        const double al = sc->alpha_, om = sc->omega_;
        const bool ee = sc->exit_early != 0;
        for (int64_t i = i0; i < n; i += st) {
            const double ap = al * V.phat[i];
            const double sv = V.s[i];
            const double xn = ee ? V.x[i] + ap : V.x[i] + (ap + om * V.shat[i]);
            const double rn = ee ? sv : sv - om * V.t[i];
            V.x[i] = xn;
            V.r[i] = rn;
            acc[0] += rn * rn;
            acc[1] += V.rhat[i] * rn;
        }
    }
    double tot[2];
    if (!grid_reduce<2>(acc, partials, &sc->ticket, tot)) return;
    if (threadIdx.x == 0) {
        if (MODE == 1) {
            sc->ss = tot[0];
            sc->exit_early = (tot[0] < sc->atol2) ? 1 : 0;
        } else {
            long long k = (sc->omega_ == 0.0 || sc->alpha_ == 0.0) ? -11 : sc->k + 1;
            if (sc->rho_ == 0.0) k = -10;
            sc->k = k;
            sc->rho = sc->rho_; sc->alpha = sc->alpha_; sc->omega = sc->omega_;
            sc->rho_ = tot[1];
            sc->rs = tot[0];
            sc->done = !((tot[0] > sc->atol2) && (k < sc->maxiter) && (k >= 0));
        }
    }
}

__global__ void k_csr_diag(const int64_t* __restrict__ nbr_ptr, const int32_t* __restrict__ nbr, const double* __restrict__ data,
                           int64_t nn, double* __restrict__ diag, int invert) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * nn) return;
    const int64_t n = t / 3;
    const int i = (int)(t - 3 * n);
    const int64_t b0 = nbr_ptr[n];
    const int m = (int)(nbr_ptr[n + 1] - b0);
    int lo = 0, hi = m - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (nbr[b0 + mid] < (int32_t)n) lo = mid + 1; else hi = mid;
    }
    const double d = data[9 * b0 + (int64_t)i * 3 * m + 3 * lo + i];
    diag[t] = invert ? 1.0 / d : d;
}

// ||a - b||^2 -> sc->ss (deterministic grid reduction)
__global__ void __launch_bounds__(RED_BLOCK)
k_norm2_diff(const double* __restrict__ a, const double* __restrict__ b, int64_t n, BicgScal* sc, double* partials) {
    double acc[1] = {0.0};
    for (int64_t i = (int64_t)blockIdx.x * RED_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * RED_BLOCK) {
        const double d = a[i] - b[i];
        acc[0] += d * d;
    }
    double tot[1];
    if (!grid_reduce<1>(acc, partials, &sc->ticket, tot)) return;
    if (threadIdx.x == 0) sc->ss = tot[0];
}

// -----------------------------------------------------------------------------------------------
// workspace (plan-owned, allocated on first use)
// -----------------------------------------------------------------------------------------------
#define BICG_GRAPH_ITERS 8
struct cpfem_solver_ws {
    int64_t n = 0, pitch = 0;
    double* vec = nullptr;         // 8 vectors + minv, `pitch` doubles apart
    BicgScal* sc = nullptr;
    double* partials = nullptr;
    BicgScal* host_sc = nullptr;   // pinned
    BicgVecs* dV = nullptr;        // device copy of the pointer block
    BicgVecs* hV = nullptr;        // pinned staging copy
    cudaStream_t cap_stream = nullptr;
    cudaGraphExec_t iter_graph = nullptr;    // BICG_GRAPH_ITERS iterations = 5 x BICG_GRAPH_ITERS kernel nodes
    bool graph_failed = false;
};

static int ws_get(cpfem_plan* p, cpfem_solver_ws** out) {
    if (!p->solver_ws) {
        cpfem_solver_ws* w = new (std::nothrow) cpfem_solver_ws();
        if (!w) return set_err(-3, "solver workspace: out of host memory");
        w->n = 3 * p->nn;
        // vector pitch: 256-byte aligned rows + an odd number of KiB of skew, so that the nine vectors neither start
        // misaligned nor sit a power of two (or the same DRAM channel phase) apart
        w->pitch = ((w->n + 31) / 32) * 32 + 32 * 37;
        cudaError_t e = cudaMalloc((void**)&w->vec, sizeof(double) * 9 * (size_t)w->pitch);
        if (e == cudaSuccess) e = cudaMalloc((void**)&w->sc, sizeof(BicgScal));
        if (e == cudaSuccess) e = cudaMalloc((void**)&w->partials, sizeof(double) * 2 * MAX_PARTIAL_BLOCKS);
        if (e == cudaSuccess) e = cudaMallocHost((void**)&w->host_sc, sizeof(BicgScal));
        if (e == cudaSuccess) e = cudaMalloc((void**)&w->dV, sizeof(BicgVecs));
        if (e == cudaSuccess) e = cudaMallocHost((void**)&w->hV, sizeof(BicgVecs));
        if (e == cudaSuccess) e = cudaMemset(w->sc, 0, sizeof(BicgScal));
        if (e != cudaSuccess) {
            cudaFree(w->vec); cudaFree(w->sc); cudaFree(w->partials); cudaFree(w->dV);
            if (w->host_sc) cudaFreeHost(w->host_sc);
            if (w->hV) cudaFreeHost(w->hV);
            delete w;
            return set_err(-2, "solver workspace allocation", e);
        }
        p->solver_ws = w;
    }
    *out = (cpfem_solver_ws*)p->solver_ws;
    return 0;
}
void cpfem_solver_ws_free(void* ws) {
    cpfem_solver_ws* w = (cpfem_solver_ws*)ws;
    if (!w) return;
    cudaFree(w->vec); cudaFree(w->sc); cudaFree(w->partials); cudaFree(w->dV);
    if (w->host_sc) cudaFreeHost(w->host_sc);
    if (w->hV) cudaFreeHost(w->hV);
    if (w->iter_graph) cudaGraphExecDestroy(w->iter_graph);
    if (w->cap_stream) cudaStreamDestroy(w->cap_stream);
    delete w;
}

static unsigned spmv_grid(const cpfem_plan* p) {
    int64_t g = (p->nn + (RED_BLOCK / 32) - 1) / (RED_BLOCK / 32);
    const int64_t cap = (int64_t)p->sm_count * 8;
    if (g > cap) g = cap;
    if (g > MAX_PARTIAL_BLOCKS) g = MAX_PARTIAL_BLOCKS;
    return (unsigned)(g < 1 ? 1 : g);
}
static unsigned vec_grid(const cpfem_plan* p, int64_t n) {
    int64_t g = (n + RED_BLOCK - 1) / RED_BLOCK;
    const int64_t cap = (int64_t)p->sm_count * 8;
    if (g > cap) g = cap;
    if (g > MAX_PARTIAL_BLOCKS) g = MAX_PARTIAL_BLOCKS;
    return (unsigned)(g < 1 ? 1 : g);
}

extern "C" int cpfem_spmv(const cpfem_plan* plan, const double* csr_data, const double* x, double* y, void* stream_) {
    if (!plan || !csr_data || !x || !y) return set_err(-1, "cpfem_spmv: null argument");
    cpfem_count_launches(1);
    k_bicg_spmv<3><<<spmv_grid(plan), RED_BLOCK, 0, (cudaStream_t)stream_>>>(plan->nbr_ptr, plan->nbr, csr_data, plan->nn, nullptr,
                                                                           nullptr, nullptr, x, y, 0.0, 0.0, 0);
    CU_TRY(cudaGetLastError());
    return 0;
}

extern "C" int cpfem_csr_diagonal(const cpfem_plan* plan, const double* csr_data, double* diag, int32_t invert, void* stream_) {
    if (!plan || !csr_data || !diag) return set_err(-1, "cpfem_csr_diagonal: null argument");
    const int64_t n = 3 * plan->nn;
    cpfem_count_launches(1);
    k_csr_diag<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(plan->nbr_ptr, plan->nbr, csr_data, plan->nn, diag, invert);
    CU_TRY(cudaGetLastError());
    return 0;
}

// one BiCGStab iteration = five kernels (body_fun of _bicgstab_solve)
static void launch_iteration(const cpfem_plan* plan, cpfem_solver_ws* w, unsigned gs, unsigned gv, cudaStream_t stream) {
    cpfem_count_launches(5);
    k_bicg_vec<0><<<gv, RED_BLOCK, 0, stream>>>(w->n, w->sc, w->partials, w->dV);
    k_bicg_spmv<1><<<gs, RED_BLOCK, 0, stream>>>(plan->nbr_ptr, plan->nbr, nullptr, plan->nn, w->sc, w->partials, w->dV, nullptr, nullptr,
                                                0.0, 0.0, 0);
    k_bicg_vec<1><<<gv, RED_BLOCK, 0, stream>>>(w->n, w->sc, w->partials, w->dV);
    k_bicg_spmv<2><<<gs, RED_BLOCK, 0, stream>>>(plan->nbr_ptr, plan->nbr, nullptr, plan->nn, w->sc, w->partials, w->dV, nullptr, nullptr,
                                                0.0, 0.0, 0);
    k_bicg_vec<2><<<gv, RED_BLOCK, 0, stream>>>(w->n, w->sc, w->partials, w->dV);
}

// The iteration is launch-bound on the meshes of the reference's own drivers (10^3 ... 25^3 cells: tens of microseconds
// of work per kernel), so BICG_GRAPH_ITERS iterations are captured once per plan into a CUDA graph (40 kernel nodes)
// and replayed; every kernel checks the device-side `done` flag first, so replaying past convergence changes nothing.
static cudaGraphExec_t iteration_graph(const cpfem_plan* plan, cpfem_solver_ws* w, unsigned gs, unsigned gv) {
    if (w->iter_graph || w->graph_failed) return w->iter_graph;
    cudaGraph_t g = nullptr;
    bool ok = true;
    if (!w->cap_stream) ok = cudaStreamCreateWithFlags(&w->cap_stream, cudaStreamNonBlocking) == cudaSuccess;
    if (ok) ok = cudaStreamBeginCapture(w->cap_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
        for (int it = 0; it < BICG_GRAPH_ITERS; ++it) launch_iteration(plan, w, gs, gv, w->cap_stream);
        ok = cudaStreamEndCapture(w->cap_stream, &g) == cudaSuccess && g != nullptr;
    }
    if (ok) ok = cudaGraphInstantiate(&w->iter_graph, g, 0) == cudaSuccess;
    if (g) cudaGraphDestroy(g);
    if (!ok) {
        cudaGetLastError();            // clear; fall back to plain launches
        w->iter_graph = nullptr;
        w->graph_failed = true;
    }
    return w->iter_graph;
}

extern "C" int cpfem_bicgstab(cpfem_plan* plan, const double* csr_data, const double* b, double* x, int32_t precond,
                              double tol, double atol, int64_t maxiter, int64_t* info, double* resid, void* stream_) {
    if (!plan || !csr_data || !b || !x) return set_err(-1, "cpfem_bicgstab: null argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    cpfem_solver_ws* w = nullptr;
    int rc = ws_get(plan, &w);
    if (rc) return rc;
    const int64_t n = w->n;
    BicgVecs V;
    V.data = csr_data; V.b = b; V.x = x;
    const int64_t vp = w->pitch;
    V.r = w->vec; V.rhat = w->vec + vp; V.p = w->vec + 2 * vp; V.q = w->vec + 3 * vp; V.phat = w->vec + 4 * vp;
    V.s = w->vec + 5 * vp; V.shat = w->vec + 6 * vp; V.t = w->vec + 7 * vp;
    double* minv = w->vec + 8 * vp;
    V.minv = precond ? minv : nullptr;
    *w->hV = V;
    CU_TRY(cudaMemcpyAsync(w->dV, w->hV, sizeof(BicgVecs), cudaMemcpyHostToDevice, stream));
    const unsigned gs = spmv_grid(plan), gv = vec_grid(plan, n);
    if (precond) {
        cpfem_count_launches(1);
        k_csr_diag<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(plan->nbr_ptr, plan->nbr, csr_data, plan->nn, minv, 1);
    }
    cpfem_count_launches(1);
    k_bicg_spmv<0><<<gs, RED_BLOCK, 0, stream>>>(plan->nbr_ptr, plan->nbr, nullptr, plan->nn, w->sc, w->partials, w->dV, nullptr, nullptr,
                                                tol, atol, (long long)maxiter);
    CU_TRY(cudaGetLastError());
    cudaGraphExec_t graph = iteration_graph(plan, w, gs, gv);
    // the host polls the device-side flag between batches of iterations (the flag also covers k >= maxiter)
    int64_t launched = 0;
    int batches = 1;
    for (;;) {
        CU_TRY(cudaMemcpyAsync(w->host_sc, w->sc, sizeof(BicgScal), cudaMemcpyDeviceToHost, stream));
        CU_TRY(cudaStreamSynchronize(stream));      // also guarantees hV was consumed before the next call rewrites it
        if (w->host_sc->done || launched >= maxiter + BICG_GRAPH_ITERS) break;
        for (int bi = 0; bi < batches; ++bi, launched += BICG_GRAPH_ITERS) {
            if (graph) {
                CU_TRY(cudaGraphLaunch(graph, stream));
                cpfem_count_launches(5 * BICG_GRAPH_ITERS);
            } else {
                for (int it = 0; it < BICG_GRAPH_ITERS; ++it) launch_iteration(plan, w, gs, gv, stream);
            }
        }
        CU_TRY(cudaGetLastError());
        if (batches < 8) batches *= 2;
    }
    if (w->host_sc->k < 0 && getenv("CPFEM_DEBUG_BICG"))       // diagnostic: the scalars at a breakdown (JAX codes -10 / -11)
        fprintf(stderr, "bicgstab breakdown k=%lld rho=%.17g rho_=%.17g alpha=%.17g alpha_=%.17g omega=%.17g omega_=%.17g ss=%.17g rs=%.17g atol2=%.17g\n",
                w->host_sc->k, w->host_sc->rho, w->host_sc->rho_, w->host_sc->alpha, w->host_sc->alpha_, w->host_sc->omega,
                w->host_sc->omega_, w->host_sc->ss, w->host_sc->rs, w->host_sc->atol2);
    if (info) {
        info[0] = w->host_sc->k;                                   // iterations taken (negative: breakdown code of JAX)
        info[1] = (w->host_sc->rs > w->host_sc->atol2) ? 1 : 0;    // 1 = stopped without reaching the tolerance
    }
    if (resid) {
        // ||A x - b|| as jax_solve checks it (solver.py:43-45): one more SpMV into the t vector
        cpfem_count_launches(2);
        k_bicg_spmv<3><<<gs, RED_BLOCK, 0, stream>>>(plan->nbr_ptr, plan->nbr, csr_data, plan->nn, nullptr, nullptr, nullptr, x, V.t, 0.0, 0.0, 0);
        k_norm2_diff<<<gv, RED_BLOCK, 0, stream>>>(V.t, b, n, w->sc, w->partials);
        CU_TRY(cudaMemcpyAsync(w->host_sc, w->sc, sizeof(BicgScal), cudaMemcpyDeviceToHost, stream));
        CU_TRY(cudaStreamSynchronize(stream));
        *resid = sqrt(w->host_sc->ss);
    }
    return 0;
}

// -----------------------------------------------------------------------------------------------
// Enqueue-only BiCGStab (for XLA-FFI handlers, SURVEY 8(b) "re-entrant, enqueue-only"): nothing here waits for the
// device.  The pointer block travels as a kernel argument (no pinned staging buffer that a second call could overwrite
// before the first copy ran), a FIXED number of iterations is enqueued - iterations after convergence are no-ops on the
// device - and the outcome is written to DEVICE memory by a last one-thread kernel.
// -----------------------------------------------------------------------------------------------
__global__ void k_bicg_set_vecs(BicgVecs* dst, const BicgVecs v) { *dst = v; }
__global__ void k_bicg_finish(const BicgScal* sc, int64_t* info, double* resid, int with_resid) {
    if (info) {
        info[0] = sc->k;                                  // iterations taken (negative: breakdown code of JAX)
        info[1] = (sc->rs > sc->atol2) ? 1 : 0;           // 1 = stopped / ran out of enqueued iterations above the tolerance
    }
    if (resid && with_resid) *resid = sqrt(sc->ss);
}

extern "C" int cpfem_bicgstab_enqueue(cpfem_plan* plan, const double* csr_data, const double* b, double* x, int32_t precond,
                                      double tol, double atol, int64_t maxiter, int64_t iters_to_enqueue, int64_t* info_dev,
                                      double* resid_dev, void* stream_) {
    if (!plan || !csr_data || !b || !x) return set_err(-1, "cpfem_bicgstab_enqueue: null argument");
    if (iters_to_enqueue <= 0 || maxiter <= 0) return set_err(-1, "cpfem_bicgstab_enqueue: iteration counts must be positive");
    cudaStream_t stream = (cudaStream_t)stream_;
    cpfem_solver_ws* w = nullptr;
    int rc = ws_get(plan, &w);            // first call on a plan allocates the workspace (not enqueue-only: warm it up once)
    if (rc) return rc;
    const int64_t n = w->n;
    BicgVecs V;
    V.data = csr_data; V.b = b; V.x = x;
    const int64_t vp = w->pitch;
    V.r = w->vec; V.rhat = w->vec + vp; V.p = w->vec + 2 * vp; V.q = w->vec + 3 * vp; V.phat = w->vec + 4 * vp;
    V.s = w->vec + 5 * vp; V.shat = w->vec + 6 * vp; V.t = w->vec + 7 * vp;
    double* minv = w->vec + 8 * vp;
    V.minv = precond ? minv : nullptr;
    const unsigned gs = spmv_grid(plan), gv = vec_grid(plan, n);
    k_bicg_set_vecs<<<1, 1, 0, stream>>>(w->dV, V);
    if (precond) k_csr_diag<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(plan->nbr_ptr, plan->nbr, csr_data, plan->nn, minv, 1);
    k_bicg_spmv<0><<<gs, RED_BLOCK, 0, stream>>>(plan->nbr_ptr, plan->nbr, nullptr, plan->nn, w->sc, w->partials, w->dV, nullptr, nullptr,
                                                tol, atol, (long long)maxiter);
    cpfem_count_launches(precond ? 3 : 2);
    CU_TRY(cudaGetLastError());
    cudaGraphExec_t graph = iteration_graph(plan, w, gs, gv);      // captured once per plan, on a plan-owned stream
    const int64_t want = iters_to_enqueue < maxiter ? iters_to_enqueue : maxiter;
    for (int64_t done = 0; done < want; done += BICG_GRAPH_ITERS) {
        if (graph) {
            CU_TRY(cudaGraphLaunch(graph, stream));
            cpfem_count_launches(5 * BICG_GRAPH_ITERS);
        } else {
            for (int it = 0; it < BICG_GRAPH_ITERS; ++it) launch_iteration(plan, w, gs, gv, stream);
        }
    }
    if (resid_dev) {
        k_bicg_spmv<3><<<gs, RED_BLOCK, 0, stream>>>(plan->nbr_ptr, plan->nbr, csr_data, plan->nn, nullptr, nullptr, nullptr, x, V.t, 0.0, 0.0, 0);
        k_norm2_diff<<<gv, RED_BLOCK, 0, stream>>>(V.t, b, n, w->sc, w->partials);
        cpfem_count_launches(2);
    }
    k_bicg_finish<<<1, 1, 0, stream>>>(w->sc, info_dev, resid_dev, resid_dev != nullptr);
    cpfem_count_launches(1);
    CU_TRY(cudaGetLastError());
    return 0;
}

// -----------------------------------------------------------------------------------------------
// A^T on the plan's pattern (structurally symmetric: node adjacency is).  The adjoint solve of implicit_vjp is
// linear_solver(A.transpose(), v) (crystal_plasticity_OR_design/solver.py:844): the transposed values are written once and
// the ordinary node-block BiCGStab runs on them.  One thread per (node n, neighbour slot j): block (n, m) of A^T is the
// transpose of block (m, n) of A, found by a binary search of n in m's sorted neighbour list.
// -----------------------------------------------------------------------------------------------
__global__ void k_csr_transpose(const int64_t* __restrict__ nbr_ptr, const int32_t* __restrict__ nbr, int64_t nn, int64_t nblocks,
                                const double* __restrict__ in, double* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nblocks) return;
    // node n that owns block slot t: binary search in nbr_ptr
    int64_t lo = 0, hi = nn - 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (nbr_ptr[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const int64_t n = lo, b0 = nbr_ptr[n], mn = nbr_ptr[n + 1] - b0, j = t - b0;
    const int64_t m = nbr[t], c0 = nbr_ptr[m], mm = nbr_ptr[m + 1] - c0;
    int64_t a = 0, b = mm - 1;
    while (a < b) {
        const int64_t mid = (a + b) >> 1;
        if (nbr[c0 + mid] < (int32_t)n) a = mid + 1; else b = mid;
    }
    const int64_t r = a;                          // rank of n in m's list (present: the adjacency is symmetric)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) out[9 * b0 + i * 3 * mn + 3 * j + k] = in[9 * c0 + k * 3 * mm + 3 * r + i];
}

extern "C" int cpfem_csr_transpose(const cpfem_plan* plan, const double* csr_data, double* csr_data_T, void* stream_) {
    if (!plan || !csr_data || !csr_data_T || csr_data == csr_data_T) return set_err(-1, "cpfem_csr_transpose: bad argument");
    const int64_t nblocks = plan->nnz / 9;
    cpfem_count_launches(1);
    k_csr_transpose<<<(unsigned)((nblocks + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(plan->nbr_ptr, plan->nbr, plan->nn, nblocks,
                                                                                         csr_data, csr_data_T);
    CU_TRY(cudaGetLastError());
    return 0;
}
