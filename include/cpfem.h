/* cpfem.h - C ABI of the B200-native JAX-CPFEM hot path (libcpfem_b200.so).
 *
 * Drop-in boundary: every entry point takes plain DEVICE pointers, sizes and a CUDA stream (as an opaque
 * void*), is enqueue-only (no device synchronisation, no allocation on the hot path) and returns an int
 * status (0 = OK, <0 = hard error, text via cpfem_last_error).  These are the functions an XLA-FFI /
 * ctypes / cffi binding of the reference's hot path would bind (INTEGRATION.md shows the stubs).
 *
 * Reference interfaces replaced (paths relative to the JAX-CPFEM tree; jax_fem = deepmodeling/jax-fem,
 * imported by the reference at singlecrystal_copper/models_copper.py:9 but not vendored):
 *
 *   cpfem_plan_create / cpfem_plan_csr   jax_fem Problem.__post_init__ I/J construction (consumed at
 *                                        crystal_plasticity_OR_design/solver.py:281) + the scipy COO->CSR
 *                                        canonicalisation of get_A (solver.py:279-288)
 *   cpfem_update_state                   CrystalPlasticity.update_int_vars_gp   (models_copper.py:273-282)
 *   cpfem_residual                       Problem.compute_residual               (solver.py:244,715,807)
 *   cpfem_newton_update                  Problem.newton_update -> res, V        (solver.py:392,281)
 *                                        + get_A's CSR data                     (solver.py:281)
 *   cpfem_avg_stress                     CrystalPlasticity.compute_avg_stress   (models_copper.py:297-319)
 *   cpfem_update_state_avg_stress        the two calls above fused (driver order singlecrystal_copper.py:205,227)
 *   cpfem_point_stress_tangent           get_tensor_map()'s tensor_map under vmap, and its jacfwd
 *                                        (models_copper.py:135-137,155-162,251-265)
 *   cpfem_point_update_state             get_maps()'s update_int_vars_map under vmap (models_copper.py:164-169,267-269)
 *   cpfem_apply_dirichlet                apply_bc_vec + zeroRows                (solver.py:119-133,290-293)
 *   cpfem_point_jac_x / cpfem_point_vjp  f_jvp's jac_x, jac_y (models_copper.py:251-259) and their reverse mode
 *   cpfem_vjp_params                     vjp_linear_fn of implicit_vjp          (solver.py:832-848)
 */
#ifndef CPFEM_H
#define CPFEM_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPFEM_MAX_SLIP 24

/* Uniform material constants (field meaning: models_copper.py:54-56,94-96,141-149,212,231). */
typedef struct cpfem_material {
    double C11, C12, C44;
    double h, t_sat, gss_a, ao, xm, r;
    double tol;          /* local Newton tolerance (1e-8)                     */
    int32_t max_sub_step; /* 5 (Cu, Ta, DP) or 8 (304 steel)                  */
    int32_t max_iter;     /* safety cap on local Newton iterations (e.g. 200) */
} cpfem_material;

/* State layout of the (point, component) arrays. */
enum { CPFEM_LAYOUT_AOS = 0,   /* reference layout: (nc, 8, comps) C-contiguous        */
       CPFEM_LAYOUT_SOA = 1 }; /* native layout:    (comps, nc*8)   C-contiguous        */

/* Quadrature-point state, same order as the reference's internal_vars list
 * (models_copper.py:133; models_DPsteel_inhomo.py:229).  The six trailing pointers are the per-point
 * material arrays of the DP-steel / calibration variants; pass NULL to use cpfem_material instead.
 * C, when given, is the (nc,8,3,3,3,3) elastic tensor array; it must be cubic in the crystal frame
 * (only C[0,0,0,0], C[0,0,1,1], C[1,2,1,2] are read), which is what the reference builds. */
typedef struct cpfem_state {
    const double* Fp_inv;   /* (np, 9)  */
    const double* g;        /* (np, ns) slip resistance */
    const double* slip;     /* (np, ns) accumulated slip */
    const double* rot;      /* (np, 9)  */
    const double* gss_a;    /* (np) or NULL */
    const double* h;        /* (np) or NULL */
    const double* t_sat;    /* (np) or NULL */
    const double* xm;       /* (np) or NULL */
    const double* r;        /* (np) or NULL */
    const double* C;        /* (np, 81) or NULL */
    int32_t layout;         /* CPFEM_LAYOUT_* for Fp_inv, g, slip, rot */
} cpfem_state;

typedef struct cpfem_state_out {
    double* Fp_inv;         /* (np, 9)  */
    double* g;              /* (np, ns) */
    double* slip;           /* (np, ns) */
    int32_t layout;
} cpfem_state_out;

/* status words written by the kernels (device memory, 4 x int64, accumulated with atomics; zero it first):
 *   [0] points that hit max_iter   [1] points with a non-finite residual
 *   [2] max local Newton iterations seen   [3] sum of local Newton iterations */
#define CPFEM_STATUS_WORDS 4

typedef struct cpfem_plan cpfem_plan;

/* Build the per-mesh plan on the current device: keeps device copies of the connectivity and node
 * coordinates, the CSR pattern (bit-identical to scipy.sparse.csr_array((V,(I,J))) with the jax_fem I/J
 * rule: columns sorted, duplicates merged, explicit zeros kept) and the per-cell slot map.
 *   cells   device int32 (nc, 8), hex8 in meshio/Gmsh node order
 *   points  device double (nnodes, 3)
 *   slip    HOST double (ns, 6): rows "normal(3) direction(3)" as in data/csv/input_slip_sys*.txt
 *           (normalised internally, models_copper.py:62-66) */
int cpfem_plan_create(const int32_t* cells, int64_t nc, const double* points, int64_t nnodes,
                      const double* slip, int32_t ns, void* stream, cpfem_plan** out);
int cpfem_plan_destroy(cpfem_plan* plan);

/* CSR pattern owned by the plan (device pointers). indptr has 3*nnodes+1 entries. */
int cpfem_plan_csr(const cpfem_plan* plan, const int64_t** indptr, const int32_t** indices, int64_t* nnz);
/* Copy the pattern into caller-owned buffers (device or host; cudaMemcpyDefault): indptr_out int64
 * (3*nnodes+1), indices_out int32 (nnz).  Either may be NULL. */
int cpfem_plan_csr_copy(const cpfem_plan* plan, int64_t* indptr_out, int32_t* indices_out, void* stream);
/* Element-partitioned runs: the plan's mesh is "owned cells followed by ghost cells" (ghost cells only contribute
 * sparsity); restrict the kernels to the first n_active cells.  State arrays then have n_active*8 points.
 * n_active = 0 (a rank that owns no cell) is legal: the state pointers may then be NULL, residual / CSR values come
 * back zeroed and nothing is launched. */
int cpfem_plan_set_active_cells(cpfem_plan* plan, int64_t n_active);
/* Element-partitioned runs, overlap of the interface exchange with the assembly: `event` is a caller-owned cudaEvent_t
 * (NULL switches the feature off).  Every later cpfem_newton_update records it once the contributions of the cells
 * [0, cell_prefix) - and the zero-fill of res / csr_data - are complete (at the end of the first assembly chunk that
 * covers them), so that a communication stream waiting on it can ship the interface rows those cells feed (the
 * reduction of SURVEY 8(e), which the reference - single device - does not have) while the remaining chunks compute.
 * Contributions that arrive from peers may be added to res / csr_data with atomics from that point on. */
int cpfem_plan_set_progress_event(cpfem_plan* plan, int64_t cell_prefix, void* event);
/* Sizes: nc, nnodes, ns, nnz, max node valence, cells per assembly chunk. out[6]. */
int cpfem_plan_info(const cpfem_plan* plan, int64_t* out);

/* update_int_vars_gp: sol (nnodes,3) + old state -> new state.  in/out may alias array-wise. */
int cpfem_update_state(const cpfem_plan* plan, const cpfem_material* mat, const double* sol, const cpfem_state* in,
                       const cpfem_state_out* out, double dt, int64_t* status, void* stream);

/* Same for the cells [cell0, cell0 + ncells) only: the state arrays of `in` / `out` hold just those ncells*8 points
 * (chunk-local, any layout).  Lets a host-resident state stream through the device chunk by chunk (H2D copy of chunk
 * k+1 and D2H copy of chunk k-1 overlap the update of chunk k). */
int cpfem_update_state_cells(const cpfem_plan* plan, const cpfem_material* mat, const double* sol, const cpfem_state* in,
                             const cpfem_state_out* out, double dt, int64_t cell0, int64_t ncells, int64_t* status,
                             void* stream);

/* update_int_vars_gp fused with compute_avg_stress (models_copper.py:273-282 + 297-319): the reference drivers call
 * compute_avg_stress(sol, params) and then update_int_vars_gp(sol, params) with the same arguments
 * (singlecrystal_copper.py:205,227), i.e. the same converged local solve twice; this entry point does it once and
 * writes both the new state and sigma_cell (nc, 9). */
int cpfem_update_state_avg_stress(const cpfem_plan* plan, const cpfem_material* mat, const double* sol,
                                  const cpfem_state* in, const cpfem_state_out* out, double dt, double* sigma_cell,
                                  int64_t* status, void* stream);

/* compute_residual: res (nnodes,3) is OVERWRITTEN (zeroed inside, then accumulated). */
int cpfem_residual(const cpfem_plan* plan, const cpfem_material* mat, const double* sol, const cpfem_state* st,
                   double dt, double* res, int64_t* status, void* stream);

/* newton_update: res (nnodes,3) and csr_data (nnz) are overwritten with the residual and the assembled
 * tangent on the plan's pattern.  coo_V, if not NULL, receives the reference's problem.V layout
 * (nc, 24, 24): V[c, 3a+i, 3b+k].  Either of csr_data / coo_V may be NULL. */
int cpfem_newton_update(const cpfem_plan* plan, const cpfem_material* mat, const double* sol, const cpfem_state* st,
                        double dt, double* res, double* csr_data, double* coo_V, int64_t* status, void* stream);

/* compute_avg_stress: sigma_cell (nc, 9) = JxW-weighted mean Cauchy stress per cell. */
int cpfem_avg_stress(const cpfem_plan* plan, const cpfem_material* mat, const double* sol, const cpfem_state* st,
                     double dt, double* sigma_cell, int64_t* status, void* stream);

/* tensor_map under vmap: u_grads (np, 9) given explicitly -> P (np, 9) and, if tangent != NULL,
 * dP_ij/dH_kl (np, 81).  np need not be a multiple of 8; the plan only supplies the slip table. */
int cpfem_point_stress_tangent(const cpfem_plan* plan, const cpfem_material* mat, const double* u_grads,
                               int64_t np, const cpfem_state* st, double dt, double* P, double* tangent,
                               int64_t* status, void* stream);

/* update_int_vars_map under vmap (models_copper.py:164-169,267-269): u_grads (np, 9) given explicitly -> new state of
 * every point (AoS arrays of np points).  The point-wise counterpart of cpfem_update_state. */
int cpfem_point_update_state(const cpfem_plan* plan, const cpfem_material* mat, const double* u_grads, int64_t np,
                             const cpfem_state* st, const cpfem_state_out* out, double dt, int64_t* status, void* stream);

/* Everything the two calls above produce, from ONE local solve per point, plus the per-point account of that solve:
 * any of P (np, 9), tangent (np, 81), out (new state) may be NULL; point_info, if not NULL, is a device int32 (np, 3)
 * array receiving [local Newton iterations, residual evaluations, status bits (1: hit max_iter, 2: non-finite)] of every
 * point - what the parity tests compare with the reference's control flow (models_copper.py:204-249), point by point. */
int cpfem_point_eval(const cpfem_plan* plan, const cpfem_material* mat, const double* u_grads, int64_t np,
                     const cpfem_state* st, double dt, double* P, double* tangent, const cpfem_state_out* out,
                     int32_t* point_info, int64_t* status, void* stream);

/* ---- adjoint columns --------------------------------------------------------------------
 * x = ravel([u_grad, Fp_inv_old, slip_resistance_old, slip_old, rot_mat]) (models_copper.py:156; nx = 27 + 2 ns), followed by
 * [gss_a, h, t_sat, xm, r] when nextra >= 5 (calibration form, calibration/case1.py:153) and by C (81) when nextra == 6 (DP
 * form, models_DPsteel_inhomo.py:245; nx = 161 for ns = 24).  y = the nine entries of S.  Derivatives follow the reference's
 * literal formulation (rot_mat's nine entries independent, like jax.jacfwd sees them). */
/* f_jvp's Jacobians at the converged local solution (models_copper.py:256-257): jac_x (np, 9, nx) = d implicit_residual/dx,
 * and, if not NULL, jac_y (np, 9, 9) = d implicit_residual/dy and S (np, 9) = y itself. */
int cpfem_point_jac_x(const cpfem_plan* plan, const cpfem_material* mat, const double* u_grads, int64_t np,
                      const cpfem_state* st, double dt, int32_t nextra, double* jac_x, double* jac_y, double* S,
                      int64_t* status, void* stream);
/* Reverse mode of tensor_map through the local solve: grad (np, nx) = W : dP/dx with dP/dx = dP/dx|_S - dP/dS J_y^-1 J_x,
 * W (np, 9) the cotangent of P.  (The first nine columns are W : the consistent tangent.) */
int cpfem_point_vjp(const cpfem_plan* plan, const cpfem_material* mat, const double* u_grads, int64_t np,
                    const cpfem_state* st, double dt, int32_t nextra, const double* W, double* grad, int64_t* status, void* stream);
/* vjp_linear_fn of implicit_vjp (crystal_plasticity_OR_design/solver.py:832-848): adjoint (nnodes, 3) contracted with
 * d(residual vector)/d(internal_vars), i.e. per point W_ij = sum_a adjoint[node_a, i] dN_a/dX_j JxW pushed through
 * cpfem_point_vjp, written into arrays shaped like the state (any pointer may be NULL).  Zero the adjoint on Dirichlet dofs
 * first (their residual rows do not depend on the state, solver.py:119-133).  The caller applies implicit_vjp's final
 * minus sign (solver.py:849). */
typedef struct cpfem_state_grad {
    double* Fp_inv;   /* (np, 9)  */
    double* g;        /* (np, ns) */
    double* slip;     /* (np, ns) - identically zero: P does not depend on the accumulated slip */
    double* rot;      /* (np, 9)  */
    double* gss_a;    /* (np) - zero: the hardening law does not enter P */
    double* h;        /* (np) - zero */
    double* t_sat;    /* (np) - zero */
    double* xm;       /* (np) */
    double* r;        /* (np) - zero */
    double* C;        /* (np, 81) */
} cpfem_state_grad;
int cpfem_vjp_params(const cpfem_plan* plan, const cpfem_material* mat, const double* sol, const cpfem_state* st, double dt,
                     const double* adjoint, const cpfem_state_grad* out, int64_t* status, void* stream);

/* Input validation for the DP-steel form of the state (models_DPsteel_inhomo.py:121-147,185-186): counts the points of a
 * (np, 81) elastic-tensor array that are NOT of the cubic pattern in the crystal frame (C11 on iiii, C12 on iijj, C44 on
 * ijij / ijji, zero elsewhere; tolerance rtol x the largest constant).  *bad_count is a device int64 that is accumulated
 * atomically (zero it first); callers treat a non-zero count as a hard error - the kernels read only C[0,0,0,0],
 * C[0,0,1,1], C[1,2,1,2]. */
int cpfem_check_cubic(const double* C, int64_t np, double rtol, int64_t* bad_count, void* stream);

/* Row-elimination Dirichlet conditions on device: res[row] = sol[row] - val  (apply_bc_vec) and, if
 * csr_data != NULL, row := unit row (zeroRows with diag 1).  rows: device int64 (nbc) dof indices. */
int cpfem_apply_dirichlet(const cpfem_plan* plan, const int64_t* rows, const double* vals, int64_t nbc,
                          const double* sol, double* res, double* csr_data, void* stream);

/* ---- device linear solver on the plan's pattern ---------------------------------------
 * Replaces jax_solve (crystal_plasticity_OR_design/solver.py:19-48): get_A's host CSR -> BCOO round trip disappears,
 * the matrix assembled by cpfem_newton_update (+ cpfem_apply_dirichlet) is used where it lies. */
/* y = A x, A = (plan pattern, csr_data).  Node-block SpMV (three rows per node share one neighbour list). */
int cpfem_spmv(const cpfem_plan* plan, const double* csr_data, const double* x, double* y, void* stream);
/* diag(A) (solver.py:32 `A_sp_scipy.diagonal()`), or 1/diag(A) when invert != 0. */
int cpfem_csr_diagonal(const cpfem_plan* plan, const double* csr_data, double* diag, int32_t invert, void* stream);
/* Values of A^T on the same pattern (the pattern is structurally symmetric): what linear_solver(A.transpose(), ...) of
 * implicit_vjp needs (solver.py:844).  csr_data_T must not alias csr_data. */
int cpfem_csr_transpose(const cpfem_plan* plan, const double* csr_data, double* csr_data_T, void* stream);
/* BiCGStab with the semantics of jax.scipy.sparse.linalg.bicgstab(A, b, x0=x, M=Jacobi if precond, tol, atol, maxiter)
 * as called at solver.py:34-40: x holds x0 on entry and the solution on return.  info[0] = iterations taken (JAX's
 * negative breakdown codes -10 / -11 are passed through), info[1] = 1 if the tolerance was not reached; *resid, if not
 * NULL, receives ||A x - b||_2 (the check of solver.py:43-45).  Unlike the assembly entry points this call
 * synchronises the stream (it polls the device-side convergence flag) and owns a workspace inside the plan. */
int cpfem_bicgstab(cpfem_plan* plan, const double* csr_data, const double* b, double* x, int32_t precond, double tol,
                   double atol, int64_t maxiter, int64_t* info, double* resid, void* stream);

/* The same solve, enqueue-only (what an XLA-FFI handler may call: no stream synchronisation, no host read-back):
 * `iters_to_enqueue` iterations (rounded up to the plan's CUDA-graph batch of 8, capped by maxiter) are enqueued on the
 * stream - iterations after convergence / breakdown are no-ops on the device - and the outcome is written to DEVICE
 * memory: info_dev[0] = iterations taken (negative: JAX breakdown code), info_dev[1] = 1 if the tolerance was not reached
 * within the enqueued iterations (the caller may call again with x as the start vector), *resid_dev = ||A x - b||_2 (may
 * be NULL).  The first call on a plan allocates the plan's solver workspace and instantiates its CUDA graph: warm it up
 * once outside any stream capture. */
int cpfem_bicgstab_enqueue(cpfem_plan* plan, const double* csr_data, const double* b, double* x, int32_t precond, double tol,
                           double atol, int64_t maxiter, int64_t iters_to_enqueue, int64_t* info_dev, double* resid_dev,
                           void* stream);

/* Interface exchange helper for element-partitioned runs: dst[map[i]] += src[i]. */
int cpfem_scatter_add(const double* src, const int64_t* map, int64_t n, double* dst, void* stream);
/* Pack helper: dst[i] = src[map[i]]. */
int cpfem_gather(const double* src, const int64_t* map, int64_t n, double* dst, void* stream);
/* sum of squares of a device vector into out[0] (device double, accumulated atomically; zero it first). */
int cpfem_sumsq(const double* x, int64_t n, double* out, void* stream);

/* ---- peer-memory mailbox for the interface exchange (one process per GPU, one node; csrc/cpfem_peer.cu) ------------------
 * The reduction of SURVEY 8(e) - the reference is single-device and has no counterpart.  A mailbox is device memory of the
 * RECEIVER that senders map through CUDA IPC and write over NVLink; flags are 64-bit epochs that only grow.
 *   owner:   cpfem_peer_alloc -> handle (CPFEM_PEER_HANDLE_BYTES, ship it to the peers by any host channel)
 *   sender:  cpfem_peer_open(handle) -> pointer valid in the sender's process; cpfem_peer_put copies (gathers if map != NULL)
 *            n doubles into it and then stores `epoch` to *flag_peer (NULL: no flag) with system-scope release semantics
 *   owner:   cpfem_peer_wait(flag, epoch) makes the stream wait until the flag has reached `epoch` (bounded by timeout_s,
 *            default 20 s: a miss is counted in status[1] instead of hanging the device), then consumes the data with
 *            cpfem_scatter_add and acknowledges with cpfem_peer_signal into a flag in the sender's own mailbox.
 * All calls but alloc / open / close / free are enqueue-only. */
#define CPFEM_PEER_HANDLE_BYTES 64
int cpfem_peer_alloc(int64_t bytes, void** ptr, uint8_t* handle);
int cpfem_peer_open(const uint8_t* handle, void** ptr);
int cpfem_peer_close(void* ptr);
int cpfem_peer_free(void* ptr);
int cpfem_peer_put(double* dst_peer, const double* src, const int64_t* map, int64_t n, uint64_t* flag_peer, uint64_t epoch,
                   void* stream);
int cpfem_peer_signal(uint64_t* flag_peer, uint64_t epoch, void* stream);
int cpfem_peer_wait(const uint64_t* flag_local, uint64_t epoch, double timeout_s, int64_t* status, void* stream);

/* Layout conversion (np, comps) <-> (comps, np). */
int cpfem_aos_to_soa(const double* aos, int64_t np, int32_t comps, double* soa, void* stream);
int cpfem_soa_to_aos(const double* soa, int64_t np, int32_t comps, double* aos, void* stream);

/* FP64 FMA throughput microbenchmark used for the roofline denominator: runs `iters` dependent-chain
 * DFMA rounds on every SM and returns the flop count in *flops (time it with CUDA events). */
int cpfem_dfma_peak_kernel(int64_t iters, double* sink, double* flops, void* stream);

/* Number of kernels this library has launched since it was loaded (hot-path entry points; memsets / copies excluded):
 * bench.py reads it around its timed region for the `gpu_launches` figure. */
int64_t cpfem_launch_count(void);

const char* cpfem_last_error(void);
int cpfem_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CPFEM_H */
