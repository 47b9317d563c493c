// cpfem_ffi.cc - XLA FFI custom-call handlers over the C ABI of include/cpfem.h (north_star: "a thin C-ABI exposed as
// JAX FFI custom calls").  One handler per hot-path entry point; each one only unpacks XLA buffers into the plain
// pointers of the C ABI and enqueues on XLA's stream - no synchronisation, no allocation, re-entrant (the plan is
// read-only during these calls; the one exception, the solver workspace of cpfem_bicgstab_enqueue, is documented there).
//
// Built by cpfem_b200.jax_ffi.build() when JAX is importable (needs the headers of `jax.ffi.include_dir()`):
//   g++ -std=c++17 -O2 -shared -fPIC -I$(python -c "import jax; print(jax.ffi.include_dir())") -I/usr/local/cuda/include
//       -Iinclude jax-cpfem_b200/csrc/cpfem_ffi.cc -L<dir of libcpfem_b200.so> -lcpfem_b200 -Wl,-rpath,'$ORIGIN'
//       -o jax-cpfem_b200/cpfem_b200/libcpfem_ffi.so
// JAX is not installable in the build container of this repo (no index), so this file is compiled only on a box that
// has it; tests/test_jax_ffi.py skips otherwise.  tests/ffi_stub/ holds a minimal stand-in for the two XLA headers that is
// used for ONE thing: checking that this file parses and type-checks with g++ in the container (tests/test_abi.py).
//
// Reference call sites replaced (paths relative to the JAX-CPFEM tree):
//   cpfem_update_state_ffi      CrystalPlasticity.update_int_vars_gp     singlecrystal_copper/models_copper.py:273-282
//   cpfem_avg_stress_ffi        CrystalPlasticity.compute_avg_stress     models_copper.py:297-319
//   cpfem_update_avg_ffi        both of the above from one local solve   (driver order singlecrystal_copper.py:205,227)
//   cpfem_residual_ffi          Problem.compute_residual                 crystal_plasticity_OR_design/solver.py:244
//   cpfem_newton_update_ffi     Problem.newton_update (+ get_A's CSR)    solver.py:392, 279-288
//   cpfem_point_eval_ffi        get_tensor_map() / update_int_vars_map under vmap, jacfwd(tensor_map)
//                                                                        models_copper.py:135-137,155-169,251-271
//   cpfem_dirichlet_ffi         apply_bc_vec + zeroRows                  solver.py:119-133,290-293
//   cpfem_bicgstab_ffi          jax_solve                                solver.py:19-48
//   cpfem_point_jac_x_ffi       f_jvp's jac_x / jac_y / y                models_copper.py:251-259
//   cpfem_vjp_params_ffi        vjp_linear_fn of implicit_vjp            solver.py:832-848
//   cpfem_csr_transpose_ffi     A.transpose() of implicit_vjp            solver.py:844
// State arrays arrive in the reference's internal_vars order (models_copper.py:133: Fp_inv, slip resistance, slip,
// rot_mats; models_DPsteel_inhomo.py:229 adds gss_a, h, t_sat, xm, r, C) as the trailing ("remaining") arguments: 4, 9
// or 10 buffers.  Attributes: `plan` = the cpfem_plan* as int64 (created once per mesh through ctypes,
// cpfem_b200/jax_ffi.py), `dt`, and the material as a dictionary decoded into cpfem_material.
#include <cuda_runtime_api.h>

#include <cstdint>
#include <string>

#include "xla/ffi/api/c_api.h"
#include "xla/ffi/api/ffi.h"

#include "cpfem.h"

namespace ffi = xla::ffi;

XLA_FFI_REGISTER_STRUCT_ATTR_DECODING(cpfem_material, ffi::StructMember<double>("C11"), ffi::StructMember<double>("C12"),
                                      ffi::StructMember<double>("C44"), ffi::StructMember<double>("h"),
                                      ffi::StructMember<double>("t_sat"), ffi::StructMember<double>("gss_a"),
                                      ffi::StructMember<double>("ao"), ffi::StructMember<double>("xm"),
                                      ffi::StructMember<double>("r"), ffi::StructMember<double>("tol"),
                                      ffi::StructMember<int32_t>("max_sub_step"), ffi::StructMember<int32_t>("max_iter"));

namespace {

using F64 = ffi::Buffer<ffi::F64>;
using F64Out = ffi::ResultBuffer<ffi::F64>;
using S64Out = ffi::ResultBuffer<ffi::S64>;

ffi::Error Fail(const char* who, int rc) {
    std::string msg = std::string(who) + " failed (" + std::to_string(rc) + "): " + cpfem_last_error();
    return rc == -1 ? ffi::Error(ffi::ErrorCode::kInvalidArgument, msg) : ffi::Error(ffi::ErrorCode::kInternal, msg);
}
ffi::Error Done(const char* who, int rc) { return rc == 0 ? ffi::Error::Success() : Fail(who, rc); }

cpfem_plan* PlanOf(int64_t handle) { return reinterpret_cast<cpfem_plan*>(static_cast<intptr_t>(handle)); }

// internal_vars (4, 9 or 10 trailing buffers) -> cpfem_state, reference (AoS) layout
ffi::Error StateOf(ffi::RemainingArgs vars, cpfem_state* st) {
    const size_t n = vars.size();
    if (n != 4 && n != 9 && n != 10)
        return ffi::Error(ffi::ErrorCode::kInvalidArgument, "internal_vars must have 4, 9 or 10 arrays");
    const double* p[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    for (size_t i = 0; i < n; ++i) {
        auto b = vars.get<F64>(i);
        if (!b.has_value()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "internal_vars must be float64 arrays");
        p[i] = b.value().typed_data();
    }
    st->Fp_inv = p[0]; st->g = p[1]; st->slip = p[2]; st->rot = p[3];
    st->gss_a = p[4]; st->h = p[5]; st->t_sat = p[6]; st->xm = p[7]; st->r = p[8]; st->C = p[9];
    st->layout = CPFEM_LAYOUT_AOS;
    return ffi::Error::Success();
}

ffi::Error ZeroStatus(cudaStream_t stream, S64Out& status) {
    if (cudaMemsetAsync(status->typed_data(), 0, CPFEM_STATUS_WORDS * sizeof(int64_t), stream) != cudaSuccess)
        return ffi::Error(ffi::ErrorCode::kInternal, "cudaMemsetAsync(status) failed");
    return ffi::Error::Success();
}

// ---- update_int_vars_gp -----------------------------------------------------------------------------------------
ffi::Error UpdateState(cudaStream_t stream, int64_t plan, double dt, cpfem_material mat, F64 sol, F64Out Fp_new, F64Out g_new,
                       F64Out slip_new, S64Out status, ffi::RemainingArgs vars) {
    cpfem_state in;
    if (ffi::Error e = StateOf(vars, &in); e.failure()) return e;
    if (ffi::Error e = ZeroStatus(stream, status); e.failure()) return e;
    cpfem_state_out out = {Fp_new->typed_data(), g_new->typed_data(), slip_new->typed_data(), CPFEM_LAYOUT_AOS};
    return Done("cpfem_update_state", cpfem_update_state(PlanOf(plan), &mat, sol.typed_data(), &in, &out, dt,
                                                          status->typed_data(), stream));
}

// ---- compute_avg_stress -----------------------------------------------------------------------------------------
ffi::Error AvgStress(cudaStream_t stream, int64_t plan, double dt, cpfem_material mat, F64 sol, F64Out sigma, S64Out status,
                     ffi::RemainingArgs vars) {
    cpfem_state in;
    if (ffi::Error e = StateOf(vars, &in); e.failure()) return e;
    if (ffi::Error e = ZeroStatus(stream, status); e.failure()) return e;
    return Done("cpfem_avg_stress", cpfem_avg_stress(PlanOf(plan), &mat, sol.typed_data(), &in, dt, sigma->typed_data(),
                                                      status->typed_data(), stream));
}

// ---- update_int_vars_gp + compute_avg_stress from one local solve ---------------------------------------------------
ffi::Error UpdateAvg(cudaStream_t stream, int64_t plan, double dt, cpfem_material mat, F64 sol, F64Out Fp_new, F64Out g_new,
                     F64Out slip_new, F64Out sigma, S64Out status, ffi::RemainingArgs vars) {
    cpfem_state in;
    if (ffi::Error e = StateOf(vars, &in); e.failure()) return e;
    if (ffi::Error e = ZeroStatus(stream, status); e.failure()) return e;
    cpfem_state_out out = {Fp_new->typed_data(), g_new->typed_data(), slip_new->typed_data(), CPFEM_LAYOUT_AOS};
    return Done("cpfem_update_state_avg_stress",
                cpfem_update_state_avg_stress(PlanOf(plan), &mat, sol.typed_data(), &in, &out, dt, sigma->typed_data(),
                                              status->typed_data(), stream));
}

// ---- compute_residual ---------------------------------------------------------------------------------------------
ffi::Error Residual(cudaStream_t stream, int64_t plan, double dt, cpfem_material mat, F64 sol, F64Out res, S64Out status,
                    ffi::RemainingArgs vars) {
    cpfem_state in;
    if (ffi::Error e = StateOf(vars, &in); e.failure()) return e;
    if (ffi::Error e = ZeroStatus(stream, status); e.failure()) return e;
    return Done("cpfem_residual", cpfem_residual(PlanOf(plan), &mat, sol.typed_data(), &in, dt, res->typed_data(),
                                                  status->typed_data(), stream));
}

// ---- newton_update: residual + CSR data (+ the reference's V when want_V != 0; V is a zero-size buffer otherwise) --------
ffi::Error NewtonUpdate(cudaStream_t stream, int64_t plan, double dt, cpfem_material mat, int64_t want_V, F64 sol, F64Out res,
                        F64Out csr_data, F64Out V, S64Out status, ffi::RemainingArgs vars) {
    cpfem_state in;
    if (ffi::Error e = StateOf(vars, &in); e.failure()) return e;
    if (ffi::Error e = ZeroStatus(stream, status); e.failure()) return e;
    double* v = (want_V != 0 && V->element_count() > 0) ? V->typed_data() : nullptr;
    return Done("cpfem_newton_update", cpfem_newton_update(PlanOf(plan), &mat, sol.typed_data(), &in, dt, res->typed_data(),
                                                            csr_data->typed_data(), v, status->typed_data(), stream));
}

// ---- tensor_map / jacfwd(tensor_map) / update_int_vars_map on explicit u_grads ------------------------------------------
// `what` bit 0: tangent wanted, bit 1: new state wanted (unused outputs are zero-size buffers on the JAX side)
ffi::Error PointEval(cudaStream_t stream, int64_t plan, double dt, cpfem_material mat, int64_t what, F64 u_grads, F64Out P,
                     F64Out tangent, F64Out Fp_new, F64Out g_new, F64Out slip_new, ffi::ResultBuffer<ffi::S32> info, S64Out status,
                     ffi::RemainingArgs vars) {
    cpfem_state in;
    if (ffi::Error e = StateOf(vars, &in); e.failure()) return e;
    if (ffi::Error e = ZeroStatus(stream, status); e.failure()) return e;
    const int64_t np = static_cast<int64_t>(u_grads.element_count() / 9);
    cpfem_state_out out = {Fp_new->typed_data(), g_new->typed_data(), slip_new->typed_data(), CPFEM_LAYOUT_AOS};
    return Done("cpfem_point_eval",
                cpfem_point_eval(PlanOf(plan), &mat, u_grads.typed_data(), np, &in, dt, P->typed_data(),
                                 (what & 1) ? tangent->typed_data() : nullptr, (what & 2) ? &out : nullptr, info->typed_data(),
                                 status->typed_data(), stream));
}

// ---- apply_bc_vec + zeroRows: res and csr_data are updated in place (input_output_aliases on the JAX side) --------------
ffi::Error Dirichlet(cudaStream_t stream, int64_t plan, ffi::Buffer<ffi::S64> rows, F64 vals, F64 sol, F64 res_in, F64 csr_in,
                     F64Out res, F64Out csr_data) {
    const int64_t nbc = static_cast<int64_t>(rows.element_count());
    // aliased buffers: nothing to copy; without aliasing XLA hands out fresh outputs, which are filled from the inputs first
    if (res->typed_data() != res_in.typed_data())
        cudaMemcpyAsync(res->typed_data(), res_in.typed_data(), res_in.size_bytes(), cudaMemcpyDeviceToDevice, stream);
    if (csr_data->typed_data() != csr_in.typed_data())
        cudaMemcpyAsync(csr_data->typed_data(), csr_in.typed_data(), csr_in.size_bytes(), cudaMemcpyDeviceToDevice, stream);
    return Done("cpfem_apply_dirichlet", cpfem_apply_dirichlet(PlanOf(plan), rows.typed_data(), vals.typed_data(), nbc,
                                                                sol.typed_data(), res->typed_data(), csr_data->typed_data(), stream));
}

// ---- jax_solve: Jacobi-BiCGStab on the device-resident CSR, enqueue-only -------------------------------------------------
ffi::Error Bicgstab(cudaStream_t stream, int64_t plan, int64_t precond, double tol, double atol, int64_t maxiter,
                    int64_t iters_to_enqueue, F64 csr_data, F64 b, F64 x0, F64Out x, S64Out info, F64Out resid) {
    if (x->typed_data() != x0.typed_data())
        cudaMemcpyAsync(x->typed_data(), x0.typed_data(), x0.size_bytes(), cudaMemcpyDeviceToDevice, stream);
    return Done("cpfem_bicgstab_enqueue",
                cpfem_bicgstab_enqueue(PlanOf(plan), csr_data.typed_data(), b.typed_data(), x->typed_data(), static_cast<int32_t>(precond),
                                       tol, atol, maxiter, iters_to_enqueue, info->typed_data(), resid->typed_data(), stream));
}

// ---- adjoint row: f_jvp's Jacobians at the converged local solution ---------------------------------------------------
// nextra = 0 / 5 / 6 selects x's trailing parameter columns (include/cpfem.h); jac_y may be a zero-size buffer
ffi::Error PointJacX(cudaStream_t stream, int64_t plan, double dt, cpfem_material mat, int64_t nextra, F64 u_grads, F64Out jac_x,
                     F64Out jac_y, F64Out S, S64Out status, ffi::RemainingArgs vars) {
    cpfem_state in;
    if (ffi::Error e = StateOf(vars, &in); e.failure()) return e;
    if (ffi::Error e = ZeroStatus(stream, status); e.failure()) return e;
    const int64_t np = static_cast<int64_t>(u_grads.element_count() / 9);
    return Done("cpfem_point_jac_x",
                cpfem_point_jac_x(PlanOf(plan), &mat, u_grads.typed_data(), np, &in, dt, static_cast<int32_t>(nextra),
                                  jac_x->typed_data(), jac_y->element_count() > 0 ? jac_y->typed_data() : nullptr,
                                  S->element_count() > 0 ? S->typed_data() : nullptr, status->typed_data(), stream));
}

// ---- adjoint row: nodal adjoint . d(residual)/d(internal_vars); ten result buffers in internal_vars order, zero-size
// where the state has no such array (or the caller does not want it)
ffi::Error VjpParams(cudaStream_t stream, int64_t plan, double dt, cpfem_material mat, F64 sol, F64 adjoint, F64Out g0, F64Out g1,
                     F64Out g2, F64Out g3, F64Out g4, F64Out g5, F64Out g6, F64Out g7, F64Out g8, F64Out g9, S64Out status,
                     ffi::RemainingArgs vars) {
    cpfem_state in;
    if (ffi::Error e = StateOf(vars, &in); e.failure()) return e;
    if (ffi::Error e = ZeroStatus(stream, status); e.failure()) return e;
    auto ptr = [](F64Out& b) -> double* { return b->element_count() > 0 ? b->typed_data() : nullptr; };
    cpfem_state_grad out = {ptr(g0), ptr(g1), ptr(g2), ptr(g3), ptr(g4), ptr(g5), ptr(g6), ptr(g7), ptr(g8), ptr(g9)};
    return Done("cpfem_vjp_params", cpfem_vjp_params(PlanOf(plan), &mat, sol.typed_data(), &in, dt, adjoint.typed_data(), &out,
                                                      status->typed_data(), stream));
}

// ---- values of A^T on the plan's pattern -------------------------------------------------------------------------------
ffi::Error CsrTranspose(cudaStream_t stream, int64_t plan, F64 csr_data, F64Out csr_data_T) {
    return Done("cpfem_csr_transpose", cpfem_csr_transpose(PlanOf(plan), csr_data.typed_data(), csr_data_T->typed_data(), stream));
}

}  // namespace

#define CPFEM_COMMON_ATTRS() \
    Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("plan").Attr<double>("dt").Attr<cpfem_material>("mat")

XLA_FFI_DEFINE_HANDLER_SYMBOL(cpfem_update_state_ffi, UpdateState,
                              ffi::Ffi::Bind().CPFEM_COMMON_ATTRS().Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>()
                                  .Ret<ffi::Buffer<ffi::S64>>().RemainingArgs());
XLA_FFI_DEFINE_HANDLER_SYMBOL(cpfem_avg_stress_ffi, AvgStress,
                              ffi::Ffi::Bind().CPFEM_COMMON_ATTRS().Arg<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::S64>>().RemainingArgs());
XLA_FFI_DEFINE_HANDLER_SYMBOL(cpfem_update_avg_ffi, UpdateAvg,
                              ffi::Ffi::Bind().CPFEM_COMMON_ATTRS().Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>()
                                  .Ret<ffi::Buffer<ffi::S64>>().RemainingArgs());
XLA_FFI_DEFINE_HANDLER_SYMBOL(cpfem_residual_ffi, Residual,
                              ffi::Ffi::Bind().CPFEM_COMMON_ATTRS().Arg<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::S64>>().RemainingArgs());
XLA_FFI_DEFINE_HANDLER_SYMBOL(cpfem_newton_update_ffi, NewtonUpdate,
                              ffi::Ffi::Bind().CPFEM_COMMON_ATTRS().Attr<int64_t>("want_V").Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>()
                                  .Ret<ffi::Buffer<ffi::S64>>().RemainingArgs());
XLA_FFI_DEFINE_HANDLER_SYMBOL(cpfem_point_eval_ffi, PointEval,
                              ffi::Ffi::Bind().CPFEM_COMMON_ATTRS().Attr<int64_t>("what").Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::S64>>().RemainingArgs());
XLA_FFI_DEFINE_HANDLER_SYMBOL(cpfem_dirichlet_ffi, Dirichlet,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("plan")
                                  .Arg<ffi::Buffer<ffi::S64>>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(cpfem_bicgstab_ffi, Bicgstab,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("plan").Attr<int64_t>("precond")
                                  .Attr<double>("tol").Attr<double>("atol").Attr<int64_t>("maxiter").Attr<int64_t>("iters_to_enqueue")
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::S64>>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(cpfem_point_jac_x_ffi, PointJacX,
                              ffi::Ffi::Bind().CPFEM_COMMON_ATTRS().Attr<int64_t>("nextra").Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>()
                                  .Ret<ffi::Buffer<ffi::S64>>().RemainingArgs());
XLA_FFI_DEFINE_HANDLER_SYMBOL(cpfem_vjp_params_ffi, VjpParams,
                              ffi::Ffi::Bind().CPFEM_COMMON_ATTRS().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::S64>>()
                                  .RemainingArgs());
XLA_FFI_DEFINE_HANDLER_SYMBOL(cpfem_csr_transpose_ffi, CsrTranspose,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("plan").Arg<F64>().Ret<F64>());
