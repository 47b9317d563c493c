"""Device-resident mirror of the reference's row-elimination Newton solver for the hot path
(crystal_plasticity_OR_design/solver.py): same function names, arguments and control flow, but the residual, the CSR
matrix and every vector stay on the GPU - `get_A` hands the matrix assembled by `problem.newton_update` to the device
BiCGStab (`cpfem_bicgstab`) where it lies instead of building a scipy/PETSc matrix on the host.

    solver(problem, solver_options)          solver.py:310-437
    linear_incremental_solver                solver.py:213-237
    linear_solver / jax_solve                solver.py:19-48, 92-116  ('jax_solver' branch; 'umfpack_solver' goes through scipy
                                                                       on the host, for tests)
    line_search                              solver.py:240-276
    apply_bc_vec / get_A                     solver.py:119-133, 279-293  (one kernel: cpfem_apply_dirichlet)

    implicit_vjp                             solver.py:801-853  (adjoint solve on A^T + the parameter VJP kernel, row F5)
    ad_wrapper                               solver.py:856-874  (torch.autograd.Function instead of jax.custom_vjp)

Out of scope here: arc-length, dynamic relaxation, PETSc, P_mat constraints.
"""
from __future__ import annotations

import logging
import time

import numpy as onp
import torch

from . import api

logger = logging.getLogger('cpfem_b200.solver')


def _bc_rows_vals(problem):
    """Flat Dirichlet dof indices and values of problem.fes[0] (solver.py:125-131 / 290-292), cached on the problem
    until update_Dirichlet_boundary_conditions replaces the lists."""
    fe = problem.fes[0]
    key = (getattr(fe, 'bc_version', 0), len(fe.node_inds_list))
    hit = getattr(problem, '_bc_cache', None)
    if hit is None or hit[0] != key:
        if len(fe.node_inds_list):
            rows = onp.concatenate([onp.asarray(n, dtype=onp.int64) * fe.vec + onp.asarray(v, dtype=onp.int64) + problem.offset[0]
                                    for n, v in zip(fe.node_inds_list, fe.vec_inds_list)])
            vals = onp.concatenate([onp.asarray(v, dtype=onp.float64) for v in fe.vals_list])
            # A dof named by several Dirichlet sets: the reference applies the sets one after the other (solver.py:125-131),
            # so the LAST one wins.  One kernel thread per entry would race on such a dof - keep its last occurrence only.
            _, first_rev = onp.unique(rows[::-1], return_index=True)
            if len(first_rev) != len(rows):
                keep = onp.sort(len(rows) - 1 - first_rev)
                rows, vals = rows[keep], vals[keep]
        else:
            rows, vals = onp.zeros(0, onp.int64), onp.zeros(0, onp.float64)
        dev = problem.device
        hit = (key, torch.as_tensor(rows, device=dev), torch.as_tensor(vals, device=dev))
        problem._bc_cache = hit
    return hit[1], hit[2]


def apply_bc_vec(res_vec, dofs, problem, scale=1.):
    """solver.py:119-133: res[bc] = dofs[bc] - scale * vals (in place on the device vector)."""
    rows, vals = _bc_rows_vals(problem)
    if rows.numel():
        problem.plan.apply_dirichlet(rows, vals * scale if scale != 1. else vals, dofs, res=res_vec, csr_data=None)
    return res_vec


def get_A(problem, solver_options=None):
    """solver.py:279-293: the reference builds scipy CSR from (V, I, J), copies it into PETSc and zeroes the Dirichlet
    rows (diag = 1).  Here the CSR data was assembled on the device by newton_update; only the rows are rewritten."""
    rows, vals = _bc_rows_vals(problem)
    if rows.numel():
        problem.plan.apply_dirichlet(rows, vals, problem._last_sol.reshape(-1), res=None, csr_data=problem.csr_data)
    return problem.csr_data


def jax_solve(problem, A, b, x0, precond, restarts=0):
    """solver.py:19-48: Jacobi-preconditioned BiCGStab (tol = atol = 1e-10, maxiter = 10000) + acceptance test.

    `restarts` (default 0 = the reference's behaviour): BiCGStab can break down (JAX's codes -10 / -11: rho or omega
    vanish - it happens on the tangents of the drivers whose boundary conditions leave a rigid rotation free, SURVEY
    App. H.1).  The reference then returns the stagnated iterate, which passes its `err < 0.1` test however poor it is, and
    the Newton loop absorbs the damage.  With `restarts` > 0 (solver_options['jax_solver']['restarts']) the solve is
    restarted from that iterate (at most `restarts` times) while it is still above the requested tolerance, so the
    increment - and with it the run-to-run reproducibility of the Newton path - does not depend on where a breakdown
    happens to strike.  problem.last_linear_restarts records how many restarts a solve took."""
    x, k, err = problem.plan.bicgstab(A, b, x0=x0, precond=precond, tol=1e-10, atol=1e-10, maxiter=10000)
    total = max(k, 0)
    tries = 0
    while k < 0 and tries < restarts and err > 1e-10 * max(float(torch.linalg.norm(b)), 1.0):
        logger.debug('device BiCGStab: breakdown code %d at res = %g, restarting from the current iterate', k, err)
        x, k, err = problem.plan.bicgstab(A, b, x0=x, precond=precond, tol=1e-10, atol=1e-10, maxiter=10000)
        total += max(k, 0)
        tries += 1
    logger.debug('device BiCGStab: %d iterations, res = %g', total, err)
    problem.last_linear_iterations = total if k >= 0 else k
    problem.last_linear_restarts = tries
    assert err < 0.1, f'linear solver failed to converge with err = {err}'
    return x


def umfpack_solve(problem, A, b):
    """solver.py:50-61 (host, scipy): kept for tests and tiny problems."""
    import scipy.sparse.linalg
    Asp = problem.csr_scipy(A)                   # the values handed in (get_A's, or their transpose in implicit_vjp)
    x = scipy.sparse.linalg.spsolve(Asp.tocsc(), b.cpu().numpy())
    return torch.as_tensor(x, device=problem.device)


def linear_solver(problem, A, b, x0, solver_options):
    """solver.py:92-116."""
    if len(solver_options.keys() & {'jax_solver', 'umfpack_solver', 'petsc_solver', 'custom_solver'}) == 0:
        solver_options['jax_solver'] = {}
    if 'jax_solver' in solver_options:
        precond = solver_options['jax_solver'].get('precond', True)
        return jax_solve(problem, A, b, x0, precond, restarts=solver_options['jax_solver'].get('restarts', 0))
    if 'umfpack_solver' in solver_options:
        return umfpack_solve(problem, A, b)
    if 'custom_solver' in solver_options:
        return solver_options['custom_solver'](A, b, x0, solver_options)
    raise NotImplementedError('petsc_solver is not part of the device path (SURVEY section 8: out of scope)')


def line_search(problem, dofs, inc):
    """solver.py:240-276: up to three halvings of the step while the residual norm decreases."""
    def res_norm_fn(alpha):
        d = dofs + alpha * inc
        res = problem.compute_residual(problem.unflatten_fn_sol_list(d))[0].reshape(-1)
        return float(torch.linalg.norm(apply_bc_vec(res, d, problem)))
    alpha = 1.
    res_norm = res_norm_fn(alpha)
    for _ in range(3):
        alpha *= 0.5
        res_norm_half = res_norm_fn(alpha)
        if res_norm_half > res_norm:
            alpha *= 2.
            break
        res_norm = res_norm_half
    return dofs + alpha * inc


def linear_incremental_solver(problem, res_vec, A, dofs, solver_options):
    """solver.py:213-237.  x0 is exact on the Dirichlet dofs: assign_bc(0) - copy_bc(dofs)."""
    b = -res_vec
    rows, vals = _bc_rows_vals(problem)
    x0 = torch.zeros_like(dofs)
    if rows.numel():
        x0[rows] = vals - dofs[rows]
    inc = linear_solver(problem, A, b, x0, solver_options)
    if solver_options.get('line_search_flag', False):
        return line_search(problem, dofs, inc)
    return dofs + inc


def solver(problem, solver_options=None):
    """solver.py:310-437: Newton iteration with row elimination; returns sol_list (device tensors)."""
    solver_options = {} if solver_options is None else solver_options
    start = time.time()
    dev = problem.device
    if 'initial_guess' in solver_options:
        dofs = torch.cat([api._dev_f64(s, dev).reshape(-1) for s in solver_options['initial_guess']]).clone()
    else:
        dofs = torch.zeros(problem.num_total_dofs_all_vars, dtype=torch.float64, device=dev)
    rel_tol = solver_options.get('rel_tol', 1e-8)
    tol = solver_options.get('tol', 1e-6)

    def newton_update_helper(dofs):
        sol_list = problem.unflatten_fn_sol_list(dofs)
        res_vec = problem.newton_update(sol_list)[0].reshape(-1)
        res_vec = apply_bc_vec(res_vec, dofs, problem)
        A = get_A(problem, solver_options)
        return res_vec, A

    res_vec, A = newton_update_helper(dofs)
    res_val = float(torch.linalg.norm(res_vec))
    res_val_initial = res_val
    rel_res_val = res_val / res_val_initial if res_val_initial > 0 else 0.0
    logger.debug('Before, l_2 res = %g, relative l_2 res = %g', res_val, rel_res_val)
    its = 0
    while (rel_res_val > rel_tol) and (res_val > tol):
        dofs = linear_incremental_solver(problem, res_vec, A, dofs, solver_options)
        res_vec, A = newton_update_helper(dofs)
        res_val = float(torch.linalg.norm(res_vec))
        rel_res_val = res_val / res_val_initial
        its += 1
        logger.debug('l_2 res = %g, relative l_2 res = %g', res_val, rel_res_val)
    assert onp.isfinite(res_val), 'res_val contains NaN, stop the program!'
    assert bool(torch.isfinite(dofs).all()), 'dofs contains NaN, stop the program!'
    problem.last_newton_iterations = its
    logger.info('Solve took %g [s]', time.time() - start)
    return problem.unflatten_fn_sol_list(dofs)


def implicit_vjp(problem, sol_list, params, v_list, adjoint_solver_options=None):
    """solver.py:801-853: the adjoint method.  Solves A^T lambda = v on the device (the tangent assembled at sol_list with
    its Dirichlet rows, transposed values on the same pattern), then contracts lambda with d(constraint)/d(params) by the
    per-point VJP kernel (cpfem_vjp_params) and returns MINUS that, as a list shaped like `params` - what the reference's
    jax.vjp(partial_params_c_fn)(adjoint) followed by tree_map(-x) gives.  The constraint rows of Dirichlet dofs
    (u - u_bc, apply_bc) do not depend on the parameters, so lambda is zeroed there before the contraction."""
    adjoint_solver_options = {} if adjoint_solver_options is None else adjoint_solver_options
    problem.set_params(params)
    problem.newton_update(sol_list)
    A = get_A(problem, adjoint_solver_options)
    v_vec = torch.cat([api._dev_f64(v, problem.device).reshape(-1) for v in v_list])
    AT = problem.plan.csr_transpose(A)
    adjoint_vec = linear_solver(problem, AT, v_vec, torch.zeros_like(v_vec), adjoint_solver_options)
    rows, _ = _bc_rows_vals(problem)
    lam = adjoint_vec.clone()
    if rows.numel():
        lam[rows] = 0.0
    grads = problem.vjp_params(sol_list[0], params, problem.unflatten_fn_sol_list(lam)[0])
    return [-g for g in grads]


def ad_wrapper(problem, solver_options=None, adjoint_solver_options=None):
    """solver.py:856-874: `fwd_pred(params) -> sol_list`, differentiable with respect to `params` by the adjoint method.
    The reference wraps the forward solve in `jax.custom_vjp`; the mirror wraps it in a `torch.autograd.Function` whose
    backward is `implicit_vjp` (one transposed linear solve + the per-point parameter VJP kernel), so that
    `torch.autograd.grad(objective(fwd_pred(params)[0]), params)` works on the device tensors.  Like the reference's, the
    wrapper differentiates the SOLVE only: what a driver does with `sol` afterwards (compute_avg_stress,
    update_int_vars_gp) runs on the kernels and is outside torch's tape."""
    solver_options = {} if solver_options is None else solver_options
    adjoint_solver_options = {} if adjoint_solver_options is None else adjoint_solver_options

    class _FwdPred(torch.autograd.Function):
        @staticmethod
        def forward(ctx, *params):
            plist = [p.detach() for p in params]
            problem.set_params(plist)
            sol_list = solver(problem, dict(solver_options))
            ctx.plist, ctx.sol_list = plist, [s.detach() for s in sol_list]
            return tuple(sol_list)

        @staticmethod
        def backward(ctx, *v_list):
            logger.info('Running backward and solving the adjoint problem...')
            v = [torch.zeros_like(s) if g is None else g for g, s in zip(v_list, ctx.sol_list)]
            grads = implicit_vjp(problem, ctx.sol_list, ctx.plist, v, dict(adjoint_solver_options))
            return tuple(grads)

    def fwd_pred(params):
        params = [api._dev_f64(p, problem.device) for p in params]
        return list(_FwdPred.apply(*params))

    return fwd_pred
