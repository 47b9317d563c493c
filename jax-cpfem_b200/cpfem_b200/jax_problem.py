"""Drop-in for the reference's `CrystalPlasticity(Problem)` classes ON JAX: the hot-path methods are replaced by XLA-FFI
custom calls into libcpfem_b200.so (csrc/cpfem_ffi.cc), everything else - `custom_init`, the driver loop, the L3 solver
- stays the reference's own code.

    from applications.singlecrystal_copper.models_copper import CrystalPlasticity as Reference
    from cpfem_b200.jax_problem import accelerate
    CrystalPlasticity = accelerate(Reference, 'copper')          # same name, same constructor, same methods
    problem = CrystalPlasticity(mesh, vec=3, dim=3, ele_type='HEX8', dirichlet_bc_info=..., additional_info=(quat, cell_ori_inds))

Overridden (reference lines; jax_fem = deepmodeling/jax-fem, imported at models_copper.py:9):
    newton_update(sol_list)          jax_fem Problem.newton_update, consumed at crystal_plasticity_OR_design/solver.py:392
    compute_residual(sol_list)       jax_fem Problem.compute_residual, consumed at solver.py:244
    update_int_vars_gp(sol, params)  models_copper.py:273-282
    compute_avg_stress(sol, params)  models_copper.py:297-319
    get_tensor_map() / get_maps()    models_copper.py:135-137, 263-271 (batched: the device kernel is the vmap)
Added for the adjoint (implicit_vjp, solver.py:801-853): point_jacobians (f_jvp's jac_x / jac_y), vjp_params, csr_transpose.
Added: `csr_data` (device) on `csr_indptr` / `csr_indices` (the pattern scipy would build at solver.py:281, bit for bit),
`csr_scipy()` for L3's `get_A`, `last_status`, `keep_V` (materialise the reference's `problem.V` as well).

The parameter values live inside the reference's `get_maps` closures (models_copper.py:141-151), out of reach of a
subclass, so `accelerate` takes them from `param_sets.PRESETS[name]` or from an explicit dictionary.

JAX and jax_fem are not installable in the container this repo is built in: this module is exercised by
tests/test_jax_ffi.py only where `import jax` works (skipped otherwise) and by tools/verify_upstream.py on a JAX box.
"""
from __future__ import annotations

import numpy as onp

from . import jax_ffi
from .param_sets import PRESETS


def _jax():
    import jax
    import jax.numpy as jnp
    jax.config.update('jax_enable_x64', True)
    return jax, jnp


class B200HotPath:
    """Mixin placed in front of the reference's CrystalPlasticity class (see `accelerate`)."""
    b200_preset = None          # dict(material=..., slip=..., gss_initial=...)
    keep_V = False              # also return the reference's COO values problem.V (nc*576): 4.6 kB per cell

    # ---- set-up: after the reference's custom_init built internal_vars, create the plan ---------------------------
    def custom_init(self, *args, **kw):
        super().custom_init(*args, **kw)
        jax_ffi.register()
        fe = self.fes[0]
        self._b200_plan = jax_ffi.JaxPlan(onp.asarray(fe.cells), onp.asarray(fe.points), self.b200_preset['slip'])
        self._b200_mat = jax_ffi.material_attr(self.b200_preset['material'])
        self.csr_indptr, self.csr_indices = self._b200_plan.indptr, self._b200_plan.indices
        self.csr_data = None
        self.last_status = None

    def _call(self, target, out_types, *arrays, **attrs):
        ffi = jax_ffi.jax_ffi_module()
        fn = ffi.ffi_call(target, out_types, vmap_method='sequential')
        return fn(*arrays, plan=self._b200_plan.handle, **attrs)

    def _common(self):
        return dict(dt=onp.float64(self.dt), mat=self._b200_mat)

    # ---- jax_fem Problem.newton_update (solver.py:392) ------------------------------------------------------------------
    def newton_update(self, sol_list):
        jax, jnp = _jax()
        sol = jnp.asarray(sol_list[0], dtype=jnp.float64)
        p = self._b200_plan
        nV = p.nc * 576 if self.keep_V else 0
        out = (jax.ShapeDtypeStruct(sol.shape, jnp.float64), jax.ShapeDtypeStruct((p.nnz,), jnp.float64),
               jax.ShapeDtypeStruct((nV,), jnp.float64), jax.ShapeDtypeStruct((4,), jnp.int64))
        res, self.csr_data, V, self.last_status = self._call('cpfem_newton_update_ffi', out, sol, *self.internal_vars,
                                                             want_V=onp.int64(1 if self.keep_V else 0), **self._common())
        if self.keep_V:
            self.V = onp.asarray(V)                  # what jax_fem leaves for get_A (solver.py:281)
        return [res]

    def csr_scipy(self):
        """The matrix get_A builds from (V, I, J) at solver.py:281, taken from the device-assembled CSR instead."""
        import scipy.sparse
        n = self._b200_plan.ndof
        return scipy.sparse.csr_array((onp.asarray(self.csr_data), self.csr_indices, self.csr_indptr), shape=(n, n))

    # ---- jax_fem Problem.compute_residual (solver.py:244) ---------------------------------------------------------------
    def compute_residual(self, sol_list):
        jax, jnp = _jax()
        sol = jnp.asarray(sol_list[0], dtype=jnp.float64)
        out = (jax.ShapeDtypeStruct(sol.shape, jnp.float64), jax.ShapeDtypeStruct((4,), jnp.int64))
        res, self.last_status = self._call('cpfem_residual_ffi', out, sol, *self.internal_vars, **self._common())
        return [res]

    # ---- models_copper.py:273-282 -----------------------------------------------------------------------------------------
    def update_int_vars_gp(self, sol, params):
        jax, jnp = _jax()
        sol = jnp.asarray(sol, dtype=jnp.float64)
        out = tuple(jax.ShapeDtypeStruct(params[k].shape, jnp.float64) for k in range(3)) + (jax.ShapeDtypeStruct((4,), jnp.int64),)
        Fp, g, slip, self.last_status = self._call('cpfem_update_state_ffi', out, sol, *params, **self._common())
        return [Fp, g, slip] + list(params[3:])      # rot_mats (and the DP parameter arrays) pass through (:282)

    # ---- models_copper.py:297-319 -----------------------------------------------------------------------------------------
    def compute_avg_stress(self, sol, params):
        jax, jnp = _jax()
        sol = jnp.asarray(sol, dtype=jnp.float64)
        out = (jax.ShapeDtypeStruct((self._b200_plan.nc, 3, 3), jnp.float64), jax.ShapeDtypeStruct((4,), jnp.int64))
        sigma, self.last_status = self._call('cpfem_avg_stress_ffi', out, sol, *params, **self._common())
        return sigma

    def update_and_avg_stress(self, sol, params):
        """Both of the above from one local solve per point (the drivers call them back to back with the same arguments,
        singlecrystal_copper.py:205,227)."""
        jax, jnp = _jax()
        sol = jnp.asarray(sol, dtype=jnp.float64)
        out = tuple(jax.ShapeDtypeStruct(params[k].shape, jnp.float64) for k in range(3)) + \
            (jax.ShapeDtypeStruct((self._b200_plan.nc, 3, 3), jnp.float64), jax.ShapeDtypeStruct((4,), jnp.int64))
        Fp, g, slip, sigma, self.last_status = self._call('cpfem_update_avg_ffi', out, sol, *params, **self._common())
        return [Fp, g, slip] + list(params[3:]), sigma

    # ---- models_copper.py:135-137, 263-271 ----------------------------------------------------------------------------------
    def _point_eval(self, u_grad, state, what):
        jax, jnp = _jax()
        ug = jnp.asarray(u_grad, dtype=jnp.float64)
        lead = ug.shape[:-2]
        n = int(onp.prod(lead)) if lead else 1
        flat = [jnp.asarray(s, dtype=jnp.float64).reshape((n,) + tuple(jnp.shape(s)[len(lead):])) for s in state]
        ns = flat[1].shape[-1]
        z = lambda want, shape: shape if want else (0,)
        out = (jax.ShapeDtypeStruct((n, 3, 3), jnp.float64), jax.ShapeDtypeStruct(z(what & 1, (n, 3, 3, 3, 3)), jnp.float64),
               jax.ShapeDtypeStruct(z(what & 2, (n, 3, 3)), jnp.float64), jax.ShapeDtypeStruct(z(what & 2, (n, ns)), jnp.float64),
               jax.ShapeDtypeStruct(z(what & 2, (n, ns)), jnp.float64), jax.ShapeDtypeStruct((n, 3), jnp.int32),
               jax.ShapeDtypeStruct((4,), jnp.int64))
        P, A, Fp, g, slip, info, self.last_status = self._call('cpfem_point_eval_ffi', out, ug.reshape(n, 3, 3), *flat,
                                                               what=onp.int64(what), **self._common())
        rs = lambda a: a.reshape(tuple(lead) + tuple(a.shape[1:]))
        return rs(P), (rs(A) if what & 1 else None), ((rs(Fp), rs(g), rs(slip)) if what & 2 else None), rs(info)

    def get_maps(self):
        """(tensor_map, update_int_vars_map) with the reference's signatures (u_grad, *state); they accept one point or
        any batch of points - jax_fem's vmap over (cell, quad) is replaced by the kernel's own grid."""
        def tensor_map(u_grad, *state):
            return self._point_eval(u_grad, state, 0)[0]

        def update_int_vars_map(u_grad, *state):
            return self._point_eval(u_grad, state, 2)[2]
        return tensor_map, update_int_vars_map

    def get_tensor_map(self):
        return self.get_maps()[0]

    def tensor_map_jacobian(self, u_grad, *state):
        """P and jax.jacfwd(tensor_map)(u_grad) = dP/dH (..., 3, 3, 3, 3) from the hand-derived implicit-function tangent
        (models_copper.py:251-259)."""
        P, A, _, _ = self._point_eval(u_grad, state, 1)
        return P, A


    # ---- adjoint row: f_jvp's Jacobians and the pieces of implicit_vjp ----------------------------------
    def point_jacobians(self, u_grad, *state):
        """jac_x (n, 9, nx), jac_y (n, 9, 9) and y = S (n, 9) of f_jvp (models_copper.py:251-259) at the converged local
        solution; x in the reference's ravel order (51 columns, 56 / 161 in the calibration / DP forms)."""
        jax, jnp = _jax()
        ug = jnp.asarray(u_grad, dtype=jnp.float64)
        lead = ug.shape[:-2]                                   # one point, or any batch of points (like _point_eval)
        n = int(onp.prod(lead)) if lead else 1
        ug = ug.reshape(n, 3, 3)
        flat = [jnp.asarray(s, dtype=jnp.float64).reshape((n,) + tuple(jnp.shape(s)[len(lead):])) for s in state]
        ns = flat[1].shape[-1]
        nextra = {4: 0, 9: 5, 10: 6}[len(state)]
        nx = 27 + 2 * ns + (5 if nextra >= 5 else 0) + (81 if nextra >= 6 else 0)
        out = (jax.ShapeDtypeStruct((n, 9, nx), jnp.float64), jax.ShapeDtypeStruct((n, 9, 9), jnp.float64),
               jax.ShapeDtypeStruct((n, 9), jnp.float64), jax.ShapeDtypeStruct((4,), jnp.int64))
        jx, jy, S, self.last_status = self._call('cpfem_point_jac_x_ffi', out, ug, *flat, nextra=onp.int64(nextra), **self._common())
        return jx, jy, S

    def vjp_params(self, sol, params, adjoint):
        """adjoint (nnodes, 3) . d(compute_residual)/d(internal_vars): what `jax.vjp(partial_params_c_fn)(adjoint)` yields
        inside implicit_vjp (crystal_plasticity_OR_design/solver.py:832-848), as a list shaped like `params`.  Zero the
        adjoint on the Dirichlet dofs first; the caller applies implicit_vjp's final minus sign (:849)."""
        jax, jnp = _jax()
        sol = jnp.asarray(sol, dtype=jnp.float64)
        adj = jnp.asarray(adjoint, dtype=jnp.float64).reshape(sol.shape)
        shapes = [tuple(jnp.shape(p)) for p in params] + [(0,)] * (10 - len(params))
        out = tuple(jax.ShapeDtypeStruct(s, jnp.float64) for s in shapes) + (jax.ShapeDtypeStruct((4,), jnp.int64),)
        res = self._call('cpfem_vjp_params_ffi', out, sol, adj, *params, **self._common())
        self.last_status = res[10]
        return list(res[:len(params)])

    def csr_transpose(self, csr_data=None):
        """Values of A^T on the same pattern (`A.transpose()` of implicit_vjp, solver.py:844)."""
        jax, jnp = _jax()
        data = self.csr_data if csr_data is None else jnp.asarray(csr_data, dtype=jnp.float64)
        ffi = jax_ffi.jax_ffi_module()
        fn = ffi.ffi_call('cpfem_csr_transpose_ffi', jax.ShapeDtypeStruct(data.shape, jnp.float64), vmap_method='sequential')
        return fn(data, plan=self._b200_plan.handle)


def accelerate(reference_cls, preset):
    """`reference_cls` with its hot path on the B200 kernels.  `preset`: a key of param_sets.PRESETS ('copper', 'tantalum',
    '304steel', 'dpsteel') or a dict(material={cpfem_material fields}, slip=(ns, 6) table)."""
    p = PRESETS[preset] if isinstance(preset, str) else preset
    return type(reference_cls.__name__, (B200HotPath, reference_cls), {'b200_preset': p, '__module__': reference_cls.__module__})
