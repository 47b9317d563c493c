"""Host-side mirror of the reference's plugin interface for the hot path.

`Problem` / `FiniteElement` carry the attribute and method names the reference consumes from
`jax_fem.problem.Problem` (singlecrystal_copper/models_copper.py:9,84,277,315;
crystal_plasticity_OR_design/solver.py:119-133,244,281,290-293,391-392):

    problem.fes[0].cells / points / num_quads / num_total_nodes / vec / shape_grads / JxW
    problem.fes[0].node_inds_list / vec_inds_list / vals_list / update_Dirichlet_boundary_conditions
    problem.newton_update(sol_list) -> res_list   (side effect: problem.V / problem.csr_data)
    problem.compute_residual(sol_list) -> res_list
    problem.I, problem.J, problem.V, problem.unflatten_fn_sol_list, num_total_dofs_all_vars, offset

`CrystalPlasticityBase` carries the model-level methods of `CrystalPlasticity(Problem)`
(models_copper.py:51-319): custom_init, get_tensor_map, get_maps, set_params, update_int_vars_gp,
compute_avg_stress, inspect_interval_vars, attribute `dt`, attribute `internal_vars`.

Arrays are torch CUDA tensors where the reference uses JAX arrays; numpy inputs are accepted and copied
to the device.  All arithmetic is done by the CUDA kernels behind the C ABI (api.Plan); nothing here
computes stresses, tangents or state on the host.
"""
from __future__ import annotations

import os
from typing import List, Optional

import numpy as onp
import torch

from . import api
from .generate_mesh import Mesh

SLIP_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')


def get_rot_mat(q):
    """models_copper.py:37-45 (quaternion (w,x,y,z) -> rotation matrix), numpy, batched."""
    q = onp.asarray(q, dtype=onp.float64)
    q0, q1, q2, q3 = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    return onp.stack([
        onp.stack([q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, 2 * q1 * q2 - 2 * q0 * q3, 2 * q1 * q3 + 2 * q0 * q2], -1),
        onp.stack([2 * q1 * q2 + 2 * q0 * q3, q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, 2 * q2 * q3 - 2 * q0 * q1], -1),
        onp.stack([2 * q1 * q3 - 2 * q0 * q2, 2 * q2 * q3 + 2 * q0 * q1, q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3], -1)], -2)


get_rot_mat_vmap = get_rot_mat


def _hex8_ref_grads():
    g = onp.array([(1 - 1 / onp.sqrt(3)) / 2, (1 + 1 / onp.sqrt(3)) / 2])
    nodes = onp.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]])
    quad = onp.array([[g[i], g[j], g[k]] for i in range(2) for j in range(2) for k in range(2)])
    dN = onp.zeros((8, 8, 3))
    for q in range(8):
        f = onp.where(nodes == 1, quad[q][None, :], 1 - quad[q][None, :])        # (a, 3)
        s = onp.where(nodes == 1, 1.0, -1.0)
        dN[q, :, 0] = s[:, 0] * f[:, 1] * f[:, 2]
        dN[q, :, 1] = s[:, 1] * f[:, 0] * f[:, 2]
        dN[q, :, 2] = s[:, 2] * f[:, 0] * f[:, 1]
    return dN


class FiniteElement:
    """The slice of jax_fem.fe.FiniteElement the reference touches."""

    def __init__(self, mesh: Mesh, vec, dim, ele_type, dirichlet_bc_info):
        if ele_type != 'HEX8' or vec != 3 or dim != 3:
            raise NotImplementedError('the crystal-plasticity hot path is HEX8, vec = dim = 3')
        self.mesh = mesh
        self.points = mesh.points
        self.cells = mesh.cells
        self.vec, self.dim, self.ele_type = vec, dim, ele_type
        self.num_cells = len(self.cells)
        self.num_total_nodes = len(self.points)
        self.num_total_dofs = self.num_total_nodes * vec
        self.num_quads = 8
        self.num_nodes = 8
        self._shape_grads = None
        self._JxW = None
        self.dirichlet_bc_info = dirichlet_bc_info
        self.node_inds_list, self.vec_inds_list, self.vals_list = [], [], []
        if dirichlet_bc_info is not None:
            self.update_Dirichlet_boundary_conditions(dirichlet_bc_info)

    # geometry arrays kept for API parity (the kernels recompute them from points/cells on the fly)
    def _geometry(self):
        dN = _hex8_ref_grads()
        X = self.points[self.cells]
        jac = onp.einsum('cai,qaj->cqij', X, dN)
        self._shape_grads = onp.einsum('qaj,cqji->cqai', dN, onp.linalg.inv(jac))
        self._JxW = onp.linalg.det(jac) * 0.125

    @property
    def shape_grads(self):
        if self._shape_grads is None:
            self._geometry()
        return self._shape_grads

    @property
    def JxW(self):
        if self._JxW is None:
            self._geometry()
        return self._JxW

    @staticmethod
    def _eval_location(fn, points):
        try:
            m = onp.asarray(fn(points.T))
            if m.shape == (len(points),):
                return m.astype(bool)
        except Exception:
            pass
        return onp.array([bool(fn(p)) for p in points])

    @staticmethod
    def _eval_value(fn, pts):
        try:
            v = onp.asarray(fn(pts.T), dtype=onp.float64)
            if v.shape == (len(pts),):
                return v
            if v.shape == ():
                return onp.full(len(pts), float(v))
        except Exception:
            pass
        return onp.array([float(fn(p)) for p in pts], dtype=onp.float64)

    def update_Dirichlet_boundary_conditions(self, dirichlet_bc_info):
        """jax_fem: node_inds_list[i] = argwhere(location_fn(points)); vec_inds_list[i]; vals_list[i]."""
        self.dirichlet_bc_info = dirichlet_bc_info
        self.bc_version = getattr(self, 'bc_version', 0) + 1        # lets the device solver cache the flat dof lists
        location_fns, vecs, value_fns = dirichlet_bc_info
        self.node_inds_list, self.vec_inds_list, self.vals_list = [], [], []
        for loc, v, val in zip(location_fns, vecs, value_fns):
            inds = onp.argwhere(self._eval_location(loc, self.points)).reshape(-1)
            self.node_inds_list.append(inds)
            self.vec_inds_list.append(onp.full(len(inds), v, dtype=onp.int32))
            self.vals_list.append(self._eval_value(val, self.points[inds]))


class Problem:
    """The slice of jax_fem.problem.Problem on the hot path (see module docstring)."""

    def __init__(self, mesh, vec=3, dim=3, ele_type='HEX8', dirichlet_bc_info=None, additional_info=(), device=None,
                 keep_V=False):
        self.mesh = [mesh]
        self.vec, self.dim, self.ele_type = [vec], dim, [ele_type]
        self.fes = [FiniteElement(mesh, vec, dim, ele_type, dirichlet_bc_info)]
        self.num_vars = 1
        self.offset = [0]
        self.num_total_dofs_all_vars = self.fes[0].num_total_dofs
        self.num_cells = self.fes[0].num_cells
        self.additional_info = additional_info
        self.keep_V = keep_V
        self.dt = None
        self.internal_vars = []
        self._I = self._J = None
        self._V = None
        self.csr_data = None
        self.last_status = None
        self._last_sol = None
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.custom_init(*additional_info)
        self.plan = api.Plan(self.fes[0].cells, self.fes[0].points, self.slip_table, device=self.device)
        self.internal_vars = [api._dev_f64(v, self.device) for v in self.internal_vars]

    def custom_init(self, *args):
        raise NotImplementedError

    # ---- dof bookkeeping (solver.py:119-133,391) ---------------------------------------------
    def unflatten_fn_sol_list(self, dofs):
        return [dofs.reshape(self.fes[0].num_total_nodes, self.fes[0].vec)]

    # ---- COO indices (jax_fem rule, consumed at solver.py:281) -------------------------------
    def _coo(self):
        cells = self.fes[0].cells.astype(onp.int64)
        inds = (3 * cells[:, :, None] + onp.arange(3)[None, None, :]).reshape(len(cells), -1)
        self._I = onp.repeat(inds[:, :, None], 24, axis=2).reshape(-1)
        self._J = onp.repeat(inds[:, None, :], 24, axis=1).reshape(-1)

    @property
    def I(self):
        if self._I is None:
            self._coo()
        return self._I

    @property
    def J(self):
        if self._J is None:
            self._coo()
        return self._J

    @property
    def V(self):
        """problem.V (nc*576,) in the reference layout.  Materialised only on request (keep_V) - the CSR data
        assembled on the device (problem.csr_data on problem.csr_pattern()) is the product's hand-off."""
        if self._V is None:
            if self._last_sol is None:
                raise RuntimeError('problem.V: call newton_update first')
            keep, self.keep_V = self.keep_V, True
            self.newton_update([self._last_sol])
            self.keep_V = keep
        return self._V

    def csr_pattern(self):
        return self.plan.csr_pattern()

    def csr_scipy(self, data=None):
        """Assembled tangent as a scipy CSR (host copy) - what get_A builds at solver.py:281.  `data`: other values on the
        same pattern (e.g. the transposed ones of implicit_vjp)."""
        import scipy.sparse
        ip, ix = self.plan.csr_pattern()
        n = self.num_total_dofs_all_vars
        data = self.csr_data if data is None else data
        return scipy.sparse.csr_array((data.cpu().numpy(), ix.cpu().numpy(), ip.cpu().numpy()), shape=(n, n))

    # ---- the two assembly entry points ---------------------------------------------------------
    def newton_update(self, sol_list):
        sol = api._dev_f64(sol_list[0], self.device)
        self._last_sol = sol
        self._drop_fused()
        st = self.plan.new_status()
        res, self.csr_data, V = self.plan.newton_update(self.material, sol, self.internal_vars, self.dt,
                                                        want_csr=True, want_V=self.keep_V, status=st)
        self._V = V
        self.last_status = st
        return [res]

    def compute_residual(self, sol_list):
        sol = api._dev_f64(sol_list[0], self.device)
        st = self.plan.new_status()
        res = self.plan.residual(self.material, sol, self.internal_vars, self.dt, status=st)
        self.last_status = st
        return [res]

    def set_params(self, params):
        """models_copper.py:284-285."""
        self.internal_vars = [api._dev_f64(v, self.device) for v in params]
        self._drop_fused()

    def _drop_fused(self):
        pass


class CrystalPlasticityBase(Problem):
    """Model-level methods shared by the four `models_*.py` files of the reference; subclasses only supply
    the parameter set (class attributes below) - the reference files differ in nothing else."""
    slip_file = None          # (ns, 6) numpy table
    gss_initial = None
    C11 = C12 = C44 = None
    h = t_sat = gss_a = xm = None
    r = 1.0
    ao = 0.001
    max_sub_step = 5
    tol = 1e-8

    def custom_init(self, quat, cell_ori_inds):
        """models_copper.py:52-133."""
        self.slip_table = onp.asarray(self.slip_file, dtype=onp.float64)
        ns = len(self.slip_table)
        nc, nq = self.fes[0].num_cells, self.fes[0].num_quads
        quat = onp.asarray(quat, dtype=onp.float64)
        # JAX clamps out-of-range gathers; numpy would raise
        ori = onp.clip(onp.asarray(cell_ori_inds, dtype=onp.int64), 0, len(quat) - 1)
        rot_mats = get_rot_mat(quat)[ori]
        Fp_inv_gp = onp.tile(onp.eye(3)[None, None], (nc, nq, 1, 1))
        slip_resistance_gp = self.gss_initial * onp.ones((nc, nq, ns))
        slip_gp = onp.zeros_like(slip_resistance_gp)
        rot_mats_gp = onp.repeat(rot_mats[:, None, :, :], nq, axis=1)
        self.material = api.make_material(self.C11, self.C12, self.C44, self.h, self.t_sat, self.gss_a, self.xm, self.r,
                                          self.ao, self.tol, self.max_sub_step)
        self.internal_vars = [Fp_inv_gp, slip_resistance_gp, slip_gp, rot_mats_gp]

    # ---- models_copper.py:135-137 --------------------------------------------------------------
    def get_tensor_map(self):
        tensor_map, _ = self.get_maps()
        return tensor_map

    def get_maps(self):
        """Returns (tensor_map, update_int_vars_map) acting on BATCHES of points: the reference returns scalar
        functions that jax_fem vmaps over (cell, quad); here the batch axis is explicit (leading axes are
        flattened) because the device kernel is the vmap."""
        def tensor_map(u_grad, *state):
            ug = api._dev_f64(u_grad, self.device)
            lead = ug.shape[:-2]
            flat = []
            for s in state:
                t = api._dev_f64(s, self.device)
                flat.append(t.reshape((-1,) + tuple(t.shape[len(lead):])))
            P, _ = self.plan.point_stress_tangent(self.material, ug.reshape(-1, 3, 3), flat, self.dt, want_tangent=False)
            return P.reshape(*lead, 3, 3)

        def update_int_vars_map(u_grad, *state):
            """models_copper.py:164-169,267-269: (u_grad, *state) -> (Fp_inv_new, slip_resistance_new, slip_new); a single
            point (u_grad (3,3), state arrays without batch axes) or any batch of points."""
            ug = api._dev_f64(u_grad, self.device)
            lead = ug.shape[:-2]
            flat = []
            for k, s in enumerate(state):
                t = api._dev_f64(s, self.device)
                tail = t.shape[len(lead):]
                flat.append(t.reshape((-1,) + tuple(tail)))
            new = self.plan.point_update_state(self.material, ug.reshape(-1, 3, 3), flat, self.dt)
            return tuple(n.reshape(tuple(lead) + tuple(n.shape[1:])) for n in new)
        return tensor_map, update_int_vars_map

    def tensor_map_jacobian(self, u_grad, *state):
        """jax.jacfwd(tensor_map) w.r.t. u_grad on a batch: (..., 3, 3, 3, 3)."""
        ug = api._dev_f64(u_grad, self.device)
        lead = ug.shape[:-2]
        flat = [api._dev_f64(s, self.device) for s in state]
        P, A = self.plan.point_stress_tangent(self.material, ug.reshape(-1, 3, 3), flat, self.dt, want_tangent=True)
        return P.reshape(*lead, 3, 3), A.reshape(*lead, 3, 3, 3, 3)

    # The drivers call compute_avg_stress(sol, params) and then update_int_vars_gp(sol, params) with the SAME arguments
    # (singlecrystal_copper.py:205,227): both need the same converged local solve.  With fuse_avg_stress (default) the
    # first of the two calls runs the fused kernel and keeps the other result for the second call; the key is the
    # identity + version counter of every argument tensor, dt and the bytes of the material, so any change of the inputs
    # falls back to a fresh run.  The kept result is dropped by the second call of the pair and by set_params /
    # newton_update (the next things a driver loop does), so a lone call does not pin a second copy of the state for long.
    # The library's own kernels write through raw pointers without bumping torch's version counter: callers that update a
    # params tensor IN PLACE through api.Plan(..., out=...) between the two calls must set fuse_avg_stress = False.
    fuse_avg_stress = True

    @staticmethod
    def _fuse_key(sol, params, dt):
        ts = [sol] + list(params)
        if not all(isinstance(t, torch.Tensor) and t.is_cuda for t in ts):
            return None
        return (float(dt),) + tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in ts)

    def _drop_fused(self):
        self._fuse_cache = None

    def _fused(self, sol, params):
        """Runs (or recalls) the fused update + average stress for these arguments: returns (new_state, sigma) or None."""
        if not self.fuse_avg_stress:
            return None
        key = self._fuse_key(sol, params, self.dt)
        if key is None:
            return None
        key = key + (bytes(self.material),)       # a changed parameter set is a different computation
        hit = getattr(self, '_fuse_cache', None)
        if hit is not None and hit[0] == key:
            self._fuse_cache = None               # second call of the pair: hand over and forget (holds no extra memory)
            return hit[1], hit[2]
        st = self.plan.new_status()
        new, sigma = self.plan.update_state_avg_stress(self.material, sol, list(params), self.dt, status=st)
        self.last_status = st
        self._fuse_cache = (key, new, sigma, (sol, list(params)))   # the inputs stay alive, so data_ptr cannot be recycled
        return new, sigma

    # ---- models_copper.py:273-282 --------------------------------------------------------------
    def update_int_vars_gp(self, sol, params):
        if all(isinstance(v, torch.Tensor) and not v.is_cuda for v in params):
            # host-resident state (torch CPU tensors, ideally pinned): stream it through the device, return host tensors
            st = self.plan.new_status()
            new = self.plan.update_state_host(self.material, sol, list(params), self.dt, status=st)
            self.last_status = st
            return [new[0], new[1], new[2]] + list(params[3:])
        f = self._fused(sol, params)
        if f is not None:
            new = f[0]
            return [new[0], new[1], new[2]] + list(params[3:])
        params = [api._dev_f64(v, self.device) for v in params]
        st = self.plan.new_status()
        new = self.plan.update_state(self.material, sol, params, self.dt, status=st)
        self.last_status = st
        return [new[0], new[1], new[2]] + list(params[3:])

    # ---- models_copper.py:297-319 --------------------------------------------------------------
    def compute_avg_stress(self, sol, params):
        f = self._fused(sol, params)
        if f is not None:
            return f[1]
        params = [api._dev_f64(v, self.device) for v in params]
        return self.plan.avg_stress(self.material, sol, params, self.dt)

    def vjp_params(self, sol, params, adjoint):
        """adjoint . d(compute_residual)/d(params): what jax.vjp of the constraint function gives inside implicit_vjp
        (crystal_plasticity_OR_design/solver.py:832-848), as a list shaped like `params`."""
        params = [api._dev_f64(v, self.device) for v in params]
        st = self.plan.new_status()
        out = self.plan.vjp_params(self.material, sol, params, self.dt, adjoint, status=st)
        self.last_status = st
        return out

    def inspect_interval_vars(self, params):
        """models_copper.py:287-295 (post-processing only)."""
        Fp_inv_gp, slip_resistance_gp, slip_gp = params[0], params[1], params[2]
        F_p = torch.linalg.inv(torch.as_tensor(Fp_inv_gp)[0, 0])
        return float(F_p[2, 2]), float(slip_resistance_gp[0, 0, 0]), float(slip_gp[0, 0, 0])
