"""Smaller pieces of the drop-in surface on the GPU: the scalar maps of get_maps (models_copper.py:263-271), input
validation of C_gp, duplicate Dirichlet dofs, the fused-call cache, the launch counter."""
import numpy as np
import pytest

import cases
import cpfem_oracle as O

pytestmark = pytest.mark.gpu


def _problem(N=2):
    from cpfem_b200.generate_mesh import Mesh
    from cpfem_b200.models_304steel import CrystalPlasticity
    pts, cells = O.box_mesh(N, N, N)
    rng = np.random.default_rng(0)
    quat = cases.rand_quat(rng, 3)
    ori = rng.integers(0, 3, size=len(cells))
    p = CrystalPlasticity(Mesh(pts, cells), vec=3, dim=3, ele_type='HEX8', dirichlet_bc_info=None, additional_info=(quat, ori))
    p.dt = 2e-3
    return p, pts, cells


def test_get_maps_scalar_and_batched():
    """tensor_map / update_int_vars_map on ONE point (the reference's signature: u_grad (3,3), state without batch axes)
    and on a (cell, quad) batch give what update_int_vars_gp gives for the same points."""
    import torch
    p, pts, cells = _problem()
    eps = 2.5e-3
    sol = torch.as_tensor(np.stack([-0.3 * eps * pts[:, 0], -0.3 * eps * pts[:, 1], eps * pts[:, 2]], 1), device='cuda')
    params = p.internal_vars
    new = p.update_int_vars_gp(sol, params)
    tensor_map, update_map = p.get_maps()
    ug = torch.einsum('cai,cqaj->cqij', sol[torch.as_tensor(cells.astype(np.int64), device='cuda')],
                      torch.as_tensor(p.fes[0].shape_grads, device='cuda'))
    nb = update_map(ug, *params)                                   # batched over (cell, quad)
    for a, b in zip(nb, new[:3]):
        assert a.shape == b.shape and float((a - b).abs().max()) <= 1e-12 * max(float(b.abs().max()), 1e-6)
    c, q = 3, 5
    one = update_map(ug[c, q], *[v[c, q] for v in params])          # a single point, reference signature
    assert one[0].shape == (3, 3) and one[1].shape == (12,) and one[2].shape == (12,)
    for a, b in zip(one, new[:3]):
        assert float((a - b[c, q]).abs().max()) <= 1e-12 * max(float(b[c, q].abs().max()), 1e-6)
    P1 = tensor_map(ug[c, q], *[v[c, q] for v in params])
    Pb = tensor_map(ug, *params)
    assert P1.shape == (3, 3) and Pb.shape == (len(cells), 8, 3, 3) and float((P1 - Pb[c, q]).abs().max()) == 0.0
    assert float(Pb.abs().max()) > 100.0


def test_c_gp_must_be_cubic():
    """A per-point elastic tensor that is not of the cubic crystal-frame form is refused (the kernels read three entries)."""
    import torch
    from cpfem_b200 import Plan, make_material
    pts, cells = O.box_mesh(2, 2, 2)
    nc = len(cells)
    params, ph, quat, ori = cases.dp_params(nc, seed=1)
    plan = Plan(cells, pts, O.SLIP_BCC24)
    f = O.dp_ferrite()
    m = make_material(f.C11, f.C12, f.C44, f.h, f.t_sat, f.gss_a, f.xm, f.r, f.ao, f.tol, f.max_sub_step)
    sol = np.zeros((len(pts), 3))
    plan.update_state(m, sol, params, 0.2)                          # the reference's own construction passes
    bad = [p.copy() for p in params]
    bad[9][3, 2, 0, 1, 0, 2] = 5.0e3                                # one non-cubic entry at one point
    with pytest.raises(ValueError, match='not cubic'):
        plan.update_state(m, sol, bad, 0.2)
    rot = [p.copy() for p in params]
    R = O.get_rot_mat(cases.rand_quat(np.random.default_rng(2), 1))[0]
    rot[9][:] = np.einsum('ia,jb,kc,ld,abcd->ijkl', R, R, R, R, rot[9][0, 0])       # pre-rotated to the lab frame
    with pytest.raises(ValueError, match='not cubic'):
        plan.newton_update(m, sol, rot, 0.2)
    host = [torch.as_tensor(p) for p in bad]
    with pytest.raises(ValueError, match='not cubic'):
        plan.update_state_host(m, torch.as_tensor(sol), host, 0.2)


def test_duplicate_dirichlet_dofs_last_one_wins():
    """A dof named by two Dirichlet sets takes the value of the LAST set, like the reference's sequential loop
    (solver.py:125-131) - and the device kernel sees every dof once."""
    import torch
    from cpfem_b200 import solver as S
    p, pts, cells = _problem()
    bottom = lambda x: np.isclose(x[2], 0.)
    p.fes[0].update_Dirichlet_boundary_conditions([[bottom, bottom, bottom], [2, 2, 0], [lambda x: 1.0, lambda x: 2.0, lambda x: 0.5]])
    rows, vals = S._bc_rows_vals(p)
    rows, vals = rows.cpu().numpy(), vals.cpu().numpy()
    assert len(np.unique(rows)) == len(rows)
    nb = int(np.isclose(pts[:, 2], 0.).sum())
    assert len(rows) == 2 * nb
    z = rows % 3 == 2
    assert (vals[z] == 2.0).all() and (vals[~z] == 0.5).all()
    dofs = torch.zeros(3 * len(pts), dtype=torch.float64, device='cuda')
    res = S.apply_bc_vec(torch.ones_like(dofs), dofs, p)
    assert set(np.unique(res.cpu().numpy()[rows[z]])) == {-2.0}


def test_fused_cache_sees_material_and_is_dropped():
    import torch
    p, pts, cells = _problem()
    eps = 2.5e-3
    sol = torch.as_tensor(np.stack([-0.3 * eps * pts[:, 0], -0.3 * eps * pts[:, 1], eps * pts[:, 2]], 1), device='cuda')
    params = p.internal_vars
    s1 = p.compute_avg_stress(sol, params)
    assert p._fuse_cache is not None
    p.material.h = p.material.h * 2.0                                # another parameter set: the kept state is stale
    new = p.update_int_vars_gp(sol, params)
    ref = p.plan.update_state(p.material, sol, params, p.dt)
    assert all(torch.equal(a, b) for a, b in zip(new[:3], ref))
    p.compute_avg_stress(sol, params)
    p.set_params(params)
    assert p._fuse_cache is None
    p.compute_avg_stress(sol, params)
    p.newton_update([sol])
    assert p._fuse_cache is None


def test_launch_counter_counts_kernels():
    import torch
    import cpfem_b200
    p, pts, cells = _problem()
    sol = torch.zeros(len(pts), 3, dtype=torch.float64, device='cuda')
    L = cpfem_b200.lib()
    n0 = int(L.cpfem_launch_count())
    p.plan.update_state(p.material, sol, p.internal_vars, p.dt)
    n1 = int(L.cpfem_launch_count())
    p.plan.newton_update(p.material, sol, p.internal_vars, p.dt)
    n2 = int(L.cpfem_launch_count())
    assert n1 - n0 == 1 and n2 - n1 == 2                              # one chunk: point kernel + element kernel
