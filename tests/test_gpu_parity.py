"""GPU suite (-m gpu): the CUDA path, called through the C ABI (cpfem_b200.Plan / the Problem mirror), against the
CPU oracle on the same seeded inputs.  Tolerance: 1e-10 relative to the field maximum for fp64 values (north_star),
bit-exact for the CSR pattern."""
import os

import numpy as np
import pytest
import scipy.sparse
import torch

import cases
import cpfem_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _mat(m):
    from cpfem_b200 import make_material
    return make_material(m.C11, m.C12, m.C44, m.h, m.t_sat, m.gss_a, m.xm, m.r, m.ao, m.tol, m.max_sub_step)


def _dummy_plan(slip):
    from cpfem_b200 import Plan
    pts, cells = O.box_mesh(1, 1, 1)
    return Plan(cells, pts, slip)


@pytest.mark.parametrize('name', list(cases.MATERIALS))
def test_point_stress_tangent(name):
    """tensor_map and jacfwd(tensor_map) on explicit u_grads (cpfem_point_stress_tangent) vs oracle."""
    plan = None
    for step, mat, dt, H, A, g, sl, R in cases.point_history(name, n=75, steps=8):
        if plan is None:
            plan = _dummy_plan(mat.slip)
        pb = O.PointBatch(A, g, sl, R, mat)
        y, it_o, _ = pb.newton_solver(H, dt, True)
        P_o = pb.first_PK_stress(H, dt, y).numpy()
        T_o = pb.tangent(H, dt, y).numpy()
        st = plan.new_status()
        P, T = plan.point_stress_tangent(_mat(mat), H, [A, g, sl, R], dt, status=st)
        st = st.cpu().numpy()
        assert st[0] == 0 and st[1] == 0
        assert st[2] == it_o.max().item() and st[3] == it_o.sum().item()      # identical iteration counts
        assert cases.relerr(P.cpu().numpy(), P_o) < TOL
        assert cases.relerr(T.cpu().numpy(), T_o) < TOL


@pytest.mark.parametrize('name', ['304steel', 'tantalum'])
def test_fe_update_residual_tangent(name):
    """update_int_vars_gp / compute_residual / newton_update (V and CSR) / compute_avg_stress on a small distorted
    polycrystal mesh vs the oracle's FE layer."""
    from cpfem_b200 import Plan
    fe, mat, dt, sol, params, quat, ori = cases.small_fe_case(name, N=3, steps=6)
    plan = Plan(fe.cells, fe.points, mat.slip)
    m = _mat(mat)
    # state update
    new_o = fe.update_int_vars_gp(sol, params, dt)
    new = plan.update_state(m, sol, params, dt)
    for k in range(2):
        assert cases.relerr(new[k].cpu().numpy(), new_o[k]) < TOL
    assert np.abs(new[2].cpu().numpy() - new_o[2]).max() < TOL * max(np.abs(new_o[2]).max(), mat.ao * dt)
    # residual only
    res_o = fe.compute_residual(sol, params, dt)
    res = plan.residual(m, sol, params, dt)
    scale = np.abs(fe.cell_residual(sol, params, dt)).max()          # nodal sums cancel in the interior
    assert np.abs(res.cpu().numpy() - res_o).max() < TOL * scale
    # residual + tangent
    res_o2, V_o = fe.newton_update(sol, params, dt)
    res2, data, V = plan.newton_update(m, sol, params, dt, want_V=True)
    assert np.abs(res2.cpu().numpy() - res_o2).max() < TOL * scale
    assert cases.relerr(V.cpu().numpy(), V_o) < TOL
    A_o = O.csr_from_coo(V_o, fe.I, fe.J, fe.nn * 3)
    indptr, indices = plan.csr_pattern()
    assert np.array_equal(indptr.cpu().numpy(), A_o.indptr.astype(np.int64))          # bit-exact pattern
    assert np.array_equal(indices.cpu().numpy(), A_o.indices.astype(np.int32))
    assert cases.relerr(data.cpu().numpy(), A_o.data) < TOL
    # CSR data must equal scipy's canonicalisation of OUR V up to atomic summation order
    A_v = O.csr_from_coo(V.cpu().numpy(), fe.I, fe.J, fe.nn * 3)
    assert np.abs(data.cpu().numpy() - A_v.data).max() < 1e-13 * np.abs(A_v.data).max()
    # average stress
    sg_o = fe.compute_avg_stress(sol, params, dt)
    sg = plan.avg_stress(m, sol, params, dt)
    assert cases.relerr(sg.cpu().numpy(), sg_o) < TOL


def test_dp_steel_per_point_parameters():
    """DP-steel form: 10 internal_vars arrays (per-point a, h, t_sat, xm, r, C), 24 slip systems."""
    from cpfem_b200 import Plan
    N = 2
    pts, cells = O.box_mesh(N, N, N)
    nc = len(cells)
    params, ph, quat, ori = cases.dp_params(nc)
    fe = O.FEOracle(pts, cells, O.make_dp_batch_factory())
    plan = Plan(cells, pts, O.SLIP_BCC24)
    m = _mat(O.dp_ferrite())
    dt = 0.2
    rng = np.random.default_rng(5)
    for s in range(1, 8):
        eps = 5e-4 * s
        sol = np.stack([-0.3 * eps * pts[:, 0], -0.3 * eps * pts[:, 1], eps * pts[:, 2]], 1) + rng.uniform(-1, 1, pts.shape) * 2e-5
        if s == 7:
            break
        params = fe.update_int_vars_gp(sol, params, dt)
    res_o, V_o = fe.newton_update(sol, params, dt)
    res, data, V = plan.newton_update(m, sol, params, dt, want_V=True)
    assert cases.relerr(V.cpu().numpy(), V_o) < TOL
    new_o = fe.update_int_vars_gp(sol, params, dt)
    new = plan.update_state(m, sol, params, dt)
    assert cases.relerr(new[0].cpu().numpy(), new_o[0]) < TOL and cases.relerr(new[1].cpu().numpy(), new_o[1]) < TOL


def test_calibration_form_uniform_arrays_are_demoted():
    """9-array 'calibration' form of the state (calibration/case4.py:88-92: per-point gss_a, h, t_sat, xm, r).  Arrays that
    hold one value everywhere are demoted to scalars by the binding (uniform-parameter kernels, compile-time exponent);
    arrays that vary take the per-point kernels.  Both against the oracle; an in-place edit of an array must be noticed."""
    from cpfem_b200 import Plan
    fe, mat, dt, sol, params, quat, ori = cases.small_fe_case('304steel', N=3, steps=6)
    plan = Plan(fe.cells, fe.points, mat.slip)
    m = _mat(mat)
    nc = len(fe.cells)
    full = lambda v: torch.full((nc, 8), v, dtype=torch.float64, device='cuda')
    dev4 = [torch.as_tensor(np.ascontiguousarray(p), device='cuda') for p in params]
    extra = [full(mat.gss_a), full(mat.h), full(mat.t_sat), full(mat.xm), full(mat.r)]
    # material struct with WRONG scalars: the arrays must win, demoted or not
    wrong = _mat(O.copper())
    wrong.C11, wrong.C12, wrong.C44, wrong.max_sub_step = mat.C11, mat.C12, mat.C44, mat.max_sub_step
    new_o = fe.update_int_vars_gp(sol, params, dt)
    res_o, V_o = fe.newton_update(sol, params, dt)
    new_u = plan.update_state(wrong, sol, dev4 + extra, dt)
    assert len(plan._uniform_cache) == 5 and all(v[1] is not None for v in plan._uniform_cache.values())
    _, _, V_u = plan.newton_update(wrong, sol, dev4 + extra, dt, want_V=True)
    for k in range(2):
        assert cases.relerr(new_u[k].cpu().numpy(), new_o[k]) < TOL
    assert cases.relerr(V_u.cpu().numpy(), V_o) < TOL
    # one entry of h moved by an ulp-scale amount: no longer uniform -> per-point kernels, same answers to 1e-10
    extra[1][0, 0] *= (1.0 + 4e-16)
    new_p = plan.update_state(wrong, sol, dev4 + extra, dt)
    assert any(v[1] is None for v in plan._uniform_cache.values())
    for k in range(2):
        assert cases.relerr(new_p[k].cpu().numpy(), new_o[k]) < TOL
    # a NEW array object (the calibration loop builds `coeff * array` every evaluation; the allocator may hand it the
    # address of a freed one) is looked at afresh
    h2 = torch.full((nc, 8), 2.0 * mat.h, dtype=torch.float64, device='cuda')
    new_2h = plan.update_state(wrong, sol, dev4 + [extra[0], h2, extra[2], extra[3], extra[4]], dt)
    del h2
    h3 = torch.full((nc, 8), mat.h, dtype=torch.float64, device='cuda')          # likely the same address as h2
    new_1h = plan.update_state(wrong, sol, dev4 + [extra[0], h3, extra[2], extra[3], extra[4]], dt)
    assert not torch.equal(new_2h[1], new_1h[1]) and cases.relerr(new_1h[1].cpu().numpy(), new_o[1]) < TOL
    # a genuinely different hardening modulus on half of the cells changes g there and only there
    extra[1][: nc // 2] *= 2.0
    new_h = plan.update_state(wrong, sol, dev4 + extra, dt)
    d = (new_h[1] - new_p[1]).abs().amax(dim=(1, 2))
    assert float(d[: nc // 2].max()) > 0 and float(d[nc // 2:].max()) == 0.0


def test_soa_layout_and_transposes():
    """Native SoA state layout gives the same bits as the reference AoS layout; the transposes round-trip."""
    from cpfem_b200 import Plan, api
    fe, mat, dt, sol, params, quat, ori = cases.small_fe_case('copper', N=3, steps=4)
    plan = Plan(fe.cells, fe.points, mat.slip)
    m = _mat(mat)
    dev = [torch.as_tensor(p).cuda().contiguous() for p in params]
    comps = [9, 12, 12, 9]
    soa = [api.aos_to_soa(t, c) for t, c in zip(dev, comps)]
    for t, s, c in zip(dev, soa, comps):
        assert torch.equal(s, t.reshape(-1, c).T.contiguous())
        assert torch.equal(api.soa_to_aos(s, c).reshape(t.shape), t)
    a = plan.update_state(m, sol, dev, dt)
    b = plan.update_state(m, sol, soa, dt, layout=api.LAYOUT_SOA)
    for x, y, c in zip(a, b, [9, 12, 12]):
        assert torch.equal(api.soa_to_aos(y, c).reshape(x.shape), x)
    r_a = plan.residual(m, sol, dev, dt)
    r_b = plan.residual(m, sol, soa, dt, layout=api.LAYOUT_SOA)
    assert (r_a - r_b).abs().max().item() < 1e-13 * r_a.abs().max().item() + 1e-12


def test_host_streaming_state_update():
    """Plan.update_state_host (chunk-pipelined H2D / update / D2H of a host-resident state, cpfem_update_state_cells)
    gives the same bits as the device-resident call, for chunk sizes that do and do not divide the mesh."""
    from cpfem_b200 import Plan
    fe, mat, dt, sol, params, quat, ori = cases.small_fe_case('304steel', N=4, steps=6)
    plan = Plan(fe.cells, fe.points, mat.slip)
    m = _mat(mat)
    ref = [t.cpu() for t in plan.update_state(m, sol, params, dt)]
    host = [torch.as_tensor(np.ascontiguousarray(p)).pin_memory() for p in params]
    for cc in (64, 24, 7, 1000):
        st = plan.new_status()
        new = plan.update_state_host(m, torch.as_tensor(sol), host, dt, status=st, chunk_cells=cc)
        for a, b in zip(new, ref):
            assert not a.is_cuda and torch.equal(a, b)
        assert int(st[0]) == 0 and int(st[3]) > 0
    # rot_mats is kept on the device between calls (it never changes in a simulation); an edited tensor must be re-read
    host[3].copy_(host[3][torch.randperm(host[3].shape[0])])          # in-place edit: version counter moves
    ref2 = [t.cpu() for t in plan.update_state(m, sol, [host[0], host[1], host[2], host[3]], dt)]
    new2 = plan.update_state_host(m, torch.as_tensor(sol), host, dt, chunk_cells=24)
    assert all(torch.equal(a, b) for a, b in zip(new2, ref2)) and not torch.equal(new2[0], ref[0])
    new3 = plan.update_state_host(m, torch.as_tensor(sol), host, dt, chunk_cells=24, cache_rot=False)
    assert all(torch.equal(a, b) for a, b in zip(new3, ref2))


def test_csr_pattern_ragged_and_unstructured():
    """Pattern vs scipy on meshes that are not boxes: an L-shaped cell subset with a permuted node numbering."""
    from cpfem_b200 import Plan
    pts, cells = O.box_mesh(4, 3, 3)
    keep = np.array([c for c in range(len(cells)) if not (c % 4 >= 2 and (c // 4) % 3 >= 1)])
    cells = cells[keep]
    used = np.unique(cells)
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(used))
    remap = -np.ones(len(pts), dtype=np.int64)
    remap[used] = perm
    cells2 = remap[cells].astype(np.int32)
    pts2 = np.zeros((len(used), 3))
    pts2[perm] = pts[used]
    plan = Plan(cells2, pts2, O.SLIP_FCC12)
    I, J = O.coo_indices(cells2)
    A = scipy.sparse.csr_array((np.ones(len(I)), (I, J)), shape=(3 * len(used),) * 2)
    indptr, indices = plan.csr_pattern()
    assert np.array_equal(indptr.cpu().numpy(), A.indptr.astype(np.int64))
    assert np.array_equal(indices.cpu().numpy(), A.indices.astype(np.int32))
    assert plan.nnz == A.nnz


def test_dirichlet_rows():
    """apply_bc_vec + zeroRows on device (solver.py:119-133,290-293)."""
    from cpfem_b200 import Plan
    fe, mat, dt, sol, params, quat, ori = cases.small_fe_case('copper', N=2, steps=3)
    plan = Plan(fe.cells, fe.points, mat.slip)
    res, data, _ = plan.newton_update(_mat(mat), sol, params, dt)
    A0 = scipy.sparse.csr_array((data.cpu().numpy(), plan.csr_pattern()[1].cpu().numpy(), plan.csr_pattern()[0].cpu().numpy()))
    rows = np.array([0, 5, 13, 40], dtype=np.int64)
    vals = np.array([0., 0.1, -0.2, 0.3])
    sol_t = torch.as_tensor(sol).cuda().contiguous()
    plan.apply_dirichlet(torch.as_tensor(rows).cuda(), torch.as_tensor(vals).cuda(), sol_t, res, data)
    r = res.cpu().numpy().reshape(-1)
    assert np.allclose(r[rows], sol.reshape(-1)[rows] - vals, rtol=0, atol=0)
    A1 = scipy.sparse.csr_array((data.cpu().numpy(), A0.indices, A0.indptr)).toarray()
    A0 = A0.toarray()
    for i in range(A0.shape[0]):
        if i in rows:
            e = np.zeros(A0.shape[0]); e[i] = 1
            assert np.array_equal(A1[i], e)
        else:
            assert np.array_equal(A1[i], A0[i])


def test_problem_mirror_driver_loop():
    """The reference's load-step call sequence (singlecrystal_copper.py:179-233) through the Problem mirror:
    set_params -> newton_update / compute_residual -> compute_avg_stress -> update_int_vars_gp, vs the oracle."""
    from cpfem_b200.generate_mesh import Mesh, box_mesh
    from cpfem_b200.models_copper import CrystalPlasticity
    mm = box_mesh(2, 2, 2, 0.1, 0.1, 0.1)
    mesh = Mesh(mm.points, mm.cells_dict['hexahedron'])
    quat = np.array([[1., 0, 0, 0], [0.5, 0.5, 0.5, 0.5]])
    ori = np.arange(8) % 2
    bottom = lambda p: np.isclose(p[2], 0., atol=1e-5)
    top = lambda p: np.isclose(p[2], 0.1, atol=1e-5)
    bc = [[bottom, top], [2, 2], [lambda p: 0., lambda p: 1e-4]]
    problem = CrystalPlasticity(mesh, vec=3, dim=3, ele_type='HEX8', dirichlet_bc_info=bc, additional_info=(quat, ori))
    assert len(problem.fes[0].node_inds_list[0]) == 9 and problem.fes[0].vals_list[1][0] == 1e-4
    mat = O.copper()
    fe = O.FEOracle(mesh.points, mesh.cells, O.make_uniform_batch_factory(mat))
    params_o = O.initial_internal_vars(8, mat, O.get_rot_mat(quat)[ori])
    params = problem.internal_vars
    rng = np.random.default_rng(2)
    for step in range(1, 5):
        problem.dt = 0.01
        eps = 1e-3 * step
        sol = np.stack([-0.3 * eps * mesh.points[:, 0], -0.3 * eps * mesh.points[:, 1], eps * mesh.points[:, 2]], 1)
        sol += rng.uniform(-1, 1, sol.shape) * 1e-6
        problem.set_params(params)
        res = problem.newton_update([sol])[0]
        res_o, V_o = fe.newton_update(sol, params_o, 0.01)
        A_o = O.csr_from_coo(V_o, fe.I, fe.J, fe.nn * 3)
        assert cases.relerr(problem.csr_data.cpu().numpy(), A_o.data) < TOL
        assert np.array_equal(problem.I, fe.I) and np.array_equal(problem.J, fe.J)
        assert cases.relerr(problem.V.cpu().numpy(), V_o) < TOL
        sg = problem.compute_avg_stress(sol, params)
        assert cases.relerr(sg.cpu().numpy(), fe.compute_avg_stress(sol, params_o, 0.01)) < TOL
        P = problem.get_tensor_map()(fe.u_grads(sol).reshape(8, 8, 3, 3), *params)
        assert cases.relerr(P.cpu().numpy(), fe.point_stress(sol, params_o, 0.01)) < TOL
        params = problem.update_int_vars_gp(sol, params)
        params_o = fe.update_int_vars_gp(sol, params_o, 0.01)
        for k in range(2):
            assert cases.relerr(params[k].cpu().numpy(), params_o[k]) < TOL
    assert int(problem.last_status[2]) > 3


def test_fused_update_avg_stress():
    """cpfem_update_state_avg_stress (F3: one local solve for compute_avg_stress + update_int_vars_gp) vs the two separate
    entry points (bitwise state, sigma to rounding) and vs the oracle; the Problem mirror pairs the two driver calls
    through it and drops the pairing as soon as an argument changes."""
    import torch
    from cpfem_b200 import Plan
    from cpfem_b200.generate_mesh import Mesh
    from cpfem_b200.models_304steel import CrystalPlasticity
    fe, mat, dt, sol, params, quat, ori = cases.small_fe_case('304steel', N=3, steps=6)
    plan = Plan(fe.cells, fe.points, mat.slip)
    m = _mat(mat)
    st = plan.new_status()
    new_f, sg_f = plan.update_state_avg_stress(m, sol, params, dt, status=st)
    new_s = plan.update_state(m, sol, params, dt)
    sg_s = plan.avg_stress(m, sol, params, dt)
    for a, b in zip(new_f, new_s):
        assert torch.equal(a, b)
    assert cases.relerr(sg_f.cpu().numpy(), sg_s.cpu().numpy()) < 1e-14
    assert cases.relerr(sg_f.cpu().numpy(), fe.compute_avg_stress(sol, params, dt)) < TOL
    new_o = fe.update_int_vars_gp(sol, params, dt)
    for k in range(2):
        assert cases.relerr(new_f[k].cpu().numpy(), new_o[k]) < TOL
    assert int(st[0]) == 0 and int(st[1]) == 0
    # Problem mirror: compute_avg_stress then update_int_vars_gp on the same device tensors -> one kernel launch
    problem = CrystalPlasticity(Mesh(fe.points, fe.cells), vec=3, dim=3, ele_type='HEX8', dirichlet_bc_info=None,
                                additional_info=(quat, ori))
    problem.dt = dt
    dsol = torch.as_tensor(sol, device='cuda')
    dpar = [torch.as_tensor(np.ascontiguousarray(v), device='cuda') for v in params]
    sg = problem.compute_avg_stress(dsol, dpar)
    assert problem._fuse_cache is not None
    new = problem.update_int_vars_gp(dsol, dpar)
    assert problem._fuse_cache is None                      # handed over
    assert torch.equal(sg, sg_f) and all(torch.equal(a, b) for a, b in zip(new[:3], new_f))
    assert new[3] is dpar[3]                                # rot_mats passed through (models_copper.py:282)
    # a changed argument (in-place edit bumps the version counter) must not be served from the pairing
    sg = problem.compute_avg_stress(dsol, dpar)
    dsol.mul_(1.01)
    new2 = problem.update_int_vars_gp(dsol, dpar)
    ref2 = plan.update_state(m, dsol, dpar, dt)
    assert all(torch.equal(a, b) for a, b in zip(new2[:3], ref2)) and not torch.equal(new2[0], new_f[0])


def test_chunked_assembly_matches_single_chunk():
    """The 200^3 benchmark assembles in 16 chunks of 2^19 cells through one scratch buffer; the multi-chunk path (chunk
    boundaries inside the bulk-copy pipeline of the element kernel, partial last chunk, partial last quad of cells) is
    exercised here on a small mesh by lowering the chunk limit: V is bitwise the single-chunk V, the CSR data and the
    residual agree to atomic-summation order, and both match the oracle."""
    import torch
    from cpfem_b200 import Plan
    fe, mat, dt, sol, params, quat, ori = cases.small_fe_case('304steel', N=5, steps=6)     # 125 cells
    m = _mat(mat)
    ref = Plan(fe.cells, fe.points, mat.slip)
    assert ref.chunk_cells == 125
    res0, data0, V0 = ref.newton_update(m, sol, params, dt, want_V=True)
    try:
        for lim in (16, 48):
            os.environ['CPFEM_CHUNK_CELLS'] = str(lim)
            plan = Plan(fe.cells, fe.points, mat.slip)
            assert plan.chunk_cells == lim
            res, data, V = plan.newton_update(m, sol, params, dt, want_V=True)
            assert torch.equal(V, V0)
            assert float((data - data0).abs().max()) < 1e-13 * float(data0.abs().max())
            assert float((res - res0).abs().max()) < 1e-12 * float(V0.abs().max())
    finally:
        os.environ.pop('CPFEM_CHUNK_CELLS', None)
    res_o, V_o = fe.newton_update(sol, params, dt)
    assert cases.relerr(V0.cpu().numpy(), V_o) < TOL


def test_partitioned_assembly_sums_to_global():
    """Element-partitioned assembly on the device (what every rank of a multi-GPU run does): the slabs of a 2- and a
    3-way partition are assembled one after the other on this GPU - owned cells active, ghost cells pattern only -
    and their matrices / residuals, mapped to global dofs, add up to the single-plan assembly.  (The exchange itself is
    index plumbing over torch.distributed and is covered by the gloo tests in test_partition.py.)"""
    from cpfem_b200 import Plan
    from cpfem_b200.partition import slab_partition_structured, partition_cells
    name, N = '304steel', 4
    fac, deps, dt = cases.MATERIALS[name]
    mat = fac()
    m = _mat(mat)
    pts, cells = O.box_mesh(N, N, N)
    rng = np.random.default_rng(7)
    quat = cases.rand_quat(rng, 6)
    ori = rng.integers(0, 6, size=len(cells))
    fe = O.FEOracle(pts, cells, O.make_uniform_batch_factory(mat))
    params = O.initial_internal_vars(len(cells), mat, O.get_rot_mat(quat)[ori])
    disp = lambda e: np.stack([-0.3 * e * pts[:, 0], -0.3 * e * pts[:, 1], e * pts[:, 2]], 1) + rng.uniform(-1, 1, pts.shape) * 1e-6
    whole = Plan(cells, pts, mat.slip)
    for s_ in range(1, 8):                                  # into plastic flow, on the device
        new = whole.update_state(m, disp(deps * s_), params, dt)
        params = [new[0].cpu().numpy(), new[1].cpu().numpy(), new[2].cpu().numpy(), params[3]]
    sol = disp(deps * 8)
    res_g, data_g, _ = whole.newton_update(m, sol, params, dt)
    ip, ix = whole.csr_pattern()
    ndof = whole.ndof
    A_g = scipy.sparse.csr_array((data_g.cpu().numpy(), ix.cpu().numpy(), ip.cpu().numpy()), shape=(ndof, ndof))
    for world, structured in ((2, True), (3, False)):
        A_sum = scipy.sparse.csr_array((ndof, ndof))
        res_sum = np.zeros((len(pts), 3))
        owners = np.zeros(len(pts), int)
        for r in range(world):
            rm = slab_partition_structured(N, world, r) if structured else partition_cells(cells, pts, world, r)
            plan = Plan(rm.cells, rm.points, mat.slip)
            plan.set_active_cells(rm.n_owned_cells)
            own_c = rm.cell_gid[:rm.n_owned_cells]
            st = [v[own_c] for v in params]
            res, data, _ = plan.newton_update(m, sol[rm.node_gid], st, dt)
            lp, lx = plan.csr_pattern()
            lp, lx, d = lp.cpu().numpy(), lx.cpu().numpy().astype(np.int64), data.cpu().numpy()
            rows = np.repeat(np.arange(plan.ndof), np.diff(lp))
            gdof = lambda l: 3 * rm.node_gid[l // 3] + l % 3
            A_sum = A_sum + scipy.sparse.csr_array((d, (gdof(rows), gdof(lx))), shape=(ndof, ndof))
            res_sum[rm.node_gid] += res.cpu().numpy()
            owners[rm.node_gid[rm.owned_node_mask]] += 1
        assert (owners == 1).all()
        diff = (A_sum - A_g)
        assert np.abs(diff.data).max() < 1e-12 * np.abs(A_g.data).max()
        assert np.abs(res_sum - res_g.cpu().numpy()).max() < 1e-12 * np.abs(A_g.data).max()


def test_error_and_status_conventions():
    """Boundary behaviour (SURVEY 8(b) 'Error conventions'): hard errors are negative return codes with a message and no
    abort; soft conditions are counted in the 4-word status (points at the iteration cap, non-finite results, max and
    total local Newton iterations)."""
    import ctypes
    from cpfem_b200 import Plan, _lib, make_material
    from cpfem_b200._lib import CpfemError
    fe, mat, dt, sol, params, quat, ori = cases.small_fe_case('304steel', N=2, steps=6)
    plan = Plan(fe.cells, fe.points, mat.slip)
    m = _mat(mat)
    L = _lib.lib()
    # hard errors
    assert L.cpfem_newton_update(plan._h, None, None, None, 0.1, None, None, None, None, None) < 0
    assert b'null' in L.cpfem_last_error()
    with pytest.raises(CpfemError):
        Plan(fe.cells, fe.points, mat.slip[:5])                       # ns must be 12 or 24
    bad = fe.cells.copy()
    bad[0, 0] = len(fe.points) + 3
    with pytest.raises(CpfemError):
        Plan(bad, fe.points, mat.slip)                                # node index out of range
    with pytest.raises(CpfemError):
        plan.set_active_cells(plan.nc + 1)
    with pytest.raises(CpfemError):
        plan.update_state(make_material(mat.C11, mat.C12, mat.C44, mat.h, mat.t_sat, mat.gss_a, mat.xm, mat.r, mat.ao, mat.tol, 0),
                          sol, params, dt)                            # max_sub_step < 1
    # soft conditions: iteration cap
    st = plan.new_status()
    capped = make_material(mat.C11, mat.C12, mat.C44, mat.h, mat.t_sat, mat.gss_a, mat.xm, mat.r, mat.ao, mat.tol, mat.max_sub_step, 3)
    plan.update_state(capped, sol, params, dt, status=st)
    assert int(st[0]) > 0 and int(st[2]) == 3 and int(st[1]) == 0
    # ... and non-finite input: one poisoned node -> its 8 x 8 quadrature points report non-finite, nobody else
    st = plan.new_status()
    bad_sol = np.array(sol)
    bad_sol[13, 1] = np.nan                                            # centre node of the 2^3 mesh touches all 8 cells
    new = plan.update_state(m, bad_sol, params, dt, status=st)
    assert int(st[1]) == 64 and int(st[0]) == 0
    st = plan.new_status()
    ok = plan.update_state(m, sol, params, dt, status=st)
    assert int(st[0]) == 0 and int(st[1]) == 0 and int(st[2]) > 3 and int(st[3]) >= int(st[2])
    assert bool(torch.isfinite(ok[0]).all())


def test_single_cell_mesh_and_zero_active_cells():
    """Smallest inputs: a one-cell copper mesh (8 points: a quarter of one warp, one element of an 8-warp block) against
    the oracle, and a plan restricted to zero active cells (every kernel launch is skipped or empty: outputs are zero,
    no error, status untouched)."""
    from cpfem_b200 import Plan
    fe, mat, dt, sol, params, quat, ori = cases.small_fe_case('copper', N=1, steps=5)
    plan = Plan(fe.cells, fe.points, mat.slip)
    m = _mat(mat)
    assert plan.nc == 1 and plan.nnz == 24 * 24
    new_o = fe.update_int_vars_gp(sol, params, dt)
    new = plan.update_state(m, sol, params, dt)
    for k in range(2):
        assert cases.relerr(new[k].cpu().numpy(), new_o[k]) < TOL
    res_o, V_o = fe.newton_update(sol, params, dt)
    res, data, V = plan.newton_update(m, sol, params, dt, want_V=True)
    assert cases.relerr(V.cpu().numpy(), V_o) < TOL
    assert np.abs(res.cpu().numpy() - res_o).max() < TOL * np.abs(res_o).max()
    A_o = O.csr_from_coo(V_o, fe.I, fe.J, fe.nn * 3)
    indptr, indices = plan.csr_pattern()
    assert np.array_equal(indptr.cpu().numpy(), A_o.indptr.astype(np.int64))
    assert np.array_equal(indices.cpu().numpy(), A_o.indices.astype(np.int32))
    assert cases.relerr(data.cpu().numpy(), A_o.data) < TOL
    assert cases.relerr(plan.avg_stress(m, sol, params, dt).cpu().numpy(), fe.compute_avg_stress(sol, params, dt)) < TOL
    # zero active cells: the pattern stays, the values are all zero
    plan.set_active_cells(0)
    empty = [torch.empty((0,) + tuple(np.shape(p)[1:]), dtype=torch.float64, device='cuda') for p in params]
    st = plan.new_status()
    res0, data0, _ = plan.newton_update(m, sol, empty, dt, status=st)
    assert float(res0.abs().max()) == 0.0 and float(data0.abs().max()) == 0.0 and data0.numel() == 576
    out = plan.update_state(m, sol, empty, dt, status=st)
    assert out[0].shape[0] == 0 and int(st.abs().sum()) == 0


def test_full_size_properties():
    """Size-independent properties on a mesh the oracle cannot follow (64^3, 2.1 M points): the residual of a rigid
    translation of the converged field is unchanged, the tangent annihilates rigid translations, CSR row sums of the
    x-translation vanish, nnz = 9(3N+1)^3, run-to-run determinism of the non-atomic outputs."""
    from cpfem_b200 import Plan, synthetic
    N = 64
    mesh, quat, gid = synthetic.polycrystal(N)
    mat = O.steel304()
    plan = Plan(mesh.cells, mesh.points, mat.slip)
    assert plan.nnz == 9 * (3 * N + 1) ** 3
    m = _mat(mat)
    R = torch.as_tensor(O.get_rot_mat(quat)).cuda()[torch.as_tensor(gid).cuda()]
    nc = plan.nc
    params = [torch.eye(3, dtype=torch.float64, device='cuda').expand(nc, 8, 3, 3).contiguous(),
              torch.full((nc, 8, 12), mat.gss_initial, dtype=torch.float64, device='cuda'),
              torch.zeros(nc, 8, 12, dtype=torch.float64, device='cuda'),
              R[:, None].expand(nc, 8, 3, 3).contiguous()]
    dt = 2e-3
    for s in range(1, 8):
        sol = torch.as_tensor(synthetic.displacement(mesh.points, 2e-4 * s, N)).cuda()
        new = plan.update_state(m, sol, params, dt)
        params = [new[0], new[1], new[2], params[3]]
    sol = torch.as_tensor(synthetic.displacement(mesh.points, 2e-4 * 8, N)).cuda()
    st = plan.new_status()
    res, data, _ = plan.newton_update(m, sol, params, dt, status=st)
    st = st.cpu().numpy()
    assert st[0] == 0 and st[1] == 0 and st[2] >= 5
    shift = torch.tensor([0.3, -0.2, 0.1], dtype=torch.float64, device='cuda')
    res2 = plan.residual(m, sol + shift, params, dt)
    assert (res - res2).abs().max().item() < 1e-9 * res.abs().max().item()
    indptr, indices = plan.csr_pattern()
    A = torch.sparse_csr_tensor(indptr, indices.to(torch.int64), data, size=(plan.ndof, plan.ndof))
    t = torch.zeros(plan.nn, 3, dtype=torch.float64, device='cuda')
    t[:, 0] = 1.0
    y = A @ t.reshape(-1)
    assert y.abs().max().item() < 1e-9 * data.abs().max().item()
    a = plan.update_state(m, sol, params, dt)
    b = plan.update_state(m, sol, params, dt)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
