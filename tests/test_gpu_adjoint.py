"""GPU suite for the adjoint row: cpfem_point_jac_x / cpfem_point_vjp / cpfem_vjp_params /
cpfem_csr_transpose and the implicit_vjp mirror, called through the C ABI, against the oracle's autodiff and against
finite differences of the GPU forward solve.

Reference: f_jvp's jac_x, jac_y (singlecrystal_copper/models_copper.py:251-259; polycrystal_DPsteel/models_DPsteel_inhomo.py:245)
and implicit_vjp (crystal_plasticity_OR_design/solver.py:801-853).  Tolerance: 1e-10 of each block's largest entry for the
Jacobians (the local solution S itself is only converged to the reference's 1e-8 on both sides, but both sides stop at the
same iterate - the iteration counts are identical); 1e-9 for the products that go through a 9x9 solve."""
import os

import numpy as np
import pytest
import torch

import cases
import cpfem_oracle as O
from test_adjoint_oracle import assert_blocks
from test_solver_oracle import _clamped_case

pytestmark = pytest.mark.gpu


def _mat(m, tol=None):
    from cpfem_b200 import make_material
    return make_material(m.C11, m.C12, m.C44, m.h, m.t_sat, m.gss_a, m.xm, m.r, m.ao, m.tol if tol is None else tol, m.max_sub_step)


def _dummy_plan(slip):
    from cpfem_b200 import Plan
    pts, cells = O.box_mesh(1, 1, 1)
    return Plan(cells, pts, slip)


@pytest.mark.parametrize('name', list(cases.MATERIALS))
def test_point_jac_x_vs_oracle(name):
    """All columns of jac_x (51, or 75 for BCC24), jac_y and y = S from the kernel's own local solve."""
    plan = None
    for step, mat, dt, H, A, g, sl, R in cases.point_history(name, n=40, steps=6, seed=2):
        if step not in (1, 4, 6):
            continue
        plan = plan or _dummy_plan(mat.slip)
        pb = O.PointBatch(A, g, sl, R, mat)
        y = pb.newton_solver(H, dt)
        st = plan.new_status()
        jx, jy, S = plan.point_jac_x(_mat(mat), H, [A, g, sl, R], dt, status=st)
        assert int(st[0]) == 0 and int(st[1]) == 0
        ns = g.shape[1]
        assert jx.shape == (len(H), 9, 27 + 2 * ns)
        assert cases.relerr(S.cpu().numpy(), y.numpy()) < 1e-10
        Hh = torch.as_tensor(H)
        assert_blocks(jx.cpu().numpy(), pb.jac_x(Hh, y, dt).numpy(), ns, 0, what=f'{name} step {step}')
        jy_o = pb.jac_y(Hh, y, dt).numpy()
        assert np.abs(jy.cpu().numpy() - jy_o).max() < 1e-10 * np.abs(jy_o).max()


def _dp_points(n, seed=5):
    params, ph, quat, ori = cases.dp_params(n, seed=seed)
    A, g, sl, R, ga, h, ts, xm, r, C = [a[:, 0] for a in params]
    f = O.dp_ferrite()
    mk = lambda A_, g_, s_: O.PointBatch(A_, g_, s_, R, gss_a=ga, h=h, t_sat=ts, xm=xm, r=r, C=C, slip_table=O.SLIP_BCC24, ao=f.ao,
                                         max_sub_step=f.max_sub_step, tol=f.tol)
    rng = np.random.default_rng(seed + 1)
    H = np.zeros((n, 3, 3))
    H[:, 2, 2] = 4e-3
    H[:, 0, 0] = H[:, 1, 1] = -1.2e-3
    H += rng.uniform(-1, 1, size=H.shape) * 2e-4
    An, gn, sn = mk(A, g, sl).update_int_vars(H, 0.2)
    An, gn, sn = An.numpy(), gn.numpy(), sn.numpy()
    return f, mk(An, gn, sn), 1.5 * H, [An, gn, sn, R, ga, h, ts, xm, r, C]


def test_point_jac_x_dp_form_161_columns():
    """DP form of the state (10 arrays, per-point parameters and elastic tensor): nx = 161."""
    n = 24
    f, pb, H, params = _dp_points(n)
    plan = _dummy_plan(O.SLIP_BCC24)
    y = pb.newton_solver(H, 0.2)
    jx, jy, S = plan.point_jac_x(_mat(f), H, params, 0.2)
    assert jx.shape == (n, 9, 161)
    assert cases.relerr(S.cpu().numpy(), y.numpy()) < 1e-10
    assert_blocks(jx.cpu().numpy(), pb.jac_x(torch.as_tensor(H), y, 0.2, nextra=6).numpy(), 24, 6, what='dp jac_x')
    # reverse mode over the same 161 columns
    W = np.random.default_rng(3).normal(size=(n, 9))
    grad = plan.point_vjp(_mat(f), H, params, 0.2, W)
    want = np.einsum('pi,pic->pc', W, pb.dP_dx(torch.as_tensor(H), 0.2, y, nextra=6).numpy())
    assert_blocks(grad.cpu().numpy(), want, 24, 6, tol=1e-9, what='dp vjp')


@pytest.mark.parametrize('name', ['copper', '304steel', 'tantalum'])
def test_point_vjp_vs_oracle(name):
    """W : d tensor_map / dx through the local solve; the first nine columns are W : (the consistent tangent)."""
    rng = np.random.default_rng(11)
    plan = None
    for step, mat, dt, H, A, g, sl, R in cases.point_history(name, n=40, steps=5, seed=4):
        if step != 5:
            continue
        plan = plan or _dummy_plan(mat.slip)
        pb = O.PointBatch(A, g, sl, R, mat)
        y = pb.newton_solver(H, dt)
        W = rng.normal(size=(len(H), 9))
        grad = plan.point_vjp(_mat(mat), H, [A, g, sl, R], dt, W).cpu().numpy()
        D = pb.dP_dx(torch.as_tensor(H), dt, y).numpy()
        assert_blocks(grad, np.einsum('pi,pic->pc', W, D), g.shape[1], 0, tol=1e-9, what=f'{name} vjp')
        T = pb.tangent(H, dt, y).numpy().reshape(len(H), 9, 9)
        assert np.abs(grad[:, :9] - np.einsum('pi,pic->pc', W, T)).max() < 1e-9 * np.abs(T).max()


@pytest.mark.parametrize('name', ['304steel', 'copper'])
def test_vjp_params_vs_oracle(name):
    """vjp_linear_fn of implicit_vjp on a small distorted polycrystal mesh: nodal adjoint . d(residual)/d(internal_vars)."""
    from cpfem_b200 import Plan
    fe, mat, dt, sol, params, quat, ori = cases.small_fe_case(name, N=2, steps=5)
    plan = Plan(fe.cells, fe.points, mat.slip)
    lam = np.random.default_rng(5).normal(size=(fe.nn, 3))
    got = plan.vjp_params(_mat(mat), sol, params, dt, lam)
    want = fe.vjp_params(sol, params, dt, lam)
    for k, (a, b) in enumerate(zip(got, want)):
        a = a.cpu().numpy().reshape(b.shape)
        if np.abs(b).max() == 0.0:
            assert np.abs(a).max() == 0.0, k
        else:
            assert np.abs(a - b).max() < 1e-9 * np.abs(b).max(), (k, np.abs(a - b).max() / np.abs(b).max())


def test_vjp_params_dp_form():
    """The same with the 10-array DP form: gradients with respect to xm and C per point, zeros for gss_a, h, t_sat, r."""
    from cpfem_b200 import Plan
    pts, cells = O.box_mesh(2, 2, 2)
    rng = np.random.default_rng(8)
    pts = pts + rng.uniform(-1, 1, size=pts.shape) * 0.02
    params, ph, quat, ori = cases.dp_params(len(cells), seed=2)
    f = O.dp_ferrite()
    fe = O.FEOracle(pts, cells, O.make_dp_batch_factory(f.max_sub_step))
    u = lambda e: np.stack([-0.3 * e * pts[:, 0], -0.3 * e * pts[:, 1], e * pts[:, 2]], 1)
    params = fe.update_int_vars_gp(u(4e-3), params, 0.2)
    sol = u(6e-3) + rng.uniform(-1, 1, size=pts.shape) * 1e-5
    plan = Plan(cells, pts, O.SLIP_BCC24)
    lam = rng.normal(size=(fe.nn, 3))
    got = plan.vjp_params(_mat(f), sol, params, 0.2, lam)
    want = fe.vjp_params(sol, params, 0.2, lam)
    assert len(got) == 10
    for k, (a, b) in enumerate(zip(got, want)):
        a = a.cpu().numpy().reshape(b.shape)
        if np.abs(b).max() == 0.0:
            assert np.abs(a).max() == 0.0, k
        else:
            assert np.abs(a - b).max() < 1e-9 * np.abs(b).max(), (k, np.abs(a - b).max() / np.abs(b).max())


def test_csr_transpose_vs_scipy():
    import scipy.sparse
    from cpfem_b200 import Plan
    fe, mat, dt, sol, params, quat, ori = cases.small_fe_case('304steel', N=3, steps=5)
    plan = Plan(fe.cells, fe.points, mat.slip)
    res, data, _ = plan.newton_update(_mat(mat), sol, params, dt)
    rows = torch.arange(0, 12, device='cuda')                                 # a few Dirichlet rows make A unsymmetric
    plan.apply_dirichlet(rows, torch.zeros(12, dtype=torch.float64, device='cuda'), torch.as_tensor(sol, device='cuda').reshape(-1),
                         res=res.reshape(-1), csr_data=data)
    ip, ix = plan.csr_pattern()
    A = scipy.sparse.csr_array((data.cpu().numpy(), ix.cpu().numpy(), ip.cpu().numpy()), shape=(plan.ndof, plan.ndof))
    AT = scipy.sparse.csr_array(A.T)
    AT.sort_indices()
    dT = plan.csr_transpose(data).cpu().numpy()
    assert np.array_equal(AT.indices, ix.cpu().numpy()) and np.array_equal(dT, AT.data)      # a permutation: bit-exact


def test_implicit_vjp_vs_finite_differences():
    """implicit_vjp (solver.py:801-853) end to end on the GPU path: d(v . sol)/d(internal_vars) from one adjoint solve
    against central differences of the full forward solve (Newton + BiCGStab) along random directions in g, Fp_inv and
    rot_mats.  Local and global tolerances are tightened so that the forward map is smooth at the finite-difference step."""
    from cpfem_b200.generate_mesh import Mesh
    from cpfem_b200.models_copper import CrystalPlasticity
    from cpfem_b200.solver import implicit_vjp, solver

    class Cu(CrystalPlasticity):
        tol = 5e-10

    fe, mat, dt, deps, params_o, pts, nodes, comps, bottom, top, quat, ori = _clamped_case(2)
    zb = lambda p: np.isclose(p[2], 0., atol=1e-9)
    zt = lambda p: np.isclose(p[2], 1., atol=1e-9)
    d = deps * 6
    bc = [[zb, zb, zb, zt, zt, zt], [0, 1, 2, 0, 1, 2], [lambda p: 0.] * 5 + [lambda p: d]]
    problem = Cu(Mesh(pts, fe.cells), vec=3, dim=3, ele_type='HEX8', dirichlet_bc_info=bc, additional_info=(quat, ori))
    problem.dt = dt
    opts = lambda g: {'jax_solver': {}, 'initial_guess': [g], 'tol': 1e-10, 'rel_tol': 1e-12}
    zero = torch.zeros(fe.nn, 3, dtype=torch.float64, device='cuda')
    # a plastically deformed state to differentiate at: two load steps
    params = [torch.as_tensor(p, dtype=torch.float64, device='cuda') for p in problem.internal_vars]
    problem.set_params(params)
    sol = solver(problem, opts(zero))[0]
    params = problem.update_int_vars_gp(sol, params)
    bc[2][5] = lambda p: 1.5 * d
    problem.fes[0].update_Dirichlet_boundary_conditions(bc)

    def forward(p):
        problem.set_params(p)
        return solver(problem, opts(sol))[0]
    sol1 = forward(params)
    assert int(problem.last_status[2]) > 2                                    # plastic flow
    rng = np.random.default_rng(1)
    v = torch.as_tensor(rng.normal(size=(fe.nn, 3)), device='cuda')
    grads = implicit_vjp(problem, [sol1], params, [v], {'jax_solver': {}})
    assert len(grads) == 4 and all(g.shape == p.shape for g, p in zip(grads, params))
    assert float(grads[2].abs().max()) == 0.0                                  # accumulated slip does not enter the residual
    for k, eps in ((1, 1e-4), (0, 1e-7), (3, 1e-7)):                           # g (MPa), Fp_inv, rot_mats
        dirn = torch.as_tensor(rng.normal(size=tuple(params[k].shape)), device='cuda')
        up, dn = list(params), list(params)
        if k == 3:
            # The forward kernels work in the crystal frame and use R^T R = I, so they agree with the reference's literal
            # formulation ON the rotation group only (cp_adjoint.cuh header): differentiate along it, R(+-eps) = exp(+-eps W) R
            # with a random skew W per point; the direction is W R.
            W = dirn - dirn.transpose(-1, -2)
            up[k] = torch.linalg.matrix_exp(eps * W) @ params[k]
            dn[k] = torch.linalg.matrix_exp(-eps * W) @ params[k]
            dirn = W @ params[k]
        else:
            up[k] = params[k] + eps * dirn
            dn[k] = params[k] - eps * dirn
        fd = float(((forward(up) - forward(dn)) * v).sum()) / (2 * eps)
        an = float((grads[k] * dirn).sum())
        print(f'implicit_vjp block {k}: adjoint {an:.8e}  central difference {fd:.8e}')
        assert abs(fd - an) < 2e-5 * abs(an), (k, fd, an)
    # ad_wrapper (solver.py:856-874): the same gradients through torch's tape
    from cpfem_b200.solver import ad_wrapper
    fwd_pred = ad_wrapper(problem, opts(sol), {'jax_solver': {}})
    leaves = [p.clone().requires_grad_(True) for p in params]
    J = (fwd_pred(leaves)[0] * v).sum()
    auto = torch.autograd.grad(J, leaves)
    for k in (0, 1, 3):
        assert float((auto[k] - grads[k]).abs().max()) < 1e-9 * float(grads[k].abs().max()), k
    assert float(auto[2].abs().max()) == 0.0
    # the reference's own choice for this step (calibration_case4_...1D_GB.py:90: adjoint_solver_options={'umfpack_solver': {}}):
    # a direct solve of the TRANSPOSED system must give the same gradients as the device BiCGStab
    problem.set_params(params)
    grads_d = implicit_vjp(problem, [sol1], params, [v], {'umfpack_solver': {}})
    for k in (0, 1, 3):
        assert float((grads_d[k] - grads[k]).abs().max()) < 1e-6 * float(grads[k].abs().max()), k


def test_adjoint_kernels_vs_host_header_large():
    """5 x 10^4 points of the 304-steel set driven into plastic flow on the GPU, then f_jvp's Jacobians and the per-point VJP
    from the kernels (cpfem_point_jac_x / cpfem_point_vjp: the kernel's own local solve, S back to the lab frame, dual-number
    columns) against the host build of the same header evaluated at the kernel's S (tests/hostcheck, all host cores).  Same
    algebra, different arithmetic details: every block to 1e-10 of its largest entry.  A check of the kernels' indexing and
    per-thread plumbing at scale; the algebra itself is pinned on the oracle by the tests above."""
    import hostcheck_build
    from test_adjoint_oracle import host_jac, host_vjp
    lib = hostcheck_build.load()
    hostcheck_build.set_threads(lib, os.cpu_count() or 1)
    n = 50_000
    mat = O.steel304()
    plan, m = _dummy_plan(mat.slip), _mat(mat)
    rng = np.random.default_rng(21)
    R = O.get_rot_mat(cases.rand_quat(rng, n))
    A = torch.eye(3, dtype=torch.float64, device='cuda').repeat(n, 1, 1)
    g = torch.full((n, 12), mat.gss_initial, dtype=torch.float64, device='cuda')
    sl = torch.zeros(n, 12, dtype=torch.float64, device='cuda')
    Rd = torch.as_tensor(R, device='cuda')
    dt = 2e-3
    mkH = lambda s: (np.diag([-0.3, -0.3, 1.0])[None] * (2e-4 * s) + rng.uniform(-1, 1, size=(n, 3, 3)) * 2e-5)
    for s in range(1, 9):
        A, g, sl = plan.point_update_state(m, mkH(s), [A, g, sl, Rd], dt)
    H = mkH(9)
    st = plan.new_status()
    jx, jy, S = plan.point_jac_x(m, H, [A, g, sl, Rd], dt, status=st)
    W = rng.normal(size=(n, 9))
    grad = plan.point_vjp(m, H, [A, g, sl, Rd], dt, W)
    torch.cuda.synchronize()
    assert int(st[0]) == 0 and int(st[1]) == 0 and int(st[3]) > 6 * n          # plastic: > 6 local iterations per point
    Ah, gh, Sh = A.cpu().numpy(), g.cpu().numpy(), S.cpu().numpy()
    pp = np.tile([mat.C11, mat.C12, mat.C44, mat.xm], (n, 1))
    jx_h, jy_h, dPx, dPs = host_jac(lib, mat, dt, H, Ah, gh, R, Sh, pp, 0, 0)
    assert_blocks(jx.cpu().numpy(), jx_h, 12, 0, what='jac_x vs host header')
    assert np.abs(jy.cpu().numpy() - jy_h).max() < 1e-10 * np.abs(jy_h).max()
    grad_h = host_vjp(lib, mat, dt, H, Ah, gh, R, Sh, pp, W, 0)
    assert_blocks(grad.cpu().numpy(), grad_h, 12, 0, tol=1e-9, what='vjp vs host header')
