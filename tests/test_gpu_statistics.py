"""GPU-side parity at scale (B200, `-m gpu`): the CUDA kernels - not the host build of their header - against the
oracle on large seeded samples, with the same accounting as tests/test_oracle_golden.py::test_point_algebra_statistics:

  * every point whose local-Newton iteration AND residual-evaluation counts equal the oracle's must agree to 1e-10 of the
    field maximum in stress, tangent and new state (north_star tolerance);
  * points where the counts differ - `||r|| > tol` or a line-search comparison decided in the last bit - must be rare
    (<= 1e-3 of the sample) and within 1e-8; the disputed points are then taken to the mpmath arbiter (oracle/mp_arbiter.py),
    which solves the same residual at 50 digits: both fp64 answers have to sit within the reference's own stopping
    tolerance of the exact root.

Sample size: CPFEM_STAT_POINTS points per material x 8 load steps (default 12500 -> 1e5 point-evaluations per material,
about 25 s of oracle time each on the GPU box's host cores; profiles/r2/*_statistics*.txt keep the printed accounts).
"""
import os

import numpy as np
import pytest

import cases
import cpfem_oracle as O

N_POINTS = int(os.environ.get('CPFEM_STAT_POINTS', '12500'))
STEPS = 8


def _plan(slip):
    from cpfem_b200 import Plan
    pts, cells = O.box_mesh(1, 1, 1)
    return Plan(cells, pts, slip)                 # point-wise entry points only use the plan's slip table


def _material(mat):
    from cpfem_b200 import make_material
    return make_material(mat.C11, mat.C12, mat.C44, mat.h, mat.t_sat, mat.gss_a, mat.xm, mat.r, mat.ao, mat.tol, mat.max_sub_step)


class _Account:
    def __init__(self):
        self.worst_same = self.worst_diff = 0.0
        self.mism = self.tot = 0
        self.disputed = []                         # (label, inputs..., P_gpu, P_oracle)

    def add(self, pairs, info, it_o, ev_o):
        diff = (info[:, 0] != it_o) | (info[:, 1] != ev_o)
        self.mism += int(diff.sum())
        self.tot += len(diff)
        for a, b, *sc in pairs:
            e = np.abs(a - b).reshape(len(a), -1).max(1) / (sc[0] if sc else np.abs(b).max())
            if (~diff).any():
                self.worst_same = max(self.worst_same, float(e[~diff].max()))
            if diff.any():
                self.worst_diff = max(self.worst_diff, float(e[diff].max()))
        return np.where(diff)[0]

    def check(self, label):
        print(f'{label}: {self.tot} point-evaluations, {self.mism} with different iteration / evaluation counts, '
              f'worst rel. difference {self.worst_same:.2e} (same counts) / {self.worst_diff:.2e} (different counts)')
        assert self.worst_same < 1e-10, label
        assert self.worst_diff < 1e-8, label
        assert self.mism <= max(2, self.tot // 1000), label


def _arbitrate(label, rows, slip, consts, dt):
    """rows: (H, A, g, R, y_oracle, P_gpu, P_oracle, xm, C11, C12, C44) of disputed points.  Both answers must be within
    5e-10 of the exact first Piola-Kirchhoff stress (what a residual of 1e-8 MPa leaves on a stress of 1e2..1e3 MPa)."""
    import mp_arbiter as MP
    worst = 0.0
    for H, A, g, R, y, P_g, P_o, xm, C11, C12, C44 in rows[:8]:
        S_e, P_e, prob = MP.exact_point(H, A, g, R, slip, C11, C12, C44, xm, consts['ao'], dt, S_start=np.asarray(y).reshape(3, 3))
        scale = np.abs(P_e).max()
        e_g, e_o = np.abs(P_g - P_e).max() / scale, np.abs(P_o - P_e).max() / scale
        print(f'{label}: disputed point: |P_gpu - P_exact| = {e_g:.2e}, |P_oracle - P_exact| = {e_o:.2e} (relative), '
              f'oracle residual at its S (50 digits) = {float(prob.residual_norm(np.asarray(y).reshape(3, 3))):.2e}')
        worst = max(worst, e_g, e_o)
    assert worst < 5e-10, label


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['304steel', 'copper', 'tantalum', 'dp_ferrite'])
def test_gpu_point_statistics_uniform(name):
    """The four uniform parameter sets (FCC12 x2, BCC12 with the run-time pow path, BCC24) through cpfem_point_eval."""
    import torch
    plan, acc, rows = None, _Account(), []
    for step, mat, dt, H, A, g, sl, R in cases.point_history(name, n=N_POINTS, steps=STEPS, seed=2):
        if plan is None:
            plan, m = _plan(mat.slip), _material(mat)
        pb = O.PointBatch(A, g, sl, R, mat)
        y, it_o, ev_o = pb.newton_solver(H, dt, return_iters=True)
        P_o, T_o = pb.first_PK_stress(H, dt, y).numpy(), pb.tangent(H, dt, y).numpy()
        An_o, gn_o, sn_o = [v.numpy() for v in pb.update_int_vars(H, dt, y)]
        P, T, new, info = plan.point_eval(m, H, [A, g, sl, R], dt)
        torch.cuda.synchronize()
        P, T, info = P.cpu().numpy(), T.cpu().numpy(), info.cpu().numpy()
        An, gn, sn = [v.cpu().numpy() for v in new]
        assert (info[:, 2] == 0).all()
        bad = acc.add(((P, P_o), (T, T_o), (An, An_o), (gn, gn_o), (sn, sn_o, max(np.abs(sn_o).max(), mat.ao * dt))), info, it_o.numpy(), ev_o.numpy())
        for k in bad:
            rows.append((H[k], A[k], g[k], R[k], y[k].numpy(), P[k], P_o[k], mat.xm, mat.C11, mat.C12, mat.C44))
    acc.check(name)
    assert acc.tot == N_POINTS * STEPS
    if rows:
        _arbitrate(name, rows, mat.slip, {'ao': mat.ao}, dt)


@pytest.mark.gpu
def test_gpu_point_statistics_dp_per_point():
    """DP-steel form of the state (models_DPsteel_inhomo.py:229): BCC24, per-point gss_a, h, t_sat, xm, r and elastic
    tensor (two phases, 40 % martensite) through the kernels' per-point-parameter path."""
    import torch
    nc = max(8, N_POINTS // 8)
    n = nc * 8
    params, ph, quat, ori = cases.dp_params(nc, seed=5)
    Fp, g, sl, R, a, h, ts, xm, r, C = params
    A_, g_, sl_, R_ = Fp.reshape(n, 3, 3), g.reshape(n, 24), sl.reshape(n, 24), R.reshape(n, 3, 3)
    flat = [a.reshape(n), h.reshape(n), ts.reshape(n), xm.reshape(n), r.reshape(n), C.reshape(n, 3, 3, 3, 3)]
    f = O.dp_ferrite()
    plan, m = _plan(O.SLIP_BCC24), _material(f)
    rng = np.random.default_rng(7)
    acc, rows, dt = _Account(), [], 0.2
    for step in range(1, STEPS + 1):
        eps = 4e-4 * step
        H = np.zeros((n, 3, 3)); H[:, 2, 2] = eps; H[:, 0, 0] = H[:, 1, 1] = -0.3 * eps
        H += rng.uniform(-1, 1, size=H.shape) * 4e-5
        pb = O.PointBatch(A_, g_, sl_, R_, gss_a=flat[0], h=flat[1], t_sat=flat[2], xm=flat[3], r=flat[4], C=flat[5],
                          slip_table=O.SLIP_BCC24)
        y, it_o, ev_o = pb.newton_solver(H, dt, True)
        P_o, T_o = pb.first_PK_stress(H, dt, y).numpy(), pb.tangent(H, dt, y).numpy()
        An_o, gn_o, sn_o = [v.numpy() for v in pb.update_int_vars(H, dt, y)]
        P, T, new, info = plan.point_eval(m, H, [A_, g_, sl_, R_] + flat, dt)
        torch.cuda.synchronize()
        P, T, info = P.cpu().numpy(), T.cpu().numpy(), info.cpu().numpy()
        An, gn, sn = [v.cpu().numpy() for v in new]
        bad = acc.add(((P, P_o), (T, T_o), (An, An_o), (gn, gn_o), (sn, sn_o, max(np.abs(sn_o).max(), 0.001 * dt))), info, it_o.numpy(), ev_o.numpy())
        Cf = flat[5].reshape(n, 81)
        for k in bad:
            rows.append((H[k], A_[k], g_[k], R_[k], y[k].numpy(), P[k], P_o[k], flat[3][k], Cf[k, 0], Cf[k, 4], Cf[k, 50]))
        A_, g_, sl_ = An_o, gn_o, sn_o
    acc.check('dp steel (per-point parameters)')
    assert int(it_o.max()) > 3
    if rows:
        _arbitrate('dp steel', rows, O.SLIP_BCC24, {'ao': 0.001}, dt)


@pytest.mark.gpu
def test_gpu_vs_host_header_large(hostcheck):
    """2 x 10^5 points of the 304-steel set driven into plastic flow by the GPU path itself, then one evaluation compared
    with the host build of the same header (all host cores): the two share the algebra but not the arithmetic details
    (FMA contraction, the device's reciprocal sequence, warp-united active sets), so they must agree to 1e-11 wherever the
    iteration counts coincide - a cheap check that the kernels' warp-level machinery does not leak between points."""
    import torch
    import hostcheck_build
    n = 200_000
    mat = O.steel304()
    plan, m = _plan(mat.slip), _material(mat)
    rng = np.random.default_rng(11)
    R = O.get_rot_mat(cases.rand_quat(rng, n))
    dev = 'cuda'
    A = torch.eye(3, dtype=torch.float64, device=dev).repeat(n, 1, 1)
    g = torch.full((n, 12), mat.gss_initial, dtype=torch.float64, device=dev)
    sl = torch.zeros(n, 12, dtype=torch.float64, device=dev)
    Rd = torch.as_tensor(R, device=dev)
    dt = 2e-3
    mkH = lambda s: (np.diag([-0.3, -0.3, 1.0])[None] * (2e-4 * s) + rng.uniform(-1, 1, size=(n, 3, 3)) * 2e-5)
    for s in range(1, 10):
        A, g, sl = plan.point_update_state(m, mkH(s), [A, g, sl, Rd], dt)
    H = mkH(10)
    P, T, new, info = plan.point_eval(m, H, [A, g, sl, Rd], dt)
    torch.cuda.synchronize()
    hostcheck_build.set_threads(hostcheck, os.cpu_count() or 1)
    P_h, T_h, An_h, gn_h, sn_h, info_h = hostcheck_build.evaluate(hostcheck, mat, dt, H, A.cpu().numpy(), g.cpu().numpy(),
                                                                 sl.cpu().numpy(), R, pown=119)
    info = info.cpu().numpy()
    same = (info[:, 0] == info_h[:, 0]) & (info[:, 1] == info_h[:, 1])
    assert info[:, 0].mean() > 6 and (info[:, 2] == 0).all()
    assert (~same).sum() <= n // 1000, int((~same).sum())
    for a, b in ((P.cpu().numpy(), P_h), (T.cpu().numpy(), T_h), (new[0].cpu().numpy(), An_h), (new[1].cpu().numpy(), gn_h)):
        e = np.abs(a - b).reshape(n, -1).max(1) / np.abs(b).max()
        assert e[same].max() < 1e-11, float(e[same].max())
        assert e.max() < 1e-8


@pytest.mark.gpu
def test_dp_steel_10cubed_newton_update_vs_oracle():
    """BASELINE config 4 at its own size: the committed DP-steel mesh (10^3 cells, BCC24, two phases with per-point
    parameters and elastic tensors, polycrystal_DPsteel_inhomo.py:78-229) advanced into plastic flow; newton_update's
    residual, V (reference layout, 576 per cell) AND the assembled CSR data against the oracle's restatement of
    jax_fem's kernel_jac + scipy's COO->CSR (solver.py:281), plus the state update and the average stress."""
    import torch
    from cpfem_b200 import Plan
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    gd = np.load(os.path.join(gold, 'dpsteel_vtu.npz'))
    quat = np.loadtxt(os.path.join(gold, 'quat_dp.txt'))[:20, 1:]
    pts, cells = gd['points'], gd['cells']
    nc = len(cells)
    ph = gd['phase_inds'].astype(int)
    f, mm = O.dp_ferrite(), O.dp_martensite()
    pick = lambda a, b: np.array([a, b])[ph]
    rep = lambda v: np.repeat(v[:, None], 8, axis=1)
    ori = np.clip(gd['cell_ori_inds'].astype(int), 0, len(quat) - 1)
    R = np.repeat(O.get_rot_mat(quat)[ori][:, None], 8, axis=1)
    C = np.stack([O.cubic_C(a, b, c) for a, b, c in zip(gd['C11'], gd['C12'], gd['C44'])])
    params = [np.tile(np.eye(3)[None, None], (nc, 8, 1, 1)), np.repeat(rep(pick(f.gss_initial, mm.gss_initial))[:, :, None], 24, axis=2),
              np.zeros((nc, 8, 24)), R, rep(pick(f.gss_a, mm.gss_a)), rep(pick(f.h, mm.h)), rep(pick(f.t_sat, mm.t_sat)),
              rep(pick(f.xm, mm.xm)), rep(pick(f.r, mm.r)), np.repeat(C[:, None], 8, axis=1)]
    fe = O.FEOracle(pts, cells, O.make_dp_batch_factory())
    plan, m = Plan(cells, pts, O.SLIP_BCC24), _material(f)
    rng = np.random.default_rng(4)
    L = pts.max(axis=0)
    dt = 0.2

    def disp(eps):
        u = np.stack([-0.3 * eps * pts[:, 0], -0.3 * eps * pts[:, 1], eps * pts[:, 2]], axis=1)
        return u + rng.uniform(-1, 1, size=u.shape) * 2e-5 * L[0] / 10
    # drive the state with the GPU update (its parity is the subject of the other tests), then hand it to both sides
    dparams = [torch.as_tensor(p, device='cuda') for p in params]
    for s in range(1, 8):
        new = plan.update_state(m, disp(4e-4 * s), dparams, dt)
        dparams = [new[0], new[1], new[2]] + dparams[3:]
    params = [p.cpu().numpy() for p in dparams]
    sol = disp(4e-4 * 8)
    st = plan.new_status()
    res, data, V = plan.newton_update(m, sol, dparams, dt, want_V=True, status=st)
    new = plan.update_state(m, sol, dparams, dt)
    sig = plan.avg_stress(m, sol, dparams, dt)
    torch.cuda.synchronize()
    assert int(st[0]) == 0 and int(st[1]) == 0 and int(st[2]) > 3          # no caps / NaNs, plastic flow reached
    res_o, V_o = fe.newton_update(sol, params, dt)
    A_o = O.csr_from_coo(V_o, fe.I, fe.J, fe.nn * 3)
    ip, ix = plan.csr_pattern()
    assert np.array_equal(ip.cpu().numpy(), A_o.indptr) and np.array_equal(ix.cpu().numpy(), A_o.indices)
    assert cases.relerr(V.cpu().numpy(), V_o) < 1e-10
    assert cases.relerr(data.cpu().numpy(), A_o.data) < 1e-10
    scale = np.abs(fe.cell_residual(sol, params, dt)).max()          # nodal sums cancel in the interior
    assert np.abs(res.cpu().numpy() - res_o).max() < 1e-10 * scale
    new_o = fe.update_int_vars_gp(sol, params, dt)
    assert cases.relerr(new[0].cpu().numpy(), new_o[0]) < 1e-10 and cases.relerr(new[1].cpu().numpy(), new_o[1]) < 1e-10
    assert np.abs(new[2].cpu().numpy() - new_o[2]).max() < 1e-10 * max(np.abs(new_o[2]).max(), 0.001 * dt)
    assert cases.relerr(sig.cpu().numpy(), fe.compute_avg_stress(sol, params, dt)) < 1e-10
