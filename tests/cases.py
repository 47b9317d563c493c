"""Seeded test cases shared by the CPU and GPU suites: random orientations, a short uniaxial load history that
drives the points from elastic into plastic flow, for the four parameter sets of the reference."""
import numpy as np

import cpfem_oracle as O

MATERIALS = {
    # name: (oracle Material factory, d_eps per step, dt per step)
    '304steel': (O.steel304, 2e-4, 2e-3),
    'copper': (O.copper, 1e-3, 1e-2),
    'tantalum': (O.tantalum, -2.5e-4, 0.25),
    'dp_ferrite': (O.dp_ferrite, 2e-4, 0.2),
}


# compile-time rate exponent n - 1 the kernels instantiate for each set (0 = run-time path only)
RATE_POWN = {'304steel': 119, 'copper': 9, 'tantalum': 0, 'dp_ferrite': 19}


def rand_quat(rng, n):
    q = rng.normal(size=(n, 4))
    return q / np.linalg.norm(q, axis=1)[:, None]


def point_history(name, n=48, steps=8, seed=0):
    """Yields (step, mat, dt, H, A, g, slip, R) along an oracle-advanced load history of n independent points."""
    fac, deps, dt = MATERIALS[name]
    mat = fac()
    rng = np.random.default_rng(seed)
    ns = len(mat.slip)
    R = O.get_rot_mat(rand_quat(rng, n))
    A = np.tile(np.eye(3), (n, 1, 1))
    g = mat.gss_initial * np.ones((n, ns))
    sl = np.zeros((n, ns))
    for step in range(1, steps + 1):
        eps = deps * step
        H = np.zeros((n, 3, 3))
        H[:, 2, 2] = eps
        H[:, 0, 0] = H[:, 1, 1] = -0.3 * eps
        H += rng.uniform(-1, 1, size=(n, 3, 3)) * abs(deps) * 0.1
        yield step, mat, dt, H, A, g, sl, R
        pb = O.PointBatch(A, g, sl, R, mat)
        An, gn, sn = pb.update_int_vars(H, dt)
        A, g, sl = An.numpy(), gn.numpy(), sn.numpy()


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def dp_params(nc, seed=0, n_quat=8):
    """DP-steel style internal_vars (10 arrays) with a seeded 40 % martensite draw and random orientations."""
    rng = np.random.default_rng(seed)
    ph = np.zeros(nc, dtype=int)
    ph[:int(nc * 0.4)] = 1
    rng.shuffle(ph)
    f, m = O.dp_ferrite(), O.dp_martensite()
    pick = lambda a, b: np.array([a, b])[ph]
    rep = lambda v: np.repeat(v[:, None], 8, axis=1)
    quat = rand_quat(rng, n_quat)
    ori = rng.integers(0, n_quat, size=nc)
    R = np.repeat(O.get_rot_mat(quat)[ori][:, None], 8, axis=1)
    Fp = np.tile(np.eye(3)[None, None], (nc, 8, 1, 1))
    g = np.repeat(rep(pick(f.gss_initial, m.gss_initial))[:, :, None], 24, axis=2)
    sl = np.zeros_like(g)
    C = np.array([O.cubic_C(f.C11, f.C12, f.C44), O.cubic_C(m.C11, m.C12, m.C44)])[ph]
    C = np.repeat(C[:, None], 8, axis=1)
    params = [Fp, g, sl, R, rep(pick(f.gss_a, m.gss_a)), rep(pick(f.h, m.h)), rep(pick(f.t_sat, m.t_sat)),
              rep(pick(f.xm, m.xm)), rep(pick(f.r, m.r)), C]
    return params, ph, quat, ori


def small_fe_case(name, N=3, seed=0, steps=6):
    """A small polycrystal FE case advanced `steps` load steps with the oracle: returns the FEOracle, the material,
    dt, the displacement field of the next step and the current params (reference layout)."""
    fac, deps, dt = MATERIALS[name]
    mat = fac()
    rng = np.random.default_rng(seed)
    pts, cells = O.box_mesh(N, N, N)
    # distort the mesh slightly so that shape gradients differ between cells
    pts = pts + rng.uniform(-1, 1, size=pts.shape) * 0.05 / N
    nc = len(cells)
    quat = rand_quat(rng, 5)
    ori = rng.integers(0, 5, size=nc)
    fe = O.FEOracle(pts, cells, O.make_uniform_batch_factory(mat))
    params = O.initial_internal_vars(nc, mat, O.get_rot_mat(quat)[ori])

    def disp(eps):
        u = np.stack([-0.3 * eps * pts[:, 0], -0.3 * eps * pts[:, 1], eps * pts[:, 2]], axis=1)
        return u + rng.uniform(-1, 1, size=u.shape) * abs(deps) * 0.02 / N
    for s in range(1, steps + 1):
        params = fe.update_int_vars_gp(disp(deps * s), params, dt)
    sol = disp(deps * (steps + 1))
    return fe, mat, dt, sol, params, quat, ori
