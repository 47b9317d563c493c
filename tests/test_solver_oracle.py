"""CPU suite for row F2 (device linear solver hand-off): the oracle's restatement of jax.scipy.sparse.linalg.bicgstab
and of the reference's jax_solve path (solver.py:19-48, 213-237) against a direct solve, on assembled FE matrices."""
import numpy as np
import scipy.sparse.linalg

import cases
import cpfem_oracle as O


def _clamped_case(N=2, name='copper'):
    fac, deps, dt = cases.MATERIALS[name]
    mat = fac()
    pts, cells = O.box_mesh(N, N, N)
    rng = np.random.default_rng(3)
    quat = cases.rand_quat(rng, 3)
    ori = rng.integers(0, 3, size=len(cells))
    fe = O.FEOracle(pts, cells, O.make_uniform_batch_factory(mat))
    params = O.initial_internal_vars(len(cells), mat, O.get_rot_mat(quat)[ori])
    bottom = np.where(np.abs(pts[:, 2]) < 1e-9)[0]
    top = np.where(np.abs(pts[:, 2] - 1) < 1e-9)[0]
    # singlecrystal_copper.py:155-157: bottom x, y, z = 0; top x, y = 0, z = disp
    nodes = np.concatenate([bottom, bottom, bottom, top, top, top])
    comps = np.concatenate([0 * bottom, 0 * bottom + 1, 0 * bottom + 2, 0 * top, 0 * top + 1, 0 * top + 2])
    return fe, mat, dt, deps, params, pts, nodes, comps, bottom, top, quat, ori


def test_bicgstab_ref_vs_direct():
    fe, mat, dt, deps, params, pts, nodes, comps, bottom, top, quat, ori = _clamped_case()
    sol = np.stack([0 * pts[:, 0], 0 * pts[:, 1], 4 * deps * pts[:, 2]], 1)
    res, V = fe.newton_update(sol, params, dt)
    A = O.csr_from_coo(V, fe.I, fe.J, fe.nn * 3).tolil()
    rows = nodes * 3 + comps
    for r in rows:
        A.rows[r] = [int(r)]
        A.data[r] = [1.0]
    A = A.tocsr()
    b = -res.reshape(-1)
    b[rows] = 0.0
    x_d = scipy.sparse.linalg.spsolve(A.tocsc(), b)
    x, k, err = O.jax_solve_ref(A, b, np.zeros_like(b), True)
    assert k > 0 and err < 1e-8 * np.linalg.norm(b) + 1e-9
    assert np.abs(x - x_d).max() < 1e-8 * np.abs(x_d).max()
    # unpreconditioned variant and a warm start that already solves the system (zero iterations)
    x2, k2 = O.bicgstab_ref(A, b, tol=1e-10, atol=1e-10, maxiter=10000)
    assert np.abs(x2 - x_d).max() < 1e-7 * np.abs(x_d).max()
    x3, k3 = O.bicgstab_ref(A, b, x0=x_d, tol=1e-8, atol=1e-8, maxiter=10)
    assert k3 == 0 and np.array_equal(x3, x_d)


def test_load_step_bicgstab_vs_direct():
    """Two load steps into plastic flow: the reference's solver path (Jacobi-BiCGStab with its x0) lands on the same
    displacement field as the direct solve of the oracle."""
    fe, mat, dt, deps, params, pts, nodes, comps, bottom, top, quat, ori = _clamped_case()
    sol_a = np.zeros((fe.nn, 3))
    sol_b = np.zeros((fe.nn, 3))
    for step in (4, 8):
        vals = np.concatenate([0. * bottom, 0. * bottom, 0. * bottom, 0. * top, 0. * top, 0. * top + deps * step])
        sol_a, it_a = O.solve_load_step(fe, sol_a, params, dt, nodes, comps, vals)
        sol_b, it_b, lin = O.solve_load_step_bicgstab(fe, sol_b, params, dt, nodes, comps, vals)
        assert it_a == it_b and all(k > 0 for k in lin)
        assert np.abs(sol_a - sol_b).max() < 1e-8 * np.abs(sol_a).max()
        params = fe.update_int_vars_gp(sol_a, params, dt)
