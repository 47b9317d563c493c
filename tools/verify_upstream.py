#!/usr/bin/env python
"""Checks, on a box that HAS jax + jax_fem (+ basix) and a B200, the [UPSTREAM] assumptions this repo makes about
jax_fem - the facts no committed artefact of the reference can pin:

  1. order of the 8 Gauss points inside the (nc, 8, ...) arrays and of the 8 nodes inside `shape_grads (nc,8,8,3)`
     (consumed at singlecrystal_copper/models_copper.py:277);
  2. `JxW (nc,8)` (models_copper.py:315);
  3. the COO index rule `I, J` and the flattening of `problem.V` = cells_jac.reshape(-1) with V[c, 3a+i, 3b+k]
     (consumed at crystal_plasticity_OR_design/solver.py:281);
  4. the values of V and of the residual for the reference's own copper model on its own 2x2x2 mesh.

usage:  python tools/verify_upstream.py /path/to/JAX-CPFEM [--n 2]
Exit code 0 = every assumption holds; the report says which one fails otherwise.  (Cannot run in the build container of
this repo: neither JAX nor jax_fem is installable there.)
"""
import argparse
import os
import sys

import numpy as onp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'jax-cpfem_b200'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('reference', help='checkout of SuperkakaSCU/JAX-CPFEM')
    ap.add_argument('--n', type=int, default=2)
    args = ap.parse_args()
    import jax
    jax.config.update('jax_enable_x64', True)
    import jax.numpy as np
    from jax_fem.generate_mesh import Mesh, box_mesh, get_meshio_cell_type
    sys.path.insert(0, os.path.join(args.reference, 'singlecrystal_copper'))
    os.chdir(os.path.join(args.reference, 'singlecrystal_copper'))     # models_copper.py reads data/csv/input_slip_sys.txt relatively
    from models_copper import CrystalPlasticity as Reference

    import torch
    from cpfem_b200 import Plan, make_material
    from cpfem_b200.param_sets import PRESETS
    from cpfem_b200.problem import FiniteElement as MirrorFE
    from cpfem_b200.generate_mesh import Mesh as MirrorMesh

    N = args.n
    mm = box_mesh(N, N, N, 0.1, 0.1, 0.1)
    mesh = Mesh(mm.points, mm.cells_dict[get_meshio_cell_type('HEX8')])
    rng = onp.random.default_rng(0)
    q = rng.normal(size=(3, 4)); q /= onp.linalg.norm(q, axis=1)[:, None]
    ori = rng.integers(0, 3, size=len(mesh.cells))
    problem = Reference(mesh, vec=3, dim=3, ele_type='HEX8', dirichlet_bc_info=None, additional_info=(q, ori))
    problem.dt = 0.05
    fe = problem.fes[0]
    ok = True

    def report(name, err, tol):
        nonlocal ok
        good = err <= tol
        ok &= good
        print('%-62s max |diff| = %.3e  (tol %.1e)  %s' % (name, err, tol, 'OK' if good else 'FAIL'))

    mirror = MirrorFE(MirrorMesh(onp.asarray(mesh.points), onp.asarray(mesh.cells)), 3, 3, 'HEX8', None)
    report('1. shape_grads (nc,8,8,3): quad-point and node order', float(onp.abs(onp.asarray(fe.shape_grads) - mirror.shape_grads).max()), 1e-12)
    report('2. JxW (nc,8)', float(onp.abs(onp.asarray(fe.JxW) - mirror.JxW).max()), 1e-15)
    cells = onp.asarray(mesh.cells).astype(onp.int64)
    inds = (3 * cells[:, :, None] + onp.arange(3)[None, None, :]).reshape(len(cells), -1)
    I = onp.repeat(inds[:, :, None], 24, axis=2).reshape(-1)
    J = onp.repeat(inds[:, None, :], 24, axis=1).reshape(-1)
    report('3a. I (rows of the COO triplets)', float(onp.abs(onp.asarray(problem.I) - I).max()), 0)
    report('3b. J (columns of the COO triplets)', float(onp.abs(onp.asarray(problem.J) - J).max()), 0)

    # a plastic state: a few load steps with the reference's own update, then one newton_update on both sides
    pts = onp.asarray(mesh.points)
    params = problem.internal_vars
    for s in range(1, 6):
        eps = 1e-3 * s
        sol = np.asarray(onp.stack([-0.3 * eps * pts[:, 0], -0.3 * eps * pts[:, 1], eps * pts[:, 2]], 1))
        problem.set_params(params)
        params = problem.update_int_vars_gp(sol, params)
    eps = 6e-3
    sol = onp.stack([-0.3 * eps * pts[:, 0], -0.3 * eps * pts[:, 1], eps * pts[:, 2]], 1)
    sol += rng.uniform(-1, 1, size=sol.shape) * 1e-6
    problem.set_params(params)
    res_ref = onp.asarray(problem.newton_update([np.asarray(sol)])[0])
    V_ref = onp.asarray(problem.V)

    ps = PRESETS['copper']
    plan = Plan(onp.asarray(mesh.cells), pts, ps['slip'])
    m = ps['material']
    mat = make_material(m['C11'], m['C12'], m['C44'], m['h'], m['t_sat'], m['gss_a'], m['xm'], m['r'], m['ao'], m['tol'], m['max_sub_step'])
    dparams = [torch.as_tensor(onp.asarray(p), device='cuda') for p in params]
    res, data, V = plan.newton_update(mat, sol, dparams, problem.dt, want_V=True)
    V = V.cpu().numpy()
    report('4a. problem.V[:576] (first cell, layout V[c, 3a+i, 3b+k])', float(onp.abs(V[:576] - V_ref[:576]).max() / onp.abs(V_ref).max()), 1e-10)
    report('4b. problem.V (all cells)', float(onp.abs(V - V_ref).max() / onp.abs(V_ref).max()), 1e-10)
    report('4c. residual', float(onp.abs(res.cpu().numpy() - res_ref).max() / max(onp.abs(res_ref).max(), 1e-300)), 1e-8)
    new_ref = problem.update_int_vars_gp(np.asarray(sol), params)
    new = plan.update_state(mat, sol, dparams, problem.dt)
    for k, name in enumerate(('Fp_inv', 'slip resistance', 'slip')):
        a, b = new[k].cpu().numpy(), onp.asarray(new_ref[k])
        report('4d. update_int_vars_gp: ' + name, float(onp.abs(a - b).max() / max(onp.abs(b).max(), 1e-12)), 1e-10)
    import scipy.sparse
    A_ref = scipy.sparse.csr_array((V_ref, (onp.asarray(problem.I), onp.asarray(problem.J))), shape=(3 * len(pts),) * 2)
    ip, ix = plan.csr_pattern()
    report('5. CSR pattern vs scipy on the reference triplets', float(onp.abs(ip.cpu().numpy() - A_ref.indptr).max() + onp.abs(ix.cpu().numpy() - A_ref.indices).max()), 0)
    report('6. CSR data', float(onp.abs(data.cpu().numpy() - A_ref.data).max() / onp.abs(A_ref.data).max()), 1e-10)
    print('ALL UPSTREAM ASSUMPTIONS HOLD' if ok else 'SOME ASSUMPTIONS FAILED - see SURVEY.md Appendix D')
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
