#!/usr/bin/env python
"""The reference's forward drivers on the B200 path.

Mirrors the load-step loop shared by singlecrystal_copper.py / singlecrystal_tantalum.py / polycrystal_304steel.py /
polycrystal_DPsteel_inhomo.py of JAX-CPFEM (e.g. singlecrystal_copper/singlecrystal_copper.py:162-233) with the imports
swapped:

    from jax_fem.solver import solver                 ->  from cpfem_b200.solver import solver
    from jax_fem.generate_mesh import Mesh, box_mesh  ->  from cpfem_b200.generate_mesh import Mesh, box_mesh
    from jax_fem.utils import save_sol                ->  from cpfem_b200.utils import save_sol
    from applications.<case>.models_<case> import CrystalPlasticity  ->  from cpfem_b200.models_<case> import CrystalPlasticity

    python examples/run_driver.py --case copper --n 16 --steps 10 [--mesh path/to/mesh16.msh --quat path/to/quat.txt] [--vtk out_dir]

Without --mesh a structured box of n^3 cells with the case's domain size is generated (the Neper meshes the reference
ships are exactly such boxes); without --quat random orientations are drawn (seeded)."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'jax-cpfem_b200'))

CASES = {
    # case: (model module, domain size, total displacement factor, total time, number of steps, number of orientations)
    'copper': ('models_copper', 1.0, 0.05, 0.5, 50, 1),            # singlecrystal_copper.py:63-95
    'tantalum': ('models_tantalum', 1.0, -0.0125, 12.5, 50, 1),    # singlecrystal_tantalum.py:68-101
    '304steel': ('models_304steel', 0.016, 0.01, 0.1, 50, 8),      # polycrystal_304steel.py:75-123
    'dpsteel': ('models_DPsteel_inhomo', 2.0, 0.01, 10.0, 50, 20),  # polycrystal_DPsteel_inhomo.py:68-111
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--case', default='copper', choices=list(CASES))
    ap.add_argument('--n', type=int, default=16)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--mesh', default=None)
    ap.add_argument('--quat', default=None)
    ap.add_argument('--vtk', default=None)
    args = ap.parse_args()

    import importlib
    import torch
    from cpfem_b200.generate_mesh import Mesh, box_mesh, read_gmsh22_hex
    from cpfem_b200.solver import solver
    from cpfem_b200.utils import save_sol
    mod, L, dfac, ttot, nsteps, noris = CASES[args.case]
    CrystalPlasticity = importlib.import_module('cpfem_b200.' + mod).CrystalPlasticity

    rng = np.random.default_rng(0)
    if args.mesh:
        mm = read_gmsh22_hex(args.mesh)
        cell_grain = mm.cell_data['gmsh:physical'][0] - 1
    else:
        mm = box_mesh(args.n, args.n, args.n, L, L, L)
        g = max(1, args.n // 2)                                   # 2 x 2 x 2 block "grains"
        k, j, i = np.meshgrid(*(np.arange(args.n),) * 3, indexing='ij')
        cell_grain = ((i // g) + 2 * (j // g) + 4 * (k // g)).ravel()
    mesh = Mesh(mm.points, mm.cells_dict['hexahedron'])
    if args.quat:
        quat = np.loadtxt(args.quat)[:noris, 1:]
    else:
        quat = rng.normal(size=(noris, 4))
        quat /= np.linalg.norm(quat, axis=1)[:, None]
    if args.case == 'dpsteel':
        cell_ori_inds = rng.integers(0, noris, size=len(mesh.cells))       # polycrystal_DPsteel_inhomo.py:83
    else:
        cell_ori_inds = np.arange(noris)[cell_grain % noris]

    Lx, Ly, Lz = mesh.points.max(0)
    disps = np.linspace(0., dfac * Lx, nsteps + 1)
    ts = np.linspace(0., ttot, nsteps + 1)
    close = lambda a, b: np.isclose(a, b, atol=1e-5)
    bottom = lambda p: close(p[2], 0.)
    top = lambda p: close(p[2], Lz)
    left = lambda p: close(p[0], 0.)
    front = lambda p: close(p[1], 0.)
    corner = lambda p: close(p[0], 0.) & close(p[1], 0.) & close(p[2], Lz if args.case == '304steel' else 0.)
    zero = lambda p: 0.
    val = lambda d: (lambda p: d)
    if args.case == 'copper':        # singlecrystal_copper.py:155-157: clamped bottom, top x = y = 0, z = disp
        mk = lambda d: [[bottom, bottom, bottom, top, top, top], [0, 1, 2, 0, 1, 2], [zero] * 5 + [val(d)]]
    elif args.case == 'dpsteel':     # polycrystal_DPsteel_inhomo.py:180-182
        mk = lambda d: [[left, front, bottom, top], [0, 1, 2, 2], [zero, zero, zero, val(d)]]
    else:                            # singlecrystal_tantalum.py:163-165 / polycrystal_304steel.py:184-186
        mk = lambda d: [[corner, corner, bottom, top], [0, 1, 2, 2], [zero, zero, zero, val(d)]]
    options = {'jax_solver': {}}
    if args.case == 'dpsteel':
        options['line_search_flag'] = True                                   # polycrystal_DPsteel_inhomo.py:227

    problem = CrystalPlasticity(mesh, vec=3, dim=3, ele_type='HEX8', dirichlet_bc_info=mk(disps[0]),
                                additional_info=(quat, cell_ori_inds))
    sol_list = [torch.zeros(problem.fes[0].num_total_nodes, 3, dtype=torch.float64, device=problem.device)]
    params = problem.internal_vars
    if args.vtk:
        os.makedirs(args.vtk, exist_ok=True)
    print(f'{args.case}: {len(mesh.cells)} cells, {8 * len(mesh.cells)} quadrature points, {problem.num_total_dofs_all_vars} dofs')
    for i in range(min(args.steps, nsteps)):
        t0 = time.time()
        problem.dt = ts[i + 1] - ts[i]
        problem.fes[0].update_Dirichlet_boundary_conditions(mk(disps[i + 1]))
        problem.set_params(params)
        sol_list = solver(problem, dict(options, initial_guess=sol_list))
        sigma = problem.compute_avg_stress(sol_list[0], params)
        params = problem.update_int_vars_gp(sol_list[0], params)
        torch.cuda.synchronize()
        szz = sigma[:, 2, 2]
        print(f'step {i + 1:3d}  disp {disps[i + 1]: .4e}  mean sigma_zz {float(szz.mean()): .6f}  Newton its '
              f'{problem.last_newton_iterations}  BiCGStab its (last) {getattr(problem, "last_linear_iterations", 0)}  '
              f'local Newton max {int(problem.last_status[2])}  {time.time() - t0:.2f} s')
        if args.vtk:
            save_sol(problem.fes[0], sol_list[0], os.path.join(args.vtk, f'u_{i:03d}.vtu'),
                     cell_infos=[('cell_ori_inds', cell_ori_inds), ('sigma_xx', sigma[:, 0, 0]), ('sigma_yy', sigma[:, 1, 1]),
                                 ('sigma_zz', szz)])


if __name__ == '__main__':
    main()
