#!/usr/bin/env python
"""Update / assembly time of the four uniform-material parameter sets at n^3 (load step 11, CUDA events).
usage: material_probe.py [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'jax-cpfem_b200'))
import numpy as np, torch
from cpfem_b200 import Plan, make_material, synthetic, slip_systems
from cpfem_b200.problem import get_rot_mat
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device('cuda', 0)
mesh, quat, gid = synthetic.polycrystal(N)
SETS = {  # name: (slip, g0, material args, d_eps, dt)
    'copper (FCC12, n = 10)': (slip_systems.FCC12, 60.8, (1.684e5, 1.214e5, 0.754e5, 541.5, 109.8, 2.5, 0.1, 1.0, 0.001, 1e-8, 5), 1e-3, 1e-2),
    'tantalum (BCC12, n = 45.2726, pow())': (slip_systems.BCC12, 67.4641, (2.670e5, 1.610e5, 0.825e5, 1959.132, 7295.1754, 200.0, 1.0 / 45.2726, 1.0, 0.001, 1e-8, 5), -2.5e-4, 0.25),
    '304 steel (FCC12, n = 120)': (slip_systems.FCC12, 90.0, (2.622e5, 1.120e5, 0.746e5, 392.9772, 7295.1754, 8.0, 1.0 / 120.0, 1.0, 0.001, 1e-8, 8), 2e-4, 2e-3),
}
pts = torch.as_tensor(mesh.points, device=dev)
noise = torch.as_tensor(synthetic.noise_field(N), device=dev)
def timeit(f, n=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n
for name, (slip, g0, margs, deps, dt) in SETS.items():
    plan = Plan(mesh.cells, mesh.points, slip)
    nc = plan.nc
    mat = make_material(*margs)
    rot = torch.as_tensor(get_rot_mat(quat)[gid], device=dev)[:, None].expand(nc, 8, 3, 3).contiguous()
    cur = [torch.eye(3, dtype=torch.float64, device=dev).expand(nc, 8, 3, 3).contiguous(), torch.full((nc, 8, 12), g0, dtype=torch.float64, device=dev),
           torch.zeros(nc, 8, 12, dtype=torch.float64, device=dev), rot]
    nxt = [torch.empty_like(cur[0]), torch.empty_like(cur[1]), torch.empty_like(cur[2])]
    disp = lambda s: (pts * torch.tensor([-0.3, -0.3, 1.0], dtype=torch.float64, device=dev) * (deps * s) + noise).contiguous()
    for s in range(1, 11):
        plan.update_state(mat, disp(s), cur, dt, out=nxt)
        cur, nxt = [nxt[0], nxt[1], nxt[2], rot], [cur[0], cur[1], cur[2]]
    sol = disp(11)
    st = plan.new_status()
    t_u = timeit(lambda: plan.update_state(mat, sol, cur, dt, out=nxt, status=st))
    iters = float(st[3]) / (4 * nc * 8)
    t_a = timeit(lambda: plan.newton_update(mat, sol, cur, dt))
    print('%-40s n = %d: update %.3f ms (%.3g updates/s), assembly %.3f ms, local Newton iterations %.2f' % (name, N, t_u, nc * 8 / t_u * 1e3, t_a, iters))
    del plan
