#!/usr/bin/env python
"""SpMV / BiCGStab timing at n^3 (CUDA events / wall clock).  usage: solver_probe.py [n]; honours CPFEM_B200_LIB."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'jax-cpfem_b200'))
import numpy as np, torch
from cpfem_b200 import Plan, make_material, synthetic, slip_systems
from cpfem_b200.problem import get_rot_mat
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device('cuda', 0)
mesh, quat, gid = synthetic.polycrystal(N)
plan = Plan(mesh.cells, mesh.points, slip_systems.FCC12)
nc = plan.nc
mat = make_material(2.622e5, 1.120e5, 0.746e5, 392.9772, 7295.1754, 8.0, 1.0 / 120.0, 1.0, 0.001, 1e-8, 8)
rot = torch.as_tensor(get_rot_mat(quat)[gid], device=dev)[:, None].expand(nc, 8, 3, 3).contiguous()
cur = [torch.eye(3, dtype=torch.float64, device=dev).expand(nc, 8, 3, 3).contiguous(), torch.full((nc, 8, 12), 90.0, dtype=torch.float64, device=dev),
       torch.zeros(nc, 8, 12, dtype=torch.float64, device=dev), rot]
pts = torch.as_tensor(mesh.points, device=dev)
sol = (pts * torch.tensor([-0.3, -0.3, 1.0], dtype=torch.float64, device=dev) * 2e-4).contiguous()
res, csr, _ = plan.newton_update(mat, sol, cur, 2e-3)
x = torch.randn(plan.ndof, dtype=torch.float64, device=dev)
y = torch.empty_like(x)
for _ in range(3): plan.spmv(csr, x, out=y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): plan.spmv(csr, x, out=y)
e1.record(); e1.synchronize()
t = e0.elapsed_time(e1) / 10
nbytes = plan.nnz * 8 + (plan.nnz // 9) * 4 + (plan.nn + 1) * 8 + 2 * plan.ndof * 8
plan.bicgstab(csr, res.reshape(-1), tol=0.0, atol=0.0, maxiter=4)
torch.cuda.synchronize()
t0 = time.perf_counter()
plan.bicgstab(csr, res.reshape(-1), tol=0.0, atol=0.0, maxiter=40)
torch.cuda.synchronize()
tb = (time.perf_counter() - t0) / 40
print(os.environ.get('CPFEM_B200_LIB', 'default'), 'n', N, 'spmv %.4f ms  %.0f GB/s | bicgstab %.4f ms/iteration' % (t, nbytes / t / 1e6, tb * 1e3))
