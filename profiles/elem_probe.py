#!/usr/bin/env python
"""Where does the assembly time go?  Times newton_update at n^3 with the CSR scatter on / off (CUDA events)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'jax-cpfem_b200'))
import numpy as np, torch
from cpfem_b200 import Plan, make_material, synthetic, slip_systems
from cpfem_b200.problem import get_rot_mat
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device('cuda', 0)
mesh, quat, gid = synthetic.polycrystal(N)
plan = Plan(mesh.cells, mesh.points, slip_systems.FCC12)
nc = plan.nc
mat = make_material(2.622e5, 1.120e5, 0.746e5, 392.9772, 7295.1754, 8.0, 1.0 / 120.0, 1.0, 0.001, 1e-8, 8)
rot = torch.as_tensor(get_rot_mat(quat)[gid], device=dev)[:, None].expand(nc, 8, 3, 3).contiguous()
cur = [torch.eye(3, dtype=torch.float64, device=dev).expand(nc, 8, 3, 3).contiguous(), torch.full((nc, 8, 12), 90.0, dtype=torch.float64, device=dev),
       torch.zeros(nc, 8, 12, dtype=torch.float64, device=dev), rot]
pts = torch.as_tensor(mesh.points, device=dev)
noise = torch.as_tensor(synthetic.noise_field(N), device=dev)
disp = lambda s: (pts * torch.tensor([-0.3, -0.3, 1.0], dtype=torch.float64, device=dev) * (2e-4 * s) + noise).contiguous()
for s in range(1, 11):
    new = plan.update_state(mat, disp(s), cur, 2e-3)
    cur = [new[0], new[1], new[2], rot]
sol = disp(11)
res = torch.empty(plan.nn, 3, dtype=torch.float64, device=dev)
csr = torch.empty(plan.nnz, dtype=torch.float64, device=dev)
def timeit(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n
print('n', N, 'assembly with CSR scatter   %.3f ms' % timeit(lambda: plan.newton_update(mat, sol, cur, 2e-3, res=res, csr_data=csr)))
print('n', N, 'assembly without CSR scatter %.3f ms' % timeit(lambda: plan.newton_update(mat, sol, cur, 2e-3, res=res, want_csr=False)))
print('n', N, 'memset of csr_data           %.3f ms' % timeit(lambda: csr.zero_()))
print('n', N, 'update                       %.3f ms' % timeit(lambda: plan.update_state(mat, sol, cur, 2e-3, out=new)))
