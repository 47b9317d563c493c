#!/usr/bin/env python
"""Secondary workload of SURVEY 8(d): the DP-steel set (BCC24, per-point parameters gss_a, h, t_sat, xm, r and elastic
tensor C_gp, 40 % martensite) on the synthetic n^3 polycrystal - update pass and assembly per Newton iteration at load
step 11 (CUDA events).  usage: dp_probe.py [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'jax-cpfem_b200'))
import numpy as np, torch
from cpfem_b200 import Plan, make_material, synthetic, slip_systems
from cpfem_b200.problem import get_rot_mat
from cpfem_b200.models_DPsteel_inhomo import CrystalPlasticity as DP, cubic_C
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device('cuda', 0)
mesh, quat, gid = synthetic.polycrystal(N)
plan = Plan(mesh.cells, mesh.points, slip_systems.BCC24)
nc = plan.nc
ph = DP.phase
rng = np.random.default_rng(0)
phase = np.zeros(nc, dtype=int); phase[:int(0.4 * nc)] = 1; rng.shuffle(phase)
pt = torch.as_tensor(phase, device=dev)
pick = lambda k: torch.tensor(ph[k], dtype=torch.float64, device=dev)[pt]
rep = lambda v: v[:, None].expand(nc, 8).contiguous()
Cph = torch.as_tensor(np.array([cubic_C(ph['C11'][k], ph['C12'][k], ph['C44'][k]) for k in (0, 1)]), device=dev)
C_gp = Cph[pt][:, None].expand(nc, 8, 3, 3, 3, 3).contiguous()
rot = torch.as_tensor(get_rot_mat(quat)[gid], device=dev)[:, None].expand(nc, 8, 3, 3).contiguous()
g0 = rep(pick('gss_initial'))[:, :, None].expand(nc, 8, 24).contiguous()
cur = [torch.eye(3, dtype=torch.float64, device=dev).expand(nc, 8, 3, 3).contiguous(), g0, torch.zeros(nc, 8, 24, dtype=torch.float64, device=dev), rot,
       rep(pick('gss_a0')), rep(pick('h0')), rep(pick('t_sat0')), rep(pick('xm0')), rep(pick('r0')), C_gp]
mat = make_material(ph['C11'][0], ph['C12'][0], ph['C44'][0], ph['h0'][0], ph['t_sat0'][0], ph['gss_a0'][0], ph['xm0'][0], 1.0, 0.001, 1e-8, 5)
pts = torch.as_tensor(mesh.points, device=dev)
noise = torch.as_tensor(synthetic.noise_field(N), device=dev)
disp = lambda s: (pts * torch.tensor([-0.3, -0.3, 1.0], dtype=torch.float64, device=dev) * (4e-4 * s) + noise).contiguous()
nxt = [torch.empty_like(cur[0]), torch.empty_like(cur[1]), torch.empty_like(cur[2])]
for s in range(1, 11):
    plan.update_state(mat, disp(s), cur, 0.2, out=nxt)
    cur, nxt = [nxt[0], nxt[1], nxt[2]] + cur[3:], [cur[0], cur[1], cur[2]]
sol = disp(11)
res = torch.empty(plan.nn, 3, dtype=torch.float64, device=dev)
csr = torch.empty(plan.nnz, dtype=torch.float64, device=dev)
st = plan.new_status()
def timeit(f, n=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n
st.zero_()
t_u = timeit(lambda: plan.update_state(mat, sol, cur, 0.2, out=nxt, status=st))
iters = float(st[3]) / (4 * nc * 8)
t_a = timeit(lambda: plan.newton_update(mat, sol, cur, 0.2, res=res, csr_data=csr))
print('DP steel BCC24 per-point parameters, n = %d (%d points): update %.3f ms = %.3g updates/s, assembly %.3f ms, mean local Newton iterations %.2f, cap hits %d'
      % (N, nc * 8, t_u, nc * 8 / t_u * 1e3, t_a, iters, int(st[0])))
