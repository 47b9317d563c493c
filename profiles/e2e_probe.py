#!/usr/bin/env python
"""Where the end-to-end (host-buffer) update pass stands against the PCIe floor: raw pinned H2D / D2H / duplex copy rates
of this box, then Plan.update_state_host at several chunk sizes (virgin state, elastic step: the pass is bound by the bus
either way).  usage: e2e_probe.py [n] [chunk_cells ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'jax-cpfem_b200'))
import numpy as np, torch
from cpfem_b200 import Plan, make_material, synthetic, slip_systems
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
chunks = [int(a) for a in sys.argv[2:]] or [1 << 14, 1 << 15, 1 << 16, 1 << 17, 1 << 18]
dev = torch.device('cuda', 0)

def ev():
    return torch.cuda.Event(enable_timing=True)

# ---- raw copy rates (1 GiB blocks) -------------------------------------------------------------------------
nb = 1 << 27
h1 = torch.empty(nb, dtype=torch.float64, pin_memory=True); h1.fill_(1.0)
h2 = torch.empty(nb, dtype=torch.float64, pin_memory=True)
d1 = torch.empty(nb, dtype=torch.float64, device=dev)
d2 = torch.ones(nb, dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def rate(f, nbytes, reps=4):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both():
    h2d(); d2h()
print('raw pinned copies, 1 GiB blocks: H2D %.1f GB/s, D2H %.1f GB/s, duplex %.1f GB/s per direction' %
      (rate(h2d, nb * 8), rate(d2h, nb * 8), rate(both, nb * 8)), flush=True)
del h1, h2, d1, d2

# ---- the host-streamed update pass ---------------------------------------------------------------------------
mesh, quat, gid = synthetic.polycrystal(N)
plan = Plan(mesh.cells, mesh.points, slip_systems.FCC12)
nc = plan.nc
mat = make_material(2.622e5, 1.120e5, 0.746e5, 392.9772, 7295.1754, 8.0, 1.0 / 120.0, 1.0, 0.001, 1e-8, 8)
from cpfem_b200.problem import get_rot_mat
hs = [torch.empty((nc, 8, 3, 3), dtype=torch.float64, pin_memory=True), torch.empty((nc, 8, 12), dtype=torch.float64, pin_memory=True),
      torch.empty((nc, 8, 12), dtype=torch.float64, pin_memory=True), torch.empty((nc, 8, 3, 3), dtype=torch.float64, pin_memory=True)]
hs[0].copy_(torch.eye(3, dtype=torch.float64).expand(nc, 8, 3, 3)); hs[1].fill_(90.0); hs[2].zero_()
hs[3].copy_(torch.as_tensor(get_rot_mat(quat)[gid])[:, None].expand(nc, 8, 3, 3))
out = [torch.empty(h.shape, dtype=torch.float64, pin_memory=True) for h in hs[:3]]
sol = torch.empty((plan.nn, 3), dtype=torch.float64, pin_memory=True)
sol.copy_(torch.as_tensor(synthetic.displacement(mesh.points, 2e-4, N)))
nbytes = sum(h.numel() * 8 for h in hs[:3])
print('n = %d: %d points, %.2f GB each way per pass' % (N, nc * 8, nbytes / 1e9), flush=True)
for cc in chunks:
    plan.update_state_host(mat, sol, hs, 2e-3, out=out, chunk_cells=cc)
    t0 = time.perf_counter()
    for _ in range(3):
        plan.update_state_host(mat, sol, hs, 2e-3, out=out, chunk_cells=cc)
    dt = (time.perf_counter() - t0) / 3
    print('chunk %7d cells (%4d chunks): %.1f ms per pass = %.3g updates/s, %.1f GB/s per direction' %
          (cc, -(-nc // cc), dt * 1e3, nc * 8 / dt, nbytes / dt / 1e9), flush=True)
