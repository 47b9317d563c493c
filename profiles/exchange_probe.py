#!/usr/bin/env python
"""Cost of the multi-GPU steps that follow an assembly, in isolation (torchrun, >= 2 GPUs, 200^3 z-slabs by default):
the interface exchange of ExchangePlan (NCCL send/recv of residual + CSR rows, unpack-add) and the owned-row norm +
all-reduce.  Device time by CUDA events around 10 repetitions, max over ranks.
usage: torchrun ... profiles/exchange_probe.py [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'jax-cpfem_b200'))
import torch
import torch.distributed as dist
from cpfem_b200 import Plan, slip_systems
from cpfem_b200.partition import slab_partition_structured, ExchangePlan

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rm = slab_partition_structured(N, world, rank)
plan = Plan(rm.cells, rm.points, slip_systems.FCC12)
plan.set_active_cells(rm.n_owned_cells)
ip, ix = plan.csr_pattern()
ex = ExchangePlan(rm, ip, ix)
ex.prepare()
res = torch.randn(plan.nn, 3, dtype=torch.float64, device=dev)
csr = torch.randn(plan.nnz, dtype=torch.float64, device=dev)
norm = torch.zeros(1, dtype=torch.float64, device=dev)
sent = sum(ex.send_slots[p].numel() + ex.send_rows[p].numel() for p in ex.send_rows) * 8
recv = sum(ex.recv_slots[p].numel() + ex.recv_rows[p].numel() for p in ex.recv_rows) * 8

def timed(fn, reps=10):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

def norm_step():
    s = ex.owned_sumsq(res, norm)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)

t_ex = timed(lambda: ex.exchange(res, csr))
ex.attach_peer(with_csr=True)
t_peer = timed(lambda: ex.exchange_peer(res, csr))
t_peer_res = timed(lambda: ex.exchange_peer(res, None))
n_to = ex.peer_timeouts()
ex.detach_peer()
t_res = timed(lambda: ex.exchange(res, None))
t_nrm = timed(norm_step)
t_memset = timed(lambda: csr.zero_())
mx = torch.tensor([sent, recv], dtype=torch.float64, device=dev)
dist.all_reduce(mx, op=dist.ReduceOp.MAX)
if rank == 0:
    print('n = %d, %d ranks, NCCL env: %s' % (N, world, {k: v for k, v in os.environ.items() if k.startswith('NCCL_')}))
    print('exchange (residual + CSR rows): %.3f ms for up to %.1f MB sent / %.1f MB received per rank = %.0f GB/s per direction'
          % (t_ex, mx[0].item() / 1e6, mx[1].item() / 1e6, mx[0].item() / t_ex / 1e6))
    print('peer-memory exchange (cpfem_peer_put into the owner\'s mailbox over NVLink + flag, unpack-add): %.3f ms = %.0f GB/s per direction; '
          'residual alone %.3f ms; wait timeouts %d' % (t_peer, mx[0].item() / t_peer / 1e6, t_peer_res, n_to))
    print('exchange of the residual alone: %.3f ms; owned-row norm + all-reduce: %.3f ms; CSR zero-fill (%.2f GB): %.3f ms'
          % (t_res, t_nrm, csr.numel() * 8 / 1e9, t_memset))
dist.barrier()
dist.destroy_process_group()
