#!/usr/bin/env python
"""Host <-> device bandwidth of the box with all ranks copying at once (torchrun, one rank per GPU): what bounds the `e2e`
leg of bench.py (host-resident state, 264 B/point each way) when N GPUs share one host.  Every rank copies 1 GiB pinned
buffers H2D, D2H and both at once (two streams); rank 0 prints per-rank and aggregate GB/s and what the OS says about
NUMA nodes and CPU affinity.

usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29551 profiles/host_bw_probe.py"""
import glob
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
GB = 1 << 30
h_in = torch.empty(GB, dtype=torch.uint8).pin_memory()
h_out = torch.empty(GB, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
d_in = torch.empty(GB, dtype=torch.uint8, device=dev)
d_out = torch.ones(GB, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=6):
    for it in range(reps + 2):
        if it == 2:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s1.wait_event(e0)
            s2.wait_event(e0)
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    e1.synchronize()
    return reps * GB / (e0.elapsed_time(e1) * 1e-3) / 1e9      # GB/s per direction


res = []
for name, a, b in (('H2D alone', True, False), ('D2H alone', False, True), ('H2D + D2H at once (per direction)', True, True)):
    v = torch.tensor([run(a, b)], dtype=torch.float64, device=dev)
    if world > 1:
        allv = [torch.zeros_like(v) for _ in range(world)]
        dist.all_gather(allv, v)
        vals = [float(x) for x in allv]
    else:
        vals = [float(v)]
    res.append((name, vals))
if rank == 0:
    nodes = sorted(glob.glob('/sys/devices/system/node/node[0-9]*'))
    print(f'{world} rank(s); host: {os.cpu_count()} logical CPUs, NUMA nodes visible: {len(nodes)}, affinity of rank 0: {len(os.sched_getaffinity(0))} CPUs')
    try:
        with open('/proc/meminfo') as f:
            print('  ' + f.readline().strip())
    except OSError:
        pass
    for name, vals in res:
        print(f'{name:36s} per rank {" ".join("%5.1f" % x for x in vals)}  | aggregate {sum(vals):6.1f} GB/s, slowest {min(vals):5.1f}')
    print('bench.py e2e moves 264 B/point each way: at S GB/s per direction (slowest rank, both directions busy) the floor of a '
          '200^3 / N slab is 64e6 * 264 / N / S.')
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
