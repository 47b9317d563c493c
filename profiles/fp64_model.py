#!/usr/bin/env python
"""Per-point work of the three hot kernels from an `ncu --set full` summary (profiles/ncu_summary.py output) taken at
the bench state: FP64 thread instructions (DFMA / DMUL / DADD), DRAM bytes, duration.  Writes profiles/fp64_instr.json,
which bench.py reads for `roofline` (executed FP64 issue slots and true flops per point, traffic per launch).

usage: fp64_model.py profiles/r2/<tag>_ncu_summary_n128.txt 128 > profiles/fp64_instr.json
"""
import json
import re
import sys

path, n = sys.argv[1], int(sys.argv[2])
chunk_cells = min(n ** 3, 1 << 19)
points = {'k_update_state': 8 * n ** 3, 'k_point_tangent': 8 * chunk_cells, 'k_element_tangent': 8 * chunk_cells}
kern, cur = {}, None
for line in open(path):
    m = re.match(r'^(?:void )?(k_\w+)', line)
    if m:
        cur = kern.setdefault(m.group(1), {})
        continue
    t = line.split()
    if cur is not None and len(t) >= 2:
        try:
            cur[t[0]] = float(t[1])
        except ValueError:
            pass
out = {'source': path, 'mesh_n': n, 'what': 'ncu --set full --clock-control none of the timed step of bench.py --n %d (load step 11 of the '
       '304-steel workload, all points plastic); per-point figures = launch totals / points of the launch' % n, 'kernels': {}}
for name, v in kern.items():
    if name not in points:
        continue
    cyc = v['sm__cycles_elapsed.max']
    p = points[name]
    d = {k: v['smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed' % k] * cyc / p for k in ('dfma', 'dmul', 'dadd')}
    out['kernels'][name] = {
        'points_per_launch': p, 'ms': v['gpu__time_duration.sum'],
        'dfma_per_point': d['dfma'], 'dmul_per_point': d['dmul'], 'dadd_per_point': d['dadd'],
        'fp64_instr_per_point': d['dfma'] + d['dmul'] + d['dadd'],
        'flops_per_point': 2 * d['dfma'] + d['dmul'] + d['dadd'],
        'dram_bytes_per_point': (v['dram__bytes_read.sum'] + v['dram__bytes_write.sum']) * 1e9 / p,
        'pipe_fp64_cycles_active_pct': v['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'],
        'registers': v['launch__registers_per_thread']}
print(json.dumps(out, indent=1))
