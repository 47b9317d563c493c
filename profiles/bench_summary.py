#!/usr/bin/env python
"""Key numbers of a bench.py JSON line.  usage: bench_summary.py file.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
print('n_gpus %d  value %.4g %s  update %.3f ms  assembly %.3f ms  fused update+avg %.3f ms  avg stress %.3f ms' % (
    d['n_gpus'], d['value'], d['unit'], d['update_ms'], d['assembly_ms'], d.get('update_avg_stress_fused_ms', float('nan')),
    d.get('avg_stress_ms', float('nan'))))
r, ra = d['roofline'], d['roofline_assembly']
if 'executed' in r:       # round-1 lines: frac = contract model, executed.frac = FP64 slots
    print('roofline (contract) %.3f  executed %.3f | assembly %.3f / %.3f | local Newton iters %.9f' % (
        r['frac'], r['executed']['frac'], ra['frac'], ra['executed']['frac'], d['mean_local_newton_iters']))
else:                     # round 2: frac = executed FP64 issue slots, true_flops and contract beside it
    print('roofline.frac (FP64 slots) %.3f  true flops %.3f  contract %.3f | assembly %.3f / %.3f / %.3f | local Newton iters %.9f' % (
        r['frac'], r['true_flops']['frac'], r['contract']['frac_of_peak'], ra['frac'], ra['true_flops']['frac'],
        ra['contract']['frac_of_peak'], d['mean_local_newton_iters']))
print('gpu_launches', d.get('gpu_launches'))
print('elastic step', d.get('elastic_step'))
print('solver', {k: v for k, v in (d.get('solver') or {}).items() if k != 'what'})
if d.get('e2e'):
    print('e2e', {k: d['e2e'][k] for k in ('value', 'ms_per_step', 'assembly_ms')})
if d.get('cpu_baseline'):
    print('cpu baseline %.1f %s on %d cores' % (d['cpu_baseline']['value'], d['cpu_baseline']['unit'], d['cpu_baseline']['cores']))
    sa = d['cpu_baseline'].get('same_algorithm')
    if sa:
        print('cpu baseline, same algorithm as the GPU path: %.4g %s on %d cores (with tangent %.4g points/s)' % (
            sa['value'], sa['unit'], sa['cores'], sa['with_tangent_points_per_s']))
print('clocks', d.get('clocks'))
