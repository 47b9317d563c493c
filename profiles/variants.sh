#!/bin/bash
# kernel-tuning A/B: runs the n^3 bench (device-resident numbers only) once per library build under build/variants/
N=${1:-64}
mkdir -p gpurun_out/variants
shopt -s nullglob
for lib in default build/variants/*.so; do
  if [ "$lib" = default ]; then unset CPFEM_B200_LIB; else export CPFEM_B200_LIB=$PWD/$lib; fi
  python bench.py --n $N --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$lib', 'update_ms', round(d['update_ms'],3), 'assembly_ms', round(d['assembly_ms'],3), 'iters', round(d['mean_local_newton_iters'],3), 'resnorm', d['residual_norm'])
    else: print(l.rstrip())
" | tee -a gpurun_out/variants/results_n$N.txt
done
