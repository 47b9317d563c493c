#!/bin/bash
# One gpurun session: GPU parity suite, bench lines, ncu launch list + full captures.  Outputs -> gpurun_out/
# usage: gpurun --timeout 1500 -- 'bash profiles/gpu_session.sh <tag> [n_small]'
TAG=${1:-r1}
NS=${2:-64}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> $OUT/nvidia-smi.txt; free -g >> $OUT/nvidia-smi.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
tail -5 $OUT/pytest_gpu.log
timeout 600 python bench.py --n $NS --steps 5 --warmup 3 > $OUT/bench_n$NS.json 2> $OUT/bench_n$NS.err
cat $OUT/bench_n$NS.json
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e > $OUT/bench_n200.json 2> $OUT/bench_n200.err
cat $OUT/bench_n200.json; tail -3 $OUT/bench_n200.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_n$NS.csv \
    python bench.py --n $NS --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_update_state|k_point_tangent|k_element_tangent' -s 17 -c 3 \
    -f -o $OUT/prof_n$NS python bench.py --n $NS --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_bicg_spmv' -s 1 -c 2 \
    -f -o $OUT/prof_spmv_n$NS python bench.py --n $NS --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/ncu_spmv.log 2>&1
ls -la $OUT
