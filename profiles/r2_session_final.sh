#!/bin/bash
# round 2, final single-GPU evidence: smoke, GPU suite, default bench line + reference arm (as the driver runs them),
# ncu launch list of the bench command and --set full captures of the hot kernels at 128^3
TAG=${1:-r2k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
( time timeout 1800 python -m pytest tests -m gpu -x -q -s --durations=8 ) > $OUT/pytest_gpu.log 2>&1
tail -4 $OUT/pytest_gpu.log
grep -E "point-evaluations|disputed|copper curve|implicit_vjp block" $OUT/pytest_gpu.log > $OUT/statistics.txt
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > $OUT/bench_n200.json 2> $OUT/bench_n200.err
tail -3 $OUT/bench_n200.err
python profiles/bench_summary.py $OUT/bench_n200.json 2>/dev/null || head -c 600 $OUT/bench_n200.json
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $OUT/bench_reference.json 2> $OUT/bench_reference.err
head -c 400 $OUT/bench_reference.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_n128.csv \
    python bench.py --n 128 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_update_state|k_point_tangent|k_element_tangent' -s 23 -c 3 \
    -f -o $OUT/prof_n128 python bench.py --n 128 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log
ls -la $OUT
