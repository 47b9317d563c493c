#!/bin/bash
# round 2: GPU tests + default bench line (with e2e, CPU baselines) + reference arm, as the driver runs them
TAG=${1:-r2g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -x -q -s --durations=8 ) > $OUT/pytest_gpu.log 2>&1
tail -4 $OUT/pytest_gpu.log
grep -E "point-evaluations|disputed|copper curve" $OUT/pytest_gpu.log > $OUT/statistics.txt
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > $OUT/bench_n200.json 2> $OUT/bench_n200.err
tail -3 $OUT/bench_n200.err
python profiles/bench_summary.py $OUT/bench_n200.json 2>/dev/null || head -c 600 $OUT/bench_n200.json
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $OUT/bench_reference.json 2> $OUT/bench_reference.err
tail -3 $OUT/bench_reference.err; head -c 700 $OUT/bench_reference.json
