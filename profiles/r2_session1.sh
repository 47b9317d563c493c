#!/bin/bash
# round 2, GPU session 1: baseline at bench size (ncu --set full at 128^3), layout A/B, compile-flag variants, RED probe
OUT=gpurun_out/r2a
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> $OUT/nvidia-smi.txt; free -g >> $OUT/nvidia-smi.txt
./build/red_probe > $OUT/red_probe.txt 2>&1; cat $OUT/red_probe.txt
rm -f gpurun_out/variants/results_n200.txt
bash profiles/variants.sh 200
cp gpurun_out/variants/results_n200.txt $OUT/variants_n200.txt
for lib in default all3; do
  if [ "$lib" = default ]; then unset CPFEM_B200_LIB; else export CPFEM_B200_LIB=$PWD/build/variants/$lib.so; fi
  python bench.py --n 200 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --layout soa > $OUT/bench_soa_$lib.json 2> $OUT/bench_soa_$lib.err
  python -c "
import json; d=json.load(open('$OUT/bench_soa_$lib.json')); print('SOA $lib update_ms', d['update_ms'], 'assembly_ms', d['assembly_ms'], 'elastic', d['elastic_step']['update_ms'])"
done
unset CPFEM_B200_LIB
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
tail -3 $OUT/pytest_gpu.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_update_state|k_point_tangent|k_element_tangent' -s 23 -c 3 \
    -f -o $OUT/prof_n128 python bench.py --n 128 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log
timeout 600 ncu --set full --clock-control none -k regex:'k_update_state' -s 15 -c 1 \
    -f -o $OUT/prof_soa_n128 python bench.py --n 128 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --layout soa > $OUT/ncu_soa.log 2>&1
ls -la $OUT
