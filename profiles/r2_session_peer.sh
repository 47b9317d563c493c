#!/bin/bash
# round 2: peer-memory interface exchange on N GPUs (N = number of visible GPUs): correctness check, exchange probe, bench A/B
OUT=gpurun_out/r2i
mkdir -p $OUT
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
nvidia-smi topo -m > $OUT/topo_${N}gpu.txt 2>&1
( timeout 600 $TR tests/multigpu_check.py ) > $OUT/multigpu_check_${N}gpu.txt 2>&1; echo "multigpu_check rc=$?"; grep -E "rank|Error|error" $OUT/multigpu_check_${N}gpu.txt | head -12
( timeout 600 $TR profiles/exchange_probe.py 200 ) > $OUT/exchange_probe_${N}gpu.txt 2>&1; echo "probe rc=$?"; grep -vE "^W|warn" $OUT/exchange_probe_${N}gpu.txt | tail -6
( timeout 300 $TR profiles/host_bw_probe.py ) > $OUT/host_bw_probe_${N}gpu.txt 2>&1; echo "host_bw rc=$?"; grep -vE "^W|warn|\*\*\*" $OUT/host_bw_probe_${N}gpu.txt | tail -7
for ex in peer nccl; do
  E2E="--no-e2e"; if [ "$ex" = peer ] && [ "${WITH_E2E:-0}" = 1 ]; then E2E=""; fi
  ( timeout 900 $TR bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 $E2E --no-cpu-baseline --exchange $ex ) > $OUT/bench_${N}gpu_$ex.json 2> $OUT/bench_${N}gpu_$ex.err
  python - <<PY
import json
for l in open('$OUT/bench_${N}gpu_$ex.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$ex', 'n_gpus', d['n_gpus'], 'update_ms', round(d['update_ms'],3), 'assembly_ms', round(d['assembly_ms'],3), 'ms_per_step', round(d['ms_per_step'],3), d['config'].get('exchange'), 'e2e', d.get('e2e'))
PY
done
