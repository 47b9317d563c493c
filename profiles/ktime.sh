#!/bin/bash
# per-kernel device times (ncu launch list, cold-cache/serialised) of the timed step, once per library under build/variants/
N=${1:-64}
mkdir -p gpurun_out/ktime
shopt -s nullglob
for lib in default build/variants/*.so; do
  if [ "$lib" = default ]; then unset CPFEM_B200_LIB; else export CPFEM_B200_LIB=$PWD/$lib; fi
  tag=$(basename $lib .so)
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_update_state|k_point_tangent|k_element_tangent|k_residual' -s 12 -c 9 --csv \
      --log-file gpurun_out/ktime/$tag.csv python bench.py --n $N --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
  python - <<PY
import csv,collections
rows=list(csv.reader(open('gpurun_out/ktime/$tag.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; start=i; break
kn=h.index('Kernel Name'); mv=h.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[start+2:]:
    if len(r)>mv: agg[r[kn].split('(')[0]].append(float(r[mv].replace(',',''))/1e6)
print('$tag', {k: round(min(v),4) for k,v in agg.items()})
PY
done
