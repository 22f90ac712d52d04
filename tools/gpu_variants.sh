#!/bin/bash
# kernel variants (MGS_VARIANT + MGS_NVCC_DEFINES, prebuilt here with the same pair so the stamp matches on the box):
# stage timings per variant; VARIANTS="name:defines;name:defines"
mkdir -p gpurun_out
IFS=';' read -ra VS <<< "$VARIANTS"
for spec in "${VS[@]}"; do
  v="${spec%%:*}"; defs="${spec#*:}"
  export MGS_VARIANT=$v MGS_NVCC_DEFINES="$defs"
  if [ -n "$PARITY" ]; then echo "== parity $v"; timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -3 | cut -c1-300; fi
  for w in ${WORKLOADS:-c3}; do
    timeout 300 python bench.py --workload $w --no-e2e --no-cpu-baseline --steps 30 > gpurun_out/var_${v}_$w.json 2> gpurun_out/var_${v}_$w.err || tail -3 gpurun_out/var_${v}_$w.err
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/var_${v}_$w.json'))
    s=(d.get('measured') or d['config']).get('stage_ms_per_launch') or {}
    print('variant $v [$defs] $w: %.4g G/s  %.3f ms/step  fwd %.4f bwd %.4f'%(d['value'],d['ms_per_step'],s.get('blend_fwd',0),s.get('blend_bwd',0)))
except Exception as e: print('variant $v $w: no json', e)
PY
  done
done
