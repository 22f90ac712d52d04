#!/bin/bash
# round 2: parity of the tensor-core forward blend + A/B timing against the round-1 SIMT kernels and chunk-size variants
mkdir -p gpurun_out
echo "== parity (default build)"; timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -15 | cut -c1-400
run() { tag=$1; shift; timeout 300 python bench.py --workload c3 --no-e2e --no-cpu-baseline --steps 30 "$@" > gpurun_out/r2f_$tag.json 2> gpurun_out/r2f_$tag.err || tail -3 gpurun_out/r2f_$tag.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2f_$tag.json')); s=(d.get('measured') or d['config']).get('stage_ms_per_launch') or {}
    print('$tag: %.4g G/s  %.3f ms/step  fwd %.4f bwd %.4f'%(d['value'],d['ms_per_step'],s.get('blend_fwd',0),s.get('blend_bwd',0)))
except Exception as e: print('$tag: no json', e)
PY
}
run mma
MGS_BLEND=simt run simt
for v in ${VARS:-CH16 B64 CH16M20}; do MGS_VARIANT=$v MGS_NVCC_DEFINES="$(python - <<PY
import os
d={'CH16':'-DMGS_FWD_CH=16','B64':'-DMGS_FWD_BATCH=64','CH16M20':'-DMGS_FWD_CH=16 -DMGS_FWD_MIN_CTAS=20'}
print(d.get('$v',''))
PY
)" run $v; done
