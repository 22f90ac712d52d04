#!/bin/bash
# round 2: parity of the tensor-core blends (isolating forward / backward on failure) + A/B timing against the round-1
# SIMT kernels and chunk-size variants (prebuilt with tools/build_variants.sh and the same VARIANTS string)
mkdir -p gpurun_out
par() { echo "== parity MGS_BLEND=$1"; MGS_BLEND=$1 timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -${2:-12} | cut -c1-600; }
par mma 25 | tee gpurun_out/r2_parity_mma.txt
if ! grep -q " passed" gpurun_out/r2_parity_mma.txt || grep -q "failed" gpurun_out/r2_parity_mma.txt; then par simt_bwd 25; par simt_fwd 25; fi
run() { tag=$1; shift; timeout 300 python bench.py --workload ${WL:-c3} --no-e2e --no-cpu-baseline --steps 30 "$@" > gpurun_out/r2b_$tag.json 2> gpurun_out/r2b_$tag.err || tail -3 gpurun_out/r2b_$tag.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2b_$tag.json')); s=d['config'].get('stage_ms_per_launch') or {}
    print('$tag: %.4g G/s  %.3f ms/step  fwd %.4f bwd %.4f'%(d['value'],d['ms_per_step'],s.get('blend_fwd',0),s.get('blend_bwd',0)))
except Exception as e: print('$tag: no json', e)
PY
}
run mma
MGS_BLEND=simt run simt
IFS=';' read -ra VS <<< "$VARIANTS"
for spec in "${VS[@]}"; do v="${spec%%:*}"; defs="${spec#*:}"; MGS_VARIANT=$v MGS_NVCC_DEFINES="$defs" run $v; done
