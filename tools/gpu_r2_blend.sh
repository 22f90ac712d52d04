#!/bin/bash
# round 2: parity of the tensor-core blends + timing of the default build and of prebuilt variants
# (tools/build_variants.sh with the same VARIANTS string); FULL=1 also runs the BASELINE-size parity tests
mkdir -p gpurun_out
echo "== parity"; timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -12 | cut -c1-600
if [ -n "$FULL" ]; then echo "== baseline sizes"; timeout 1500 python -m pytest tests/test_baseline_sizes_gpu.py -m gpu -q --tb=short -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -30 | cut -c1-700; fi
run() { tag=$1; shift; timeout 300 python bench.py --workload ${WL:-c3} --no-e2e --no-cpu-baseline --steps 30 "$@" > gpurun_out/r2b_$tag.json 2> gpurun_out/r2b_$tag.err || tail -3 gpurun_out/r2b_$tag.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2b_$tag.json')); s=(d.get('measured') or d['config']).get('stage_ms_per_launch') or {}
    print('$tag: %.4g G/s  %.3f ms/step  fwd %.4f bwd %.4f'%(d['value'],d['ms_per_step'],s.get('blend_fwd',0),s.get('blend_bwd',0)), {k:v for k,v in s.items() if not k.startswith('blend') and v})
except Exception as e: print('$tag: no json', e)
PY
}
run default
IFS=';' read -ra VS <<< "$VARIANTS"
for spec in "${VS[@]}"; do v="${spec%%:*}"; defs="${spec#*:}"; MGS_VARIANT=$v MGS_NVCC_DEFINES="$defs" run $v; done
