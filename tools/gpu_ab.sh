#!/bin/bash
# generic A/B: parity of the default build, then bench c3 for default, each VARIANTS entry, default again
mkdir -p gpurun_out
export MGS_NO_BUILD=1
echo "== parity"; timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -2 | cut -c1-300
run() { tag=$1; wl=$2; shift 2; timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-c5 --steps 40 "$@" > gpurun_out/ab_${tag}_$wl.json 2> gpurun_out/ab_${tag}_$wl.err || tail -3 gpurun_out/ab_${tag}_$wl.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/ab_${tag}_$wl.json')); m=d.get('measured') or {}; s=m.get('stage_ms_per_launch') or {}
    print('$tag $wl: %.4g G/s  %.3f ms/step  fwd %.4f bwd %.4f pbwd_views %s'%(d['value'],d['ms_per_step'],s.get('blend_fwd',0),s.get('blend_bwd',0),m.get('project_bwd_views_ms')))
except Exception as e: print('$tag $wl: no json', e)
PY
}
for wl in ${WLS:-c3}; do
run default $wl
IFS=';' read -ra VS <<< "$VARIANTS"
for spec in "${VS[@]}"; do v="${spec%%:*}"; defs="${spec#*:}"; MGS_VARIANT=$v MGS_NVCC_DEFINES="$defs" run $v $wl; done
run default2 $wl
done
