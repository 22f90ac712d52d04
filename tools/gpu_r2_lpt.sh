#!/bin/bash
# round 2, late: longest-list-first launch order, conflict-free forward weight tile, product-form transmittance,
# project_bwd occupancy -- parity of the default build, then A/B timings (runtime switch + prebuilt variants)
mkdir -p gpurun_out
export MGS_NO_BUILD=1
echo "== parity"; timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -8 | cut -c1-600
run() { tag=$1; wl=$2; shift 2; timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-c5 --steps 40 "$@" > gpurun_out/lpt_${tag}_$wl.json 2> gpurun_out/lpt_${tag}_$wl.err || tail -3 gpurun_out/lpt_${tag}_$wl.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/lpt_${tag}_$wl.json')); m=d.get('measured') or {}; s=m.get('stage_ms_per_launch') or {}
    print('$tag $wl: %.4g G/s  %.3f ms/step  fwd %.4f bwd %.4f pbwd %.4f pbwd_views %s ranges %.4f'%(d['value'],d['ms_per_step'],s.get('blend_fwd',0),s.get('blend_bwd',0),s.get('project_bwd',0),m.get('project_bwd_views_ms'),s.get('ranges_pack',0)))
except Exception as e: print('$tag $wl: no json', e)
PY
}
for wl in c3; do
  run default $wl
  MGS_TILE_ORDER=0 run noorder $wl
done
IFS=';' read -ra VS <<< "$VARIANTS"
for spec in "${VS[@]}"; do v="${spec%%:*}"; defs="${spec#*:}"; MGS_VARIANT=$v MGS_NVCC_DEFINES="$defs" run $v c3; done
run default2 c3
MGS_ONE_STREAM=1 run onestream c3
echo "== tile stats"; timeout 200 python tools/tile_stats.py c3 > gpurun_out/tile_stats_c3.json 2> gpurun_out/tile_stats.err || tail -5 gpurun_out/tile_stats.err
python - <<'PY'
import json
try:
    for r in json.load(open('gpurun_out/tile_stats_c3.json')): print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items()})
except Exception as e: print('no stats', e)
PY
