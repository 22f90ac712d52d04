#!/bin/bash
# parity + timing of the current build, then one ncu --set full capture of the two blend kernels (c3 view, 1 stream)
mkdir -p gpurun_out
echo "== parity"; timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -8 | cut -c1-600
run() { tag=$1; shift; timeout 300 python bench.py --workload ${WL:-c3} --no-e2e --no-cpu-baseline --steps 30 "$@" > gpurun_out/r2b_$tag.json 2> gpurun_out/r2b_$tag.err || tail -3 gpurun_out/r2b_$tag.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2b_$tag.json')); s=(d.get('measured') or d['config']).get('stage_ms_per_launch') or {}
    print('$tag: %.4g G/s  %.3f ms/step  fwd %.4f bwd %.4f'%(d['value'],d['ms_per_step'],s.get('blend_fwd',0),s.get('blend_bwd',0)), {k:v for k,v in s.items() if not k.startswith('blend')})
except Exception as e: print('$tag: no json', e)
PY
}
run ${TAG:-cur}
echo "== ncu full"; timeout 1000 ncu --set full --clock-control none --import-source on -k regex:blend -s 8 -c 2 -o gpurun_out/prof_blend_${TAG:-cur} -f python bench.py --steps 2 --warmup 1 --settle 0 --streams 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_${TAG:-cur}.log 2>&1; echo rc=$?; tail -2 gpurun_out/ncu_full_${TAG:-cur}.log
