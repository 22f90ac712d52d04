#!/bin/bash
# whole GPU suite + the bench lines of both arms (c3 headline with e2e and the c5 block; the 16k-Gaussian ManiGaussian call)
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu_r2.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_gpu_r2.log | cut -c1-500
show() { python - <<PY
import json
try:
    d=json.load(open('$1')); m=d.get('measured') or {}
    print('$1: value %.4g %s  ms/step %.3f  e2e %s  launches %s'%(d['value'],d['unit'],d['ms_per_step'],(d.get('e2e') or {}).get('value'),d.get('gpu_launches')))
    print('   stages',m.get('stage_ms_per_launch'),'pbv',m.get('project_bwd_views_ms'),'host',m.get('host_step_ms'),'cap',m.get('binning_capacity_per_view'),'overflow',m.get('binning_overflow'))
    print('   roofline',d.get('roofline')); print('   c5',d.get('c5')); print('   e2e',d.get('e2e'))
except Exception as e: print('$1: no json', e)
PY
}
for w in ${WLS:-c3 mg}; do for impl in ours reference; do
  extra=""; [ "$w" != "c3" ] && extra="--no-c5"
  timeout 900 python bench.py --workload $w --impl $impl --no-cpu-baseline $extra > gpurun_out/r2_bench_${w}_${impl}.json 2> gpurun_out/r2_bench_${w}_${impl}.err; echo "bench $w $impl rc=$?"; grep -v Warning gpurun_out/r2_bench_${w}_${impl}.err | tail -4
  show gpurun_out/r2_bench_${w}_${impl}.json
done; done
