#!/bin/bash
# rows f1-f3: GPU tests of the new layers, then the c3 line (e2e now through render_views) and the c4 dyna line for both arms
mkdir -p gpurun_out
echo "== pytest new"; timeout 900 python -m pytest tests/test_render_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_render.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_render.log | cut -c1-400
for w in c3 c4; do for impl in ours reference; do
  echo "== bench $w $impl"; timeout 600 python bench.py --workload $w --impl $impl --no-cpu-baseline > gpurun_out/bench_${w}_${impl}.json 2> gpurun_out/bench_${w}_${impl}.err; echo "rc=$?"; grep -v Warning gpurun_out/bench_${w}_${impl}.err | tail -3
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${w}_${impl}.json'))
    print('$w $impl value %.4g ms/step %.3f'%(d['value'],d['ms_per_step']),'e2e',d.get('e2e'),'launches',d.get('gpu_launches'))
    print('   ', {k:d['config'].get(k) for k in ('loss','grads_checked','stage_ms_per_launch')})
except Exception as e: print('no json', e)
PY
done; done
