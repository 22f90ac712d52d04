#!/bin/bash
# new parity cases + c3 (e2e with lagged loss read) + c4 (activation roofline)
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "sweep or degenerate or opaque" > gpurun_out/pytest_new.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_new.log | cut -c1-300
for spec in "c3 ours" "c4 ours" "mg ours" "mg reference"; do set -- $spec
  timeout 600 python bench.py --workload $1 --impl $2 --no-cpu-baseline > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err; echo "bench $1 $2 rc=$?"; grep -v Warning gpurun_out/bench_$1_$2.err | tail -2
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$1_$2.json'))
    print('$1 $2 value %.4g ms/step %.3f'%(d['value'],d['ms_per_step']),'e2e %.4g'%d.get('e2e',{}).get('value',0),'launches',d.get('gpu_launches'))
    print('   roofline', d.get('roofline'))
    print('   stages', (d.get('measured') or d['config']).get('stage_ms_per_launch'))
except Exception as e: print('no json', e)
PY
done
