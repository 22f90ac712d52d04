#!/bin/bash
# the driver's single-GPU invocations: GPU tests, smoke, default bench of both arms
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench default ours"; ( time timeout 900 python bench.py > gpurun_out/bench_default_ours.json 2> gpurun_out/bench_default_ours.err ) 2>&1 | grep real; grep -v Warning gpurun_out/bench_default_ours.err | tail -3
echo "== bench default reference"; ( time timeout 900 python bench.py --impl reference > gpurun_out/bench_default_ref.json 2> gpurun_out/bench_default_ref.err ) 2>&1 | grep real; grep -v Warning gpurun_out/bench_default_ref.err | tail -3
python - <<'PY'
import json
for t in ('ours','ref'):
    d=json.load(open('gpurun_out/bench_default_%s.json'%t))
    print(t,'value %.4g ms/step %.3f'%(d['value'],d['ms_per_step']),'e2e',d.get('e2e',{}).get('value'),'launches',d.get('gpu_launches'))
    print('   roofline',d.get('roofline')); print('   cpu',d.get('cpu_baseline')); print('   clocks',d.get('clocks')); print('   stages',(d.get('measured') or d['config']).get('stage_ms_per_launch'))
PY
