#!/bin/bash
# the driver's invocations: default bench of both arms, smoke
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo rc=$?; tail -2 gpurun_out/smoke.log
echo "== bench default ours"; ( time timeout 900 python bench.py > gpurun_out/bench_default_ours.json 2> gpurun_out/bench_default_ours.err ) 2>&1 | grep real; tail -3 gpurun_out/bench_default_ours.err; python -c "
import json;d=json.load(open('gpurun_out/bench_default_ours.json'));print({k:(v if not isinstance(v,dict) else '...') for k,v in d.items()}); print('e2e',d.get('e2e')); print('cpu',d.get('cpu_baseline')); print('roof',d.get('roofline')); print('clocks',d.get('clocks'))"
echo "== bench default reference"; ( time timeout 900 python bench.py --impl reference > gpurun_out/bench_default_ref.json 2> gpurun_out/bench_default_ref.err ) 2>&1 | grep real; tail -3 gpurun_out/bench_default_ref.err; python -c "
import json;d=json.load(open('gpurun_out/bench_default_ref.json'));print({k:(v if not isinstance(v,dict) else '...') for k,v in d.items()}); print('e2e',d.get('e2e')); print('cpu',d.get('cpu_baseline'))"
