#!/bin/bash
# quick regression + timing loop: GPU tests, then bench at 1 and 4 streams
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
run() { tag=$1; shift; timeout 600 python bench.py --no-e2e --no-cpu-baseline "$@" > gpurun_out/b_$tag.json 2> gpurun_out/b_$tag.err; python -c "
import json;d=json.load(open('gpurun_out/b_$tag.json'));h=(d.get('measured') or d['config']).get('host_step_ms');print('$tag: value %.4g ms/step %.3f max %.1f'%(d['value'],d['ms_per_step'],h['max']), (d.get('measured') or d['config']).get('stage_ms_per_launch'))"; tail -2 gpurun_out/b_$tag.err; }
run s1 --streams 1
run s4 --streams 4
