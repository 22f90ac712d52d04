#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python bench.py --no-e2e --no-cpu-baseline "$@" > gpurun_out/b_$tag.json 2> gpurun_out/b_$tag.err; python -c "
import json;d=json.load(open('gpurun_out/b_$tag.json'));h=d['config'].get('host_step_ms');print('$tag: value %.4g ms/step %.3f max %.1f'%(d['value'],d['ms_per_step'],max(h)), d['config'].get('stage_ms_per_launch'))"; grep -v Warning gpurun_out/b_$tag.err | tail -2; }
run s1 --streams 1
run s4 --streams 4
