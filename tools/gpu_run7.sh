#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline "$@" > gpurun_out/b_$tag.json 2> gpurun_out/b_$tag.err; python -c "
import json;d=json.load(open('gpurun_out/b_$tag.json'));print('$tag: ms/step %.3f wall %.3f'%(d['ms_per_step'],d['config']['wall_ms_per_step']), d['config'].get('host_step_ms'))"; tail -2 gpurun_out/b_$tag.err; }
run s1_full --streams 1
run s1_noclk --streams 1 --no-clocks
run s1_nost --streams 1 --no-stage-timing
run s1_none --streams 1 --no-clocks --no-stage-timing
run s4_full --streams 4
run s4_none --streams 4 --no-clocks --no-stage-timing
run s4_full2 --streams 4
run s1_full_again --streams 1
