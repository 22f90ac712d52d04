#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python bench.py --no-e2e --no-cpu-baseline "$@" > gpurun_out/b_$tag.json 2> gpurun_out/b_$tag.err; python -c "
import json;d=json.load(open('gpurun_out/b_$tag.json'));h=d['config'].get('host_step_ms');print('$tag: ms/step %.3f wall %.3f settle %s max %.1f med %.2f'%(d['ms_per_step'],d['config']['wall_ms_per_step'],d['config'].get('settle_steps'),max(h),sorted(h)[len(h)//2]), [x for x in h if x>1.5*sorted(h)[len(h)//2]])"; tail -2 gpurun_out/b_$tag.err; }
run s1_a --streams 1 --steps 100 --warmup 5
run s4_a --streams 4 --steps 100 --warmup 5
run s1_b --streams 1 --steps 100 --warmup 5 --settle 0
run s4_b --streams 4 --steps 100 --warmup 5 --settle 0
run s4_c --streams 4
run ref --impl reference --steps 20 --warmup 5
