#!/bin/bash
# round 2, final tree: smoke, the whole GPU suite, the bench lines of both arms (c3 headline with e2e and the c5 block, the same
# with the loss heads, the 16k-Gaussian ManiGaussian call with and without heads), then the ncu evidence of the final kernels
mkdir -p gpurun_out
export MGS_NO_BUILD=1
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | cut -c1-400
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu_r2.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu_r2.log | cut -c1-300
show() { python - <<PY
import json
try:
    d=json.load(open('$1')); m=d.get('measured') or {}
    print('$1: value %.4g %s  ms/step %.3f  e2e %s  launches %s'%(d['value'],d['unit'],d['ms_per_step'],(d.get('e2e') or {}).get('value'),d.get('gpu_launches')))
    print('   stages',m.get('stage_ms_per_launch'),'pbv',m.get('project_bwd_views_ms'),'host',m.get('host_step_ms'))
    print('   roofline',d.get('roofline')); print('   c5',d.get('c5')); print('   e2e',{k:v for k,v in (d.get('e2e') or {}).items() if k!='note'})
except Exception as e: print('$1: no json', e)
PY
}
for w in c3 mg c2; do for impl in ours reference; do
  extra=""; [ "$w" != "c3" ] && extra="--no-c5"
  timeout 900 python bench.py --workload $w --impl $impl --no-cpu-baseline $extra > gpurun_out/r2_bench_${w}_${impl}.json 2> gpurun_out/r2_bench_${w}_${impl}.err; echo "bench $w $impl rc=$?"; grep -v Warning gpurun_out/r2_bench_${w}_${impl}.err | tail -3
  show gpurun_out/r2_bench_${w}_${impl}.json
  timeout 600 python bench.py --workload $w --impl $impl --no-cpu-baseline --no-c5 --heads --steps 20 --no-stage-timing > gpurun_out/r2_bench_${w}_heads_${impl}.json 2> gpurun_out/r2_bench_${w}_heads_${impl}.err; echo "bench $w heads $impl rc=$?"; grep -v Warning gpurun_out/r2_bench_${w}_heads_${impl}.err | tail -3
  show gpurun_out/r2_bench_${w}_heads_${impl}.json
done; done
if [ -n "$NCU" ]; then
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --settle 0 --no-e2e --no-cpu-baseline --no-stage-timing --no-c5 --no-clocks > gpurun_out/ncu_launch_r2.log 2>&1; echo rc=$?; wc -l gpurun_out/launches_r2.csv
echo "== ncu full (our kernels, one view)"; timeout 900 ncu --set full --clock-control none --import-source on -k "regex:blend|project|emit_tiles|ranges_pack|fill_tail|tile_order" -s 8 -c 8 -o gpurun_out/prof_all_r2 -f python bench.py --steps 2 --warmup 1 --settle 0 --streams 1 --no-e2e --no-cpu-baseline --no-c5 --no-clocks > gpurun_out/ncu_full_r2.log 2>&1; echo rc=$?; tail -2 gpurun_out/ncu_full_r2.log
fi
for impl in ours reference; do
  timeout 900 python bench.py --workload c4 --impl $impl --no-cpu-baseline --no-c5 > gpurun_out/r2_bench_c4_${impl}.json 2> gpurun_out/r2_bench_c4_${impl}.err; echo "bench c4 $impl rc=$?"; grep -v Warning gpurun_out/r2_bench_c4_${impl}.err | tail -3
  show gpurun_out/r2_bench_c4_${impl}.json
done
