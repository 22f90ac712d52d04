#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== parity report"; timeout 900 python tools/parity_report.py > gpurun_out/parity_report.log 2>&1; echo rc=$?
for S in 1 4; do
echo "== bench ours c3 streams=$S"; timeout 600 python bench.py --steps 20 --warmup 5 --streams $S --no-e2e --no-cpu-baseline > gpurun_out/bench_ours_c3_s$S.json 2> gpurun_out/bench_ours_c3_s$S.err; echo "rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_ours_c3_s$S.json'));print('value %.4g  ms/step %.3f'%(d['value'],d['ms_per_step']), d['config'].get('stage_ms_per_launch'))"; tail -2 gpurun_out/bench_ours_c3_s$S.err
done
echo "== ncu full blend"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:blend -s 8 -c 2 -o gpurun_out/prof_blend_r1b -f python bench.py --steps 2 --warmup 1 --streams 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo rc=$?
