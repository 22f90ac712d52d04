#!/bin/bash
mkdir -p gpurun_out
echo "== parity report"; timeout 900 python tools/parity_report.py > gpurun_out/parity_report.log 2>&1; echo rc=$?; tail -3 gpurun_out/parity_report.log | cut -c1-400
echo "== bench ours c3 K=20"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ours_c3.json 2> gpurun_out/bench_ours_c3.err; echo "rc=$?"; cut -c1-900 gpurun_out/bench_ours_c3.json; tail -3 gpurun_out/bench_ours_c3.err
echo "== profile ours"; timeout 600 python tools/profile_step.py ours c3 > gpurun_out/profile_ours.log 2>&1; head -4 gpurun_out/profile_ours.log
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo rc=$?
echo "== ncu full blend"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:blend -s 8 -c 2 -o gpurun_out/prof_blend_r1 -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo rc=$?; ls -la gpurun_out/*.ncu-rep
