#!/bin/bash
# round-2 evidence: ncu launch list of the default bench command + one --set full capture of every hand-written kernel of one
# c3 view (single stream, so each kernel runs alone).  MGS_NO_BUILD: ncu follows child processes; no compiler may be spawned.
export MGS_NO_BUILD=1
mkdir -p gpurun_out
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --settle 0 --no-e2e --no-cpu-baseline --no-stage-timing --no-c5 --no-clocks > gpurun_out/ncu_launch_r2.log 2>&1; echo rc=$?; wc -l gpurun_out/launches_r2.csv
echo "== ncu full (our kernels, one view)"; timeout 900 ncu --set full --clock-control none --import-source on -k "regex:blend|project|emit_tiles|ranges_pack|fill_tail" -s 7 -c 7 -o gpurun_out/prof_all_r2 -f python bench.py --steps 2 --warmup 1 --settle 0 --streams 1 --no-e2e --no-cpu-baseline --no-c5 --no-clocks > gpurun_out/ncu_full_r2.log 2>&1; echo rc=$?; tail -2 gpurun_out/ncu_full_r2.log
