#!/bin/bash
# First-light GPU run: smoke, goldens from the compiled reference, parity tests, short bench of both arms.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== golden"; timeout 300 python tests/golden/make_golden.py gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "golden rc=$?"; tail -6 gpurun_out/golden.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
echo "== bench ours c3"; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ours_c3.json 2> gpurun_out/bench_ours_c3.err; echo "rc=$?"; cat gpurun_out/bench_ours_c3.json; tail -3 gpurun_out/bench_ours_c3.err
echo "== bench ref c3"; timeout 600 python bench.py --steps 5 --warmup 3 --impl reference --no-cpu-baseline > gpurun_out/bench_ref_c3.json 2> gpurun_out/bench_ref_c3.err; echo "rc=$?"; cat gpurun_out/bench_ref_c3.json; tail -3 gpurun_out/bench_ref_c3.err
