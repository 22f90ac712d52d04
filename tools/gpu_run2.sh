#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== profile ours"; timeout 600 python tools/profile_step.py ours c3 > gpurun_out/profile_ours.log 2>&1; head -60 gpurun_out/profile_ours.log
echo "== profile ref"; timeout 600 python tools/profile_step.py reference c3 > gpurun_out/profile_ref.log 2>&1; head -12 gpurun_out/profile_ref.log
