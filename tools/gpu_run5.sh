#!/bin/bash
mkdir -p gpurun_out
echo "== profile ours"; timeout 600 python tools/profile_step.py ours c3 > gpurun_out/profile_ours.log 2>&1; head -12 gpurun_out/profile_ours.log | cut -c1-700
