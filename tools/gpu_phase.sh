#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/phase_times.py c3 2>&1 | grep -v Warning | tail -6
