#!/bin/bash
# compute-sanitizer over a subset of the parity / render suites on the final kernels (memcheck, then racecheck on the blends' shared memory)
mkdir -p gpurun_out
export MGS_NO_BUILD=1
{
echo "# compute-sanitizer on the GPU box (B200), final round-2 kernels (longest-list-first launch order, product-form transmittance, per-SH-size project_bwd_views)"
echo '# command: compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_parity_gpu.py tests/test_render_gpu.py -m gpu -q -x -k "tiny_f3 or f16 or tile_launch_order or loss_heads_fused or opaque"'
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_parity_gpu.py tests/test_render_gpu.py -m gpu -q -x -p no:cacheprovider -k "tiny_f3 or f16 or tile_launch_order or loss_heads_fused or opaque" 2>&1 | grep -v Warning | tail -8
echo '# command: compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "tiny_f3 or f16"'
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -p no:cacheprovider -k "tiny_f3 or f16" 2>&1 | grep -v Warning | tail -8
} > gpurun_out/r2_compute_sanitizer.txt 2>&1
cat gpurun_out/r2_compute_sanitizer.txt | cut -c1-300
