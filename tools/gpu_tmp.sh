export MGS_VARIANT=E1 MGS_NVCC_DEFINES="-DMGS_CULL_ELLIPSE=1"
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -8 | cut -c1-300
unset MGS_VARIANT MGS_NVCC_DEFINES
WORKLOADS="c3 c2" bash tools/gpu_variants.sh
