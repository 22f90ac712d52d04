export MGS_VARIANT=L1 MGS_NVCC_DEFINES="-DMGS_FWD_ROWS_LDGSTS=1"
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -3 | cut -c1-300
unset MGS_VARIANT MGS_NVCC_DEFINES
WORKLOADS="c3 c2" bash tools/gpu_variants.sh
