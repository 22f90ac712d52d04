timeout 600 python -m pytest tests/test_render_gpu.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -5 | cut -c1-300
timeout 600 python bench.py --workload c4 --no-cpu-baseline --no-e2e > gpurun_out/bench_c4_ours.json 2> gpurun_out/bench_c4_ours.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_ours.json')); print('%.4g %.3f'%(d['value'],d['ms_per_step'])); print(d['roofline']); print(d['config']['stage_ms_per_launch'])"
