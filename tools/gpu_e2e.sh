#!/bin/bash
# new-layer GPU tests, then value + e2e of both arms on the four single-GPU workloads (no CPU baseline, no stage timing)
timeout 600 python -m pytest tests/test_render_gpu.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -4 | cut -c1-300
for spec in "mg ours" "mg reference" "c1 ours" "c1 reference" "c2 ours" "c2 reference" "c3 ours" "c3 reference"; do set -- $spec
  timeout 600 python bench.py --workload $1 --impl $2 --no-cpu-baseline --no-stage-timing > gpurun_out/e2e_$1_$2.json 2> gpurun_out/e2e_$1_$2.err || tail -3 gpurun_out/e2e_$1_$2.err
  python -c "
import json; d=json.load(open('gpurun_out/e2e_$1_$2.json')); print('$1 $2 value %.4g (%.3f ms) e2e %.4g (%.3f ms)'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step']))"
done
