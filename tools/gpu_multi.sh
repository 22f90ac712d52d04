#!/bin/bash
# multi-GPU contract path: torchrun launch exactly as the driver does (short timeouts: a hang costs N x GPU-minutes)
N=${1:-2}
mkdir -p gpurun_out
echo "== view-parallel check N=$N"; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29510 tools/check_view_parallel.py > gpurun_out/view_parallel_$N.log 2>&1; grep -v Warning gpurun_out/view_parallel_$N.log | tail -2 | cut -c1-300
echo "== multigpu tests"; MGS_NO_BUILD=1 timeout 300 python -m pytest tests/test_multigpu_gpu.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -3 | cut -c1-300
echo "== ours N=$N"; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_ours_$N.json 2> gpurun_out/scale_ours_$N.err; echo rc=$?; grep -v Warning gpurun_out/scale_ours_$N.err | tail -3 | cut -c1-300
echo "== ref N=$N"; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_ref_$N.json 2> gpurun_out/scale_ref_$N.err; echo rc=$?; grep -v Warning gpurun_out/scale_ref_$N.err | tail -3 | cut -c1-300
python - <<PY
import json
for f in ('scale_ours_$N','scale_ref_$N'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f,'value %.4g ms/step %.3f e2e %.4g'%(d['value'],d['ms_per_step'],d.get('e2e',{}).get('value',0)), d['config'].get('parallelism'), d.get('clocks'))
        print('   c5', d.get('c5')); print('   nccl', d.get('nccl')); print('   host', (d.get('measured') or {}).get('host_step_ms'))
    except Exception as ex: print(f,'ERR',ex)
PY
