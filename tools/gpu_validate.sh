#!/bin/bash
# round-end style validation on one GPU: tests, smoke, default bench of both arms, ncu evidence
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench default ours"; timeout 900 python bench.py > gpurun_out/bench_ours_c3.json 2> gpurun_out/bench_ours_c3.err; echo rc=$?; grep -v Warning gpurun_out/bench_ours_c3.err | tail -2
echo "== bench default reference"; timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_c3.json 2> gpurun_out/bench_ref_c3.err; echo rc=$?; grep -v Warning gpurun_out/bench_ref_c3.err | tail -2
for wlk in c1 c2 mg c4; do
  timeout 600 python bench.py --workload $wlk --no-cpu-baseline > gpurun_out/bench_ours_$wlk.json 2> gpurun_out/bench_ours_$wlk.err; timeout 600 python bench.py --workload $wlk --impl reference --no-cpu-baseline > gpurun_out/bench_ref_$wlk.json 2> gpurun_out/bench_ref_$wlk.err
done
python - <<'PY'
import json
for t in ('ours_c3','ref_c3','ours_c4','ref_c4','ours_c1','ref_c1','ours_c2','ref_c2','ours_mg','ref_mg'):
    try:
        d=json.load(open('gpurun_out/bench_%s.json'%t)); print(t,'value %.4g ms/step %.3f e2e %.4g'%(d['value'],d['ms_per_step'],d.get('e2e',{}).get('value',0)), 'roof', d.get('roofline',{}).get('kernel'), '%.3f'%d.get('roofline',{}).get('frac',0), d.get('clocks',{}).get('sm_mhz'), d.get('clocks',{}).get('reasons'))
    except Exception as ex: print(t,'ERR',ex)
PY
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --settle 0 --no-e2e --no-cpu-baseline --no-stage-timing > gpurun_out/ncu_launch.log 2>&1; echo rc=$?
echo "== ncu full blend"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:blend -s 8 -c 2 -o gpurun_out/prof_blend_r1_final -f python bench.py --steps 2 --warmup 1 --settle 0 --streams 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo rc=$?
