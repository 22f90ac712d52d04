#!/bin/bash
mkdir -p gpurun_out
export MGS_NO_BUILD=1 MGS_VARIANT=ctalog MGS_NVCC_DEFINES=-DMGS_CTA_LOG
for wl in c3; do
  timeout 300 python tools/cta_timeline.py $wl > gpurun_out/cta_timeline_$wl.json 2> gpurun_out/cta_timeline_$wl.err || tail -5 gpurun_out/cta_timeline_$wl.err
  true || MGS_ONE_STREAM=1 timeout 300 python tools/cta_timeline.py $wl > gpurun_out/cta_timeline_${wl}_onestream.json 2>> gpurun_out/cta_timeline_$wl.err || tail -5 gpurun_out/cta_timeline_$wl.err
done
python - <<'PY'
import json
for n in ('c3','c3_onestream'):
    try:
        d=json.load(open(f'gpurun_out/cta_timeline_{n}.json'))
        print(n,'step_ms',d['step_ms'],'logged',d['ctas_logged'])
        for k in ('blend_fwd','blend_bwd'):
            x=d[k]; print(' ',k,{a:(round(b,1) if isinstance(b,float) else b) for a,b in x.items() if a not in('longest',)})
            print('   longest',x['longest'][:5])
    except Exception as e: print(n,'failed',e)
PY
