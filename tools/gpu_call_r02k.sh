#!/bin/bash
# 1 GPU: mixed configs at full size, then the ncu evidence
mkdir -p gpurun_out
T=r02k
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 300 $B --config darcy > gpurun_out/${T}_darcy.json 2> gpurun_out/${T}_darcy.err
timeout 300 $B --config spe10 > gpurun_out/${T}_spe10.json 2> gpurun_out/${T}_spe10.err
python - <<PY
import json
for v in ('darcy','spe10'):
    try:
        d=json.loads(open('gpurun_out/${T}_%s.json'%v).read().strip().splitlines()[-1])
        a=d['roofline']['all']
        print(v, round(d['ms_per_step'],3), '%.3g'%d['value'], 'e2e %.3g'%d['e2e']['value'], d['gpu_launches'], d['pcg'], round(d['setup_s']['total'],1), round(d['setup_s']['host_peak_rss_GB'],1))
        print('   ', d['config']['workload'][:160])
        print('   ', {k[:12]:(x['launches'],round(x['ms'],1),round(x['GBs'])) for k,x in a.items() if x['launches']})
    except Exception as e: print(v,'ERR',e)
PY
grep -hE "Error|error" gpurun_out/${T}_*.err | head -5
bash tools/gpu_call_ncu.sh r02 2>&1 | tail -45
