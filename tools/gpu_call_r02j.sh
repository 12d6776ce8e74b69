#!/bin/bash
# 1 GPU: the other BASELINE configs at full size (bench.py --config ...), strict timeouts
mkdir -p gpurun_out
T=r02j
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 300 $B --config cfg1 > gpurun_out/${T}_cfg1.json 2> gpurun_out/${T}_cfg1.err
timeout 300 $B --config darcy > gpurun_out/${T}_darcy.json 2> gpurun_out/${T}_darcy.err
timeout 300 $B --config spe10 > gpurun_out/${T}_spe10.json 2> gpurun_out/${T}_spe10.err
timeout 480 $B --config hcurl > gpurun_out/${T}_hcurl.json 2> gpurun_out/${T}_hcurl.err
python - <<PY
import json
for v in ('cfg1','darcy','spe10','hcurl'):
    try:
        d=json.loads(open('gpurun_out/${T}_%s.json'%v).read().strip().splitlines()[-1])
        a=d['roofline']['all']
        print(v, round(d['ms_per_step'],3), '%.3g'%d['value'], 'e2e %.3g'%d['e2e']['value'], d['gpu_launches'], d['pcg'], round(d['setup_s']['total'],1), round(d['setup_s']['host_peak_rss_GB'],1))
        print('   ', d['config']['workload'][:160])
        print('   ', {k[:12]:(x['launches'],round(x['ms'],1),round(x['GBs'])) for k,x in a.items() if x['launches']}, d['spmv_fine_operator'])
    except Exception as e: print(v,'ERR',e)
PY
grep -hE "Error|error" gpurun_out/${T}_*.err | head
