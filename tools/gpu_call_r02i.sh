#!/bin/bash
# 1 GPU: one box of configs[4] (deformed) with the colour-ordered SELL kernel on fewer / more levels; the other configs
mkdir -p gpurun_out
T=r02i
B="python bench.py --steps 20 --warmup 3 --no-parity --no-cpu-baseline --deform"
timeout 330 $B > gpurun_out/${T}_deform.json 2> gpurun_out/${T}_deform.err
timeout 330 $B --sell-min-rows 3000000 > gpurun_out/${T}_deform_csr12.json 2> gpurun_out/${T}_deform_csr12.err
timeout 330 $B --sell-min-rows 1000000 > gpurun_out/${T}_deform_csr2.json 2> gpurun_out/${T}_deform_csr2.err
python - <<PY
import json
for v in ('deform','deform_csr12','deform_csr2'):
    try:
        d=json.loads(open('gpurun_out/${T}_%s.json'%v).read().strip().splitlines()[-1])
        a=d['roofline']['all']
        print(v, round(d['ms_per_step'],3), d['gpu_launches'], d['pcg']['iterations'], round(d['setup_s']['total'],1), round(d['setup_s']['batched_extension_stages']['h2d_s'],2), {k[:12]:(x['launches'],round(x['ms'],1),round(x['GBs'])) for k,x in a.items() if x['launches']})
        print('   levels', [(l['rows'],l['nnz']) for l in d['config']['levels']])
    except Exception as e: print(v,'ERR',e)
PY
grep -hE "Error|error" gpurun_out/${T}_*.err | head
