#!/bin/bash
# fused-sweep variants at the headline size (1 GPU): off / plain launch / cooperative launch; slab-major order
mkdir -p gpurun_out
T=r02g
B="python bench.py --steps 20 --warmup 3 --no-parity --no-cpu-baseline"
timeout 200 $B --fused-gs-mb 0 > gpurun_out/${T}_fused0.json 2> gpurun_out/${T}_fused0.err
timeout 200 $B > gpurun_out/${T}_plain.json 2> gpurun_out/${T}_plain.err
PE_FUSED_COOP=1 timeout 200 $B > gpurun_out/${T}_coop.json 2> gpurun_out/${T}_coop.err
timeout 200 $B --fused-gs-mb 40 > gpurun_out/${T}_plain40.json 2> gpurun_out/${T}_plain40.err
timeout 240 $B --gs-slabs 4 > gpurun_out/${T}_slabs4.json 2> gpurun_out/${T}_slabs4.err
timeout 240 $B --gs-slabs 8 > gpurun_out/${T}_slabs8.json 2> gpurun_out/${T}_slabs8.err
python - <<PY
import json
for v in ('fused0','plain','coop','plain40','slabs4','slabs8'):
    try:
        d=json.loads(open('gpurun_out/${T}_%s.json'%v).read().strip().splitlines()[-1])
        a=d['roofline']['all']
        print(v, round(d['ms_per_step'],3), d['gpu_launches'], d['pcg']['iterations'], round(d['setup_s']['build_solver'],1), {k[:12]:(x['launches'],round(x['ms'],1),round(x['GBs'])) for k,x in a.items() if x['launches']})
    except Exception as e: print(v,'ERR',e)
PY
grep -hE "Error|error" gpurun_out/${T}_*.err | head
