#!/bin/bash
mkdir -p gpurun_out
T=r02m
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config darcy > gpurun_out/${T}_darcy.json 2> gpurun_out/${T}_darcy.err
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config spe10 > gpurun_out/${T}_spe10.json 2> gpurun_out/${T}_spe10.err
python - <<PY
import json
for v in ('darcy','spe10'):
    try:
        d=json.loads(open('gpurun_out/${T}_%s.json'%v).read().strip().splitlines()[-1])
        print(v, round(d['ms_per_step'],3), '%.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], d.get('gpu_launches'), d.get('pcg'), d.get('setup_s',{}).get('total'))
    except Exception as e: print(v,'ERR',e)
PY
grep -hE "Error|error" gpurun_out/${T}_*.err | head -5
