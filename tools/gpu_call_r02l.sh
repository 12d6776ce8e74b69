#!/bin/bash
# 1 GPU: full single-GPU test suite, headline bench (with the weak-scaling base leg), reference arm, mixed configs with Block LDU
mkdir -p gpurun_out
T=r02l
timeout 480 python -m pytest tests -m gpu -q --timeout 150 -rfE -W ignore 2>&1 | tail -40 > gpurun_out/${T}_tests.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config darcy > gpurun_out/${T}_darcy.json 2> gpurun_out/${T}_darcy.err
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config spe10 > gpurun_out/${T}_spe10.json 2> gpurun_out/${T}_spe10.err
grep -E "passed|failed|FAILED|ERROR|Timeout" gpurun_out/${T}_tests.log | head -20
python - <<PY
import json
for v in ('bench_n1','bench_ref','darcy','spe10'):
    try:
        d=json.loads(open('gpurun_out/${T}_%s.json'%v).read().strip().splitlines()[-1])
        print(v, round(d['ms_per_step'],3), '%.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], d.get('gpu_launches'), d.get('pcg'), d.get('setup_s',{}).get('total'))
        if v=='bench_n1': print('  roofline', d['roofline']['frac'], d['roofline']['traffic'], 'parity', d['parity'], '\n  weak', d['weak_scaling_base'], '\n  cpu', d['cpu_baseline'])
        if v=='bench_ref': print('  ', d['cpu_baseline'])
    except Exception as e: print(v,'ERR',e)
PY
grep -hE "Error|error" gpurun_out/${T}_*.err | head -5
