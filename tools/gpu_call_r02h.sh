#!/bin/bash
# 2 GPUs, strict timeouts: multi-rank parity (both halo paths, deformed boxes, Darcy), bench N=2 on configs[4]
mkdir -p gpurun_out
T=r02h
timeout 540 python -m pytest tests/test_par_gpu.py -q --timeout 170 -rfE -W ignore 2>&1 | tail -60 > gpurun_out/${T}_tests.log
timeout 360 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err
grep -E "passed|failed|FAILED|ERROR|Timeout|PEError|Assertion" gpurun_out/${T}_tests.log | head -30
grep -E "PEError|Error|assert" gpurun_out/${T}_bench_n2.err | head -10
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${T}_bench_n2.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['parity'], d['setup_s']['total'], d['gpu_launches'], d['pcg'])
    print({k[:12]:(v['launches'],round(v['ms'],2),round(v['GBs'])) for k,v in d['roofline']['all'].items()})
except Exception as e: print('bench ERR',e)
PY
