#!/bin/bash
# 4 GPUs: deformed and Darcy multi-rank parity at 4 ranks, bench N=4 on configs[4]
mkdir -p gpurun_out
T=r02n
timeout 300 python -m pytest tests/test_par_gpu.py -q --timeout 140 -rfE -W ignore -k "4-p2p-1 or darcy_matches_single_domain and 4" 2>&1 | tail -30 > gpurun_out/${T}_tests.log
timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/${T}_bench_n4.json 2> gpurun_out/${T}_bench_n4.err
grep -E "passed|failed|FAILED|ERROR|Timeout|Assertion" gpurun_out/${T}_tests.log | head -10
grep -E "PEError|Error|assert" gpurun_out/${T}_bench_n4.err | head -5
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${T}_bench_n4.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['parity'], d['setup_s']['total'], d['gpu_launches'], d['pcg'])
except Exception as e: print('bench ERR',e)
PY
