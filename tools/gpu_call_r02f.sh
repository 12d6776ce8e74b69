#!/bin/bash
# 1 GPU, strict timeouts: all single-GPU tests (per-test timeout), then the headline bench
mkdir -p gpurun_out
T=${1:-r02f}
timeout 480 python -m pytest tests -m gpu -q --timeout 120 -rfE -W ignore 2>&1 | tail -60 > gpurun_out/${T}_tests.log
timeout 240 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
grep -E "passed|failed|FAILED|ERROR|Timeout" gpurun_out/${T}_tests.log | head -30
grep -E "Error|error|assert" gpurun_out/${T}_bench_n1.err | head -10
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${T}_bench_n1.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['parity'], d['setup_s']['total'], d['gpu_launches'], d['pcg'])
    print({k:(v['launches'],round(v['ms'],2),round(v['GBs'])) for k,v in d['roofline']['all'].items()})
except Exception as e: print('bench ERR',e)
PY
