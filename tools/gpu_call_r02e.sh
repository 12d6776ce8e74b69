#!/bin/bash
mkdir -p gpurun_out
T=r02e
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "sparray\|block_diag\|^$\|v1.20\|numpy arrays\|sparse" | tail -30 > gpurun_out/${T}_tests.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err
tail -n 12 gpurun_out/${T}_tests.log
grep -E "Error|error|assert" gpurun_out/${T}_bench_n1.err gpurun_out/${T}_bench_n2.err | head -20
python - <<'PY'
import json
for f in ('gpurun_out/r02e_bench_n1.json','gpurun_out/r02e_bench_n2.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['parity'], d['setup_s']['total'], d['gpu_launches'], d['pcg'])
        print({k:(v['launches'],round(v['ms'],2),round(v['GBs'])) for k,v in d['roofline']['all'].items()})
    except Exception as e: print(f,'ERR',e)
PY
