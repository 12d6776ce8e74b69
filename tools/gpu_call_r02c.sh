#!/bin/bash
# round-2 GPU visit c (2 GPUs): deformed multi-rank parity (global-workspace extension kernel), smoke, bench N=2 deformed, reference arms
mkdir -p gpurun_out
T=r02c
timeout 900 python -m pytest tests/test_par_gpu.py tests/test_deformed_gpu.py tests/test_coarsen_gpu.py -q 2>&1 | tail -25 > gpurun_out/${T}_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/${T}_bench_ref_n2.json 2> gpurun_out/${T}_bench_ref_n2.err
tail -n 5 gpurun_out/${T}_tests.log gpurun_out/${T}_smoke.log
tail -c 1500 gpurun_out/${T}_bench_n2.err gpurun_out/${T}_bench_ref.err gpurun_out/${T}_bench_ref_n2.err
