#!/bin/bash
mkdir -p gpurun_out
T=r02d
timeout 1200 python -m pytest tests/test_deformed_gpu.py tests/test_solver_gpu.py tests/test_coarsen_gpu.py tests/test_kernels_gpu.py tests/test_par_gpu.py -q 2>&1 | grep -v "Warning\|warn\|sparray\|block_diag\|^$\|sparse\|v1.20" | tail -40 > gpurun_out/${T}_tests.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/${T}_bench_ref_n2.json 2> gpurun_out/${T}_bench_ref_n2.err
tail -n 30 gpurun_out/${T}_tests.log
grep -E "Error|error|assert" gpurun_out/${T}_bench_n2.err gpurun_out/${T}_bench_ref_n2.err | head -20
