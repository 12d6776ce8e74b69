#!/bin/bash
# One GPU visit: headline bench line, ncu launch list of 2 V-cycles, ncu full capture of the fine-level GS kernel.
# usage (on the GPU box, from the repo root): bash tools/gpu_round.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
tail -c 3000 gpurun_out/${tag}_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --profile-range --no-cpu-baseline > gpurun_out/${tag}_ncu1.log 2>&1
python tools/ncu_launches.py gpurun_out/${tag}_launches.csv | tee gpurun_out/${tag}_launches_summary.txt | head -30
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_sell_gs -c 70 \
    -o gpurun_out/${tag}_gs_full -f python bench.py --steps 2 --warmup 3 --profile-range --no-cpu-baseline > gpurun_out/${tag}_ncu2.log 2>&1
ls -la gpurun_out/ | tail -12
