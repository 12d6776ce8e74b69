#!/bin/bash
# One GPU visit: headline bench line, ncu launch list of 2 V-cycles, ncu full capture of the fine-level GS kernel.
# usage (on the GPU box, from the repo root): bash tools/gpu_round.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
cat gpurun_out/${tag}_bench_reference.json
python -c "
import json,sys
d=json.load(open('gpurun_out/${tag}_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e')}, d['roofline']['frac'], d['setup_s']['total'], d['setup_s']['host_peak_rss_GB'])"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --profile-range --no-cpu-baseline > gpurun_out/${tag}_ncu1.log 2>&1
python tools/ncu_launches.py gpurun_out/${tag}_launches.csv | tee gpurun_out/${tag}_launches_summary.txt | head -24
# full capture of a few fine-level colour launches (no source import: the report has to stay small)
ncu --set full --clock-control none --profile-from-start off -k regex:k_sell_gs -c 24 \
    -o gpurun_out/${tag}_gs_full -f python bench.py --steps 2 --warmup 3 --profile-range --no-cpu-baseline > gpurun_out/${tag}_ncu2.log 2>&1
ncu -i gpurun_out/${tag}_gs_full.ncu-rep --page raw --csv > gpurun_out/${tag}_gs_full_raw.csv 2>/dev/null
ls -la gpurun_out/${tag}_gs_full.ncu-rep
sz=$(stat -c %s gpurun_out/${tag}_gs_full.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 40000000 ]; then rm -f gpurun_out/${tag}_gs_full.ncu-rep; echo "report dropped (too large), raw csv kept"; fi
du -sh gpurun_out
