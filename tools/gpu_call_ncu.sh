#!/bin/bash
# ncu evidence of round 2 (1 GPU): launch list of 2 V-cycles at the headline size; full captures of the fine-level SELL kernels
# (k_sell_gs, k_sell_spmv), the fused sweep kernels, the SpGEMM kernels (Galerkin product) and the batched local kernels
# (k_extension, k_traces).  Setup kernels are captured at 64^3 (ncu replays every kernel ~40 times).
T=${1:-r02}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv \
    --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --profile-range --no-cpu-baseline --no-parity > gpurun_out/${T}_ncu_launch.log 2>&1
python tools/ncu_launches.py gpurun_out/${T}_launches.csv > gpurun_out/${T}_launches_summary.txt
head -20 gpurun_out/${T}_launches_summary.txt
for K in k_sell_gs k_sell_spmv; do
  timeout 500 ncu --set full --clock-control none --profile-from-start off -k regex:^${K}\$ -c 12 -o gpurun_out/${T}_${K} -f \
      python bench.py --steps 2 --warmup 3 --profile-range --no-cpu-baseline --no-parity > gpurun_out/${T}_ncu_${K}.log 2>&1
done
timeout 500 ncu --set full --clock-control none --profile-from-start off -k regex:sweep -c 8 -o gpurun_out/${T}_fused_sweep -f \
      python bench.py --steps 2 --warmup 3 --profile-range --no-cpu-baseline --no-parity > gpurun_out/${T}_ncu_fused.log 2>&1
# setup kernels at 64^3, 3 levels
timeout 500 ncu --set full --clock-control none -k regex:k_extension -c 6 -o gpurun_out/${T}_k_extension -f python tools/setup_only.py 64 3 > gpurun_out/${T}_ncu_ext.log 2>&1
timeout 500 ncu --set full --clock-control none -k regex:k_traces -c 4 -o gpurun_out/${T}_k_traces -f python tools/setup_only.py 64 3 > gpurun_out/${T}_ncu_tr.log 2>&1
timeout 500 ncu --set full --clock-control none -k regex:k_spgemm -c 12 -o gpurun_out/${T}_k_spgemm -f python tools/setup_only.py 64 3 > gpurun_out/${T}_ncu_spgemm.log 2>&1
for R in k_sell_gs k_sell_spmv fused_sweep k_extension k_traces k_spgemm; do
  if [ -f gpurun_out/${T}_${R}.ncu-rep ]; then
    ncu -i gpurun_out/${T}_${R}.ncu-rep --page raw --csv > gpurun_out/${T}_${R}_raw.csv 2>/dev/null
    python tools/ncu_summary.py gpurun_out/${T}_${R}_raw.csv > gpurun_out/${T}_${R}_summary.csv
    sz=$(stat -c %s gpurun_out/${T}_${R}.ncu-rep); if [ "$sz" -gt 8000000 ]; then rm -f gpurun_out/${T}_${R}.ncu-rep; fi
    rm -f gpurun_out/${T}_${R}_raw.csv.tmp
  fi
done
du -sh gpurun_out; ls -la gpurun_out | grep ${T}_ | head -40
