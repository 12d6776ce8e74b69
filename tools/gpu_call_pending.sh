#!/bin/bash
# What was written after round 2's GPU budget was spent and still has to be run on a B200.
# One GPU:   gpurun --timeout 900 -- 'bash tools/gpu_call_pending.sh'
# Eight GPUs (the multi-rank Darcy check): gpurun --gpus 8 --timeout 600 -- 'bash tools/gpu_call_pending.sh par8'
mkdir -p gpurun_out
if [ "$1" = "par8" ]; then
  timeout 500 python -m pytest tests/test_par_gpu.py -q --timeout 300 -rfEx -W ignore -k "darcy" 2>&1 | tail -40 > gpurun_out/pending_par8.log
  grep -E "passed|failed|xfailed|xpassed|FAILED|ERROR" gpurun_out/pending_par8.log | head
  exit 0
fi
# the full suite after the one-condition change in k_extension, the new files last
timeout 800 python -m pytest tests -m gpu -q -rfEx -W ignore 2>&1 | tail -40 > gpurun_out/pending_gpu_suite.log
grep -E "passed|failed|xfailed|xpassed|FAILED|ERROR" gpurun_out/pending_gpu_suite.log | head -20
# tensor-coefficient path of configs[3] on a synthetic file in the SPE10 format
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
from oracle import spe10
rng = np.random.default_rng(10)
kh = np.exp(rng.normal(0.0, 2.0, size=(85, 220, 60))); kz = kh * np.exp(rng.normal(-2.0, 1.0, size=kh.shape))
spe10.write_permeability_file("/tmp/spe_perm_synthetic.dat", np.stack([kh, kh, kz]))
PY
timeout 300 python bench.py --config spe10 --perm-file /tmp/spe_perm_synthetic.dat --steps 10 --warmup 3 --no-cpu-baseline \
  > gpurun_out/pending_spe10_tensor.json 2> gpurun_out/pending_spe10_tensor.err
tail -c 600 gpurun_out/pending_spe10_tensor.json
# the example driver on its three default parameter lists
for form in 0 1 2; do timeout 200 python examples/multigrid_test.py --form $form > gpurun_out/pending_example_form$form.log 2>&1; grep -E "Final residual|Error|error" gpurun_out/pending_example_form$form.log | head -3; done
