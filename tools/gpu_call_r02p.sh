#!/bin/bash
mkdir -p gpurun_out
timeout 280 python -m pytest tests/test_par_gpu.py -q --timeout 130 -rfE -W ignore -k "8-p2p-1 or (darcy and 8)" 2>&1 | tail -30 > gpurun_out/r02p_tests_n8.log
grep -E "passed|failed|FAILED|ERROR|Timeout|Assertion" gpurun_out/r02p_tests_n8.log | head -10
