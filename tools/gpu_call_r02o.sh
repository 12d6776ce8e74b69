#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
timeout 240 python -m pytest tests/test_par_gpu.py -q --timeout 200 -rfE -W ignore -k "${N}-p2p-1" 2>&1 | tail -30 > gpurun_out/r02o_tests_n${N}.log
grep -E "passed|failed|FAILED|ERROR|Timeout|Assertion" gpurun_out/r02o_tests_n${N}.log | head -10
