#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -k "npt" > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log
grep -E "passed|failed|FAILED|PASSED|rc=" gpurun_out/r2x_pytest.log | tail -10 | cut -c1-300
grep -B5 -A25 "Error\|assert " gpurun_out/r2x_pytest.log | head -60 | cut -c1-250
