#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -k "set_momenta or run_loop" > gpurun_out/r2zz_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2zz_pytest.log
grep -E "passed|failed|FAILED|PASSED|rc=" gpurun_out/r2zz_pytest.log | tail -8 | cut -c1-300
grep -B5 -A25 "^E  " gpurun_out/r2zz_pytest.log | head -50 | cut -c1-250
