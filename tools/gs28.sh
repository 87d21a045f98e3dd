#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dropin.py tests/test_cabi.py -m gpu -q -rA > gpurun_out/r2z_pytest_dropin.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest_dropin.log
grep -E "passed|failed|FAILED|PASSED|rc=" gpurun_out/r2z_pytest_dropin.log | tail -14 | cut -c1-300
grep -B5 -A30 "^E  " gpurun_out/r2z_pytest_dropin.log | head -70 | cut -c1-250
