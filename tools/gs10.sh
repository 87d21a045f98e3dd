#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dropin.py -m gpu -q -x > gpurun_out/r2j_pytest_dropin.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest_dropin.log
tail -40 gpurun_out/r2j_pytest_dropin.log
