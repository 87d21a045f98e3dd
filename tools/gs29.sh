#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rA > gpurun_out/r2_pytest_1gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_1gpu.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2_pytest_1gpu.log | tail -5 | cut -c1-300
timeout 300 python __graft_entry__.py smoke nobuild 2>&1 | tail -1
