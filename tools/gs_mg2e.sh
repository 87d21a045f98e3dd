#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -rA -k "two_species or two_domains_npt or two_domains_adp" > gpurun_out/r2_pytest_mgpu_${N}e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_mgpu_${N}e.log
grep -E "passed|failed|FAILED|PASSED|rc=|overlap_nial|^npt" gpurun_out/r2_pytest_mgpu_${N}e.log | tail -14 | cut -c1-300
grep -B2 -A12 "Error\|assert" gpurun_out/r2_pytest_mgpu_${N}e.log | head -60 | cut -c1-300
