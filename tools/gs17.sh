#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "not multi_gpu" > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -n 30 gpurun_out/r2q_pytest.log | cut -c1-300
timeout 300 python tools/run_config.py nial --ncell 126 126 126 --steps 40 --thermal 40 > gpurun_out/r2q_nial4M.json 2> gpurun_out/r2q_nial4M.err
tail -c 700 gpurun_out/r2q_nial4M.json; tail -n 5 gpurun_out/r2q_nial4M.err
