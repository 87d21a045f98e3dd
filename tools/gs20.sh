#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dropin.py -m gpu -q > gpurun_out/r2s_pytest_dropin.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest_dropin.log
tail -n 12 gpurun_out/r2s_pytest_dropin.log | cut -c1-300
timeout 600 python tools/run_config.py nial --ncell 200 200 200 --steps 60 --thermal 60 > gpurun_out/r2_cfg_cfg3_nial16M_N1.json 2> gpurun_out/r2_cfg_cfg3_nial16M_N1.err
tail -c 900 gpurun_out/r2_cfg_cfg3_nial16M_N1.json; tail -n 5 gpurun_out/r2_cfg_cfg3_nial16M_N1.err
