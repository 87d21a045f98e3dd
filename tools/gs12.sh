#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -4 gpurun_out/r2l_pytest.log
bash tools/gs_cfg.sh smoke 2>&1 | grep nial
timeout 600 python tools/run_config.py nial --ncell 200 200 200 --steps 60 --thermal 60 > gpurun_out/r2l_cfg3.json 2> gpurun_out/r2l_cfg3.err; cat gpurun_out/r2l_cfg3.json | cut -c1-600
