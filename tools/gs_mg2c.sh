#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_dropin.py -m gpu -q -x -k "mpi_binding" > gpurun_out/r2_pytest_mpi_binding.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_mpi_binding.log
tail -25 gpurun_out/r2_pytest_mpi_binding.log | cut -c1-300
bash tools/gs_cfg.sh
