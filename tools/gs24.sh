#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pass" -s 60 -c 2 -f -o gpurun_out/r2w_prof_nial python tools/run_config.py nial --ncell 126 126 126 --steps 4 --thermal 30 > gpurun_out/r2w_ncu.log 2>&1
tail -3 gpurun_out/r2w_ncu.log | cut -c1-200
ls -la gpurun_out/r2w_prof_nial.ncu-rep
