#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_build_nbl2 -s 1 -c 1 -f -o gpurun_out/r2v_prof_build python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2v_ncu.log 2>&1
tail -3 gpurun_out/r2v_ncu.log | cut -c1-200
ls -la gpurun_out/r2v_prof_build.ncu-rep
