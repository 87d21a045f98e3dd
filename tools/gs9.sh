#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
IMDB200_LIB=$PWD/imd_b200/variants/libimd_b200_nogather.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pass" -s 0 -c 2 -o gpurun_out/r2i_prof_nogather -f python bench.py --thermal 0 --jitter 0.1 --steps 2 --warmup 1 --no-cpu --no-equilibrium > gpurun_out/r2i_ncu.log 2>&1
ls -la gpurun_out/r2i_prof_nogather.ncu-rep
