#!/bin/bash
cd "$(dirname "$0")/.."
IMDB200_DEBUG_REBUILD=1 timeout 600 python tools/run_config.py nial --ncell 200 200 200 --steps 40 --thermal 60 --warmup 2 > gpurun_out/r2n_nial.json 2> gpurun_out/r2n_nial.err
grep rebuild gpurun_out/r2n_nial.err | awk '$NF=="ms" && $(NF-1) > 5.0' | head -60
cut -c1-400 gpurun_out/r2n_nial.json
