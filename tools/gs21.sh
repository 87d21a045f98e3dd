#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
IMDB200_DEBUG_REBUILD=1 timeout 600 python tools/run_config.py nial --ncell 200 200 200 --steps 30 --thermal 30 > gpurun_out/r2t_nial16M_dbg.json 2> gpurun_out/r2t_nial16M_dbg.err
grep -i "rebuild\|lap\|ms" gpurun_out/r2t_nial16M_dbg.err | tail -40 | cut -c1-250
