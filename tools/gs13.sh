#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2m_launches_nial.csv python tools/run_config.py nial --ncell 126 126 126 --steps 12 --thermal 12 --warmup 2 > gpurun_out/r2m_nial.log 2>&1
python tools/launch_summary.py gpurun_out/r2m_launches_nial.csv | head -30
