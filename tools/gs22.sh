#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "not multi_gpu and not dropin" > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log
tail -n 8 gpurun_out/r2u_pytest.log | cut -c1-300
IMDB200_DEBUG_REBUILD=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
grep "list build" gpurun_out/r2u_bench.err | tail -4
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2u_bench.json").read().strip().splitlines()[-1])
print("bench", f"{d['value']:.4e}", d['ms_per_step'], d['phase_ms_per_step'], "eq", d['equilibrium_window']['ms_per_step'])
PY
IMDB200_DEBUG_REBUILD=1 timeout 300 python tools/run_config.py nial --ncell 126 126 126 --steps 40 --thermal 40 > gpurun_out/r2u_nial4M.json 2> gpurun_out/r2u_nial4M.err
grep "list build" gpurun_out/r2u_nial4M.err | tail -3
tail -c 600 gpurun_out/r2u_nial4M.json | head -c 450
