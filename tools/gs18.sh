#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "not multi_gpu" > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
tail -n 30 gpurun_out/r2r_pytest.log | cut -c1-300
timeout 300 python tools/run_config.py nial --ncell 126 126 126 --steps 40 --thermal 40 > gpurun_out/r2r_nial4M.json 2> gpurun_out/r2r_nial4M.err
tail -c 700 gpurun_out/r2r_nial4M.json; tail -n 5 gpurun_out/r2r_nial4M.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2r_bench.json").read().strip().splitlines()[-1])
print("bench", f"{d['value']:.4e}", d['ms_per_step'], d['phase_ms_per_step'], d['e2e']['value'])
PY
