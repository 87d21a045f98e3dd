#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -3 gpurun_out/r2d_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2d_bench.json").read().strip().splitlines()[-1]); p=d["phase_ms_per_step"]
print("bench", d["ms_per_step"], p, d["e2e"]["job_seconds"], d.get("equilibrium_window"))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pass" -s 324 -c 2 -o gpurun_out/r2d_prof -f python bench.py --steps 3 --warmup 10 --no-cpu --no-equilibrium > gpurun_out/r2d_ncu.log 2>&1
ls -la gpurun_out/r2d_prof.ncu-rep
