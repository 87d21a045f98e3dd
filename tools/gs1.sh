#!/bin/bash
# GPU session 1 (round 2): tests, driver-like bench, broadcast/table experiments
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpus.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
for v in base bc2 bc4 kc bc2kc; do
  lib=imd_b200/variants/libimd_b200_$v.so; [ $v = base ] && lib=imd_b200/libimd_b200.so
  IMDB200_LIB=$PWD/$lib timeout 200 python bench.py --thermal 0 --jitter 0.1 --warmup 2 --steps 8 --no-cpu --no-equilibrium > gpurun_out/r2a_exp_$v.json 2> gpurun_out/r2a_exp_$v.err
done
tail -3 gpurun_out/r2a_pytest.log
for f in gpurun_out/r2a_bench.json gpurun_out/r2a_exp_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","phase_ms_per_step")}, d.get("e2e",{}).get("value"), d.get("e2e",{}).get("job_seconds"), d.get("equilibrium_window"))
except Exception as e: print("ERR",e)
PY
done
