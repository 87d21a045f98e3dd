#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for L in 2 4; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-equilibrium --lanes $L > gpurun_out/r2y_lanes$L.json 2> gpurun_out/r2y_lanes$L.err
python - $L <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/r2y_lanes{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("lanes", sys.argv[1], f"{d['value']:.4e}", d['ms_per_step'], {k: round(v,3) for k,v in d['phase_ms_per_step'].items()})
except Exception as e: print("ERR", e); print(open(f"gpurun_out/r2y_lanes{sys.argv[1]}.err").read()[-800:])
PY
done
