#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for nc in 40 50 63 80 100 126; do
  timeout 300 python bench.py --ncell $nc --warmup 5 --steps 40 --no-cpu --no-equilibrium > gpurun_out/r2f_nc$nc.json 2> gpurun_out/r2f_nc$nc.err
  python - "$nc" gpurun_out/r2f_nc$nc.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); p=d["phase_ms_per_step"]; n=d["config"]["global_atoms"]
    print(f"ncell {sys.argv[1]:4s} atoms {n:9d} step {d['ms_per_step']:.3f} ns/atom: pass1 {1e6*p['pass1_ms']/n:.4f} pass2 {1e6*p['pass2_ms']/n:.4f} rebuild/step {1e6*p['rebuild_ms']/n:.4f} nreb {d['config']['rebuilds_in_window']}")
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
