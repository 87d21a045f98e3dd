#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in base $(ls imd_b200/variants | sed 's/libimd_b200_//; s/\.so//'); do
  lib=imd_b200/variants/libimd_b200_$v.so; [ $v = base ] && lib=imd_b200/libimd_b200.so
  IMDB200_LIB=$PWD/$lib timeout 200 python bench.py --warmup 5 --steps 40 --no-cpu --no-equilibrium > gpurun_out/r2o_exp_$v.json 2> gpurun_out/r2o_exp_$v.err
  python - "$v" gpurun_out/r2o_exp_$v.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); p=d["phase_ms_per_step"]
    print(f"{sys.argv[1]:12s} step {d['ms_per_step']:.3f} pass1 {p['pass1_ms']:.3f} pass2 {p['pass2_ms']:.3f} rebuild {p['rebuild_ms']:.3f} nreb {d['config']['rebuilds_in_window']}")
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
