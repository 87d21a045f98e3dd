#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fixture or run_loop or skin" > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -3 gpurun_out/r2c_pytest.log
for v in $(ls imd_b200/variants | sed 's/libimd_b200_//; s/\.so//'); do
  IMDB200_LIB=$PWD/imd_b200/variants/libimd_b200_$v.so timeout 200 python bench.py --warmup 5 --steps 40 --no-cpu --no-equilibrium > gpurun_out/r2c_exp_$v.json 2> gpurun_out/r2c_exp_$v.err
  python - "$v" gpurun_out/r2c_exp_$v.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); p=d["phase_ms_per_step"]
    print(f"{sys.argv[1]:12s} step {d['ms_per_step']:.3f} pass1 {p['pass1_ms']:.3f} pass2 {p['pass2_ms']:.3f} rebuild {p['rebuild_ms']:.3f} nreb {d['config']['rebuilds_in_window']}")
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
