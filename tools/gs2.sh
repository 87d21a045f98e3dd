#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large or parity_fixture" > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
for v in $(ls imd_b200/variants | sed 's/libimd_b200_//; s/\.so//'); do
  IMDB200_LIB=$PWD/imd_b200/variants/libimd_b200_$v.so timeout 200 python bench.py --thermal 0 --jitter 0.1 --warmup 2 --steps 8 --no-cpu --no-equilibrium > gpurun_out/r2b_exp_$v.json 2> gpurun_out/r2b_exp_$v.err
  python - "$v" gpurun_out/r2b_exp_$v.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); p=d["phase_ms_per_step"]
    print(f"{sys.argv[1]:12s} pass1 {p['pass1_ms']:.3f} pass2 {p['pass2_ms']:.3f} rebuild {p['rebuild_ms']:.3f}")
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
