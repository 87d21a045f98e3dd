#!/bin/bash
# multi-GPU session: pytest tests/test_multi_gpu.py + bench at N GPUs (N = number of visible GPUs)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -q -rA > gpurun_out/r2_pytest_mgpu_$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_mgpu_$N.log
grep -E "passed|failed|PASSED|FAILED|XPASS|XFAIL|SKIPPED|rc=" gpurun_out/r2_pytest_mgpu_$N.log | tail -40
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29777 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_$N.json 2> gpurun_out/r2_bench_$N.err
python - $N <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2_bench_{n}.json").read().strip().splitlines()[-1])
    print("bench", n, d["value"], d["ms_per_step"], d["phase_ms_per_step"], d.get("parity_check"), d.get("equilibrium_window"))
except Exception as e: print("ERR", e); print(open(f"gpurun_out/r2_bench_{n}.err").read()[-2000:])
PY
