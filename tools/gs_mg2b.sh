#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_dropin.py -m gpu -q -rA -x -k "overlap or nccl or migrate or two_domains_match_reference_fixture or mpi_binding or large" > gpurun_out/r2_pytest_mgpu_${N}b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_mgpu_${N}b.log
grep -E "passed|failed|FAILED|XPASS|XFAIL|rc=|Error|error" gpurun_out/r2_pytest_mgpu_${N}b.log | tail -15
for mode in p2p nccl; do
  export IMDB200_HALO_P2P=$([ $mode = p2p ] && echo 1 || echo 0)
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29777 bench.py --gpus $N --steps 40 --warmup 5 --no-parity > gpurun_out/r2b_bench_${N}_$mode.json 2> gpurun_out/r2b_bench_${N}_$mode.err
  python - $N $mode <<'PY'
import json,sys
n,mode=sys.argv[1:3]
try:
    d=json.loads(open(f"gpurun_out/r2b_bench_{n}_{mode}.json").read().strip().splitlines()[-1])
    print("bench", n, mode, f"{d['value']:.4e}", f"{d['ms_per_step']:.3f} ms", {k: round(v,3) for k,v in d["phase_ms_per_step"].items()}, "eq", (d.get("equilibrium_window") or {}).get("ms_per_step"))
except Exception as e: print("ERR", e); print(open(f"gpurun_out/r2b_bench_{n}_{mode}.err").read()[-2000:])
PY
done
