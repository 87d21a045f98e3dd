#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
K="eight_domains"; [ $N -eq 4 ] && K="four_domains"
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -rA -k "$K" > gpurun_out/r2_pytest_mgpu_${N}c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_mgpu_${N}c.log
grep -E "passed|failed|FAILED|rc=|overlap:|migration:|send_forces:" gpurun_out/r2_pytest_mgpu_${N}c.log | tail -12 | cut -c1-300
for mode in p2p nccl; do
  export IMDB200_HALO_P2P=$([ $mode = p2p ] && echo 1 || echo 0)
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29777 bench.py --gpus $N --steps 20 --warmup 5 $([ $mode = nccl ] && echo --no-parity) > gpurun_out/r2c_bench_${N}_$mode.json 2> gpurun_out/r2c_bench_${N}_$mode.err
  python - $N $mode <<'PY'
import json,sys
n,mode=sys.argv[1:3]
try:
    d=json.loads(open(f"gpurun_out/r2c_bench_{n}_{mode}.json").read().strip().splitlines()[-1])
    pc=d.get("parity_check") or {}
    print("bench", n, mode, f"{d['value']:.4e}", f"{d['ms_per_step']:.3f} ms", {k: round(v,3) for k,v in d["phase_ms_per_step"].items()}, "reb", d["config"]["rebuilds_in_window"], "parity", pc.get("ok"), pc.get("max_rel_err"), "eq", (d.get("equilibrium_window") or {}).get("ms_per_step"))
except Exception as e: print("ERR", e); print(open(f"gpurun_out/r2c_bench_{n}_{mode}.err").read()[-2000:])
PY
done
unset IMDB200_HALO_P2P
bash tools/gs_cfg.sh
