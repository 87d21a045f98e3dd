#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 1200 python -m pytest tests/test_multi_gpu.py tests/test_dropin.py -m gpu -q -rA > gpurun_out/r2_pytest_mgpu_${N}d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_mgpu_${N}d.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2_pytest_mgpu_${N}d.log | tail -8 | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29777 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2d_bench_${N}.json 2> gpurun_out/r2d_bench_${N}.err
tail -c 1200 gpurun_out/r2d_bench_${N}.json | head -c 900
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29788"
timeout 500 $L tools/run_config.py nial --ncell 200 200 200 --steps 60 --thermal 60 > gpurun_out/r2_cfg_cfg3_nial16M_N$N.json 2> gpurun_out/r2_cfg_cfg3_nial16M_N$N.err
tail -c 700 gpurun_out/r2_cfg_cfg3_nial16M_N$N.json | head -c 500
