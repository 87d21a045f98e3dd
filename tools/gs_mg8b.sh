#!/bin/bash
# trimmed 8-GPU pass: test_eight_domains, one bench line (peer-memory halo, parity_check), configs 5 (strong) and 4 (16 M/GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -rA -k "eight_domains" > gpurun_out/r2_pytest_mgpu_${N}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_mgpu_${N}.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2_pytest_mgpu_${N}.log | tail -5 | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29777 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_${N}.json 2> gpurun_out/r2_bench_${N}.err
tail -c 1500 gpurun_out/r2_bench_${N}.json
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29788"
timeout 500 $L tools/run_config.py deform --ncell 400 200 200 --strong --steps 60 --thermal 50 > gpurun_out/r2_cfg_cfg5_deform64M_N$N.json 2> gpurun_out/r2_cfg_cfg5_deform64M_N$N.err
tail -c 600 gpurun_out/r2_cfg_cfg5_deform64M_N$N.json
timeout 500 $L tools/run_config.py cu --ncell 200 200 100 --steps 60 --thermal 60 > gpurun_out/r2_cfg_cfg4_cu16M_N$N.json 2> gpurun_out/r2_cfg_cfg4_cu16M_N$N.err
tail -c 600 gpurun_out/r2_cfg_cfg4_cu16M_N$N.json
