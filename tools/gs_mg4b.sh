#!/bin/bash
# 4-GPU pass with the final code: four-domain tests, one bench line, configs 5 (strong), 4 (16 M/GPU) and 3 (16 M/GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -rA -k "four_domains" > gpurun_out/r2_pytest_mgpu_${N}b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_mgpu_${N}b.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2_pytest_mgpu_${N}b.log | tail -5 | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29777 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2e_bench_${N}.json 2> gpurun_out/r2e_bench_${N}.err
tail -c 1200 gpurun_out/r2e_bench_${N}.json | head -c 500
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29788"
timeout 500 $L tools/run_config.py deform --ncell 400 200 200 --strong --steps 60 --thermal 50 > gpurun_out/r2_cfg_cfg5_deform64M_N$N.json 2> gpurun_out/r2_cfg_cfg5_deform64M_N$N.err
tail -c 600 gpurun_out/r2_cfg_cfg5_deform64M_N$N.json | head -c 300
timeout 500 $L tools/run_config.py cu --ncell 200 200 100 --steps 60 --thermal 60 > gpurun_out/r2_cfg_cfg4_cu16M_N$N.json 2> gpurun_out/r2_cfg_cfg4_cu16M_N$N.err
tail -c 600 gpurun_out/r2_cfg_cfg4_cu16M_N$N.json | head -c 300
timeout 500 $L tools/run_config.py nial --ncell 200 200 200 --steps 60 --thermal 100 > gpurun_out/r2_cfg_cfg3_nial16M_N$N.json 2> gpurun_out/r2_cfg_cfg3_nial16M_N$N.err
tail -c 600 gpurun_out/r2_cfg_cfg3_nial16M_N$N.json | head -c 300
