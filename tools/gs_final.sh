#!/bin/bash
# final single-GPU evidence pass of the round: tests, smoke, bench (both arms), launch list, config 3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rA > gpurun_out/r2_pytest_1gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_1gpu.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2_pytest_1gpu.log | tail -5 | cut -c1-300
timeout 300 python __graft_entry__.py smoke nobuild 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_1.json 2> gpurun_out/r2_bench_1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_1.json").read().strip().splitlines()[-1])
print("bench", f"{d['value']:.4e}", d['ms_per_step'], d['phase_ms_per_step'], "e2e", f"{d['e2e']['value']:.4e}", "cpu", d.get('cpu_baseline',{}).get('value'), "roofline", d['roofline']['frac'], "eq", d['equilibrium_window']['ms_per_step'], "launches", d['gpu_launches'])
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -c 600 gpurun_out/r2_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 24 --warmup 10 --no-cpu --no-equilibrium > gpurun_out/r2_launches.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches.csv > gpurun_out/r2_launches_4M.txt; head -12 gpurun_out/r2_launches_4M.txt
timeout 600 python tools/run_config.py nial --ncell 200 200 200 --steps 60 --thermal 100 > gpurun_out/r2_cfg_cfg3_nial16M_N1.json 2> gpurun_out/r2_cfg_cfg3_nial16M_N1.err
tail -c 900 gpurun_out/r2_cfg_cfg3_nial16M_N1.json | head -c 700
