#!/bin/bash
# BASELINE configs other than configs[1] on N = visible GPUs; one JSON line each into gpurun_out/r2_cfg_<name>_N<N>.json
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
run() { # name, args...
  name=$1; shift
  if [ $N -gt 1 ]; then L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29788"; else L="python"; fi
  timeout 900 $L tools/run_config.py "$@" > gpurun_out/r2_cfg_${name}_N$N.json 2> gpurun_out/r2_cfg_${name}_N$N.err
  python - gpurun_out/r2_cfg_${name}_N$N.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], d["atoms"], f"{d['ms_per_step']:.3f} ms", f"{d['atom_steps_per_s']:.4e}", {k: round(v,3) for k,v in d["phase_ms_per_step"].items()}, "reb", d["rebuilds"], "sumF", f"{d['sum_F_rel']:.1e}", "dE", f"{d['dE_per_atom']:.1e}", "mem", round(d["mem_GB_rank0"],1))
except Exception as e:
    print(sys.argv[1], "ERR", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
}
if [ "$1" = smoke ]; then
  run smoke_cu cu --ncell 40 40 40 --steps 20 --thermal 20
  run smoke_deform deform --ncell 40 40 40 --strong --steps 20 --thermal 20
  run smoke_nial nial --ncell 40 40 40 --steps 20 --thermal 20
  exit 0
fi
if [ $N -eq 1 ]; then
  run cfg1_lj lj --ncell 20 20 20 --steps 200 --thermal 100
  run cfg3_nial16M nial --ncell 200 200 200 --steps 60 --thermal 60
  run cfg4_cu16M cu --ncell 200 200 100 --steps 60 --thermal 60
  run cfg4_cu32M cu --ncell 200 200 200 --steps 40 --thermal 60
  run cfg5_deform64M deform --ncell 400 200 200 --strong --steps 40 --thermal 50
else
  run cfg5_deform64M deform --ncell 400 200 200 --strong --steps 60 --thermal 50
  run cfg4_cu16M cu --ncell 200 200 100 --steps 60 --thermal 60
  run cfg4_cu32M cu --ncell 200 200 200 --steps 40 --thermal 60
fi
