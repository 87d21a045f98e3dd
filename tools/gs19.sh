#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/rp
python - <<'PY'
import os, subprocess
from imd_b200 import synth
tmp="/tmp/rp"
tabs = synth.make_eam_tables(tmp, "cu", nr=601, nrho=801)
extra = dict(eng_int=1, checkpt_int=20, relax_rate=0.01, relax_mode="full", bulk_module=1.0, shear_module=0.5)
p = synth.cu_param(tmp, ncell=(10,10,10), name="gpu", tables=tabs, ensemble="nve", maxsteps=20, starttemp=0.08, extra=extra)
r=subprocess.run(["stdbuf","-o0",os.path.abspath("oracle/_ref/imd_b200_dropin_full_dbg"),"-p",p],capture_output=True,text=True,cwd=tmp)
print(r.returncode); print(r.stdout[-1500:]); print(r.stderr[-3000:])
print(open(tmp+"/gpu.eng").read()[-600:] if os.path.exists(tmp+"/gpu.eng") else "no eng")
PY
