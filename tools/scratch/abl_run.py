import os, sys, json, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from imd_b200 import api, synth
lib = sys.argv[1]
if lib != "prod":
    api.LIB_PATH = os.path.join(ROOT, "tools", "scratch", "abl", lib)
tmp = tempfile.mkdtemp()
tabs = synth.make_eam_tables(tmp, "cu")
ort, box = synth.fcc_lattice((100, 100, 100), synth.CU_A0)
n = len(ort); masse = np.full(n, synth.CU_MASS)
p = synth.maxwell_momenta(n, masse, 0.05, 3)
sim = api.IMDB200(1, box, pair=tabs["core_potential_file"], embed=tabs["embedding_energy_file"], rho=tabs["atomic_e-density_file"], nbl_size=1.2, timestep=0.001)
sim.set_atoms(np.arange(n, dtype=np.int32), np.zeros(n, np.int32), masse, ort, p)
sim.run(40); sim.timers(reset=True); sim.run(60)
tm = sim.timers()
sc = sim.scalars()
print(lib, {k: round(tm[k] / tm["steps"], 4) for k in ("pass1_ms", "pass2_ms", "rebuild_ms")}, "rebuilds", tm["rebuilds"],
      "E", repr(sc["tot_pot_energy"] + sc["tot_kin_energy"]))
