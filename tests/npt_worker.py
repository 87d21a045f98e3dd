"""Worker of tests/test_gpu_parity.py::test_cuda_npt_iso_matches_reference_fixture (own process, see there)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import common  # noqa: E402
from imd_b200 import api  # noqa: E402


def main(tmp):
    g = common.load_golden("cu_npt_iso")
    paths = common.write_tables(g, tmp)
    sim = api.IMDB200(1, g["box"], pair=paths["pair"], embed=paths["embed"], rho=paths["rho"], ensemble="npt_iso",
                      timestep=float(g["timestep"]), temperature=float(g["temperature"]), eta=float(g["eta0"]),
                      isq_tau_eta=float(g["isq_tau_eta"]), isq_tau_xi=float(g["npt_start:isq_tau_xi"]),
                      pressure_ext=float(g["npt_start:pressure_ext"]))
    sim.set_atoms(g["start:nummer"], g["start:sorte"], g["start:masse"], g["start:ort"], g["start:impuls"])
    sim.set_npt_state(xi=float(g["npt_start:xi"]), Ekin_old=float(g["npt_start:Ekin_old"]),
                      pressure_ext=float(g["npt_start:pressure_ext"]))
    worst = {}

    def rel(name, got, want, tol, s):
        e = abs(got - want) / max(abs(want), 1e-300)
        worst[name] = max(worst.get(name, 0.0), e)
        assert e <= tol, (name, s, got, want, e)

    for s in range(int(g["nsteps"])):
        sim.calc_forces(s)
        tol = 1e-10 if s == 0 else 1e-8
        rel("epot", sim.scalars()["tot_pot_energy"], g["epot"][s], tol, s)
        if s == 0:
            assert common.relerr(sim.atoms()["kraft"], g["f0:kraft"]) <= 1e-10
        sim.move_atoms()
        sim.check_nblist()
        st, sc = sim.npt(), sim.scalars()
        rel("xi", st["xi"], g["npt:xi"][s], 10 * tol, s)
        rel("pressure", st["pressure"], g["npt:pressure"][s], 10 * tol, s)
        rel("volume", sc["volume"], g["npt:volume"][s], 1e-10, s)
        rel("eta", sc["eta"], g["eta"][s], 10 * tol, s)
        rel("ekin", sc["tot_kin_energy"], g["ekin"][s], tol, s)
        assert np.max(np.abs(sim.box() - g["npt:box"][s])) <= 1e-10 * np.max(np.abs(g["npt:box"][s])), s
        assert sim.have_valid_nbl == int(g["valid"][s]), f"check_nblist decision differs at step {s}"
    a = sim.atoms()
    box = sim.box()
    d = a["ort"] - g["final:ort"]
    frac = d @ np.linalg.inv(box)
    d = (frac - np.round(frac)) @ box
    assert np.max(np.abs(d)) <= 1e-8 * np.max(np.abs(box))
    sim.close()
    print("NPT_OK", {k: f"{v:.1e}" for k, v in worst.items()})


if __name__ == "__main__":
    main(sys.argv[1])
