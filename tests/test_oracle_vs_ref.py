"""CPU: the oracle (oracle/imd_oracle.c) against the UNMODIFIED reference run live (oracle/_ref/libimdref_*.so through
oracle/ref_driver.py, one subprocess per run), on configurations the committed fixtures do not contain: other sizes,
cell grids, seeds, temperatures and ensembles.  The fixtures pin the oracle on the GPU box; this pins it once more
wherever the reference libraries are present (they are built from /root/reference by oracle/Makefile and travel as
prebuilt files)."""
import os
import tempfile

import numpy as np
import pytest

from tests import common
from imd_b200 import synth
from oracle import oracle as orc
from oracle import ref_driver as rd
from oracle.oracle import canonical_pairs

CASES = {
    # name: (kind, variant, ncell, ensemble, starttemp, seed, warm, nsteps, interp, extra)
    "cu_6x5x7_nvt": ("cu", "eam", (6, 5, 7), "nvt", 0.10, 4711, 15, 10, "3point", None),
    "cu_8_nve_hot": ("cu", "eam", (8, 8, 8), "nve", 0.20, 99, 10, 8, "3point", None),
    "nial_6_nve": ("nial", "eam", (6, 6, 6), "nve", 0.08, 5, 12, 8, "3point", None),
    "cu_4point_nvt": ("cu", "eam_4point", (6, 6, 5), "nvt", 0.07, 17, 10, 8, "4point", None),
    "nial_spline_nve": ("nial", "eam_spline", (5, 6, 5), "nve", 0.05, 23, 10, 8, "spline", None),
    "cu_wire": ("cu", "eam", (5, 5, 6), "nve", 0.05, 3, 10, 8, "3point", dict(pbc_dirs=[0, 0, 1])),
    "lj_5": ("lj", "pair", (5, 5, 5), "nve", 0.008, 8, 10, 8, "3point", None),
}


def _need(variant):
    if not rd.available(variant):
        pytest.skip(f"oracle/_ref/libimdref_{variant}.so missing: run `make -C oracle ref` where /root/reference exists")


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_live_reference(name):
    kind, variant, ncell, ens, temp, seed, warm, nsteps, interp, extra = CASES[name]
    _need(variant)
    tmp = tempfile.mkdtemp(prefix="ovr_" + name)
    if kind == "cu":
        tabs = synth.make_eam_tables(tmp, "cu", nr=701, nrho=901)
        p = synth.cu_param(tmp, ncell=ncell, ensemble=ens, starttemp=temp, seed=seed, tables=tabs, extra=extra)
        nt = 1
    elif kind == "nial":
        tabs = synth.make_eam_tables(tmp, "nial", nr=701, nrho=901)
        p = synth.nial_param(tmp, ncell=ncell, ensemble=ens, starttemp=temp, seed=seed, tables=tabs, extra=extra)
        nt = 2
    else:
        tabs = synth.make_lj_table(tmp, nsteps=900)
        p = synth.lj_param(tmp, ncell=ncell, starttemp=temp, seed=seed, table=tabs, extra=extra)
        nt = 2
    spec = dict(variant=variant, paramfile=p, warm=warm, nsteps=nsteps, record_atoms=[0, nsteps - 1], record_nbl=[0],
                press=True)
    ref = rd.run_in_subprocess(spec, tmp)
    st = ref["start"]
    sc0 = ref["frames"][0]["scalars"]
    kw = dict(pbc=tuple((extra or {}).get("pbc_dirs", [1, 1, 1])), nbl_margin=0.4, interp=interp,
              pair=tabs.get("core_potential_file", tabs.get("potfile")), embed=tabs.get("embedding_energy_file"),
              rho=tabs.get("atomic_e-density_file"))
    sim = orc.OracleIMD(nt, ref["box"], **kw)
    sim.set_integrator(ensemble=ens, timestep=sc0["timestep"], temperature=sc0["temperature"], eta=sc0["eta"],
                       isq_tau_eta=1.0 / 0.1 ** 2 if ens == "nvt" else 0.0)
    sim.set_atoms(st["nummer"], st["sorte"], st["masse"], st["ort"], st["impuls"], vsorte=st["vsorte"])
    sim.set_press_calc(True)
    tol = 1e-12 if interp == "4point" else 1e-13
    for s in range(nsteps):
        sim.calc_forces(s)
        fr = ref["frames"][s]
        sc = sim.scalars()
        rtol = tol if s == 0 else 1e-10
        assert abs(sc["tot_pot_energy"] - fr["scalars"]["tot_pot_energy"]) <= rtol * abs(fr["scalars"]["tot_pot_energy"]), s
        assert abs(sc["virial"] - fr["scalars"]["virial"]) <= 10 * rtol * max(abs(fr["scalars"]["virial"]), 1.0), s
        if "atoms" in fr:
            a = sim.atoms()
            for k in ("kraft", "poteng", "rho", "dF", "presstens"):
                if np.max(np.abs(fr["atoms"][k])) == 0:
                    continue
                assert common.relerr(a[k], fr["atoms"][k]) <= rtol, (s, k, common.relerr(a[k], fr["atoms"][k]))
        if "nbl_pairs" in fr:
            got = canonical_pairs(*sim.nbl_pairs())
            want = canonical_pairs(fr["nbl_pairs"], fr["nbl_shift"])
            assert got.shape == want.shape and np.array_equal(got, want), "neighbour set differs from the reference"
        sim.move_atoms()
        sim.check_nblist()
        assert sim.have_valid_nbl == fr["valid"], f"check_nblist decision differs at step {s}"
        sc = sim.scalars()
        assert abs(sc["tot_kin_energy"] - fr["after"]["tot_kin_energy"]) <= rtol * abs(fr["after"]["tot_kin_energy"]), s
        assert abs(sc["eta"] - fr["after"]["eta"]) <= 1e-9 * max(abs(fr["after"]["eta"]), 1e-6), s
    assert np.array_equal(sim.celldims()[0], ref["celldims"][0])
    assert sim.cellsz == ref["cellsz"]
    # list builds inside the protocol: the first one plus one after every step whose check said "invalid"
    # (the reference's own counter also holds the builds of its warm-up phase)
    assert sim.nbl_count == 1 + sum(1 for fr in ref["frames"][:-1] if not fr["valid"])
