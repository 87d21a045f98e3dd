"""CPU: the oracle (oracle/imd_oracle.c) against the fixtures the unmodified reference produced."""
import numpy as np
import pytest

from tests import common
from oracle import oracle as orc

CASES = ["cu_nve", "nial_nvt", "lj_nve", "cu_slab", "cu_long", "cu_4point", "cu_spline", "nial_spline",
         "cu_lindef", "cu_frozen_nvt", "cu_frozen_nve", "cu_eeam", "nial_eeam", "cu_adp", "nial_adp"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_fixture(name, tmp_path):
    g = common.load_golden(name)
    sim = common.make_sim(orc.OracleIMD, g, str(tmp_path))
    out = common.run_protocol(sim, g)
    # PAIR_INT3 forms value and gradient as sums of four products that cancel to ~1e-3 of their size, so a 1-ulp
    # difference in rho (sum order) shows up as ~2e-13 in F'(rho): the 4point fixture gets 1e-12, the others 1e-13
    rtol = 1e-12 if name == "cu_4point" else 1e-13
    errs = common.compare(out, g, full_list=False, rtol=rtol, traj_rtol=1e-11)
    assert sim.scalars()["nactive"] == float(g["nactive"])
    if "final:box" in g:                                   # lin_deform changes the box
        assert np.array_equal(sim.box(), g["final:box"])
    else:
        assert np.array_equal(sim.celldims()[0], g["gdim"])
    assert sim.cellsz == float(g["cellsz"])
    print(name, {k: f"{v:.1e}" for k, v in errs.items()})


@pytest.mark.parametrize("name", ["nial_big", "cu_big"])
def test_oracle_matches_large_reference_fixture(name, tmp_path):
    """Parity at scale: 54 000 Ni-Al atoms (NVT) and 131 072 Cu atoms (NVE, 60 steps, 5 list builds) with the
    benchmark's full-resolution tables.  Per-atom results for a seeded sample of 4 096 atoms, the neighbour set of
    EVERY atom through its 64-bit hash, rebuild decisions and per-step scalars in full."""
    g = common.load_golden(name)
    sim = common.make_sim(orc.OracleIMD, g, str(tmp_path))
    out = common.run_protocol(sim, g)
    errs = common.compare(out, g, full_list=False, rtol=1e-12, traj_rtol=1e-10)
    assert np.array_equal(sim.celldims()[0], g["gdim"]) and sim.cellsz == float(g["cellsz"])
    assert int(g["nbl_builds"]) >= 3
    print(name, {k: f"{v:.1e}" for k, v in errs.items()})


@pytest.mark.parametrize("name", ["potaccess", "potaccess_4point", "potaccess_spline"])
def test_oracle_potaccess_known_answers(name, tmp_path):
    """PAIR_INT2 / PAIR_INT3 / PAIR_INT_SP known answers computed by the reference's own macros
    (src/potaccess.h:323-457) in the default, `4point` and `spline` builds; bit-exact."""
    g = common.load_golden(name)
    paths = common.write_tables(g, str(tmp_path))
    sim = orc.OracleIMD(2, np.eye(3) * 20.0, pair=paths["pair"], embed=paths["embed"], rho=paths["rho"],
                        interp=str(g["interp"]) if "interp" in g else "3point")
    for key in g:
        if not key.startswith("x:"):
            continue
        _, which, col = key.split(":")
        v, gr = sim.pair_int(int(which), int(col), g[key])
        assert np.array_equal(v, g[f"v:{which}:{col}"]), (which, col)
        assert np.array_equal(gr, g[f"g:{which}:{col}"]), (which, col)


def test_oracle_perfect_lattice_known_answers(tmp_path):
    """First-principles known answers: perfect fcc -> zero forces, 78 neighbours
    inside r_list (5th..6th shell), identical per-atom energies."""
    from imd_b200 import synth
    tabs = synth.make_eam_tables(str(tmp_path), "cu", nr=601, nrho=801)
    a0, nc = synth.CU_A0, 5
    base = np.array([[0, 0, 0], [.5, .5, 0], [.5, 0, .5], [0, .5, .5]]) + 0.25
    cells = np.stack(np.meshgrid(*[np.arange(nc)] * 3, indexing="ij"), -1).reshape(-1, 3)
    ort = ((cells[:, None, :] + base[None]) * a0).reshape(-1, 3)
    n = len(ort)
    sim = orc.OracleIMD(1, np.eye(3) * a0 * nc, pair=tabs["core_potential_file"],
                        embed=tabs["embedding_energy_file"], rho=tabs["atomic_e-density_file"])
    sim.set_atoms(np.arange(n), np.zeros(n, int), np.full(n, synth.CU_MASS), ort)
    sim.set_integrator("nve", 0.001)
    sim.calc_forces(0)
    a = sim.atoms()
    assert np.max(np.abs(a["kraft"])) < 1e-12
    assert np.ptp(a["poteng"]) < 1e-12 and np.ptp(a["rho"]) < 1e-12
    pairs, _ = sim.nbl_pairs()
    assert len(pairs) == 39 * n


def test_oracle_npt_iso_matches_reference_fixture(tmp_path):
    """move_atoms_npt_iso (src/imd_integrate.c:1472-1729) of the reference's `npt_iso` build: the barostat variable xi,
    the pressure it is driven by, eta, the breathing box and the trajectory.  Oracle only: the CUDA engine has no NPT
    ensemble yet, this fixture is what it will be held against."""
    g = common.load_golden("cu_npt_iso")
    paths = common.write_tables(g, str(tmp_path))
    sim = orc.OracleIMD(1, g["box"], pair=paths["pair"], embed=paths["embed"], rho=paths["rho"])
    sim.set_integrator("npt_iso", float(g["timestep"]), float(g["temperature"]), float(g["eta0"]), float(g["isq_tau_eta"]))
    sim.set_npt(xi=float(g["npt_start:xi"]), Ekin_old=float(g["npt_start:Ekin_old"]),
                pressure_ext=float(g["npt_start:pressure_ext"]), d_pressure=0.0, isq_tau_xi=float(g["npt_start:isq_tau_xi"]))
    sim.set_atoms(g["start:nummer"], g["start:sorte"], g["start:masse"], g["start:ort"], g["start:impuls"])
    n = int(g["nsteps"])
    for s in range(n):
        sim.calc_forces(s)
        sc = sim.scalars()
        tol = 1e-13 if s == 0 else 1e-10
        assert abs(sc["tot_pot_energy"] - g["epot"][s]) <= tol * abs(g["epot"][s]), s
        if s == 0:
            a = sim.atoms()
            assert common.relerr(a["kraft"], g["f0:kraft"]) <= 1e-13
        sim.move_atoms()
        sim.check_nblist()
        st, sc = sim.npt(), sim.scalars()
        assert abs(st["xi"] - g["npt:xi"][s]) <= tol * 10 * abs(g["npt:xi"][s]), s
        assert abs(st["pressure"] - g["npt:pressure"][s]) <= tol * 10 * abs(g["npt:pressure"][s]), s
        assert abs(sc["volume"] - g["npt:volume"][s]) <= 1e-13 * g["npt:volume"][s], s
        assert abs(sc["eta"] - g["eta"][s]) <= tol * 10 * abs(g["eta"][s]), s
        assert abs(sc["tot_kin_energy"] - g["ekin"][s]) <= tol * abs(g["ekin"][s]), s
        assert np.max(np.abs(sim.box() - g["npt:box"][s])) <= 1e-13 * np.max(np.abs(g["npt:box"][s])), s
        assert sim.have_valid_nbl == int(g["valid"][s]), f"check_nblist decision differs at step {s}"
    a = sim.atoms()
    box = sim.box()
    d = a["ort"] - g["final:ort"]
    frac = d @ np.linalg.inv(box)
    d = (frac - np.round(frac)) @ box
    assert np.max(np.abs(d)) <= 1e-9 * np.max(np.abs(box))
    assert np.max(np.abs(a["impuls"] - g["final:impuls"])) <= 1e-9 * np.max(np.abs(g["final:impuls"]))


@pytest.mark.parametrize("case", ["cu_npt_axial", "cu_npt_axial_xz", "cu_npt_axial_restr"])
def test_oracle_npt_axial_matches_reference_fixture(case, tmp_path):
    """move_atoms_npt_axial (src/imd_integrate.c:1747-1959) of the reference's `npt_axial` build: one barostat variable per
    box axis driven by (dyn_stress + vir)/volume of that axis (P_AXIAL builds accumulate vir_xx/yy/zz in calc_forces,
    src/imd_forces_nbl.c:548-556), a pressure ramp that differs per axis, and relax_dirs 1 0 1 holding the y axis."""
    g = common.load_golden(case)
    paths = common.write_tables(g, str(tmp_path))
    sim = orc.OracleIMD(1, g["box"], pbc=tuple(int(x) for x in g["pbc"]), pair=paths["pair"], embed=paths["embed"], rho=paths["rho"])
    sim.set_integrator("npt_axial", float(g["timestep"]), float(g["temperature"]), float(g["eta0"]), float(g["isq_tau_eta"]))
    if "restrictions" in g:          # the pinned layer: restriction vectors act on the momenta after the kick (:1859-1864)
        sim.set_restrictions(g["restrictions"])
    sim.set_npt_axial(g["npt_start:xi"], g["npt_start:pressure_ext"], g["npt_start:d_pressure"], g["npt_start:relax_dirs"],
                      Ekin_old=float(g["npt_start:Ekin_old"]), dyn_stress=g["npt_start:dyn_stress"],
                      isq_tau_xi=float(g["npt_start:isq_tau_xi"]))
    sim.set_atoms(g["start:nummer"], g["start:sorte"], g["start:masse"], g["start:ort"], g["start:impuls"], vsorte=g["start:vsorte"])
    press = bool(int(g["press"]))
    sim.set_press_calc(press)
    for s in range(int(g["nsteps"])):
        sim.calc_forces(s)
        sc = sim.scalars()
        tol = 1e-13 if s == 0 else 1e-10
        assert abs(sc["tot_pot_energy"] - g["epot"][s]) <= tol * abs(g["epot"][s]), s
        vir = np.array([sc["vir_xx"], sc["vir_yy"], sc["vir_zz"]])
        assert np.max(np.abs(vir - g["npt:vir"][s])) <= tol * 10 * np.max(np.abs(g["npt:vir"][s])), s
        if s == 0:
            assert common.relerr(sim.atoms()["kraft"], g["f0:kraft"]) <= 1e-13
        rec = press and s in [int(x) for x in g["record"]]
        if rec:      # virial part of the per-atom tensor (recorded between calc_forces and move_atoms)
            assert common.relerr(sim.atoms()["presstens"], g[f"f{s}:presstens"]) <= 10 * tol
        sim.move_atoms()
        sim.check_nblist()
        if rec:      # plus the kinetic part, which this integrator adds from the momenta BEFORE the kick (:1834-1845)
            assert common.relerr(sim.tot_presstens(), g[f"f{s}:tot_presstens"]) <= 10 * tol
        st, sc = sim.npt_axial(), sim.scalars()
        for k in ("xi", "stress", "pressure_ext", "dyn_stress"):
            assert np.max(np.abs(st[k] - g["npt:" + k][s])) <= tol * 10 * np.max(np.abs(g["npt:" + k][s])), (k, s)
        assert abs(sc["volume"] - g["npt:volume"][s]) <= 1e-13 * g["npt:volume"][s], s
        assert abs(sc["eta"] - g["eta"][s]) <= tol * 10 * abs(g["eta"][s]), s
        assert abs(sc["tot_kin_energy"] - g["ekin"][s]) <= tol * abs(g["ekin"][s]), s
        assert np.max(np.abs(sim.box() - g["npt:box"][s])) <= 1e-13 * np.max(np.abs(g["npt:box"][s])), s
        assert sim.have_valid_nbl == int(g["valid"][s]), f"check_nblist decision differs at step {s}"
    if case.endswith("_xz"):          # the held axis did not move
        assert sim.box()[1, 1] == g["box"][1, 1] and np.all(g["npt:xi"][:, 1] == 0.0)
    a = sim.atoms()
    box = sim.box()
    d = a["ort"] - g["final:ort"]
    frac = d @ np.linalg.inv(box)
    d = (frac - np.round(frac)) @ box
    assert np.max(np.abs(d)) <= 1e-9 * np.max(np.abs(box))
    assert np.max(np.abs(a["impuls"] - g["final:impuls"])) <= 1e-9 * np.max(np.abs(g["final:impuls"]))


def test_oracle_berendsen_matches_reference_fixture(tmp_path):
    """`ber` builds: Berendsen scaling of the momenta inside move_atoms_nve (src/imd_integrate.c:44-53, 341-350),
    driven by the kinetic energy of the PREVIOUS step.  Oracle only so far."""
    g = common.load_golden("cu_berendsen")
    paths = common.write_tables(g, str(tmp_path))
    sim = orc.OracleIMD(1, g["box"], pair=paths["pair"], embed=paths["embed"], rho=paths["rho"])
    sim.set_integrator("nve", float(g["timestep"]), float(g["temperature"]))
    sim.set_atoms(g["start:nummer"], g["start:sorte"], g["start:masse"], g["start:ort"], g["start:impuls"])
    sim.set_berendsen(float(g["tau_berendsen"]), float(g["ekin_start"]))
    for s in range(int(g["nsteps"])):
        sim.calc_forces(s)
        tol = 1e-13 if s == 0 else 1e-10
        assert abs(sim.scalars()["tot_pot_energy"] - g["epot"][s]) <= tol * abs(g["epot"][s]), s
        sim.move_atoms()
        sim.check_nblist()
        assert abs(sim.scalars()["tot_kin_energy"] - g["ekin"][s]) <= tol * abs(g["ekin"][s]), s
        assert sim.have_valid_nbl == int(g["valid"][s])
    a = sim.atoms()
    assert np.max(np.abs(a["impuls"] - g["final:impuls"])) <= 1e-9 * np.max(np.abs(g["final:impuls"]))
    # the thermostat did something: the kinetic energy moved towards the target faster than plain NVE would
    assert abs(g["ekin"][-1] - g["ekin"][0]) > 1e-3 * abs(g["ekin"][0])
