#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built by oracle/Makefile
from /root/reference/src).  Run in the build container only; the fixtures are committed because
/root/reference does not exist on the GPU box.

    python tools/make_golden.py            # all cases

Each fixture holds: the potential tables (as arrays, re-written to IMD format by tests/common.py),
the start state written by the reference after its own thermalisation, and what the reference
computed from it: per-atom forces/energies/densities, scalars per step, the Verlet neighbour set,
rebuild flags and the final state.  The reference ships no golden vectors of its own
(SURVEY.md section 4), so these are the pinned parity targets.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imd_b200 import synth  # noqa: E402
from oracle import ref_driver as rd  # noqa: E402
from oracle.oracle import canonical_pairs  # noqa: E402
from tools.parity_fixture import pair_hash  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

CASES = {
    # BASELINE config 2 in miniature: EAM Cu fcc, NVE, Verlet list + skin, 3x3x3 cells
    "cu_nve": dict(kind="cu", ncell=(5, 5, 5), ensemble="nve", starttemp=0.08, warm=40, nsteps=24,
                   record=[0, 23], press=True, variant="eam"),
    # BASELINE config 3 in miniature: binary EAM Ni-Al (B2), NVT, two-species tables, 2x2x2 cells
    "nial_nvt": dict(kind="nial", ncell=(5, 5, 5), ensemble="nvt", starttemp=0.06, warm=40, nsteps=16,
                     record=[0, 15], press=True, variant="eam"),
    # BASELINE config 1 in miniature: LJ Ar, tabulated pair potential (mklj layout), NVE
    "lj_nve": dict(kind="lj", ncell=(4, 4, 4), ensemble="nve", starttemp=0.006, warm=40, nsteps=16,
                   record=[0, 15], press=True, variant="pair"),
    # free surfaces: no periodic images along z, one-cell-thick ghost shell only in x,y
    "cu_slab": dict(kind="cu", ncell=(5, 5, 4), ensemble="nve", starttemp=0.05, warm=20, nsteps=8,
                    record=[0, 7], press=False, variant="eam", extra=dict(pbc_dirs=[1, 1, 0])),
    # non-cubic box, more cells, longer run with several rebuilds
    # `4point` / `spline` reference builds (cubic table interpolation, src/potaccess.h:365-457)
    "cu_4point": dict(kind="cu", ncell=(5, 5, 5), ensemble="nve", starttemp=0.08, warm=40, nsteps=16,
                      record=[0, 15], press=True, variant="eam_4point", interp="4point"),
    "nial_spline": dict(kind="nial", ncell=(5, 5, 5), ensemble="nvt", starttemp=0.06, warm=40, nsteps=16,
                        record=[0, 15], press=True, variant="eam_spline", interp="spline"),
    "cu_spline": dict(kind="cu", ncell=(5, 5, 5), ensemble="nve", starttemp=0.08, warm=40, nsteps=16,
                      record=[0, 15], press=False, variant="eam_spline", interp="spline"),
    # homogeneous deformation (lin_deform, src/imd_deform.c:35-119) every 4 steps: uniaxial strain plus two shear
    # components, so the box turns triclinic; the list stays valid across a deformation until check_nblist says no
    "cu_lindef": dict(kind="cu", ncell=(5, 5, 5), ensemble="nve", starttemp=0.08, warm=30, nsteps=24,
                      record=[0, 23], press=True, variant="eam", lindef_every=4,
                      extra=dict(lindef_interval=4, lindef_size=2.5e-3, lindef_x=[1.0, 0.2, 0.0],
                                 lindef_y=[0.0, -0.4, 0.0], lindef_z=[0.3, 0.0, 0.2])),
    # slab with free surfaces in z, NVT: the bottom layer is virtual type 1, frozen in z (restrictionvector) and pushed
    # along x every 5 steps, the rest is sheared (deform_sample, src/imd_deform.c:232-269); nactive < 3N enters eta
    "cu_frozen_nvt": dict(kind="cu_vtypes", ncell=(5, 5, 4), ensemble="nvt", starttemp=0.06, warm=20, nsteps=20,
                          record=[0, 19], press=False, variant="eam", deform_every=5,
                          extra=dict(pbc_dirs=[1, 1, 0], total_types=2, restrictionvector=[1, 1, 1, 0],
                                     max_deform_int=5, deform_size=1.0)),
    # the same slab under NVE: restrictions multiply the force before the kick (src/imd_integrate.c:193-197)
    "cu_frozen_nve": dict(kind="cu_vtypes", ncell=(5, 5, 4), ensemble="nve", starttemp=0.06, warm=20, nsteps=12,
                          record=[0, 11], press=False, variant="eam", deform_every=5,
                          extra=dict(pbc_dirs=[1, 1, 0], total_types=2, restrictionvector=[1, 1, 1, 0],
                                     max_deform_int=5, deform_size=1.0)),
    # `eeam` reference build: second embedding term M(sum rho^2), single- and two-species
    "cu_eeam": dict(kind="cu", ncell=(5, 5, 5), ensemble="nve", starttemp=0.08, warm=30, nsteps=12,
                    record=[0, 11], press=True, variant="eeam", eeam=True),
    "nial_eeam": dict(kind="nial", ncell=(5, 5, 5), ensemble="nvt", starttemp=0.06, warm=30, nsteps=12,
                      record=[0, 11], press=False, variant="eeam", eeam=True),
    # `npt_iso` reference build (Nose-Hoover thermostat + isotropic barostat): the box breathes every step.  Oracle
    # fixture only so far -- the CUDA engine does not have this ensemble yet
    "cu_npt_iso": dict(kind="cu", ncell=(6, 6, 6), ensemble="npt_iso", starttemp=0.08, warm=25, nsteps=40,
                       record=[0, 39], press=False, variant="npt",
                       extra=dict(endtemp=0.08, tau_eta=0.1, eta=0.0, tau_xi=0.5, pressure_start=0.02, pressure_end=0.02)),
    # `npt_axial` reference build: one barostat per box axis, driven by (dyn_stress + vir)/volume per axis; a pressure
    # ramp that differs per axis, and a second case with the y axis held (relax_dirs 1 0 1)
    "cu_npt_axial": dict(kind="cu", ncell=(6, 6, 6), ensemble="npt_axial", starttemp=0.08, warm=25, nsteps=40,
                         record=[0, 39], press=False, variant="npt_axial",
                         extra=dict(endtemp=0.08, tau_eta=0.1, eta=0.0, tau_xi=0.5, pressure_start=[0.02, 0.01, 0.03],
                                    pressure_end=[0.03, 0.01, 0.02], maxsteps=100)),
    "cu_npt_axial_xz": dict(kind="cu", ncell=(6, 5, 6), ensemble="npt_axial", starttemp=0.08, warm=25, nsteps=30,
                            record=[0, 29], press=True, variant="npt_axial",
                            extra=dict(endtemp=0.08, tau_eta=0.1, eta=0.0, tau_xi=0.4, pressure_start=[0.0, 0.0, 0.02],
                                       pressure_end=[0.0, 0.0, 0.02], relax_dirs=[1, 0, 1], maxsteps=100)),
    # npt_axial on a slab whose bottom layer (virtual type 1) may not move in z: the restriction vectors multiply the momenta
    # AFTER the kick in this integrator (src/imd_integrate.c:1859-1864)
    "cu_npt_axial_restr": dict(kind="cu_vtypes", ncell=(5, 5, 4), ensemble="npt_axial", starttemp=0.06, warm=20, nsteps=20,
                               record=[0, 19], press=False, variant="npt_axial", restr_only=True,
                               extra=dict(pbc_dirs=[1, 1, 0], total_types=2, restrictionvector=[1, 1, 1, 0],
                                          endtemp=0.06, tau_eta=0.1, eta=0.0, tau_xi=0.5, pressure_start=[0.0, 0.0, 0.0],
                                          pressure_end=[0.0, 0.0, 0.0], relax_dirs=[1, 1, 0], maxsteps=100)),
    # `adp` reference build (angular-dependent potential).  Oracle fixtures only so far, like cu_npt_iso
    "cu_adp": dict(kind="cu", ncell=(5, 5, 5), ensemble="nve", starttemp=0.08, warm=30, nsteps=12,
                   record=[0, 11], press=True, variant="adp", adp=True),
    "nial_adp": dict(kind="nial", ncell=(5, 5, 5), ensemble="nvt", starttemp=0.06, warm=30, nsteps=12,
                     record=[0, 11], press=False, variant="adp", adp=True),
    # `ber` reference build: Berendsen scaling inside move_atoms_nve, target temperature above the current one.
    # Oracle fixture only so far
    "cu_berendsen": dict(kind="cu", ncell=(5, 5, 5), ensemble="nve", starttemp=0.05, warm=15, nsteps=30,
                         record=[0, 29], press=False, variant="ber", extra=dict(tau_berendsen=0.05, endtemp=0.09)),
    "cu_long": dict(kind="cu", ncell=(7, 5, 6), ensemble="nve", starttemp=0.12, warm=30, nsteps=60,
                    record=[0, 59], press=False, variant="eam"),
    # Parity at scale: 131 072 Cu atoms (19^3 cells) with the benchmark's full-resolution tables
    # (2001 / 4001 rows, the 96 KB fused table of the single-species kernels) over 60 steps with several rebuilds, and
    # 54 000 Ni-Al atoms (B2, 4-column tables that do not fit in shared memory) under NVT.  `big`: the start state is
    # stored in full, per-atom results only for a seeded sample of 4 096 atoms, the neighbour set as one 64-bit hash
    # per atom (tools/parity_fixture.pair_hash) -- small files, same pinning power.
    "cu_big": dict(kind="cu", ncell=(32, 32, 32), ensemble="nve", starttemp=0.12, warm=30, nsteps=60,
                   record=[0, 59], press=False, variant="eam", big=True, fullres=True),
    "nial_big": dict(kind="nial", ncell=(30, 30, 30), ensemble="nvt", starttemp=0.06, warm=30, nsteps=40,
                     record=[0, 39], press=False, variant="eam", big=True, fullres=True),
}


def table_arrays(paths):
    out = {}
    for key in ("core_potential_file", "embedding_energy_file", "atomic_e-density_file", "potfile", "eeam_energy_file", "adp_upotfile",
                "adp_wpotfile"):
        if key in paths:
            with open(paths[key]) as f:
                out["table:" + key] = np.frombuffer(f.read().encode(), dtype=np.uint8)
    return out


def make_case(name, c):
    tmp = tempfile.mkdtemp(prefix="gold_" + name)
    if c.get("eeam"):
        emod = synth.make_eeam_table(tmp, nt=2 if c["kind"] == "nial" else 1)
        c = dict(c, extra=dict(c.get("extra") or {}, eeam_energy_file=emod))
    if c.get("adp"):
        pu, pw = synth.make_adp_tables(tmp, nt=2 if c["kind"] == "nial" else 1)
        c = dict(c, extra=dict(c.get("extra") or {}, adp_upotfile=pu, adp_wpotfile=pw))
    res = dict() if c.get("fullres") else dict(nr=601, nrho=801)
    if c["kind"] == "cu":
        tabs = synth.make_eam_tables(tmp, "cu", **res)
        p = synth.cu_param(tmp, ncell=c["ncell"], ensemble=c["ensemble"], starttemp=c["starttemp"],
                           tables=tabs, extra=c.get("extra"))
        ntypes = 1
    elif c["kind"] == "cu_vtypes":
        # our own start configuration, read by the reference: fcc Cu, virtual type 1 for the bottom layer
        tabs = synth.make_eam_tables(tmp, "cu", nr=601, nrho=801)
        ort, box = synth.fcc_lattice(c["ncell"], synth.CU_A0)
        n = len(ort)
        vs = (ort[:, 2] < 0.6 * synth.CU_A0).astype(np.int32)
        masse = np.full(n, synth.CU_MASS)
        mom = synth.maxwell_momenta(n, masse, c["starttemp"], 11)
        cfgfile = synth.write_config(os.path.join(tmp, "start.conf"), np.arange(n), vs, masse, ort, mom, box)
        extra = dict(c["extra"], box_from_header=1)
        p = synth.cu_param(tmp, ncell=c["ncell"], ensemble=c["ensemble"], starttemp=c["starttemp"], tables=tabs,
                           coordname=cfgfile, extra=extra)
        # keys that take "<vtype> x y z" appear once per virtual type: append the second lines by hand
        if not c.get("restr_only"):
            with open(p, "a") as f:
                f.write("deform_shift 0 1.0 0.0 0.0\ndeform_shear 0 0.0 0.0 4.0e-4\ndeform_base 0 0.0 0.0 0.0\n"
                        "deform_shift 1 0.012 0.0 0.0\n")
        ntypes = 1
    elif c["kind"] == "nial":
        tabs = synth.make_eam_tables(tmp, "nial", **res)
        p = synth.nial_param(tmp, ncell=c["ncell"], ensemble=c["ensemble"], starttemp=c["starttemp"],
                             tables=tabs, extra=c.get("extra"))
        ntypes = 2
    else:
        tabs = synth.make_lj_table(tmp, nsteps=1200)
        p = synth.lj_param(tmp, ncell=c["ncell"], starttemp=c["starttemp"], table=tabs, extra=c.get("extra"))
        ntypes = 2
    spec = dict(variant=c["variant"], paramfile=p, warm=c["warm"], nsteps=c["nsteps"],
                record_atoms=c["record"], record_nbl=[0], press=c["press"],
                lindef_every=c.get("lindef_every", 0), deform_every=c.get("deform_every", 0))
    out = rd.run_in_subprocess(spec, tmp)
    if c.get("eeam"):
        tabs = dict(tabs, eeam_energy_file=c["extra"]["eeam_energy_file"])
    if c.get("adp"):
        tabs = dict(tabs, adp_upotfile=c["extra"]["adp_upotfile"], adp_wpotfile=c["extra"]["adp_wpotfile"])
    g = dict(table_arrays(tabs))
    g["ntypes"] = ntypes
    g["ensemble"] = c["ensemble"]
    g["interp"] = c.get("interp", "3point")
    g["press"] = int(c["press"])
    g["nsteps"] = c["nsteps"]
    g["record"] = np.array(c["record"])
    g["pbc"] = np.array(c.get("extra", {}).get("pbc_dirs", [1, 1, 1]))
    if c.get("lindef_every"):
        e = c["extra"]
        g["lindef_every"] = c["lindef_every"]; g["lindef_size"] = e["lindef_size"]
        g["lindef_x"] = np.array(e["lindef_x"]); g["lindef_y"] = np.array(e["lindef_y"]); g["lindef_z"] = np.array(e["lindef_z"])
        g["final:box"] = out["final_box"]
    if c.get("restr_only"):
        g["total_types"] = 2
        g["restrictions"] = np.array([[1.0, 1.0, 1.0], [1.0, 1.0, 0.0]])
    if c.get("deform_every"):
        g["deform_every"] = c["deform_every"]; g["deform_size"] = c["extra"]["deform_size"]
        g["total_types"] = 2
        g["restrictions"] = np.array([[1.0, 1.0, 1.0], [1.0, 1.0, 0.0]])
        g["deform_shift"] = np.array([[1.0, 0.0, 0.0], [0.012, 0.0, 0.0]])
        g["shear_def"] = np.array([1, 0], np.int32)
        g["deform_shear"] = np.array([[0.0, 0.0, 4.0e-4], [0.0, 0.0, 0.0]])
        g["deform_base"] = np.zeros((2, 3))
    g["box"] = out["box"]
    g["cellsz"] = out["cellsz"]
    g["gdim"], g["cdim"] = out["celldims"]
    sc0 = out["frames"][0]["scalars"]
    g["timestep"] = sc0["timestep"]; g["temperature"] = sc0["temperature"]; g["eta0"] = sc0["eta"]
    g["nactive"] = sc0["nactive"]
    g["isq_tau_eta"] = 1.0 / 0.1 ** 2 if c["ensemble"] in ("nvt", "npt_iso", "npt_axial") else 0.0
    if c["variant"] == "ber":
        g["tau_berendsen"] = c["extra"]["tau_berendsen"]
        g["ekin_start"] = out["frames"][0]["scalars"]["tot_kin_energy"]      # what the last warm-up step left
        g["temperature_steps"] = np.array([f["scalars"]["temperature"] for f in out["frames"]])
    if c["variant"] == "npt":
        for k, v in out["npt_start"].items():
            g["npt_start:" + k] = v
        g["npt:xi"] = np.array([f["npt"]["xi"] for f in out["frames"]])
        g["npt:pressure"] = np.array([f["npt"]["pressure"] for f in out["frames"]])
        g["npt:volume"] = np.array([f["after"]["volume"] for f in out["frames"]])
        g["npt:box"] = np.array([f["box"] for f in out["frames"]])
    if c["variant"] == "npt_axial":
        st = out["npt_axial_start"]
        for k, v in st.items():
            g["npt_start:" + k] = v
        g["npt_start:isq_tau_xi"] = 1.0 / c["extra"]["tau_xi"] ** 2
        fr = out["frames"]
        for k in ("xi", "stress", "pressure_ext", "dyn_stress"):
            g["npt:" + k] = np.array([f["npt_axial"][k] for f in fr])
        g["npt:Ekin_old"] = np.array([f["npt_axial"]["Ekin_old"] for f in fr])
        # d_pressure is a function-local static of move_atoms_npt_axial: read it off the ramp it produces
        g["npt_start:d_pressure"] = fr[0]["npt_axial"]["pressure_ext"] - st["pressure_ext"]
        g["npt:vir"] = np.array([[f["scalars"]["vir_xx"], f["scalars"]["vir_yy"], f["scalars"]["vir_zz"]] for f in fr])
        g["npt:volume"] = np.array([f["after"]["volume"] for f in fr])
        g["npt:box"] = np.array([f["box"] for f in fr])
    for k in ("nummer", "sorte", "vsorte", "masse", "ort", "impuls"):
        g["start:" + k] = out["start"][k]
    for k in ("ort", "impuls"):
        g["final:" + k] = out["final"][k]
    keys = ["tot_pot_energy", "virial"]
    g["epot"] = np.array([f["scalars"]["tot_pot_energy"] for f in out["frames"]])
    g["virial"] = np.array([f["scalars"]["virial"] for f in out["frames"]])
    g["ekin"] = np.array([f["after"]["tot_kin_energy"] for f in out["frames"]])
    g["eta"] = np.array([f["after"]["eta"] for f in out["frames"]])
    g["valid"] = np.array([f["valid"] for f in out["frames"]])
    g["nbl_count"] = out["nbl_count"]
    g["nbl_builds"] = out["nbl_count"] - out["nbl_count0"]       # list builds of the protocol itself
    big = bool(c.get("big"))
    sel = slice(None)
    if big:
        # atoms() is sorted by NUMMER: the sample is a set of positions in that order
        sel = np.sort(np.random.default_rng(2024).choice(out["natoms"], 4096, replace=False)).astype(np.int32)
        g["sample"] = sel
        g["final:ort"] = out["final"]["ort"][sel]
        del g["final:impuls"]
    for s in c["record"]:
        a = out["frames"][s]["atoms"]
        for k in ("kraft", "poteng", "rho", "dF", "presstens", "ort") + (("eam_p", "dM") if c.get("eeam") else ()) \
                + (("adp_mu", "adp_lambda") if c.get("adp") else ()):
            if big and k in ("presstens", "ort"):
                continue
            g[f"f{s}:{k}"] = a[k][sel]
        if c["press"]:
            g[f"f{s}:tot_presstens"] = out["frames"][s]["tot_presstens"]
    fr0 = out["frames"][0]
    rows = canonical_pairs(fr0["nbl_pairs"], fr0["nbl_shift"])
    if big:
        # symmetric closure of the half list, hashed per atom: by atom numbers alone (what a domain-decomposed run can
        # report) and with the image shift of every entry folded in (single-domain runs)
        full = np.concatenate([rows, np.column_stack([rows[:, 1], rows[:, 0], -rows[:, 2:]])])
        g["nbl_len_full"] = len(full)
        g["nbl_hash_nummer"], g["nbl_hash"] = pair_hash(full[:, 0], full[:, 1])
        code = (full[:, 2] + 1) + 3 * (full[:, 3] + 1) + 9 * (full[:, 4] + 1)
        _, g["nbl_hash_shift"] = pair_hash(full[:, 0], full[:, 1] * 27 + code)
        g["nbl_hash_nummer"] = g["nbl_hash_nummer"].astype(np.int32)
    else:
        g["nbl"] = rows.astype(np.int32)
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **g)
    print(f"{name}: {out['natoms']} atoms, cells {g['gdim']}, {len(rows)} pairs, {g['nbl_builds']} builds in the protocol, "
          f"{out['nbl_count']} list builds, {os.path.getsize(path) / 1024:.0f} kB")


def make_potaccess(variant="eam", name="potaccess"):
    """Known answers of PAIR_INT2 / PAIR_INT3 / PAIR_INT_SP through the reference's own macros
    (src/potaccess.h:323-457), as selected by the build."""
    tmp = tempfile.mkdtemp(prefix="gold_pot")
    tabs = synth.make_eam_tables(tmp, "nial", nr=601, nrho=801)
    p = synth.nial_param(tmp, ncell=(5, 5, 5), tables=tabs)
    import pickle, subprocess
    code = f"""
import sys, pickle, numpy as np
sys.path.insert(0, {ROOT!r})
from oracle import ref_driver as rd
sim = rd.RefIMD({variant!r}, {p!r})
rng = np.random.default_rng(7)
out = {{}}
for which, ncol, lo, hi in ((0, 4, 0.5, 31.0), (2, 4, 0.5, 31.0), (1, 2, -1.0, 45.0)):
    for col in range(ncol):
        x = np.concatenate([rng.uniform(lo, hi, 300), [lo, hi, 1.0, 30.25, 30.249999999, 0.0, 40.0]])
        v, g = sim.pair_int(which, col, x)
        out[(which, col)] = (x, v, g)
pickle.dump(out, open({os.path.join(tmp, 'pa.pkl')!r}, 'wb'))
"""
    subprocess.check_call([sys.executable, "-c", code], cwd=tmp, stdout=subprocess.DEVNULL)
    out = pickle.load(open(os.path.join(tmp, "pa.pkl"), "rb"))
    g = dict(table_arrays(tabs))
    g["ntypes"] = 2
    g["interp"] = {"eam": "3point", "eam_4point": "4point", "eam_spline": "spline"}[variant]
    for (which, col), (x, v, gr) in out.items():
        g[f"x:{which}:{col}"] = x; g[f"v:{which}:{col}"] = v; g[f"g:{which}:{col}"] = gr
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **g)
    print(name + ": done")


if __name__ == "__main__":
    POT = {"potaccess": "eam", "potaccess_4point": "eam_4point", "potaccess_spline": "eam_spline"}
    names = sys.argv[1:] or list(CASES) + list(POT)
    for n in names:
        if n in POT:
            make_potaccess(POT[n], n)
        else:
            make_case(n, CASES[n])
