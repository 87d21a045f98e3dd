"""Shared test protocol: run the same sequence of calc_forces / move_atoms / check_nblist on any
implementation (oracle, CUDA engine) from a golden fixture's start state and compare with what the
reference recorded (tools/make_golden.py)."""
from __future__ import annotations

import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

TABLE_KEYS = {"core_potential_file": "pair", "embedding_energy_file": "embed",
              "atomic_e-density_file": "rho", "potfile": "pair", "eeam_energy_file": "emod",
              "adp_upotfile": "adp_u", "adp_wpotfile": "adp_w"}

# parity bars (BASELINE.json north_star): neighbour sets bit-exact; forces, energies and pressure
# within 1e-10 relative of IMD's CPU build.
RTOL = 1e-10


def load_golden(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def write_tables(g, outdir):
    """Re-create the IMD-format potential files stored in a fixture; returns {pair,embed,rho: path}."""
    os.makedirs(outdir, exist_ok=True)
    paths = {}
    for k, v in g.items():
        if k.startswith("table:"):
            key = k[len("table:"):]
            p = os.path.join(outdir, key.replace("/", "_") + ".pot")
            with open(p, "wb") as f:
                f.write(v.tobytes())
            paths[TABLE_KEYS[key]] = p
    return paths


def make_sim(factory, g, tabdir, **kw):
    """factory = oracle.oracle.OracleIMD or imd_b200.api.IMDB200."""
    paths = write_tables(g, tabdir)
    ens = str(g["ensemble"])
    common = dict(pbc=tuple(int(x) for x in g["pbc"]), nbl_margin=0.4, pair=paths["pair"],
                  embed=paths.get("embed"), rho=paths.get("rho"))
    if "emod" in paths:                                    # fixture of an `eeam` reference build
        common["emod"] = paths["emod"]
    if "adp_u" in paths:                                   # fixture of an `adp` reference build (oracle only so far)
        common["adp_u"] = paths["adp_u"]; common["adp_w"] = paths["adp_w"]
    if "interp" in g and str(g["interp"]) != "3point":     # fixture of a `4point` / `spline` reference build
        common["interp"] = str(g["interp"])
    integ = dict(ensemble=ens, timestep=float(g["timestep"]), temperature=float(g["temperature"]),
                 eta=float(g["eta0"]), isq_tau_eta=float(g["isq_tau_eta"]))
    if "total_types" in g and factory.__name__ != "OracleIMD":
        common["total_types"] = int(g["total_types"])
    try:
        sim = factory(int(g["ntypes"]), g["box"], **common, **integ, **kw)
    except TypeError:
        sim = factory(int(g["ntypes"]), g["box"], **common, **kw)
        sim.set_integrator(**integ)
    if "restrictions" in g:                                # restrictionvector per virtual type
        sim.set_restrictions(g["restrictions"])
    sim.set_atoms(g["start:nummer"], g["start:sorte"], g["start:masse"], g["start:ort"], g["start:impuls"],
                  vsorte=g["start:vsorte"])
    return sim


def run_protocol(sim, g):
    """Mirror of oracle/ref_driver.run_protocol for the frames stored in the fixture."""
    press = bool(int(g["press"]))
    sim.set_press_calc(press)
    nsteps = int(g["nsteps"])
    rec = {int(x) for x in g["record"]}
    out = dict(epot=[], virial=[], ekin=[], eta=[], valid=[], atoms={}, nbl=None, tot_presstens={})
    lindef = int(g["lindef_every"]) if "lindef_every" in g else 0
    deform = int(g["deform_every"]) if "deform_every" in g else 0
    for s in range(nsteps):
        # same order as main_loop (src/imd_main_3d.c:293-326)
        if s > 0 and lindef and s % lindef == 0:
            sim.lin_deform(g["lindef_x"], g["lindef_y"], g["lindef_z"], float(g["lindef_size"]))
        if s > 0 and deform and s % deform == 0:
            sim.deform_sample(float(g["deform_size"]), g["deform_shift"], g["shear_def"], g["deform_shear"], g["deform_base"])
            sim.check_nblist()
        sim.calc_forces(s)
        sc = sim.scalars()
        out["epot"].append(sc["tot_pot_energy"]); out["virial"].append(sc["virial"])
        if s in rec:
            out["atoms"][s] = sim.atoms()
        if s == 0:
            out["nbl"] = sim.nbl_pairs()
        sim.move_atoms()
        sim.check_nblist()
        sc = sim.scalars()
        out["ekin"].append(sc["tot_kin_energy"]); out["eta"].append(sc["eta"])
        out["valid"].append(sim.have_valid_nbl)
        if press and s in rec:
            out["tot_presstens"][s] = sim.tot_presstens()
    out["final"] = sim.atoms()
    out["nbl_count"] = sim.nbl_count
    return out


def relerr(a, b):
    """max |a-b| / max |b| -- error relative to the scale of the reference array."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    scale = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (scale if scale > 0 else 1.0))


# Per-component bar: |a-b| <= tol*|b| + FLOOR*tol*max|b|.  The relative part is north_star's "within 1e-10
# relative"; the absolute floor (1e-2 of the tolerance, relative to the largest component: 1e-12*max|ref| at
# tol = 1e-10) covers components that are small because large terms cancel -- a force component of 1e-6 of the
# largest one is still pinned to six digits more than relerr() alone would.
FLOOR = 1e-2


def comperr(a, b):
    """max over components of |a-b| / (|b| + FLOOR*max|b|): <= tol means every component passes the bar above."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    scale = np.max(np.abs(b))
    if scale == 0:
        return float(np.max(np.abs(a)))
    return float(np.max(np.abs(a - b) / (np.abs(b) + FLOOR * scale)))


def symmetric_closure(rows):
    """canonical half-list rows (a,b,sx,sy,sz) -> set of directed entries (i,j,sx,sy,sz)."""
    rows = np.asarray(rows, np.int64)
    fwd = rows
    bwd = np.column_stack([rows[:, 1], rows[:, 0], -rows[:, 2:]])
    allr = np.concatenate([fwd, bwd])
    return np.unique(allr, axis=0)


def compare(out, g, full_list=False, rtol=RTOL, traj_rtol=None, ignore_shift=False):
    """Assert parity of a protocol run with the reference fixture.  Returns a dict of max errors."""
    from oracle.oracle import canonical_pairs
    errs = {}
    n = len(g["start:nummer"])
    # neighbour set at the first build: bit-exact
    pairs, shift = out["nbl"]
    sel = g["sample"] if "sample" in g else slice(None)      # large fixtures: per-atom results for a sample of atoms
    if "nbl_hash" in g:
        # large fixtures store the neighbour set as one 64-bit hash per atom (tools/parity_fixture.pair_hash) of the
        # symmetric closure of the reference's half list, with and without the image shifts
        from tools.parity_fixture import pair_hash
        pairs = np.asarray(pairs, np.int64); shift = np.asarray(shift, np.int64)
        if not (full_list or ignore_shift):                   # a half list (the oracle): close it first
            rows = canonical_pairs(pairs, shift)
            pairs = np.concatenate([rows[:, :2], rows[:, 1::-1]]); shift = np.concatenate([rows[:, 2:], -rows[:, 2:]])
        assert len(pairs) == int(g["nbl_len_full"]), f"neighbour list length {len(pairs)} != {int(g['nbl_len_full'])}"
        if ignore_shift:
            u, h = pair_hash(pairs[:, 0], pairs[:, 1]); want = g["nbl_hash"]
        else:
            code = (shift[:, 0] + 1) + 3 * (shift[:, 1] + 1) + 9 * (shift[:, 2] + 1)
            u, h = pair_hash(pairs[:, 0], pairs[:, 1] * 27 + code); want = g["nbl_hash_shift"]
        assert np.array_equal(u, g["nbl_hash_nummer"]) and np.array_equal(h, want), \
            f"neighbour set differs from the reference for {int(np.sum(h != want)) if len(h) == len(want) else -1} atoms"
    elif ignore_shift:
        ref_rows = g["nbl"].astype(np.int64)
        # domain-decomposed run: a pair across an interior domain face carries no periodic shift on either
        # side, so pairs are compared by atom numbers only (as a multiset)
        want = symmetric_closure(ref_rows)[:, :2]
        got = np.asarray(pairs, np.int64)
        want = want[np.lexsort(want.T[::-1])]
        got = got[np.lexsort(got.T[::-1])]
        assert got.shape == want.shape and np.array_equal(got, want), \
            f"neighbour set differs from the reference: {got.shape} vs {want.shape}"
    elif full_list:
        ref_rows = g["nbl"].astype(np.int64)
        got = np.unique(np.column_stack([pairs.astype(np.int64), shift.astype(np.int64)]), axis=0)
        assert len(got) == len(pairs), "duplicate entries in the full neighbour list"
        want = symmetric_closure(ref_rows)
        assert got.shape == want.shape and np.array_equal(got, want), \
            f"neighbour set differs from the reference: {got.shape} vs {want.shape}"
    else:
        ref_rows = g["nbl"].astype(np.int64)
        got = canonical_pairs(pairs, shift)
        assert got.shape == ref_rows.shape and np.array_equal(got, ref_rows), "neighbour set differs"
    # rebuild decisions are part of the path's integer results
    assert list(out["valid"]) == list(g["valid"]), "check_nblist decisions differ"
    rec = sorted(int(x) for x in g["record"])
    for s in rec:
        a = out["atoms"][s]
        # error growth of a chaotic trajectory: first frame at the parity bar, later ones looser
        tol = rtol if s == rec[0] else (traj_rtol or rtol * 1e3)
        for k in ("kraft", "poteng", "rho", "dF") + (("eam_p", "dM") if f"f{s}:eam_p" in g else ()) \
                + (("adp_mu", "adp_lambda") if f"f{s}:adp_mu" in g else ()):
            if np.max(np.abs(g[f"f{s}:{k}"])) == 0 and np.max(np.abs(a[k][sel])) == 0:
                continue
            e = relerr(a[k][sel], g[f"f{s}:{k}"])
            errs[f"f{s}:{k}"] = e
            assert e <= tol, f"{k} at step {s}: rel err {e:.3e} > {tol:.1e}"
            ec = comperr(a[k][sel], g[f"f{s}:{k}"])             # every component, not only the array's scale
            errs[f"f{s}:{k}:comp"] = ec
            assert ec <= max(tol, RTOL), f"{k} at step {s}: per-component err {ec:.3e} > {max(tol, RTOL):.1e}"
        if int(g["press"]):
            e = relerr(a["presstens"], g[f"f{s}:presstens"])
            errs[f"f{s}:presstens"] = e
            assert e <= tol, f"presstens at step {s}: rel err {e:.3e}"
            e = relerr(out["tot_presstens"][s], g[f"f{s}:tot_presstens"])
            errs[f"f{s}:tot_presstens"] = e
            assert e <= tol, f"tot_presstens at step {s}: rel err {e:.3e}"
    for k in ("epot", "virial", "ekin", "eta"):
        ref = g[k]
        if np.max(np.abs(ref)) == 0:
            continue
        e0 = abs(out[k][0] - ref[0]) / max(abs(ref[0]), 1e-300)
        errs[k + "[0]"] = e0
        assert e0 <= rtol, f"{k} at step 0: rel err {e0:.3e}"
        e = relerr(out[k], ref)
        errs[k] = e
        assert e <= (traj_rtol or rtol * 1e3), f"{k} over the run: rel err {e:.3e}"
    box = g["box"]
    d = out["final"]["ort"][sel] - g["final:ort"]
    # positions are only wrapped at rebuilds: compare modulo the box
    frac = d @ np.linalg.inv(box)
    d = (frac - np.round(frac)) @ box
    e = float(np.max(np.abs(d)) / np.max(np.abs(box)))
    errs["final:ort"] = e
    assert e <= (traj_rtol or rtol * 1e3), f"final positions: rel err {e:.3e}"
    # Number of list builds over the protocol (nbl_count, src/globals.h:421).  The fixtures' own `nbl_count` also
    # counts the builds of the reference's thermalisation before the recorded start state; newer fixtures store the
    # protocol's share as `nbl_builds`, for the others it follows from the recorded decisions: one build at the first
    # calc_forces plus one after every step that ended with have_valid_nbl == 0 (runs without lin_deform /
    # deform_sample, which invalidate the list on their own, src/imd_deform.c:107-117).
    if "nbl_builds" in g:
        want_builds = int(g["nbl_builds"])
    elif "lindef_every" not in g and "deform_every" not in g:
        want_builds = 1 + int(np.sum(np.asarray(g["valid"])[:-1] == 0))
    else:
        want_builds = None
    if want_builds is not None:
        assert out["nbl_count"] == want_builds, f"nbl_count {out['nbl_count']} != {want_builds}"
    errs["nbl_builds"] = float(out["nbl_count"])
    return errs
