"""Parity of the CUDA engine on an N-rank process grid against a committed reference fixture (tests/golden/*.npz,
generated from the unmodified reference by tools/make_golden.py).  Used by bench.py --gpus N > 1 for its
`parity_check` key (outside the timed region) so that multi-GPU correctness is visible in every scaling record,
and by tests/.  No oracle code is involved: the fixture holds what the reference itself computed.

All ranks call check(); rank 0 gets the result dict, the other ranks get None.
"""
from __future__ import annotations

import os
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
TABLE_KEYS = {"core_potential_file": "pair", "embedding_energy_file": "embed", "atomic_e-density_file": "rho",
              "potfile": "pair", "eeam_energy_file": "emod", "adp_upotfile": "adp_u", "adp_wpotfile": "adp_w"}


def _tables(g, outdir):
    paths = {}
    for k in g:
        if k.startswith("table:"):
            key = k[len("table:"):]
            p = os.path.join(outdir, key.replace("/", "_") + ".pot")
            with open(p, "wb") as f:
                f.write(g[k].tobytes())
            paths[TABLE_KEYS[key]] = p
    return paths


def _rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    s = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (s if s > 0 else 1.0))


def _comp(a, b, floor=1e-2):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    s = np.max(np.abs(b))
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor * s))) if s > 0 else float(np.max(np.abs(a)))


def pair_hash(ni, nj):
    """Order-independent 64-bit hash of every atom's neighbour set: sum over its neighbours j of mix(nummer_j).
    Returns {nummer_i: hash} as two arrays (sorted nummer, hash)."""
    x = (np.asarray(nj, np.uint64) + np.uint64(0x9E3779B97F4A7C15)) * np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(31)
    x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(29)
    ni = np.asarray(ni, np.int64)
    o = np.argsort(ni, kind="stable")
    ni, x = ni[o], x[o]
    u, start = np.unique(ni, return_index=True)
    h = np.add.reduceat(x, start) if len(x) else np.zeros(0, np.uint64)
    return u, h


def check(name, grid, device, max_steps=None):
    """Run fixture `name` over the process grid `grid` (product = world size; (1,1,1) without torch.distributed)
    and compare with the reference's record: neighbour set (exact), rebuild decisions (exact), forces / energies /
    densities at the first frame, Epot / virial / Ekin per step, final positions."""
    from imd_b200 import api
    from imd_b200 import dist as idist
    world = int(np.prod(grid))
    if world > 1:
        import torch.distributed as dist
        rank = dist.get_rank()
    else:
        rank = 0
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    big = "sample" in g                       # large fixtures store per-atom data for a sample of atoms only
    tmp = tempfile.mkdtemp(prefix=f"pf{rank}_")
    paths = _tables(g, tmp)
    kw = dict(pbc=tuple(int(x) for x in g["pbc"]), nbl_margin=0.4, pair=paths["pair"], embed=paths.get("embed"),
              rho=paths.get("rho"), ensemble=str(g["ensemble"]), timestep=float(g["timestep"]),
              temperature=float(g["temperature"]), eta=float(g["eta0"]), isq_tau_eta=float(g["isq_tau_eta"]),
              interp=str(g["interp"]) if "interp" in g else "3point", emod=paths.get("emod"), device=device)
    if world > 1:
        sim = idist.create(int(g["ntypes"]), g["box"], cpu_dim=tuple(grid), **kw)
    else:
        sim = api.IMDB200(int(g["ntypes"]), g["box"], **kw)
    sim.set_atoms(g["start:nummer"], g["start:sorte"], g["start:masse"], g["start:ort"], g["start:impuls"],
                  vsorte=g["start:vsorte"])
    nsteps = int(g["nsteps"]) if max_steps is None else min(int(g["nsteps"]), max_steps)
    rec0 = int(sorted(int(x) for x in g["record"])[0])
    epot, vir, ekin, valid = [], [], [], []
    frame0 = None
    nbl = None
    for s in range(nsteps):
        sim.calc_forces(s)
        sc = sim.scalars()
        epot.append(sc["tot_pot_energy"]); vir.append(sc["virial"])
        if s == rec0:
            frame0 = idist.gather_atoms(sim) if world > 1 else sim.atoms()
        if s == 0:
            nbl = idist.gather_nbl(sim) if world > 1 else sim.nbl_pairs()[0]
        sim.move_atoms()
        sim.check_nblist()
        ekin.append(sim.scalars()["tot_kin_energy"])
        valid.append(sim.have_valid_nbl)
    final = idist.gather_atoms(sim) if world > 1 else sim.atoms()
    builds = sim.nbl_count
    sim.close()
    if rank != 0:
        return None
    res = {"fixture": name, "grid": list(grid), "atoms": int(len(g["start:nummer"])), "steps": nsteps}
    # neighbour set, by atom numbers (image shifts are per rank), exact
    if "nbl_hash" in g:
        u, h = pair_hash(nbl[:, 0], nbl[:, 1])
        res["nbl_equal"] = bool(np.array_equal(u, g["nbl_hash_nummer"]) and np.array_equal(h, g["nbl_hash"])
                                and len(nbl) == int(g["nbl_len_full"]))
    else:
        rows = g["nbl"].astype(np.int64)
        want = np.concatenate([rows[:, :2], rows[:, 1::-1]])
        want = want[np.lexsort(want.T[::-1])]
        got = np.asarray(nbl, np.int64)
        got = got[np.lexsort(got.T[::-1])]
        res["nbl_equal"] = bool(got.shape == want.shape and np.array_equal(got, want))
    res["rebuild_decisions_equal"] = bool(list(valid) == [int(x) for x in g["valid"][:nsteps]])
    res["list_builds"] = int(builds)
    errs = {}
    sel = g["sample"] if big else slice(None)        # indices into the nummer-sorted atom arrays
    for k in ("kraft", "poteng", "rho", "dF"):
        ref = g[f"f{rec0}:{k}"]
        if np.max(np.abs(ref)) == 0:
            continue
        errs[k] = max(_rel(frame0[k][sel], ref), _comp(frame0[k][sel], ref))
    errs["epot[0]"] = abs(epot[0] - g["epot"][0]) / abs(g["epot"][0])
    errs["virial[0]"] = abs(vir[0] - g["virial"][0]) / max(abs(g["virial"][0]), 1e-300)
    errs["ekin[0]"] = abs(ekin[0] - g["ekin"][0]) / abs(g["ekin"][0])
    res["first_frame_max_rel_err"] = float(max(errs.values()))
    traj = {"epot": _rel(epot, g["epot"][:nsteps]), "virial": _rel(vir, g["virial"][:nsteps]),
            "ekin": _rel(ekin, g["ekin"][:nsteps])}
    if max_steps is None or max_steps >= int(g["nsteps"]):
        box = g["box"]
        d = final["ort"][sel] - g["final:ort"]
        frac = d @ np.linalg.inv(box)
        d = (frac - np.round(frac)) @ box
        traj["final_ort"] = float(np.max(np.abs(d)) / np.max(np.abs(box)))
    res["trajectory_max_rel_err"] = float(max(traj.values()))
    res["errors"] = {k: float(v) for k, v in {**errs, **traj}.items()}
    res["ok"] = bool(res["nbl_equal"] and res["rebuild_decisions_equal"] and res["first_frame_max_rel_err"] <= 1e-10
                     and res["trajectory_max_rel_err"] <= 1e-8)
    return res
