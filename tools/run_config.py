#!/usr/bin/env python
"""Run one of BASELINE.json's other configurations on one GPU and print phase timings plus the
size-independent checks the GPU tests use (not a bench line: bench.py keeps configs[1]).

    python tools/run_config.py nial  --ncell 200   # cfg 3: binary EAM Ni-Al B2, 16M atoms, NVT
    python tools/run_config.py cu    --ncell 160   # cfg 4 sizes: 16.4M (160^3) / 32M (200^3) atoms on one GPU
    python tools/run_config.py deform --ncell 100  # cfg 5 flavour: uniaxial lin_deform every 10 steps
    python tools/run_config.py lj    --ncell 20    # cfg 1: LJ Ar 32k atoms, tabulated pair potential
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def b2_lattice(ncell, a0):
    nx, ny, nz = ncell
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cells = np.stack([ix, iy, iz], -1).reshape(-1, 1, 3).astype(np.float64)
    base = np.array([[0.25, 0.25, 0.25], [0.75, 0.75, 0.75]])
    ort = ((cells + base[None]) * a0).reshape(-1, 3)
    typ = np.tile(np.array([0, 1], np.int32), nx * ny * nz)
    return ort, typ, np.diag([nx * a0, ny * a0, nz * a0]).astype(np.float64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["nial", "cu", "deform", "lj"])
    ap.add_argument("--ncell", type=int, default=100)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=30)
    args = ap.parse_args()
    import torch
    from imd_b200 import api, synth
    tmp = tempfile.mkdtemp(prefix="imdb200_cfg_")
    nc = (args.ncell,) * 3
    kw = {}
    if args.config == "nial":
        tabs = synth.make_eam_tables(tmp, "nial")
        ort, typ, box = b2_lattice(nc, 2.88)
        masse = np.where(typ == 0, synth.NI_MASS, synth.AL_MASS)
        nt = 2
        kw = dict(ensemble="nvt", temperature=0.05, isq_tau_eta=100.0)
    elif args.config == "lj":
        tab = synth.make_lj_table(tmp)
        ort, box = synth.fcc_lattice(nc, synth.AR_A0)
        typ = np.zeros(len(ort), np.int32); masse = np.full(len(ort), synth.AR_MASS); nt = 2
        kw = dict(ensemble="nve", timestep=0.002)
    else:
        tabs = synth.make_eam_tables(tmp, "cu")
        ort, box = synth.fcc_lattice(nc, synth.CU_A0)
        typ = np.zeros(len(ort), np.int32); masse = np.full(len(ort), synth.CU_MASS); nt = 1
        kw = dict(ensemble="nve")
    n = len(ort)
    t0 = 0.0043 if args.config == "lj" else 0.05
    p = synth.maxwell_momenta(n, masse, t0, 7)
    if args.config == "lj":
        sim = api.IMDB200(nt, box, pair=tab["potfile"], nbl_size=1.2, **kw)
    else:
        sim = api.IMDB200(nt, box, pair=tabs["core_potential_file"], embed=tabs["embedding_energy_file"],
                          rho=tabs["atomic_e-density_file"], nbl_size=1.2, timestep=0.001, **kw)
    sim.set_atoms(np.arange(n, dtype=np.int32), typ, masse, ort, p)

    def run(k):
        if args.config != "deform":
            sim.run(k)
            return
        for s in range(k):                      # uniaxial x strain every 10 steps (lindef_interval 10, lindef_size 1e-4)
            if s % 10 == 0:
                sim.lin_deform([1, 0, 0], [0, 0, 0], [0, 0, 0], 1e-4)
            sim.run(1)

    run(args.warmup)
    sim.timers(reset=True)
    sc0 = sim.scalars()
    torch.cuda.synchronize()
    t = time.perf_counter()
    run(args.steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    tm = sim.timers()
    sc = sim.scalars()
    a = sim.atoms(sort=False)
    fsum = np.abs(a["kraft"].sum(axis=0)).max() / (np.abs(a["kraft"]).max() * np.sqrt(n))
    psum = np.abs(a["impuls"].sum(axis=0)).max()
    e0 = sc0["tot_pot_energy"] + sc0["tot_kin_energy"]; e1 = sc["tot_pot_energy"] + sc["tot_kin_energy"]
    steps_t = max(tm["steps"], 1)
    print(json.dumps({
        "config": args.config, "atoms": n, "steps": args.steps, "ms_per_step": 1e3 * dt / args.steps,
        "atom_steps_per_s": n * args.steps / dt,
        "phase_ms_per_step": {k: tm[k] / steps_t for k in ("rebuild_ms", "pass1_ms", "pass2_ms", "integrate_ms", "ghost_ms")},
        "rebuilds": int(tm["rebuilds"]), "nbl_len_per_atom": sim.raw_scalars().nbl_len / n,
        "sum_F_rel": fsum, "sum_p": psum, "dE_per_atom": (e1 - e0) / n, "T": 2 * sc["tot_kin_energy"] / (3 * n),
        "mem_GB": torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9}))
    sim.close()


if __name__ == "__main__":
    main()
