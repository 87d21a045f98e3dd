#!/usr/bin/env python
"""Run one of BASELINE.json's other configurations on N GPUs (one process per GPU under torchrun, N = 1 without)
and print ONE JSON line (rank 0): device-timed throughput (CUDA events on the launching stream, max over ranks),
phase timings and the size-independent checks the GPU tests use.  Not a bench line: bench.py keeps configs[1].

    python tools/run_config.py nial --ncell 200 200 200                 # cfg 3: binary EAM Ni-Al B2, 16 M atoms, NVT
    torchrun --nproc-per-node 8 tools/run_config.py cu --ncell 200 200 100   # cfg 4: weak scaling, 16 M atoms PER GPU
    torchrun --nproc-per-node 8 tools/run_config.py deform --ncell 400 200 200 --strong
                                                                        # cfg 5: 64 M atoms IN TOTAL, uniaxial lin_deform
                                                                        #        every 10 steps (lindef_size 1e-4), strong scaling
    python tools/run_config.py lj --ncell 20 20 20                      # cfg 1: LJ Ar 32k atoms, tabulated pair potential
--ncell is per GPU (weak scaling, like `size_per_cpu 1`, src/imd_generate.c:292-296) unless --strong is given.
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def b2_block(ncell, a0, origin):
    nx, ny, nz = ncell
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cells = np.stack([ix, iy, iz], -1).reshape(-1, 1, 3).astype(np.float64) + np.asarray(origin, np.float64)
    base = np.array([[0.25, 0.25, 0.25], [0.75, 0.75, 0.75]])
    ort = ((cells + base[None]) * a0).reshape(-1, 3)
    typ = np.tile(np.array([0, 1], np.int32), nx * ny * nz)
    return ort, typ


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["nial", "cu", "deform", "lj"])
    ap.add_argument("--ncell", type=int, nargs=3, default=[100, 100, 100])
    ap.add_argument("--strong", action="store_true", help="--ncell is the whole crystal, split over the ranks")
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--thermal", type=int, default=100, help="untimed thermalisation steps (setup)")
    ap.add_argument("--lindef-int", type=int, default=10)
    ap.add_argument("--lindef-size", type=float, default=1e-4)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from imd_b200 import api, synth
    from imd_b200 import dist as idist
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    grid = idist.grid_for(world)
    coord = np.array(api.cart_coords(rank, grid))
    total = np.array(args.ncell) if args.strong else np.array(args.ncell) * np.array(grid)
    if np.any(total % np.array(grid)):
        raise SystemExit(f"crystal {total} does not split over the process grid {grid}")
    mine = total // np.array(grid)                      # unit cells of this rank's block
    origin = coord * mine
    tmp = tempfile.mkdtemp(prefix="imdb200_cfg_")
    kw = {}
    if args.config == "nial":
        tabs = synth.make_eam_tables(tmp, "nial")
        a0 = 2.88
        ort, typ = b2_block(mine, a0, origin)
        masse = np.where(typ == 0, synth.NI_MASS, synth.AL_MASS)
        nt = 2
        kw = dict(ensemble="nvt", temperature=0.05, isq_tau_eta=100.0)
    elif args.config == "lj":
        tab = synth.make_lj_table(tmp)
        a0 = synth.AR_A0
        ort, _ = synth.fcc_lattice(tuple(mine), a0)
        ort = ort + origin * a0
        typ = np.zeros(len(ort), np.int32); masse = np.full(len(ort), synth.AR_MASS); nt = 2
        kw = dict(ensemble="nve", timestep=0.002)
    else:
        tabs = synth.make_eam_tables(tmp, "cu")
        a0 = synth.CU_A0
        ort, _ = synth.fcc_lattice(tuple(mine), a0)
        ort = ort + origin * a0
        typ = np.zeros(len(ort), np.int32); masse = np.full(len(ort), synth.CU_MASS); nt = 1
        kw = dict(ensemble="nve")
    box = np.diag(total * a0).astype(np.float64)
    n = len(ort)
    # NVT (nial): twice the thermostat's target, so that after equipartition the crystal sits AT the target and the timed
    # window is not a heating ramp (list lengths and rebuild cadence drift while the thermostat pumps energy in)
    t0 = 0.0043 if args.config == "lj" else (0.10 if args.config == "nial" else 0.05)
    p = synth.maxwell_momenta(n, masse, t0, 7 + rank)
    num = (np.arange(n, dtype=np.int64) + rank * n).astype(np.int32)
    if args.config == "lj":
        kw.update(pair=tab["potfile"], nbl_size=1.2)
    else:
        kw.update(pair=tabs["core_potential_file"], embed=tabs["embedding_energy_file"], rho=tabs["atomic_e-density_file"],
                  nbl_size=1.2, timestep=0.001)
    sim = idist.create(nt, box, cpu_dim=grid, device=local, **kw) if world > 1 else api.IMDB200(nt, box, device=local, **kw)
    stream = torch.cuda.current_stream()
    sim.set_stream(stream.cuda_stream)
    sim.set_atoms(num, typ, masse, ort, p)
    del ort, p
    step_no = [0]

    def run(k):
        if args.config != "deform":
            sim.run(k)
            return
        # main_loop order (src/imd_main_3d.c:293-299): lin_deform when steps % lindef_int == 0, then the step
        while k > 0:
            if step_no[0] % args.lindef_int == 0:
                sim.lin_deform([1, 0, 0], [0, 0, 0], [0, 0, 0], args.lindef_size)
            chunk = min(k, args.lindef_int - step_no[0] % args.lindef_int)
            sim.run(chunk)
            step_no[0] += chunk; k -= chunk

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    run(args.thermal)
    run(args.warmup)
    sim.timers(reset=True)
    sc0 = sim.scalars()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    run(args.steps)
    e1.record(stream)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    tm = sim.timers()
    sc = sim.scalars()
    nloc = sim.natoms
    kraft = np.zeros((nloc, 3)); impuls = np.zeros((nloc, 3))       # only what the checks need (32 M atoms per GPU)
    got = sim.L.imdb200_get_atoms(sim.h, None, None, None, None, None, impuls.ctypes.data, kraft.ctypes.data, *([None] * 5))
    assert got == nloc
    fs = torch.tensor(np.concatenate([kraft.sum(axis=0), impuls.sum(axis=0), [np.abs(kraft).max()]]), device="cuda")
    fmax = fs[6:].clone()
    if world > 1:
        dist.all_reduce(fs[:6]); dist.all_reduce(fmax, op=dist.ReduceOp.MAX)
    natoms = int(sc["nactive"] // 3) if False else world * n
    free, tot = torch.cuda.mem_get_info()
    if rank == 0:
        e_0 = sc0["tot_pot_energy"] + sc0["tot_kin_energy"]; e_1 = sc["tot_pot_energy"] + sc["tot_kin_energy"]
        steps_t = max(tm["steps"], 1)
        print(json.dumps({
            "config": args.config, "n_gpus": world, "cpu_dim": list(grid), "atoms": natoms, "atoms_per_gpu": natoms // world,
            "scaling": "strong" if args.strong else "weak", "steps": args.steps, "ms_per_step": ms / args.steps,
            "atom_steps_per_s": natoms * args.steps / (ms * 1e-3),
            "phase_ms_per_step": {k: tm[k] / steps_t for k in ("rebuild_ms", "pass1_ms", "pass2_ms", "integrate_ms", "ghost_ms")},
            "rebuilds": int(tm["rebuilds"]), "nbl_len_per_atom": sim.raw_scalars().nbl_len / max(nloc, 1),
            "lin_deform": {"interval": args.lindef_int, "size": args.lindef_size, "box_x": float(sim.box()[0, 0]),
                           "box_x_start": float(box[0, 0])} if args.config == "deform" else None,
            "sum_F_rel": float(fs[:3].abs().max() / (fmax[0] * np.sqrt(natoms))), "sum_p": float(fs[3:6].abs().max()),
            "dE_per_atom": (e_1 - e_0) / natoms, "T": 2 * sc["tot_kin_energy"] / (3 * natoms),
            "mem_GB_rank0": (tot - free) / 1e9, "timing": "CUDA events on the launching stream, max over ranks"}))
    sim.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
