"""Worker of the multi-GPU tests: launched by tests/test_multi_gpu.py as
    python -m torch.distributed.run --nproc-per-node N tests/mgpu_worker.py <case> <px> <py> <pz>
One process per GPU; torch.distributed (gloo) only carries the ncclUniqueId and gathers results for the
comparison on rank 0 -- the halo exchange itself runs inside libimd_b200.so over NCCL."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import common  # noqa: E402
from imd_b200 import api, synth  # noqa: E402
from imd_b200 import dist as idist  # noqa: E402


class DistSim:
    """Presents the N domains as one simulation to tests/common.run_protocol."""

    def __init__(self, sim):
        self.sim = sim

    def __getattr__(self, k):
        return getattr(self.sim, k)

    def atoms(self, sort=True):
        return idist.gather_atoms(self.sim)

    def nbl_pairs(self):
        pr = idist.gather_nbl(self.sim)
        return pr, np.zeros((len(pr), 3), np.int8)


def make(g, tabdir, grid, **kw):
    paths = common.write_tables(g, tabdir)
    sim = idist.create(int(g["ntypes"]), g["box"], cpu_dim=grid, device=int(os.environ.get("LOCAL_RANK", "0")),
                       pbc=tuple(int(x) for x in g["pbc"]), nbl_margin=0.4, pair=paths["pair"],
                       embed=paths.get("embed"), rho=paths.get("rho"), ensemble=str(g["ensemble"]),
                       timestep=float(g["timestep"]), temperature=float(g["temperature"]), eta=float(g["eta0"]),
                       isq_tau_eta=float(g["isq_tau_eta"]),
                       total_types=int(g["total_types"]) if "total_types" in g else None,
                       interp=str(g["interp"]) if "interp" in g else "3point", emod=paths.get("emod"),
                       adp_u=paths.get("adp_u"), adp_w=paths.get("adp_w"), **kw)
    if "restrictions" in g:
        sim.set_restrictions(g["restrictions"])
    # every rank is handed ALL atoms and keeps those of its own domain
    sim.set_atoms(g["start:nummer"], g["start:sorte"], g["start:masse"], g["start:ort"], g["start:impuls"],
                  vsorte=g["start:vsorte"])
    return sim


def case_fixture(name, grid, rank):
    g = common.load_golden(name)
    sim = make(g, tempfile.mkdtemp(prefix=f"mg{rank}_"), grid)
    out = common.run_protocol(DistSim(sim), g)
    if rank == 0:
        errs = common.compare(out, g, full_list=True, rtol=1e-10, traj_rtol=1e-8, ignore_shift=True)
        print("errors", {k: f"{v:.1e}" for k, v in errs.items()})
    return sim


def case_migration(grid, rank, pbc=(1, 1, 1)):
    """Hot crystal, 16 384 atoms, 120 steps: atoms change domains; compare with the same run on ONE GPU.
    pbc = (1, 1, 0): a slab with free surfaces in z; split along z, an atom that moves from the upper into the lower
    domain must be handed DOWN although (me + 1) % 2 names the same rank (cells.cu, k_wrap_bin)."""
    tmp = tempfile.mkdtemp(prefix=f"mig{rank}_")
    tabs = synth.make_eam_tables(tmp, "cu", nr=601, nrho=801)
    ort, box = synth.fcc_lattice((16, 16, 16), synth.CU_A0)
    if pbc[2] == 0:
        box = box.copy(); box[2, 2] *= 1.5; ort = ort + np.array([0.0, 0.0, 0.25 * 16 * synth.CU_A0])   # vacuum above and below
    n = len(ort)
    m = np.full(n, synth.CU_MASS)
    p = synth.maxwell_momenta(n, m, 0.35, 7)          # ~4000 K: diffusion across the domain faces
    num, typ = np.arange(n, dtype=np.int32), np.zeros(n, np.int32)
    kw = dict(pair=tabs["core_potential_file"], embed=tabs["embedding_energy_file"], rho=tabs["atomic_e-density_file"],
              ensemble="nve", timestep=0.001, pbc=pbc)
    sim = idist.create(1, box, cpu_dim=grid, device=int(os.environ.get("LOCAL_RANK", "0")), **kw)
    sim.set_atoms(num, typ, m, ort, p)
    counts = []
    for _ in range(6):
        sim.run(20)
        counts.append(sim.natoms)
    a = idist.gather_atoms(sim)
    sc = sim.scalars()
    allc = [None] * dist.get_world_size()
    dist.all_gather_object(allc, counts)
    if rank == 0:
        ref = api.IMDB200(1, box, device=0, **kw)
        ref.set_atoms(num, typ, m, ort, p)
        ref.run(120)
        b = ref.atoms()
        rs = ref.scalars()
        assert len(a["nummer"]) == n and np.array_equal(a["nummer"], b["nummer"]), "atoms lost or duplicated"
        assert any(len(set(c)) > 1 for c in allc), f"no atom changed its domain: {allc}"
        d = a["ort"] - b["ort"]
        frac = d @ np.linalg.inv(box)
        d = (frac - np.round(frac)) @ box
        print("migration: per-rank atom counts", allc, "max |dx|", np.abs(d).max())
        assert np.abs(d).max() < 1e-6
        assert abs(sc["tot_pot_energy"] - rs["tot_pot_energy"]) < 1e-8 * abs(rs["tot_pot_energy"])
        assert abs(sc["tot_kin_energy"] - rs["tot_kin_energy"]) < 1e-8 * abs(rs["tot_kin_energy"])
        assert sim.nbl_count == ref.nbl_count, (sim.nbl_count, ref.nbl_count)
        ref.close()
    return sim


def case_overlap(grid, rank):
    """imdb200_run over the process grid (overlapped peer-memory halo by default) against the same run on one GPU."""
    tmp = tempfile.mkdtemp(prefix=f"ov{rank}_")
    tabs = synth.make_eam_tables(tmp, "cu")
    ort, box = synth.fcc_lattice((32, 32, 32), synth.CU_A0)
    n = len(ort)
    rng = np.random.default_rng(3)
    ort = ort + rng.normal(0, 0.05, ort.shape)
    m = np.full(n, synth.CU_MASS)
    p = synth.maxwell_momenta(n, m, 0.12, 9)
    num, typ = np.arange(n, dtype=np.int32), np.zeros(n, np.int32)
    kw = dict(pair=tabs["core_potential_file"], embed=tabs["embedding_energy_file"], rho=tabs["atomic_e-density_file"],
              ensemble="nve", timestep=0.001)
    sim = idist.create(1, box, cpu_dim=grid, device=int(os.environ.get("LOCAL_RANK", "0")), **kw)
    sim.set_atoms(num, typ, m, ort, p)
    for _ in range(3):
        sim.run(20)
    a = idist.gather_atoms(sim)
    sc = sim.scalars()
    if rank == 0:
        ref = api.IMDB200(1, box, device=0, **kw)
        ref.set_atoms(num, typ, m, ort, p)
        ref.run(60)
        b = ref.atoms()
        rs = ref.scalars()
        assert np.array_equal(a["nummer"], b["nummer"])
        d = a["ort"] - b["ort"]
        frac = d @ np.linalg.inv(box)
        d = (frac - np.round(frac)) @ box
        print("overlap: max |dx|", np.abs(d).max(), "max |dp|", np.abs(a["impuls"] - b["impuls"]).max(), "builds", sim.nbl_count)
        assert np.abs(d).max() < 1e-9 and np.abs(a["impuls"] - b["impuls"]).max() < 1e-10
        assert np.max(np.abs(a["kraft"] - b["kraft"])) < 1e-8 * np.max(np.abs(b["kraft"]))
        assert abs(sc["tot_pot_energy"] - rs["tot_pot_energy"]) < 1e-10 * abs(rs["tot_pot_energy"])
        assert abs(sc["tot_kin_energy"] - rs["tot_kin_energy"]) < 1e-9 * abs(rs["tot_kin_energy"])
        assert sim.nbl_count == ref.nbl_count and sim.nbl_count >= 3
        ref.close()
    return sim


def case_overlap_nial(grid, rank):
    """Two species through imdb200_run over the process grid: pass 2 with the neighbour's type in the list entry, move_atoms
    fused into its tail, boundary-first split launches and the peer-memory halo -- against the same run on one GPU."""
    tmp = tempfile.mkdtemp(prefix=f"ovn{rank}_")
    tabs = synth.make_eam_tables(tmp, "nial", nr=601, nrho=801)
    nc, a0 = (28, 28, 28), 2.88
    ix, iy, iz = np.meshgrid(np.arange(nc[0]), np.arange(nc[1]), np.arange(nc[2]), indexing="ij")
    cells = np.stack([ix, iy, iz], -1).reshape(-1, 1, 3).astype(np.float64)
    ort = ((cells + np.array([[0.25, 0.25, 0.25], [0.75, 0.75, 0.75]])[None]) * a0).reshape(-1, 3)
    n = len(ort)
    typ = (np.arange(n) % 2).astype(np.int32)
    ort = ort + np.random.default_rng(4).normal(0, 0.04, ort.shape)
    m = np.where(typ == 0, synth.NI_MASS, synth.AL_MASS)
    p = synth.maxwell_momenta(n, m, 0.10, 11)
    box = np.diag([nc[0] * a0, nc[1] * a0, nc[2] * a0]).astype(np.float64)
    num = np.arange(n, dtype=np.int32)
    kw = dict(pair=tabs["core_potential_file"], embed=tabs["embedding_energy_file"], rho=tabs["atomic_e-density_file"],
              ensemble="nvt", timestep=0.001, temperature=0.10, eta=0.0, isq_tau_eta=100.0)
    sim = idist.create(2, box, cpu_dim=grid, device=int(os.environ.get("LOCAL_RANK", "0")), **kw)
    sim.set_atoms(num, typ, m, ort, p)
    for _ in range(3):
        sim.run(20)
    a = idist.gather_atoms(sim)
    sc = sim.scalars()
    if rank == 0:
        ref = api.IMDB200(2, box, device=0, **kw)
        ref.set_atoms(num, typ, m, ort, p)
        ref.run(60)
        b = ref.atoms()
        rs = ref.scalars()
        assert np.array_equal(a["nummer"], b["nummer"])
        d = a["ort"] - b["ort"]
        frac = d @ np.linalg.inv(box)
        d = (frac - np.round(frac)) @ box
        print("overlap_nial: max |dx|", np.abs(d).max(), "max |dp|", np.abs(a["impuls"] - b["impuls"]).max(), "builds", sim.nbl_count)
        assert np.abs(d).max() < 1e-9 and np.abs(a["impuls"] - b["impuls"]).max() < 1e-10
        assert abs(sc["tot_pot_energy"] - rs["tot_pot_energy"]) < 1e-10 * abs(rs["tot_pot_energy"])
        assert abs(sc["tot_kin_energy"] - rs["tot_kin_energy"]) < 1e-9 * abs(rs["tot_kin_energy"])
        assert abs(sc["eta"] - rs["eta"]) < 1e-9 * max(abs(rs["eta"]), 1e-12)
        assert sim.nbl_count == ref.nbl_count and sim.nbl_count >= 3
        ref.close()
    return sim


def case_npt(name, grid, rank):
    """npt_iso / npt_axial over the process grid against the reference's fixture: the barostat scalars are formed from sums
    over all domains (virial or per-axis virial, kinetic parts), every rank breathes the same box."""
    g = common.load_golden(name)
    axial = str(g["ensemble"]) == "npt_axial"
    kw = dict(isq_tau_xi=float(g["npt_start:isq_tau_xi"]))
    if not axial:
        kw["pressure_ext"] = float(g["npt_start:pressure_ext"])
    sim = make(g, tempfile.mkdtemp(prefix=f"npt{rank}_"), grid, **kw)
    if axial:
        sim.set_npt_axial(g["npt_start:xi"], g["npt_start:pressure_ext"], g["npt_start:d_pressure"], g["npt_start:relax_dirs"],
                          Ekin_old=float(g["npt_start:Ekin_old"]), dyn_stress=g["npt_start:dyn_stress"])
    else:
        sim.set_npt_state(xi=float(g["npt_start:xi"]), Ekin_old=float(g["npt_start:Ekin_old"]),
                          pressure_ext=float(g["npt_start:pressure_ext"]))
    worst = {}

    def close(k, got, want, tol, s):
        e = float(np.max(np.abs(np.asarray(got) - want)) / max(np.max(np.abs(want)), 1e-300))
        worst[k] = max(worst.get(k, 0.0), e)
        assert e <= tol, (k, s, got, want, e)

    for s in range(int(g["nsteps"])):
        tol = 1e-10 if s == 0 else 1e-8
        sim.calc_forces(s)
        close("epot", sim.scalars()["tot_pot_energy"], g["epot"][s], tol, s)
        sim.move_atoms()
        sim.check_nblist()
        sc = sim.scalars()
        if axial:
            st = sim.npt_axial()
            for k in ("xi", "stress", "pressure_ext", "dyn_stress"):
                close(k, st[k], g["npt:" + k][s], 10 * tol, s)
        else:
            st = sim.npt()
            close("xi", st["xi"], g["npt:xi"][s], 10 * tol, s)
            close("pressure", st["pressure"], g["npt:pressure"][s], 10 * tol, s)
        close("volume", sc["volume"], g["npt:volume"][s], 1e-10, s)
        close("eta", sc["eta"], g["eta"][s], 10 * tol, s)
        close("ekin", sc["tot_kin_energy"], g["ekin"][s], tol, s)
        close("box", sim.box(), g["npt:box"][s], 1e-10, s)
        assert sim.have_valid_nbl == int(g["valid"][s]), f"check_nblist decision differs at step {s}"
    a = idist.gather_atoms(sim)
    if rank == 0:
        box = sim.box()
        d = a["ort"] - g["final:ort"]
        frac = d @ np.linalg.inv(box)
        d = (frac - np.round(frac)) @ box
        assert np.max(np.abs(d)) <= 1e-8 * np.max(np.abs(box))
        print("npt", name, {k: f"{v:.1e}" for k, v in worst.items()})
    return sim


def case_send_forces(grid, rank):
    """send_forces analogue: every image carries 1.0; after the reverse exchange an owner holds the number of
    its images, which is fixed by the geometry: prod(1 + [low layer] + [high layer]) - 1 over the axes."""
    g = common.load_golden("cu_long")
    sim = make(g, tempfile.mkdtemp(prefix=f"sf{rank}_"), grid)
    sim.calc_forces(0)
    n, ng = sim.natoms, sim.nghost
    f = torch.zeros(2, n + ng, dtype=torch.float64, device="cuda")
    f[0, n:] = 1.0
    f[1, n:] = 2.0
    sim.send_forces(f.data_ptr(), 2, n + ng)
    torch.cuda.synchronize()
    got = f[:, :n].cpu().numpy()
    a = sim.atoms(sort=False)
    gd, cd = sim.celldims()
    frac = a["ort"] @ np.linalg.inv(np.asarray(g["box"]))
    cell = np.minimum((frac * gd).astype(int), gd - 1)
    per = cd - 2
    local = cell - np.array(api.cart_coords(rank, grid)) * per
    pbc = np.asarray(g["pbc"])
    npg = np.array(grid)
    my = np.array(api.cart_coords(rank, grid))
    mult = np.ones(n)
    for ax in range(3):
        lo = (local[:, ax] == 0) & ((pbc[ax] == 1) | (my[ax] > 0))
        hi = (local[:, ax] == per[ax] - 1) & ((pbc[ax] == 1) | (my[ax] < npg[ax] - 1))
        mult *= 1 + lo.astype(int) + hi.astype(int)
    want = mult - 1
    assert np.array_equal(got[0], want), (got[0][:20], want[:20])
    assert np.array_equal(got[1], 2 * want)
    tot = torch.tensor([float(got[0].sum()), float(ng)], dtype=torch.float64)
    dist.all_reduce(tot)
    assert tot[0] == tot[1]
    if rank == 0:
        print("send_forces: images summed back to owners:", int(tot[0]))
    return sim


def main():
    import signal
    signal.alarm(200)          # a rank stuck in a collective must not hold the GPU box
    case = sys.argv[1]
    grid = tuple(int(x) for x in sys.argv[2:5])
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if case.startswith("fixture:"):
        sim = case_fixture(case.split(":", 1)[1], grid, rank)
    elif case == "migration":
        sim = case_migration(grid, rank)
    elif case == "migration_slab":
        sim = case_migration(grid, rank, pbc=(1, 1, 0))
    elif case == "send_forces":
        sim = case_send_forces(grid, rank)
    elif case == "overlap":
        sim = case_overlap(grid, rank)
    elif case == "overlap_nial":
        sim = case_overlap_nial(grid, rank)
    elif case.startswith("npt:"):
        sim = case_npt(case.split(":", 1)[1], grid, rank)
    else:
        raise SystemExit(f"unknown case {case}")
    dist.barrier()
    sim.close()
    if rank == 0:
        print("MGPU_OK", case, grid)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
