"""One process per GPU: the glue between torch.distributed (plumbing: rendezvous, small host-side
collectives) and the library's own NCCL halo exchange.

The reference starts one MPI rank per domain (src/imd.c:57-58, src/imd_geom_mpi_3d.c:32-90); here the
launcher is torchrun, the rank grid is the same x-major Cartesian grid, and the data path (ghost
positions, 2F'(rho), atom migration, the scalar reductions) runs inside libimd_b200.so over NCCL.
Nothing here touches atom data except the test/IO helpers at the bottom.
"""
from __future__ import annotations

import numpy as np

from . import api


def grid_for(world_size, cpu_dim=None):
    """cpu_dim from the parameter file, or calc_cpu_dim's even factorisation (src/imd_geom_mpi_3d.c:201-266)."""
    if cpu_dim is not None and int(np.prod(cpu_dim)) == world_size:
        return tuple(int(x) for x in cpu_dim)
    return api.calc_cpu_dim(world_size, cpu_dim or (0, 0, 0))


def broadcast_unique_id(rank, group=None):
    """rank 0 creates the ncclUniqueId; torch.distributed carries the 128 bytes to the other ranks."""
    import torch.distributed as dist
    box = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    return box[0]


def create(ntypes, box, *, cpu_dim=None, device=None, **kw):
    """Build this rank's IMDB200 domain and join the communicator.  Needs an initialised
    torch.distributed process group (any backend; it only moves the 128-byte id)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    grid = grid_for(world, cpu_dim)
    coord = api.cart_coords(rank, grid)
    sim = api.IMDB200(ntypes, box, cpu_dim=grid, my_coord=coord, device=-1 if device is None else device, **kw)
    if world > 1:
        sim.comm_init(broadcast_unique_id(rank), rank, world)
    return sim


def gather_atoms(sim):
    """All ranks' atoms on every rank, sorted by NUMMER (test / output helper, like the parallel
    output collection of src/imd_io.c:174)."""
    import torch.distributed as dist
    mine = sim.atoms(sort=False)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        parts = [mine]
    else:
        parts = [None] * dist.get_world_size()
        dist.all_gather_object(parts, mine)
    out = {k: np.concatenate([p[k] for p in parts]) for k in mine}
    o = np.argsort(out["nummer"], kind="stable")
    return {k: v[o] for k, v in out.items()}


def gather_nbl(sim):
    """Union of the ranks' neighbour lists as (nummer_i, nummer_j) rows.  Image shifts are reported per
    rank (a pair across an interior domain boundary has shift 0 on both sides), so they are dropped."""
    import torch.distributed as dist
    pr, _ = sim.nbl_pairs()
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return pr
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, pr)
    return np.concatenate(parts)
