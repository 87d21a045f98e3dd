"""Worker of tests/test_dist_cpu.py: world_size-2 (or more) torch.distributed run on CPU (gloo backend)
covering the host-side logic of the N > 1 path: process grid, neighbour ranks, image-shift codes, the
ncclUniqueId hand-off, and that a rank without a GPU fails loudly instead of falling back."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imd_b200 import api  # noqa: E402
from imd_b200 import dist as idist  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    pbc = tuple(int(x) for x in sys.argv[1:4])
    grid = idist.grid_for(world)
    assert int(np.prod(grid)) == world
    coord = api.cart_coords(rank, grid)
    assert api.cart_rank(coord, grid) == rank
    peer, code = api.halo_peers(grid, coord, pbc)
    recv_order, send_order = api.halo_message_order(peer, rank)
    everyone = [None] * world
    dist.all_gather_object(everyone, dict(rank=rank, coord=coord, peer=peer, code=code, recv_order=recv_order,
                                          send_order=send_order))
    # every rank checks the whole table: what A expects from direction d, B = peer_A(d) must send towards
    # 26-d, and the image shifts the two sides apply are opposite
    for a in everyone:
        assert a["peer"][13] == -1
        for d in range(27):
            b = a["peer"][d]
            if d == 13 or b < 0:
                continue
            B = everyone[b]
            assert B["peer"][26 - d] == a["rank"], (a["rank"], d, b, B["peer"][26 - d])
            assert B["code"][26 - d] == 26 - a["code"][d]
            sg = np.array([d % 3 - 1, (d // 3) % 3 - 1, d // 9 - 1])
            want = (np.array(a["coord"]) + sg) % np.array(grid)
            assert tuple(want) == tuple(B["coord"])
            # a shift appears exactly on the axes where the step leaves the process grid
            sh = np.array([a["code"][d] % 3 - 1, (a["code"][d] // 3) % 3 - 1, a["code"][d] // 9 - 1])
            out = (np.array(a["coord"]) + sg < 0) | (np.array(a["coord"]) + sg >= np.array(grid))
            assert np.array_equal(sh, np.where(out, sg, 0))
        for ax in range(3):
            if not pbc[ax]:
                # free surface: nothing behind the first / last rank of that axis
                for d in range(27):
                    s = [d % 3 - 1, (d // 3) % 3 - 1, d // 9 - 1][ax]
                    if (a["coord"][ax] == 0 and s < 0) or (a["coord"][ax] == grid[ax] - 1 and s > 0):
                        assert a["peer"][d] == -1
    # one message per peer: for every ordered pair (A -> B) the directions A sends towards B, taken in A's send order
    # and mirrored (d -> 26-d), are exactly the directions B fills from A, in B's receive order -- so the slice is
    # contiguous and identically ordered on both sides; peers appear as contiguous groups in both orders
    for a in everyone:
        for order in (a["recv_order"], a["send_order"]):
            peers_seq = [a["peer"][d] for d in order]
            groups = [p for i, p in enumerate(peers_seq) if i == 0 or peers_seq[i - 1] != p]
            assert len(groups) == len(set(groups)), (a["rank"], peers_seq)
        assert sorted(a["recv_order"]) == [d for d in range(27) if d != 13 and a["peer"][d] >= 0]
        assert sorted(a["send_order"]) == [d for d in range(27) if d != 13 and a["peer"][d] >= 0 and a["peer"][d] != a["rank"]]
        for b in everyone:
            if b["rank"] == a["rank"]:
                continue
            sent = [26 - d for d in a["send_order"] if a["peer"][d] == b["rank"]]
            got = [d for d in b["recv_order"] if b["peer"][d] == a["rank"]]
            assert sent == got, (a["rank"], b["rank"], sent, got)
    # the 128-byte ncclUniqueId travels from rank 0 to everybody
    uid = idist.broadcast_unique_id(rank)
    ids = [None] * world
    dist.all_gather_object(ids, uid)
    assert len(uid) == 128 and all(x == ids[0] for x in ids)
    # no GPU here: creating a domain must fail loudly (no CPU fallback)
    import torch
    if not torch.cuda.is_available():
        try:
            idist.create(1, np.eye(3) * 30.0, cpu_dim=grid)
        except api.IMDError:
            pass
        else:
            raise AssertionError("a domain was created without a GPU")
    dist.barrier()
    if rank == 0:
        print("GLOO_OK", grid)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
