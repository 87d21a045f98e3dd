"""CPU (gloo, world_size 2 and 4): host-side logic of the multi-GPU path -- see tests/gloo_worker.py."""
import os
import subprocess
import sys

import pytest

from tests import common


@pytest.mark.parametrize("world,pbc,port", [(2, (1, 1, 1), 29631), (2, (0, 1, 1), 29632), (4, (1, 1, 0), 29633),
                                            (8, (1, 1, 1), 29634)])
def test_process_grid_and_halo_plan_are_consistent(built_lib, world, pbc, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(common.ROOT, "tests", "gloo_worker.py")] + [str(x) for x in pbc]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=common.ROOT)
    assert r.returncode == 0 and "GLOO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_calc_cpu_dim_matches_reference_factorisation(built_lib):
    """calc_cpu_dim (src/imd_geom_mpi_3d.c:201-266): even factorisation, largest factor on the largest axis."""
    from imd_b200 import api
    assert api.calc_cpu_dim(1) == (1, 1, 1)
    assert api.calc_cpu_dim(2) == (2, 1, 1)
    assert api.calc_cpu_dim(4) == (2, 2, 1)
    assert api.calc_cpu_dim(8) == (2, 2, 2)
    assert api.calc_cpu_dim(12) == (3, 2, 2)
    assert api.calc_cpu_dim(8, (1, 1, 4)) == (2, 2, 2)
    assert api.calc_cpu_dim(6, (1, 1, 4)) == (2, 1, 3)     # largest factor follows the largest request
    for r in range(12):
        assert api.cart_rank(api.cart_coords(r, (3, 2, 2)), (3, 2, 2)) == r
    assert api.cart_coords(5, (2, 2, 2)) == (1, 0, 1)       # z fastest, like MPI_Cart_coords
