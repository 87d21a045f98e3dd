"""GPU, N >= 2: spatial domain decomposition (one process per GPU, NCCL halo exchange inside the
library) against the fixtures of the unmodified single-process reference and against the same run on
one GPU.  Skipped on a single-GPU box; `gpurun --gpus 2 -- python -m pytest tests -m gpu` runs them."""
import os
import subprocess
import sys

import pytest

from tests import common

pytestmark = pytest.mark.gpu


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _run(case, grid, port, env=None):
    n = grid[0] * grid[1] * grid[2]
    if _ngpus() < n:
        pytest.skip(f"needs {n} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(common.ROOT, "tests", "mgpu_worker.py"), case] + [str(x) for x in grid]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=common.ROOT,
                       env=dict(os.environ, **(env or {})))
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    print(r.stdout[-1500:])


@pytest.mark.parametrize("name", ["cu_long", "nial_nvt", "lj_nve", "cu_slab", "cu_nve"])
def test_two_domains_match_reference_fixture(built_lib, name):
    _run("fixture:" + name, (2, 1, 1), 29611)


@pytest.mark.parametrize("name", ["cu_lindef", "cu_frozen_nvt", "nial_eeam"])
def test_two_domains_deformation_and_restrictions(built_lib, name):
    """lin_deform re-plans the cell grid and the halo on every rank; deform_sample and the restriction vectors act on
    virtual types that live on both sides of the domain boundary; nactive is summed over ranks."""
    _run("fixture:" + name, (2, 1, 1), 29619)


@pytest.mark.parametrize("name", ["cu_long", "nial_nvt"])
def test_two_domains_nccl_halo(built_lib, name):
    """The halo is exchanged over peer memory by default (direct stores into the neighbour's ghost region, stream memory
    operations as completion signal); IMDB200_HALO_P2P=0 keeps the NCCL messages: same fixtures."""
    _run("fixture:" + name, (2, 1, 1), 29620, env={"IMDB200_HALO_P2P": "0"})


def test_nccl_halo_in_the_device_resident_loop(built_lib):
    """imdb200_run over NCCL messages (no overlap): hot crystal, atoms migrate; against the single-GPU trajectory."""
    _run("migration", (2, 1, 1), 29625, env={"IMDB200_HALO_P2P": "0"})


def test_overlapped_halo_matches_single_gpu_run(built_lib):
    """The default multi-GPU loop: boundary warps first, their F' and new positions stored into the neighbours' ghost
    regions while the interior warps run (imdb200_run, api.cu queue_step); 131 072 thermal Cu atoms, 60 steps with list
    builds, against the same run on ONE GPU: positions, momenta, energies."""
    _run("overlap", (2, 1, 1), 29626)


def test_two_species_overlapped_halo_matches_single_gpu_run(built_lib):
    """43 904 thermal Ni-Al atoms under NVT through imdb200_run on two domains: several-species pass 2 (type in the list
    entry, one gather record), fused move_atoms, split launches, peer-memory halo; against the run on one GPU."""
    _run("overlap_nial", (2, 1, 1), 29629)


@pytest.mark.parametrize("name", ["cu_npt_iso", "cu_npt_axial", "cu_npt_axial_xz"])
def test_two_domains_npt(built_lib, name):
    """Barostat ensembles on two domains against the reference's npt fixtures (global virial / per-axis virial and kinetic
    sums over both ranks; both ranks breathe the same box and re-plan their halo)."""
    _run("npt:" + name, (2, 1, 1), 29630)


@pytest.mark.parametrize("name", ["cu_adp", "nial_adp"])
def test_two_domains_adp(built_lib, name):
    """ADP on two domains: mu and lambda travel with F' in the halo (nine components, src/imd_comm_force_3d.c:1047-1057)."""
    _run("fixture:" + name, (2, 1, 1), 29631)


@pytest.mark.parametrize("name", ["cu_big", "nial_big"])
def test_two_domains_match_large_reference_fixture(built_lib, name):
    """131 072 Cu / 54 000 Ni-Al atoms on two domains against the unmodified single-process reference."""
    _run("fixture:" + name, (2, 1, 1), 29622)


def test_two_domains_split_along_z(built_lib):
    _run("fixture:cu_long", (1, 1, 2), 29612)


def test_atoms_migrate_between_domains(built_lib):
    _run("migration", (2, 1, 1), 29613)


def test_atoms_migrate_across_a_free_axis_split(built_lib):
    """cpu_dim = (1,1,2) on a slab with free z surfaces: the leave direction must come from the coordinates."""
    _run("migration_slab", (1, 1, 2), 29621)


def test_send_forces_reverse_path(built_lib):
    _run("send_forces", (2, 1, 1), 29614)


def test_four_domains(built_lib):
    _run("fixture:cu_long", (2, 2, 1), 29615)
    _run("migration", (2, 2, 1), 29616)
    _run("overlap", (2, 2, 1), 29628)


def test_eight_domains(built_lib):
    _run("migration", (2, 2, 2), 29617)
    _run("send_forces", (2, 2, 2), 29618)
    _run("fixture:cu_big", (2, 2, 2), 29623)
    _run("fixture:nial_big", (2, 2, 2), 29624)
    _run("overlap", (2, 2, 2), 29627)
