"""CPU: the reference's own MPI build, compiled against oracle/shmpi (our shared-memory subset of MPI), against
the reference's serial build.  Both are unmodified reference code; what is on trial here is the shim, so that
`bench.py --impl reference` may call its numbers "IMD's MPI CPU build" (SURVEY.md section 8d ladder step 1,
section 8f rank 2).  Protocol of SURVEY.md section 8c: every run starts from the same reference-written
checkpoint, never from `_fcc` + maxwell (whose random stream is consumed in rank-local cell order)."""
import os
import subprocess

import numpy as np
import pytest

from tests import common
from imd_b200 import synth

REF = os.path.join(common.ROOT, "oracle", "_ref")
need = [os.path.join(REF, x) for x in ("imd_ref_mpi_eam_par", "imd_ref_serial_eam")]
need_ref = pytest.mark.skipif(not all(os.path.exists(p) for p in need),
                              reason="oracle/_ref binaries missing: run `make -C oracle ref` where /root/reference exists")


@pytest.mark.parametrize("nranks", [1, 2, 5])
def test_shmpi_selftest(tmp_path, nranks):
    """The shared-memory MPI subset on its own (oracle/shmpi/selftest.c): collectives longer than a slot, messages
    longer than a ring, out-of-order tags, MPI_ANY_SOURCE / MPI_ANY_TAG, MPI_Waitany, self-sends, the Cartesian calls;
    also with more ranks than this suite's reference runs use and with an odd rank count."""
    src = os.path.join(common.ROOT, "oracle", "shmpi")
    exe = str(tmp_path / "selftest")
    subprocess.check_call(["gcc", "-O2", "-I" + src, "-o", exe, os.path.join(src, "selftest.c"), os.path.join(src, "shmpi.c")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=dict(os.environ, SHMPI_NP=str(nranks), SHMPI_PIN="0"))
    assert r.returncode == 0 and f"SHMPI_SELFTEST_OK {nranks}" in r.stdout, r.stdout[-1000:] + r.stderr[-1000:]


def _run(exe, param, cwd, np_=1):
    r = subprocess.run([os.path.join(REF, exe), "-p", param], capture_output=True, text=True, cwd=cwd, timeout=600,
                       env=dict(os.environ, SHMPI_NP=str(np_), SHMPI_PIN="0"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def _chkpt(path):
    rows = np.loadtxt(path, comments="#")
    return rows[np.argsort(rows[:, 0])]


@pytest.fixture(scope="module")
def start(tmp_path_factory):
    tmp = str(tmp_path_factory.mktemp("refmpi"))
    tabs = synth.make_eam_tables(tmp, "cu", nr=601, nrho=801)
    p = synth.cu_param(tmp, ncell=(12, 12, 12), name="therm", maxsteps=20, starttemp=0.08, tables=tabs,
                       extra=dict(checkpt_int=20))
    _run("imd_ref_serial_eam", p, tmp)
    chk = os.path.join(tmp, "therm.00001.chkpt")
    assert os.path.exists(chk)
    return tmp, tabs, chk


def _follow(tmp, tabs, chk, name, exe, grid):
    extra = dict(eng_int=1, checkpt_int=40, box_from_header=1)
    if grid is not None:
        extra["cpu_dim"] = list(grid)
    p = synth.cu_param(tmp, ncell=(12, 12, 12), name=name, maxsteps=40, tables=tabs, coordname=chk, extra=extra)
    out = _run(exe, p, tmp, np_=int(np.prod(grid)) if grid is not None else 1)
    eng = np.loadtxt(os.path.join(tmp, name + ".eng"), comments="#", ndmin=2)
    return out, eng, _chkpt(os.path.join(tmp, name + ".00001.chkpt"))


@need_ref
@pytest.mark.parametrize("grid", [(1, 1, 1), (2, 1, 1), (1, 2, 2), (2, 2, 2)])
def test_reference_mpi_build_on_shmpi_matches_serial_reference(start, grid):
    tmp, tabs, chk = start
    _, es, cs = _follow(tmp, tabs, chk, "ser", "imd_ref_serial_eam", None)
    out, em, cm = _follow(tmp, tabs, chk, "mpi%d%d%d" % grid, "imd_ref_mpi_eam_par", grid)
    assert "MPI process array dimensions: %d %d %d" % grid in out
    assert "Starting up MPI with %d processes" % int(np.prod(grid)) in out
    assert es.shape == em.shape and len(es) == 41
    # Epot/atom and temperature per step: same arithmetic, only the order of the global sums differs
    assert np.max(np.abs(em[:, 1] - es[:, 1]) / np.abs(es[:, 1])) < 1e-10
    assert np.max(np.abs(em[:, 2] - es[:, 2]) / np.abs(es[:, 2])) < 1e-10
    assert cs.shape == cm.shape and np.array_equal(cs[:, 0], cm[:, 0])
    box = 12 * synth.CU_A0
    d = cm[:, 3:6] - cs[:, 3:6]
    d -= box * np.round(d / box)
    assert np.max(np.abs(d)) < 1e-9
    assert np.max(np.abs(cm[:, 6:9] - cs[:, 6:9])) < 1e-8
    assert np.max(np.abs(cm[:, 9] - cs[:, 9])) < 1e-9
