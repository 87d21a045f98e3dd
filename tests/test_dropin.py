"""GPU: IMD's OWN main() / parameter file / generators / writers with the force engine replaced by
integration/imd_forces_b200.c (-> libimd_b200.so), against the unmodified serial reference binary on the
same parameter file.  Both binaries are built by oracle/Makefile from the reference sources and travel to
the GPU box in oracle/_ref/ (the reference tree itself does not)."""
import os
import subprocess

import numpy as np
import pytest

from tests import common
from imd_b200 import synth

pytestmark = pytest.mark.gpu
REF = os.path.join(common.ROOT, "oracle", "_ref")


def _run(exe, param, cwd, env=None):
    r = subprocess.run([os.path.join(REF, exe), "-p", param], capture_output=True, text=True, cwd=cwd, timeout=600,
                       env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def _eng(path):
    return np.loadtxt(path, comments="#", ndmin=2)


def _chkpt(path):
    rows = np.loadtxt(path, comments="#")
    o = np.argsort(rows[:, 0])
    return rows[o]


@pytest.mark.parametrize("ensemble,sync", [("nve", "1"), ("nvt", "0")])
def test_imd_main_with_b200_engine_matches_serial_imd(built_lib, tmp_path, ensemble, sync):
    for exe in ("imd_b200_dropin", "imd_ref_serial_eam"):
        assert os.path.exists(os.path.join(REF, exe)), f"oracle/_ref/{exe} missing: run `make -C oracle ref` where /root/reference exists"
    tmp = str(tmp_path)
    tabs = synth.make_eam_tables(tmp, "cu", nr=601, nrho=801)
    extra = dict(eng_int=1, checkpt_int=30)
    outs = {}
    for name, exe in (("gpu", "imd_b200_dropin"), ("cpu", "imd_ref_serial_eam")):
        p = synth.cu_param(tmp, ncell=(10, 10, 10), name=name, ensemble=ensemble, maxsteps=30, starttemp=0.08,
                           tables=tabs, extra=extra)
        outs[name] = _run(exe, p, tmp, env={"IMD_B200_SYNC": sync})
    # the unmodified host code reports the same geometry and list cadence on both sides
    assert "Global cell array dimensions: 6 6 6" in outs["gpu"] and "Global cell array dimensions: 6 6 6" in outs["cpu"]
    eg, ec = _eng(os.path.join(tmp, "gpu.eng")), _eng(os.path.join(tmp, "cpu.eng"))
    assert eg.shape == ec.shape and len(eg) == 31
    # columns: time Epot/atom temperature pressure volume [eta*tau_eta]; pressure is printed with %e only
    assert abs(eg[0, 1] - ec[0, 1]) <= 1e-12 * abs(ec[0, 1])
    assert abs(eg[0, 2] - ec[0, 2]) <= 1e-12 * abs(ec[0, 2])
    assert np.max(np.abs(eg[:, 1] - ec[:, 1]) / np.abs(ec[:, 1])) <= 1e-9
    assert np.max(np.abs(eg[:, 2] - ec[:, 2]) / np.abs(ec[:, 2])) <= 1e-9
    assert np.max(np.abs(eg[:, 3] - ec[:, 3])) <= 2e-6 * np.max(np.abs(ec[:, 3]))
    if ensemble == "nvt":
        assert np.max(np.abs(eg[:, 5] - ec[:, 5])) <= 1e-6 * max(np.max(np.abs(ec[:, 5])), 1e-12)
    # final checkpoint written by IMD's own writer from the arrays the engine filled
    cg, cc = _chkpt(os.path.join(tmp, "gpu.00001.chkpt")), _chkpt(os.path.join(tmp, "cpu.00001.chkpt"))
    assert cg.shape == cc.shape and np.array_equal(cg[:, 0], cc[:, 0])
    box = 10 * synth.CU_A0
    d = cg[:, 3:6] - cc[:, 3:6]
    d -= box * np.round(d / box)
    assert np.max(np.abs(d)) < 1e-8                      # positions
    assert np.max(np.abs(cg[:, 6:9] - cc[:, 6:9])) < 1e-7  # velocities
    assert np.max(np.abs(cg[:, 9] - cc[:, 9])) < 1e-8      # per-atom Epot
