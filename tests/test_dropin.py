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


FULL = ("imd_b200_dropin_full", "imd_ref_serial_full")


def _pair_run(tmp, tabs, *, exes=FULL, ncell=(10, 10, 10), restart=None, **kw):
    outs = {}
    for name, exe in zip(("gpu", "cpu"), exes):
        assert os.path.exists(os.path.join(REF, exe)), f"oracle/_ref/{exe} missing: run `make -C oracle ref` where /root/reference exists"
        p = synth.cu_param(tmp, ncell=ncell, name=name, tables=tabs, **kw)
        if restart is None:
            outs[name] = _run(exe, p, tmp)
        else:
            r = subprocess.run([os.path.join(REF, exe), "-p", p, "-r", str(restart)], capture_output=True, text=True, cwd=tmp,
                               timeout=600)
            assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
            outs[name] = r.stdout
    return outs


def _force_file(path):
    rows = np.loadtxt(path, comments="#")            # type x y z fx fy fz, written BEFORE the move (src/imd_io.c:1927-1950)
    o = np.lexsort((rows[:, 3], rows[:, 2], rows[:, 1]))
    return rows[o]


def test_full_binding_homdef_stress_force_dump(built_lib, tmp_path):
    """IMD built with homdef + stress + force around the engine: lin_deform reaches the device through --wrap
    (src/imd_main_3d.c:293-299 -> imdb200_lin_deform), the box columns and Press_xx..Press_xy of the .eng file
    (src/imd_io.c:2474-2480, calc_tot_presstens on the downloaded per-atom tensor) and the .force dump
    (src/imd_io.c:1886-1950) come out of IMD's own writers and match the unmodified serial IMD."""
    tmp = str(tmp_path)
    tabs = synth.make_eam_tables(tmp, "cu", nr=601, nrho=801)
    extra = dict(eng_int=1, checkpt_int=24, force_int=12, lindef_interval=4, lindef_size=2.0e-3,
                 lindef_x=[1.0, 0.2, 0.0], lindef_y=[0.0, -0.4, 0.0], lindef_z=[0.3, 0.0, 0.2])
    _pair_run(tmp, tabs, ensemble="nve", maxsteps=24, starttemp=0.08, extra=extra)
    eg, ec = _eng(os.path.join(tmp, "gpu.eng")), _eng(os.path.join(tmp, "cpu.eng"))
    assert eg.shape == ec.shape and len(eg) == 25 and eg.shape[1] >= 11
    # columns: time Epot T fnorm fmax pressure volume eta box_x.x box_y.y box_z.z box_y.x box_x.z Press_xx yy zz yz xz xy
    assert eg.shape[1] == 19
    for col in range(1, eg.shape[1]):
        scale = max(np.max(np.abs(ec[:, col])), 1e-300)
        tol = 2e-6 if col in (5, 6) else 1e-8            # pressure and volume are printed with %e only
        assert np.max(np.abs(eg[:, col] - ec[:, col])) <= tol * scale, (col, eg[:3, col], ec[:3, col])
    # box columns moved (the deformation really happened), forces and stress columns are not empty
    assert np.ptp(ec[:, 8]) > 0 and np.max(np.abs(ec[:, -6:])) > 0 and np.min(ec[1:, 3]) > 0
    fscale = max(np.max(np.abs(_force_file(os.path.join(tmp, f"cpu.{k:05d}.force"))[:, 4:7])) for k in (1, 2))
    assert fscale > 0.1
    for k in (0, 1, 2):                                # dump 0 is the perfect lattice: forces are rounding noise there
        fg, fc = _force_file(os.path.join(tmp, f"gpu.{k:05d}.force")), _force_file(os.path.join(tmp, f"cpu.{k:05d}.force"))
        assert fg.shape == fc.shape
        tol = 1e-10 if k == 0 else 1e-7
        assert np.max(np.abs(fg[:, 1:4] - fc[:, 1:4])) <= tol * 40.0
        assert np.max(np.abs(fg[:, 4:7] - fc[:, 4:7])) <= tol * fscale


def test_full_binding_npt_iso(built_lib, tmp_path):
    """ensemble npt_iso through the binding: xi, the external pressure ramp and the breathing box live on the device and
    are mirrored into IMD's globals after every move_atoms (src/imd_integrate.c:1472-1729)."""
    tmp = str(tmp_path)
    tabs = synth.make_eam_tables(tmp, "cu", nr=601, nrho=801)
    extra = dict(eng_int=1, checkpt_int=30, endtemp=0.08, tau_eta=0.1, eta=0.0, tau_xi=0.5, pressure_start=0.02, pressure_end=0.03)
    _pair_run(tmp, tabs, ensemble="npt_iso", maxsteps=30, starttemp=0.08, extra=extra)
    eg, ec = _eng(os.path.join(tmp, "gpu.eng")), _eng(os.path.join(tmp, "cpu.eng"))
    assert eg.shape == ec.shape and len(eg) == 31
    # columns: time Epot T fnorm fmax pressure volume eta box_x.x box_y.y box_z.z Press_xx .. Press_xy
    assert eg.shape[1] == 17 and np.ptp(ec[:, 8]) > 0   # the box breathes
    for col in list(range(1, 5)) + list(range(8, 17)):
        assert np.max(np.abs(eg[:, col] - ec[:, col])) <= 1e-8 * np.max(np.abs(ec[:, col])), col
    for col in (5, 6, 7):                                 # pressure, volume, eta*tau_eta are printed with %e only
        assert np.max(np.abs(eg[:, col] - ec[:, col])) <= 2e-6 * np.max(np.abs(ec[:, col])), col
    cg, cc = _chkpt(os.path.join(tmp, "gpu.00001.chkpt")), _chkpt(os.path.join(tmp, "cpu.00001.chkpt"))
    assert cg.shape == cc.shape and np.array_equal(cg[:, 0], cc[:, 0])
    assert np.max(np.abs(cg[:, 6:9] - cc[:, 6:9])) < 1e-7


def test_full_binding_relax_pressure(built_lib, tmp_path):
    """relax_pressure (src/imd_deform.c:127-219, called from main_loop src/imd_main_3d.c:756): IMD's own host code forms the
    deformation from calc_tot_presstens() on the per-atom tensor the engine downloads every step (do_press_calc is on while
    relax_rate > 0, src/imd_main_3d.c:183-194) and applies it through lin_deform, which the binding routes to the device."""
    tmp = str(tmp_path)
    tabs = synth.make_eam_tables(tmp, "cu", nr=601, nrho=801)
    extra = dict(eng_int=1, checkpt_int=20, relax_rate=0.01, relax_mode="full", bulk_module=1.0, shear_module=0.5)
    _pair_run(tmp, tabs, ensemble="nve", maxsteps=20, starttemp=0.08, extra=extra)
    eg, ec = _eng(os.path.join(tmp, "gpu.eng")), _eng(os.path.join(tmp, "cpu.eng"))
    assert eg.shape == ec.shape and len(eg) == 21
    # columns: time Epot T fnorm fmax pressure volume eta box_x.x box_y.y box_z.z Press_xx .. Press_xy
    assert eg.shape[1] == 17 and abs(ec[-1, 8] - ec[0, 8]) > 0.1     # the box has relaxed noticeably
    for col in list(range(1, 5)) + list(range(8, 17)):
        assert np.max(np.abs(eg[:, col] - ec[:, col])) <= 1e-8 * np.max(np.abs(ec[:, col])), col
    cg, cc = _chkpt(os.path.join(tmp, "gpu.00001.chkpt")), _chkpt(os.path.join(tmp, "cpu.00001.chkpt"))
    assert cg.shape == cc.shape and np.array_equal(cg[:, 0], cc[:, 0])
    assert np.max(np.abs(cg[:, 6:9] - cc[:, 6:9])) < 1e-7


def test_axial_binding_npt_axial(built_lib, tmp_path):
    """ensemble npt_axial through the binding (make target imd_nve_nvt_npt_axial_eam_nbl_stress_hpo): per-axis xi, stress,
    pressure ramp and box live on the device and are mirrored into IMD's globals after every move_atoms
    (src/imd_integrate.c:1747-1959); stress_x/y/z, the box and Press_xx.. columns of IMD's own .eng writer."""
    tmp = str(tmp_path)
    tabs = synth.make_eam_tables(tmp, "cu", nr=601, nrho=801)
    extra = dict(eng_int=1, checkpt_int=30, endtemp=0.08, tau_eta=0.1, eta=0.0, tau_xi=0.5, pressure_start=[0.02, 0.01, 0.03],
                 pressure_end=[0.03, 0.01, 0.02], relax_dirs=[1, 0, 1])
    _pair_run(tmp, tabs, exes=("imd_b200_dropin_axial", "imd_ref_serial_axial"), ensemble="npt_axial", maxsteps=30, starttemp=0.08,
              extra=extra)
    eg, ec = _eng(os.path.join(tmp, "gpu.eng")), _eng(os.path.join(tmp, "cpu.eng"))
    assert eg.shape == ec.shape and len(eg) == 31
    # columns: time Epot T pressure volume eta stress_x/y/z box_x.x box_y.y box_z.z Press_xx .. Press_xy
    assert eg.shape[1] == 18 and np.ptp(ec[:, 9]) > 0 and np.ptp(ec[:, 10]) == 0   # x breathes, y is held
    for col in [1, 2] + list(range(6, 18)):
        assert np.max(np.abs(eg[:, col] - ec[:, col])) <= 1e-8 * np.max(np.abs(ec[:, col])), col
    for col in (3, 4, 5):                                 # pressure, volume, eta*tau_eta are printed with %e only
        assert np.max(np.abs(eg[:, col] - ec[:, col])) <= 2e-6 * np.max(np.abs(ec[:, col])), col
    cg, cc = _chkpt(os.path.join(tmp, "gpu.00001.chkpt")), _chkpt(os.path.join(tmp, "cpu.00001.chkpt"))
    assert cg.shape == cc.shape and np.array_equal(cg[:, 0], cc[:, 0])
    assert np.max(np.abs(cg[:, 6:9] - cc[:, 6:9])) < 1e-7


def test_andersen_binding(built_lib, tmp_path):
    """`and` builds (Andersen thermostat, src/imd_integrate.c:491-495): every tempintv-th move_atoms ends in IMD's own
    maxwell(temperature), which draws one drand48 triple per atom in the order of the HOST cells.  The binding keeps maxwell on
    the host, lets IMD's own fix_cells() evolve the host cells at every list build exactly as in the reference, and sends the
    new momenta to the device by atom number (imdb200_set_momenta).  Three re-draws in 60 steps of a hot crystal (atoms change
    cells in between): Epot and T per step and the final checkpoint against the unmodified serial `and` build."""
    tmp = str(tmp_path)
    tabs = synth.make_eam_tables(tmp, "cu", nr=601, nrho=801)
    extra = dict(eng_int=1, checkpt_int=60, tempintv=20, endtemp=0.12)
    _pair_run(tmp, tabs, exes=("imd_b200_dropin_and", "imd_ref_serial_and"), ensemble="nve", maxsteps=60, starttemp=0.12, extra=extra)
    eg, ec = _eng(os.path.join(tmp, "gpu.eng")), _eng(os.path.join(tmp, "cpu.eng"))
    assert eg.shape == ec.shape and len(eg) == 61
    assert ec[20, 2] > ec[19, 2] * 1.04 and ec[40, 2] > ec[39, 2] * 1.04          # the re-draws are visible in T
    assert np.max(np.abs(eg[:, 1] - ec[:, 1]) / np.abs(ec[:, 1])) <= 1e-8
    assert np.max(np.abs(eg[:, 2] - ec[:, 2]) / np.abs(ec[:, 2])) <= 1e-8
    cg, cc = _chkpt(os.path.join(tmp, "gpu.00001.chkpt")), _chkpt(os.path.join(tmp, "cpu.00001.chkpt"))
    assert cg.shape == cc.shape and np.array_equal(cg[:, 0], cc[:, 0])
    assert np.max(np.abs(cg[:, 6:9] - cc[:, 6:9])) < 1e-6                         # velocities: the same random numbers per atom


def test_restart_from_checkpoint(built_lib, tmp_path):
    """Restart `-r 1` (src/imd_param.c:3829-3872, .itr + checkpoint readers src/imd_io_3d.c:949-1086): 15 steps, checkpoint 1,
    then both binaries continue from THEIR OWN checkpoint for 15 more steps; the engine is fed by IMD's own reader."""
    tmp = str(tmp_path)
    tabs = synth.make_eam_tables(tmp, "cu", nr=601, nrho=801)
    extra = dict(eng_int=1, checkpt_int=15)
    _pair_run(tmp, tabs, exes=("imd_b200_dropin", "imd_ref_serial_eam"), ensemble="nvt", maxsteps=15, starttemp=0.08, extra=extra)
    for n in ("gpu", "cpu"):
        assert os.path.exists(os.path.join(tmp, f"{n}.00001.chkpt")) and os.path.exists(os.path.join(tmp, f"{n}.00001.itr"))
    _pair_run(tmp, tabs, exes=("imd_b200_dropin", "imd_ref_serial_eam"), restart=1, ensemble="nvt", maxsteps=30, starttemp=0.08,
              extra=extra)
    eg, ec = _eng(os.path.join(tmp, "gpu.eng")), _eng(os.path.join(tmp, "cpu.eng"))
    assert eg.shape == ec.shape and len(eg) >= 30 and eg[-1, 0] > 0.029
    assert np.max(np.abs(eg[:, 1] - ec[:, 1]) / np.abs(ec[:, 1])) <= 1e-8
    assert np.max(np.abs(eg[:, 2] - ec[:, 2]) / np.abs(ec[:, 2])) <= 1e-8
    cg, cc = _chkpt(os.path.join(tmp, "gpu.00002.chkpt")), _chkpt(os.path.join(tmp, "cpu.00002.chkpt"))
    assert cg.shape == cc.shape and np.array_equal(cg[:, 0], cc[:, 0])
    box = 10 * synth.CU_A0
    d = cg[:, 3:6] - cc[:, 3:6]
    d -= box * np.round(d / box)
    assert np.max(np.abs(d)) < 1e-7 and np.max(np.abs(cg[:, 6:9] - cc[:, 6:9])) < 1e-6


def test_mpi_binding_two_ranks(built_lib, tmp_path):
    """IMD's own MPI main() (setup_mpi_topology, per-rank atom distribution, rank-0 writers) around the engine, one rank per
    GPU on oracle/shmpi: cpu_dim / my_coord from IMD's globals, ncclUniqueId by MPI_Bcast, halo / migration / reductions
    inside the library, host cells refilled from the device at every download.  Against the unmodified SERIAL IMD."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    tmp = str(tmp_path)
    tabs = synth.make_eam_tables(tmp, "cu", nr=601, nrho=801)
    exe = os.path.join(REF, "imd_b200_dropin_mpi")
    assert os.path.exists(exe), "oracle/_ref/imd_b200_dropin_mpi missing: run `make -C oracle ref` where /root/reference exists"
    # every run starts from the same reference-written checkpoint (maxwell() draws its random numbers in rank-local cell
    # order): a hot crystal, so that atoms cross the rank boundary
    pt = synth.cu_param(tmp, ncell=(12, 10, 10), name="therm", ensemble="nvt", maxsteps=20, starttemp=0.25, tables=tabs,
                        extra=dict(checkpt_int=20))
    _run("imd_ref_serial_eam", pt, tmp)
    chk = os.path.join(tmp, "therm.00001.chkpt")
    assert os.path.exists(chk)
    common_kw = dict(ncell=(12, 10, 10), ensemble="nvt", maxsteps=40, starttemp=0.25, tables=tabs, coordname=chk)
    pg = synth.cu_param(tmp, name="gpu", extra=dict(eng_int=1, checkpt_int=40, box_from_header=1, cpu_dim=[2, 1, 1]), **common_kw)
    r = subprocess.run([exe, "-p", pg], capture_output=True, text=True, cwd=tmp, timeout=600,
                       env=dict(os.environ, SHMPI_NP="2", SHMPI_PIN="0"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MPI process array dimensions: 2 1 1" in r.stdout
    pc = synth.cu_param(tmp, name="cpu", extra=dict(eng_int=1, checkpt_int=40, box_from_header=1), **common_kw)
    _run("imd_ref_serial_eam", pc, tmp)
    eg, ec = _eng(os.path.join(tmp, "gpu.eng")), _eng(os.path.join(tmp, "cpu.eng"))
    assert eg.shape == ec.shape and len(eg) == 41
    assert abs(eg[0, 1] - ec[0, 1]) <= 1e-12 * abs(ec[0, 1])
    assert np.max(np.abs(eg[:, 1] - ec[:, 1]) / np.abs(ec[:, 1])) <= 1e-8
    assert np.max(np.abs(eg[:, 2] - ec[:, 2]) / np.abs(ec[:, 2])) <= 1e-8
    cg, cc = _chkpt(os.path.join(tmp, "gpu.00001.chkpt")), _chkpt(os.path.join(tmp, "cpu.00001.chkpt"))
    assert cg.shape == cc.shape and np.array_equal(cg[:, 0], cc[:, 0]), "atoms lost or duplicated between the ranks"
    box = np.array([12, 10, 10]) * synth.CU_A0
    d = cg[:, 3:6] - cc[:, 3:6]
    d -= box * np.round(d / box)
    assert np.max(np.abs(d)) < 1e-6 and np.max(np.abs(cg[:, 6:9] - cc[:, 6:9])) < 1e-5
