"""ctypes driver for the UNMODIFIED reference built by oracle/Makefile into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Imported by tests/, tools/make_golden.py and bench.py's
reference arm -- never by the product (imd_b200/).

The reference keeps all state in C globals and calls exit() on error, so one process can
hold exactly one simulation: use :class:`RefIMD` once per process, or go through
:func:`run_in_subprocess`.
"""
from __future__ import annotations

import ctypes as C
import os
import pickle
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")


def lib_path(variant: str) -> str:
    return os.path.join(REFDIR, f"libimdref_{variant}.so")


def available(variant: str = "eam") -> bool:
    return os.path.exists(lib_path(variant))


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class RefIMD:
    """One reference simulation (variant: 'eam', 'pair', 'eam_fast')."""

    def __init__(self, variant: str, paramfile: str, restart: int = 0, quiet: bool = True):
        self.lib = C.CDLL(lib_path(variant))
        L = self.lib
        L.ref_natoms.restype = C.c_long
        L.ref_cellsz.restype = C.c_double
        L.ref_get_atoms.restype = C.c_long
        L.ref_get_nbl_pairs.restype = C.c_long
        L.ref_get_nbl_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
        L.ref_pair_int.argtypes = [C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ref_set_eta.argtypes = [C.c_double]
        self.has_eam = variant.startswith("eam") or variant in ("eeam", "npt", "npt_axial", "adp", "ber")
        self.has_eeam = variant == "eeam"
        self.has_npt = variant == "npt"
        self.has_npt_axial = variant == "npt_axial"
        self.has_adp = variant == "adp"
        L.ref_get_adp.restype = C.c_long
        L.ref_get_eeam.restype = C.c_long
        if quiet:
            sys.stdout.flush()
            saved = os.dup(1)
            devnull = os.open(os.devnull, os.O_WRONLY)
            os.dup2(devnull, 1)
        try:
            L.ref_setup(paramfile.encode(), restart)
        finally:
            if quiet:
                C.CDLL(None).fflush(None)
                os.dup2(saved, 1)
                os.close(devnull)
                os.close(saved)

    # --- step loop pieces --------------------------------------------------------------
    def calc_forces(self, step=0):
        self.lib.ref_calc_forces(step)

    def move_atoms(self):
        self.lib.ref_move_atoms()

    def check_nblist(self):
        self.lib.ref_check_nblist()

    def step(self, n=1):
        for _ in range(n):
            self.lib.ref_step()

    def set_press_calc(self, on=True):
        self.lib.ref_set_press_calc(int(on))

    def invalidate_nbl(self):
        """have_valid_nbl = 0: the next calc_forces runs fix_cells + make_nblist."""
        self.lib.ref_set_atoms(None, None)

    def lin_deform(self):
        self.lib.ref_lin_deform()

    def deform_sample(self):
        self.lib.ref_deform_sample()

    # --- state -------------------------------------------------------------------------
    @property
    def natoms(self):
        return int(self.lib.ref_natoms())

    @property
    def nbl_count(self):
        return int(self.lib.ref_nbl_count())

    @property
    def have_valid_nbl(self):
        return int(self.lib.ref_have_valid_nbl())

    @property
    def cellsz(self):
        return float(self.lib.ref_cellsz())

    def scalars(self):
        out = np.zeros(14)
        self.lib.ref_get_scalars(_p(out, C.c_double))
        keys = ["tot_pot_energy", "tot_kin_energy", "virial", "vir_xx", "vir_yy", "vir_zz",
                "vir_yz", "vir_zx", "vir_xy", "volume", "nactive", "eta", "temperature", "timestep"]
        return dict(zip(keys, out.tolist()))

    def npt(self):
        out = np.zeros(5)
        self.lib.ref_get_npt(_p(out, C.c_double))
        return dict(zip(("xi", "Ekin_old", "pressure", "pressure_ext", "isq_tau_xi"), out.tolist()))

    def npt_axial(self):
        out = np.zeros(16)
        self.lib.ref_get_npt_axial(_p(out, C.c_double))
        return dict(xi=out[0:3].copy(), stress=out[3:6].copy(), pressure_ext=out[6:9].copy(), dyn_stress=out[9:12].copy(),
                    Ekin_old=float(out[12]), relax_dirs=out[13:16].astype(np.int32))

    def set_eta(self, eta):
        self.lib.ref_set_eta(float(eta))

    def box(self):
        out = np.zeros(9)
        self.lib.ref_get_box(_p(out, C.c_double))
        return out.reshape(3, 3)

    def celldims(self):
        out = np.zeros(6, dtype=np.int32)
        self.lib.ref_get_celldims(_p(out, C.c_int))
        return out[:3].copy(), out[3:].copy()

    def tot_presstens(self):
        out = np.zeros(6)
        self.lib.ref_calc_tot_presstens(_p(out, C.c_double))
        return out

    def atoms(self, sort=True):
        n = self.natoms
        d = dict(
            nummer=np.zeros(n, np.int32), sorte=np.zeros(n, np.int32), vsorte=np.zeros(n, np.int32),
            masse=np.zeros(n), ort=np.zeros((n, 3)), impuls=np.zeros((n, 3)), kraft=np.zeros((n, 3)),
            poteng=np.zeros(n), rho=np.zeros(n), dF=np.zeros(n), presstens=np.zeros((n, 6)),
            nblpos=np.zeros((n, 3)),
        )
        got = self.lib.ref_get_atoms(
            _p(d["nummer"], C.c_int), _p(d["sorte"], C.c_int), _p(d["vsorte"], C.c_int),
            _p(d["masse"], C.c_double), _p(d["ort"], C.c_double), _p(d["impuls"], C.c_double),
            _p(d["kraft"], C.c_double), _p(d["poteng"], C.c_double), _p(d["rho"], C.c_double),
            _p(d["dF"], C.c_double), _p(d["presstens"], C.c_double), _p(d["nblpos"], C.c_double))
        assert got == n, (got, n)
        if self.has_eeam:
            d["eam_p"] = np.zeros(n); d["dM"] = np.zeros(n)
            self.lib.ref_get_eeam(_p(d["eam_p"], C.c_double), _p(d["dM"], C.c_double))
        if self.has_adp:
            d["adp_mu"] = np.zeros((n, 3)); d["adp_lambda"] = np.zeros((n, 6))
            self.lib.ref_get_adp(_p(d["adp_mu"], C.c_double), _p(d["adp_lambda"], C.c_double))
        if sort:  # canonical order: by atom number
            o = np.argsort(d["nummer"], kind="stable")
            d = {k: v[o] for k, v in d.items()}
        return d

    def set_atoms_by_number(self, nummer, ort=None, impuls=None):
        """Overwrite positions/momenta, given arrays keyed by atom number."""
        cur = self.atoms(sort=False)["nummer"]
        idx = {int(v): i for i, v in enumerate(nummer)}
        sel = np.array([idx[int(v)] for v in cur])
        o = None if ort is None else np.ascontiguousarray(ort[sel], dtype=np.float64)
        p = None if impuls is None else np.ascontiguousarray(impuls[sel], dtype=np.float64)
        self.lib.ref_set_atoms(_p(o, C.c_double), _p(p, C.c_double))

    def nbl_pairs(self):
        """Verlet list as array (npairs, 2) of atom numbers + (npairs, 3) image shifts of j."""
        cnt = self.lib.ref_get_nbl_pairs(None, None, None, 0)
        if cnt < 0:
            raise RuntimeError("reference has no valid neighbour list")
        pi = np.zeros(cnt, np.int32); pj = np.zeros(cnt, np.int32); sh = np.zeros((cnt, 3), np.int8)
        self.lib.ref_get_nbl_pairs(pi.ctypes.data, pj.ctypes.data, sh.ctypes.data, cnt)
        return np.stack([pi, pj], 1), sh

    def pair_int(self, which, col, r2):
        """(value, 2*d/dr2) through the reference's PAIR_INT macro; which: 0 pair, 1 embed, 2 rho."""
        r2 = np.atleast_1d(np.asarray(r2, dtype=np.float64))
        v = np.zeros_like(r2); g = np.zeros_like(r2)
        a = C.c_double(); b = C.c_double()
        for i, x in enumerate(r2):
            self.lib.ref_pair_int(which, col, float(x), C.byref(a), C.byref(b))
            v[i] = a.value; g[i] = b.value
        return v, g


# -------------------------------------------------------------------------------------------
def _child(spec_path, out_path):
    with open(spec_path, "rb") as f:
        spec = pickle.load(f)
    sim = RefIMD(spec["variant"], spec["paramfile"], quiet=True)
    out = run_protocol(sim, spec)
    with open(out_path, "wb") as f:
        pickle.dump(out, f)


def run_protocol(sim, spec):
    """Shared measurement protocol (also used for our own implementations, see tests/common.py):
    optional state override, optional thermalisation, then record after each of `nsteps` steps."""
    out = {"natoms": sim.natoms}
    if spec.get("eta") is not None:
        sim.set_eta(spec["eta"])
    if spec.get("press", False):
        sim.set_press_calc(True)
    if spec.get("warm", 0):
        sim.step(spec["warm"])
    sim.invalidate_nbl()  # lists of all implementations are built from the recorded start state
    out["start"] = sim.atoms()
    if getattr(sim, "has_npt", False):
        out["npt_start"] = sim.npt()
    if getattr(sim, "has_npt_axial", False):
        out["npt_axial_start"] = sim.npt_axial()
    out["nbl_count0"] = sim.nbl_count      # builds before the protocol (thermalisation); the protocol's share = nbl_count - this
    out["box"] = sim.box()
    out["celldims"] = sim.celldims()
    out["cellsz"] = sim.cellsz
    frames = []
    for s in range(spec.get("nsteps", 1)):
        # main_loop order (src/imd_main_3d.c:293-326): lin_deform, then deform_sample + check_nblist, then the forces
        if s > 0 and spec.get("lindef_every", 0) and s % spec["lindef_every"] == 0:
            sim.lin_deform()
        if s > 0 and spec.get("deform_every", 0) and s % spec["deform_every"] == 0:
            sim.deform_sample()
            sim.check_nblist()
        sim.calc_forces(s)
        fr = {"scalars": sim.scalars()}
        if s in spec.get("record_atoms", [0]):
            fr["atoms"] = sim.atoms()
            if spec.get("press", False):
                fr["tot_presstens_virial"] = sim.tot_presstens()
        if s in spec.get("record_nbl", []):
            fr["nbl_pairs"], fr["nbl_shift"] = sim.nbl_pairs()
        sim.move_atoms()
        sim.check_nblist()
        fr["after"] = sim.scalars()
        if getattr(sim, "has_npt", False):
            fr["npt"] = sim.npt(); fr["box"] = sim.box()
        if getattr(sim, "has_npt_axial", False):
            fr["npt_axial"] = sim.npt_axial(); fr["box"] = sim.box()
        fr["valid"] = sim.have_valid_nbl
        if spec.get("press", False) and s in spec.get("record_atoms", [0]):
            fr["tot_presstens"] = sim.tot_presstens()
        frames.append(fr)
    out["frames"] = frames
    out["final"] = sim.atoms()
    out["final_box"] = sim.box()
    out["nbl_count"] = sim.nbl_count
    return out


def run_in_subprocess(spec, workdir):
    """Run run_protocol() on the reference in a fresh process; returns its result dict."""
    sp = os.path.join(workdir, "ref_spec.pkl")
    op = os.path.join(workdir, "ref_out.pkl")
    with open(sp, "wb") as f:
        pickle.dump(spec, f)
    r = subprocess.run([sys.executable, os.path.abspath(__file__), sp, op], capture_output=True, text=True,
                       cwd=workdir)
    if r.returncode != 0 or not os.path.exists(op):
        raise RuntimeError(f"reference run failed ({r.returncode}):\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    with open(op, "rb") as f:
        return pickle.load(f)


if __name__ == "__main__":
    _child(sys.argv[1], sys.argv[2])
