"""ctypes wrapper of oracle/liboracle.so (our CPU restatement, oracle/imd_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg as the checker -- never by the product package imd_b200/.
The class mirrors oracle/ref_driver.RefIMD so that both can run the same protocol.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")

PAIR, EMBED, RHO, EMOD, ADP_U, ADP_W = 0, 1, 2, 3, 4, 5
NVE, NVT = 0, 1


def build(force=False):
    src = os.path.join(HERE, "imd_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return LIB


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_double]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_read_table.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        L.orc_pair_int.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double),
                                   C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.orc_table_info.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.POINTER(C.c_double)] * 3 + [C.POINTER(C.c_int)]
        L.orc_set_atoms.argtypes = [C.c_void_p, C.c_long] + [C.c_void_p] * 6
        L.orc_set_restrictions.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_set_integrator.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
        L.orc_set_box.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_calc_forces.argtypes = [C.c_void_p, C.c_int]
        L.orc_move_atoms.argtypes = [C.c_void_p, C.c_int]
        L.orc_check_nblist.argtypes = [C.c_void_p]
        L.orc_step.argtypes = [C.c_void_p, C.c_int]
        L.orc_lin_deform.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
        L.orc_natoms.restype = C.c_long
        L.orc_natoms.argtypes = [C.c_void_p]
        L.orc_have_valid_nbl.argtypes = [C.c_void_p]
        L.orc_nbl_count.argtypes = [C.c_void_p]
        L.orc_cellsz.restype = C.c_double
        L.orc_cellsz.argtypes = [C.c_void_p]
        L.orc_get_celldims.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_scalars.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_box.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_atoms.restype = C.c_long
        L.orc_get_atoms.argtypes = [C.c_void_p] + [C.c_void_p] * 12
        L.orc_get_nbl_pairs.restype = C.c_long
        L.orc_get_nbl_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
        L.orc_tot_presstens.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_set_interpolation.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_npt.argtypes = [C.c_void_p] + [C.c_double] * 5
        L.orc_get_npt.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_set_npt_axial.argtypes = [C.c_void_p] * 5 + [C.c_double, C.c_void_p, C.c_double]
        L.orc_get_npt_axial.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_set_berendsen.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orc_get_eeam.restype = C.c_long
        L.orc_get_adp.restype = C.c_long
        L.orc_get_adp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_get_eeam.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_deform_sample.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


# table interpolation of the reference build being mirrored (src/potaccess.h:24-36)
INTERP = {"3point": 0, "4point": 1, "spline": 2}


class OracleIMD:
    def __init__(self, ntypes, box, pbc=(1, 1, 1), nbl_margin=0.4, pair=None, embed=None, rho=None,
                 default_fmt=1, interp="3point", emod=None, adp_u=None, adp_w=None):
        L = lib()
        b = np.ascontiguousarray(np.asarray(box, dtype=np.float64).reshape(9))
        p = np.ascontiguousarray(np.asarray(pbc, dtype=np.int32))
        self.h = L.orc_create(int(ntypes), _p(b, C.c_double), _p(p, C.c_int), float(nbl_margin))
        self.press = False
        L.orc_set_interpolation(self.h, INTERP[interp] if isinstance(interp, str) else int(interp))
        self.eeam = emod is not None
        self.adp = adp_u is not None
        for which, path in ((PAIR, pair), (EMBED, embed), (RHO, rho), (EMOD, emod), (ADP_U, adp_u), (ADP_W, adp_w)):
            if path:
                rc = L.orc_read_table(self.h, which, os.fspath(path).encode())
                if rc:
                    raise RuntimeError(f"oracle: cannot read table {path} (rc={rc})")

    def __del__(self):
        try:
            lib().orc_destroy(self.h)
        except Exception:
            pass

    def set_atoms(self, nummer, sorte, masse, ort, impuls=None, vsorte=None):
        n = len(nummer)
        a = [np.ascontiguousarray(nummer, np.int32), np.ascontiguousarray(sorte, np.int32),
             None if vsorte is None else np.ascontiguousarray(vsorte, np.int32),
             np.ascontiguousarray(masse, np.float64), np.ascontiguousarray(ort, np.float64),
             None if impuls is None else np.ascontiguousarray(impuls, np.float64)]
        lib().orc_set_atoms(self.h, n, *[None if x is None else x.ctypes.data for x in a])

    def set_integrator(self, ensemble="nve", timestep=0.001, temperature=0.0, eta=0.0, isq_tau_eta=0.0):
        ens = {"nve": NVE, "nvt": NVT, "npt_iso": 2, "npt_axial": 3}[str(ensemble).lower()]
        lib().orc_set_integrator(self.h, ens, timestep, temperature, eta, isq_tau_eta)

    def set_npt(self, xi=0.0, Ekin_old=-1.0, pressure_ext=0.0, d_pressure=0.0, isq_tau_xi=0.0):
        lib().orc_set_npt(self.h, float(xi), float(Ekin_old), float(pressure_ext), float(d_pressure), float(isq_tau_xi))

    def set_npt_axial(self, xi, pressure_ext, d_pressure, relax_dirs=(1, 1, 1), Ekin_old=-1.0, dyn_stress=None, isq_tau_xi=0.0):
        v = [np.ascontiguousarray(x, np.float64) for x in (xi, pressure_ext, d_pressure)]
        rd = np.ascontiguousarray(relax_dirs, np.int32)
        dy = None if dyn_stress is None else np.ascontiguousarray(dyn_stress, np.float64)
        lib().orc_set_npt_axial(self.h, v[0].ctypes.data, v[1].ctypes.data, v[2].ctypes.data, rd.ctypes.data, float(Ekin_old),
                                None if dy is None else dy.ctypes.data, float(isq_tau_xi))

    def npt_axial(self):
        out = np.zeros(13)
        lib().orc_get_npt_axial(self.h, out.ctypes.data)
        return dict(xi=out[0:3].copy(), stress=out[3:6].copy(), pressure_ext=out[6:9].copy(), dyn_stress=out[9:12].copy(),
                    Ekin_old=float(out[12]))

    def set_berendsen(self, tauber, tot_kin_energy=0.0):
        lib().orc_set_berendsen(self.h, float(tauber), float(tot_kin_energy))

    def npt(self):
        out = np.zeros(4)
        lib().orc_get_npt(self.h, out.ctypes.data)
        return dict(zip(("xi", "Ekin_old", "pressure", "pressure_ext"), out.tolist()))

    def set_restrictions(self, restr):
        r = np.ascontiguousarray(restr, np.float64).reshape(-1, 3)
        lib().orc_set_restrictions(self.h, len(r), r.ctypes.data)

    def set_box(self, box):
        b = np.ascontiguousarray(np.asarray(box, np.float64).reshape(9))
        lib().orc_set_box(self.h, b.ctypes.data)

    def set_press_calc(self, on=True):
        self.press = bool(on)

    def calc_forces(self, step=0):
        lib().orc_calc_forces(self.h, int(self.press))

    def move_atoms(self):
        lib().orc_move_atoms(self.h, int(self.press))

    def check_nblist(self):
        lib().orc_check_nblist(self.h)

    def step(self, n=1):
        lib().orc_step(self.h, int(n))

    def lin_deform(self, dx, dy, dz, scale):
        v = [np.ascontiguousarray(x, np.float64) for x in (dx, dy, dz)]
        lib().orc_lin_deform(self.h, v[0].ctypes.data, v[1].ctypes.data, v[2].ctypes.data, float(scale))

    def deform_sample(self, deform_size, deform_shift, shear_def=None, deform_shear=None, deform_base=None):
        sh = np.ascontiguousarray(deform_shift, np.float64).reshape(-1, 3)
        n = len(sh)
        sd = np.zeros(n, np.int32) if shear_def is None else np.ascontiguousarray(shear_def, np.int32)
        ss = np.zeros((n, 3)) if deform_shear is None else np.ascontiguousarray(deform_shear, np.float64)
        bs = np.zeros((n, 3)) if deform_base is None else np.ascontiguousarray(deform_base, np.float64)
        lib().orc_deform_sample(self.h, float(deform_size), sh.ctypes.data, sd.ctypes.data, ss.ctypes.data, bs.ctypes.data)

    @property
    def natoms(self):
        return int(lib().orc_natoms(self.h))

    @property
    def have_valid_nbl(self):
        return int(lib().orc_have_valid_nbl(self.h))

    @property
    def nbl_count(self):
        return int(lib().orc_nbl_count(self.h))

    @property
    def cellsz(self):
        return float(lib().orc_cellsz(self.h))

    def celldims(self):
        out = np.zeros(6, np.int32)
        lib().orc_get_celldims(self.h, out.ctypes.data)
        return out[:3].copy(), out[3:].copy()

    def scalars(self):
        out = np.zeros(14)
        lib().orc_get_scalars(self.h, out.ctypes.data)
        keys = ["tot_pot_energy", "tot_kin_energy", "virial", "vir_xx", "vir_yy", "vir_zz",
                "vir_yz", "vir_zx", "vir_xy", "volume", "nactive", "eta", "temperature", "timestep"]
        return dict(zip(keys, out.tolist()))

    def box(self):
        out = np.zeros(9)
        lib().orc_get_box(self.h, out.ctypes.data)
        return out.reshape(3, 3)

    def tot_presstens(self):
        out = np.zeros(6)
        lib().orc_tot_presstens(self.h, out.ctypes.data)
        return out

    def atoms(self, sort=True):
        n = self.natoms
        d = dict(
            nummer=np.zeros(n, np.int32), sorte=np.zeros(n, np.int32), vsorte=np.zeros(n, np.int32),
            masse=np.zeros(n), ort=np.zeros((n, 3)), impuls=np.zeros((n, 3)), kraft=np.zeros((n, 3)),
            poteng=np.zeros(n), rho=np.zeros(n), dF=np.zeros(n), presstens=np.zeros((n, 6)),
            nblpos=np.zeros((n, 3)),
        )
        order = ["nummer", "sorte", "vsorte", "masse", "ort", "impuls", "kraft", "poteng", "rho", "dF",
                 "presstens", "nblpos"]
        lib().orc_get_atoms(self.h, *[d[k].ctypes.data for k in order])
        if self.eeam:
            d["eam_p"] = np.zeros(n); d["dM"] = np.zeros(n)
            lib().orc_get_eeam(self.h, d["eam_p"].ctypes.data, d["dM"].ctypes.data)
        if self.adp:
            d["adp_mu"] = np.zeros((n, 3)); d["adp_lambda"] = np.zeros((n, 6))
            lib().orc_get_adp(self.h, d["adp_mu"].ctypes.data, d["adp_lambda"].ctypes.data)
        if sort:
            o = np.argsort(d["nummer"], kind="stable")
            d = {k: v[o] for k, v in d.items()}
        return d

    def nbl_pairs(self):
        cnt = lib().orc_get_nbl_pairs(self.h, None, None, None, 0)
        if cnt < 0:
            raise RuntimeError("oracle has no valid neighbour list")
        pi = np.zeros(cnt, np.int32); pj = np.zeros(cnt, np.int32); sh = np.zeros((cnt, 3), np.int8)
        lib().orc_get_nbl_pairs(self.h, pi.ctypes.data, pj.ctypes.data, sh.ctypes.data, cnt)
        return np.stack([pi, pj], 1), sh

    def pair_int(self, which, col, r2):
        r2 = np.atleast_1d(np.asarray(r2, np.float64))
        v = np.zeros_like(r2); g = np.zeros_like(r2)
        a = C.c_double(); b = C.c_double(); s = C.c_int()
        for i, x in enumerate(r2):
            lib().orc_pair_int(self.h, which, col, float(x), C.byref(a), C.byref(b), C.byref(s))
            v[i] = a.value; g[i] = b.value
        return v, g

    def table_info(self, which, col):
        b = C.c_double(); e = C.c_double(); s = C.c_double(); n = C.c_int()
        rc = lib().orc_table_info(self.h, which, col, C.byref(b), C.byref(e), C.byref(s), C.byref(n))
        if rc:
            return None
        return dict(begin=b.value, end=e.value, step=s.value, len=n.value)


def canonical_pairs(pairs, shift):
    """Order-independent representation of a neighbour set: sorted rows (min_id, max_id, sx, sy, sz)
    with the shift expressed for the (min -> max) direction."""
    pairs = np.asarray(pairs, np.int64); shift = np.asarray(shift, np.int64)
    swap = pairs[:, 0] > pairs[:, 1]
    a = np.where(swap, pairs[:, 1], pairs[:, 0])
    b = np.where(swap, pairs[:, 0], pairs[:, 1])
    sh = np.where(swap[:, None], -shift, shift)
    # self-image pairs (a == b): normalise the sign of the shift
    same = a == b
    if same.any():
        neg = same & ((sh[:, 0] < 0) | ((sh[:, 0] == 0) & ((sh[:, 1] < 0) | ((sh[:, 1] == 0) & (sh[:, 2] < 0)))))
        sh[neg] = -sh[neg]
    rows = np.column_stack([a, b, sh])
    o = np.lexsort(rows.T[::-1])
    return rows[o]
