"""ctypes binding of the C ABI (include/imd_b200.h) -- the call a Python user makes.

This is a thin mirror of the reference's scripted step loop (src/imd.i:37-71, src/imd.py:191-382):
same function names (calc_forces, move_atoms, check_nblist, ...), same argument meaning.
There is no fallback: if imd_b200/libimd_b200.so is missing, or no sm_100 GPU is present,
construction raises.  Nothing here imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# IMDB200_LIB: another build of the same library (tools/build_variant.sh: kernel experiments, never a fallback)
LIB_PATH = os.environ.get("IMDB200_LIB") or os.path.join(HERE, "libimd_b200.so")

NVE, NVT, NPT_ISO, NPT_AXIAL = 0, 1, 2, 3
PAIR, EMBED, RHO = 0, 1, 2


class PotTable(C.Structure):
    _fields_ = [("begin", C.POINTER(C.c_double)), ("end", C.POINTER(C.c_double)),
                ("step", C.POINTER(C.c_double)), ("invstep", C.POINTER(C.c_double)),
                ("len", C.POINTER(C.c_int)), ("ncols", C.c_int), ("maxsteps", C.c_int),
                ("table", C.POINTER(C.c_double))]


class Config(C.Structure):
    _fields_ = [("ntypes", C.c_int), ("total_types", C.c_int),
                ("box_x", C.c_double * 3), ("box_y", C.c_double * 3), ("box_z", C.c_double * 3),
                ("pbc_dirs", C.c_int * 3), ("cpu_dim", C.c_int * 3), ("my_coord", C.c_int * 3),
                ("nbl_margin", C.c_double), ("nbl_size", C.c_double), ("timestep", C.c_double),
                ("ensemble", C.c_int), ("temperature", C.c_double), ("eta", C.c_double),
                ("isq_tau_eta", C.c_double), ("device", C.c_int), ("lanes_per_atom", C.c_int),
                ("interpolation", C.c_int), ("xi", C.c_double), ("isq_tau_xi", C.c_double),
                ("pressure_ext", C.c_double), ("d_pressure", C.c_double)]


class Scalars(C.Structure):
    _fields_ = [("tot_pot_energy", C.c_double), ("tot_kin_energy", C.c_double), ("virial", C.c_double),
                ("volume", C.c_double), ("eta", C.c_double), ("max_displacement2", C.c_double),
                ("tot_presstens", C.c_double * 6), ("natoms", C.c_longlong), ("nactive", C.c_longlong),
                ("nbl_len", C.c_longlong), ("have_valid_nbl", C.c_int), ("nbl_count", C.c_int),
                ("is_short", C.c_int), ("global_cell_dim", C.c_int * 3), ("cell_dim", C.c_int * 3),
                ("cellsz", C.c_double)]


EXPORTS = [
    "imdb200_default_config", "imdb200_last_error", "imdb200_set_error_handler", "imdb200_kernel_launches",
    "imdb200_create", "imdb200_destroy", "imdb200_set_potentials", "imdb200_set_restrictions",
    "imdb200_set_atoms", "imdb200_set_stream", "imdb200_comm_unique_id", "imdb200_comm_init", "imdb200_calc_forces",
    "imdb200_move_atoms", "imdb200_check_nblist", "imdb200_fix_cells", "imdb200_make_nblist", "imdb200_run",
    "imdb200_set_press_calc", "imdb200_set_skin_skip", "imdb200_invalidate_nblist", "imdb200_set_eta", "imdb200_set_temperature",
    "imdb200_lin_deform", "imdb200_deform_sample", "imdb200_get_scalars", "imdb200_get_atoms",
    "imdb200_natoms_local", "imdb200_get_nblist", "imdb200_pair_int", "imdb200_get_timers",
    "imdb200_read_pot_table", "imdb200_free_pot_table", "imdb200_calc_cpu_dim", "imdb200_cart_rank",
    "imdb200_cart_coords", "imdb200_halo_peers", "imdb200_send_forces", "imdb200_nghost_local", "imdb200_get_box", "imdb200_set_eeam_table", "imdb200_get_eeam", "imdb200_halo_message_order", "imdb200_set_npt_state", "imdb200_get_npt_state", "imdb200_set_adp_tables", "imdb200_get_adp", "imdb200_device_count", "imdb200_set_berendsen", "imdb200_set_npt_axial", "imdb200_get_npt_axial", "imdb200_set_momenta",
]

_lib = None


def load_library():
    """Load libimd_b200.so (needs libcudart, not a GPU).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: run `make` (or __graft_entry__.build()); "
                           "imd_b200 has no CPU or PyTorch fallback")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.imdb200_last_error.restype = C.c_char_p
    L.imdb200_kernel_launches.restype = C.c_longlong
    L.imdb200_default_config.argtypes = [C.POINTER(Config)]
    L.imdb200_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.imdb200_destroy.argtypes = [vp]
    L.imdb200_set_potentials.argtypes = [vp, C.POINTER(PotTable), C.POINTER(PotTable), C.POINTER(PotTable)]
    L.imdb200_set_restrictions.argtypes = [vp, C.c_int, vp]
    L.imdb200_set_atoms.argtypes = [vp, C.c_long] + [vp] * 6
    L.imdb200_calc_forces.argtypes = [vp, C.c_int]
    L.imdb200_set_stream.argtypes = [vp, vp]
    for f in ("move_atoms", "check_nblist", "fix_cells", "make_nblist", "invalidate_nblist"):
        getattr(L, "imdb200_" + f).argtypes = [vp]
    L.imdb200_run.argtypes = [vp, C.c_int]
    L.imdb200_set_press_calc.argtypes = [vp, C.c_int]
    L.imdb200_set_skin_skip.argtypes = [vp, C.c_int]
    L.imdb200_set_eta.argtypes = [vp, C.c_double]
    L.imdb200_set_momenta.argtypes = [vp, C.c_long, vp, vp]
    L.imdb200_set_temperature.argtypes = [vp, C.c_double]
    L.imdb200_set_berendsen.argtypes = [vp, C.c_double, C.c_double]
    L.imdb200_lin_deform.argtypes = [vp, vp, vp, vp, C.c_double]
    L.imdb200_deform_sample.argtypes = [vp, C.c_double, vp, vp, vp, vp]
    L.imdb200_get_scalars.argtypes = [vp, C.POINTER(Scalars)]
    L.imdb200_get_atoms.restype = C.c_long
    L.imdb200_get_adp.restype = C.c_long
    L.imdb200_get_adp.argtypes = [vp, vp, vp]
    L.imdb200_set_adp_tables.argtypes = [vp, C.POINTER(PotTable), C.POINTER(PotTable)]
    L.imdb200_get_eeam.restype = C.c_long
    L.imdb200_get_eeam.argtypes = [vp, vp, vp]
    L.imdb200_set_eeam_table.argtypes = [vp, C.POINTER(PotTable)]
    L.imdb200_get_box.argtypes = [vp, vp]
    L.imdb200_set_npt_state.argtypes = [vp, C.c_double, C.c_double, C.c_double]
    L.imdb200_get_npt_state.argtypes = [vp, vp]
    L.imdb200_set_npt_axial.argtypes = [vp, vp, vp, vp, vp, C.c_double, vp]
    L.imdb200_get_npt_axial.argtypes = [vp, vp]
    L.imdb200_get_atoms.argtypes = [vp] + [vp] * 12
    L.imdb200_natoms_local.restype = C.c_long
    L.imdb200_natoms_local.argtypes = [vp]
    L.imdb200_get_nblist.restype = C.c_long
    L.imdb200_get_nblist.argtypes = [vp, vp, vp, vp, C.c_long]
    L.imdb200_pair_int.argtypes = [vp, C.c_int, C.c_int, C.c_long, vp, vp, vp]
    L.imdb200_get_timers.argtypes = [vp, vp, C.c_int]
    L.imdb200_read_pot_table.argtypes = [C.POINTER(PotTable), C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.POINTER(C.c_double)]
    L.imdb200_free_pot_table.argtypes = [C.POINTER(PotTable)]
    L.imdb200_comm_unique_id.argtypes = [vp]
    L.imdb200_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    i3 = C.POINTER(C.c_int)
    L.imdb200_calc_cpu_dim.argtypes = [C.c_int, i3]
    L.imdb200_calc_cpu_dim.restype = None
    L.imdb200_cart_rank.argtypes = [i3, i3]
    L.imdb200_cart_coords.argtypes = [C.c_int, i3, i3]
    L.imdb200_cart_coords.restype = None
    L.imdb200_halo_peers.argtypes = [i3, i3, i3, i3, i3]
    L.imdb200_halo_peers.restype = None
    L.imdb200_send_forces.argtypes = [vp, vp, C.c_int, C.c_long]
    L.imdb200_nghost_local.restype = C.c_long
    L.imdb200_nghost_local.argtypes = [vp]
    _lib = L
    return L


class IMDError(RuntimeError):
    pass


def _chk(rc):
    if rc != 0:
        raise IMDError(f"imd_b200 error {rc}: {load_library().imdb200_last_error().decode()}")


def read_pot_table(path, ncols, radial, ntypes, default_format=2):
    """read_pot_table (src/imd_potential.c:161-282) through the host-side C reader."""
    L = load_library()
    pt = PotTable()
    cellsz = C.c_double(0.0)
    rc = L.imdb200_read_pot_table(C.byref(pt), os.fspath(path).encode(), ncols, int(radial), ntypes,
                                  default_format, C.byref(cellsz))
    if rc:
        raise IMDError(f"cannot read potential table {path}")
    return pt, cellsz.value


def kernel_launches():
    return int(load_library().imdb200_kernel_launches())


# ---- process grid helpers (host-side C, no GPU needed) -----------------------------------------------
def _i3(v):
    return (C.c_int * 3)(*[int(x) for x in v])


def calc_cpu_dim(num_cpus, cpu_dim=(0, 0, 0)):
    """calc_cpu_dim (src/imd_geom_mpi_3d.c:201-266)."""
    cd = _i3(cpu_dim)
    load_library().imdb200_calc_cpu_dim(int(num_cpus), cd)
    return tuple(cd)


def cart_coords(rank, cpu_dim):
    out = _i3((0, 0, 0))
    load_library().imdb200_cart_coords(int(rank), _i3(cpu_dim), out)
    return tuple(out)


def cart_rank(coord, cpu_dim):
    return int(load_library().imdb200_cart_rank(_i3(coord), _i3(cpu_dim)))


def halo_peers(cpu_dim, my_coord, pbc=(1, 1, 1)):
    """26 neighbour ranks + image shift codes (setup_mpi_topology, src/imd_geom_mpi_3d.c:57-88)."""
    peer = (C.c_int * 27)(); code = (C.c_int * 27)()
    load_library().imdb200_halo_peers(_i3(cpu_dim), _i3(my_coord), _i3(pbc), peer, code)
    return list(peer), list(code)


def halo_message_order(peer, my_rank):
    """Peer-major order of the receive and send regions (one message per neighbour rank), see imd_b200.h."""
    p = (C.c_int * 27)(*peer)
    ro = (C.c_int * 26)(); so = (C.c_int * 26)(); nr = C.c_int(); ns = C.c_int()
    load_library().imdb200_halo_message_order(p, int(my_rank), ro, C.byref(nr), so, C.byref(ns))
    return list(ro)[:nr.value], list(so)[:ns.value]


def comm_unique_id():
    """rank 0: a fresh ncclUniqueId (128 bytes) to hand to every rank's IMDB200.comm_init."""
    buf = C.create_string_buffer(128)
    _chk(load_library().imdb200_comm_unique_id(buf))
    return buf.raw


# table interpolation (IMDB200_INTERP_*): what IMD's `4point` / `spline` make targets select at compile time
INTERP = {"3point": 0, "4point": 1, "spline": 2}


class IMDB200:
    """One simulation domain on one B200."""

    def __init__(self, ntypes, box, pbc=(1, 1, 1), nbl_margin=0.4, nbl_size=1.1, pair=None, embed=None,
                 rho=None, default_fmt=None, ensemble="nve", timestep=0.001, temperature=0.0, eta=0.0,
                 isq_tau_eta=0.0, device=-1, lanes_per_atom=0, total_types=None, cpu_dim=(1, 1, 1),
                 my_coord=(0, 0, 0), interp="3point", emod=None, xi=0.0, isq_tau_xi=0.0, pressure_ext=0.0,
                 d_pressure=0.0, adp_u=None, adp_w=None):
        L = load_library()
        self.L = L
        cfg = Config()
        L.imdb200_default_config(C.byref(cfg))
        cfg.ntypes = int(ntypes)
        cfg.total_types = int(total_types or ntypes)
        b = np.asarray(box, np.float64).reshape(3, 3)
        for d in range(3):
            cfg.box_x[d], cfg.box_y[d], cfg.box_z[d] = b[0, d], b[1, d], b[2, d]
            cfg.pbc_dirs[d] = int(pbc[d]); cfg.cpu_dim[d] = int(cpu_dim[d]); cfg.my_coord[d] = int(my_coord[d])
        cfg.nbl_margin = nbl_margin; cfg.nbl_size = nbl_size; cfg.timestep = timestep
        cfg.ensemble = {"nve": NVE, "nvt": NVT, "npt_iso": NPT_ISO, "npt_axial": NPT_AXIAL}[str(ensemble).lower()]
        cfg.xi = xi; cfg.isq_tau_xi = isq_tau_xi; cfg.pressure_ext = pressure_ext; cfg.d_pressure = d_pressure
        cfg.temperature = temperature; cfg.eta = eta; cfg.isq_tau_eta = isq_tau_eta
        cfg.device = device; cfg.lanes_per_atom = lanes_per_atom
        cfg.interpolation = INTERP[interp] if isinstance(interp, str) else int(interp)
        self.h = C.c_void_p()
        _chk(L.imdb200_create(C.byref(cfg), C.byref(self.h)))
        self.ntypes = int(ntypes)
        self._tabs = []
        self.eeam = False
        if pair is not None:
            self.set_potentials(pair, embed, rho, default_fmt)
        if emod is not None:
            self.set_eeam_table(emod)
        self.adp = False
        if adp_u is not None:
            self.set_adp_tables(adp_u, adp_w)

    def close(self):
        if getattr(self, "h", None):
            self.L.imdb200_destroy(self.h)
            self.h = None
            for t in self._tabs:
                self.L.imdb200_free_pot_table(C.byref(t))
            self._tabs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- setup ---------------------------------------------------------------------------
    def set_potentials(self, pair, embed=None, rho=None, default_fmt=None):
        nt = self.ntypes
        eam = rho is not None
        fmt = default_fmt if default_fmt is not None else (2 if eam else 1)
        tp, _ = read_pot_table(pair, nt * nt, 1, nt, fmt)
        self._tabs = [tp]
        te = tr = None
        if eam:
            te, _ = read_pot_table(embed, nt, 0, nt, 2)
            tr, _ = read_pot_table(rho, nt * nt, 1, nt, 2)
            self._tabs += [te, tr]
        _chk(self.L.imdb200_set_potentials(self.h, C.byref(tp), C.byref(te) if eam else None,
                                           C.byref(tr) if eam else None))

    def set_eeam_table(self, emod):
        """`eeam_energy_file` of an EEAM build: M(p), ntypes columns, not radial."""
        tm, _ = read_pot_table(emod, self.ntypes, 0, self.ntypes, 2)
        self._tabs.append(tm)
        _chk(self.L.imdb200_set_eeam_table(self.h, C.byref(tm)))
        self.eeam = True

    def set_adp_tables(self, adp_u, adp_w):
        """`adp_upotfile` / `adp_wpotfile` of an ADP build: u(r), w(r), ntypes^2 columns in r^2, radial."""
        nt = self.ntypes
        tu, _ = read_pot_table(adp_u, nt * nt, 1, nt, 2)
        tw, _ = read_pot_table(adp_w, nt * nt, 1, nt, 2)
        self._tabs += [tu, tw]
        _chk(self.L.imdb200_set_adp_tables(self.h, C.byref(tu), C.byref(tw)))
        self.adp = True

    def comm_init(self, unique_id, rank, nranks):
        """Join the NCCL communicator of the process grid (one process per GPU); see imdb200_comm_init."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        _chk(self.L.imdb200_comm_init(self.h, buf, int(rank), int(nranks)))

    def send_forces(self, dev_ptr, ncomp, stride):
        """send_forces analogue on a caller-owned device field (see imdb200_send_forces)."""
        _chk(self.L.imdb200_send_forces(self.h, C.c_void_p(int(dev_ptr)), int(ncomp), int(stride)))

    @property
    def nghost(self):
        return int(self.L.imdb200_nghost_local(self.h))

    def set_atoms(self, nummer, sorte, masse, ort, impuls=None, vsorte=None):
        a = [np.ascontiguousarray(nummer, np.int32), np.ascontiguousarray(sorte, np.int32),
             None if vsorte is None else np.ascontiguousarray(vsorte, np.int32),
             np.ascontiguousarray(masse, np.float64), np.ascontiguousarray(ort, np.float64),
             None if impuls is None else np.ascontiguousarray(impuls, np.float64)]
        _chk(self.L.imdb200_set_atoms(self.h, len(a[0]), *[None if x is None else x.ctypes.data for x in a]))

    def set_integrator(self, ensemble="nve", timestep=0.001, temperature=0.0, eta=0.0, isq_tau_eta=0.0):
        raise IMDError("integrator parameters are fixed at construction (imdb200_config)")

    def set_restrictions(self, restr):
        r = np.ascontiguousarray(restr, np.float64).reshape(-1, 3)
        _chk(self.L.imdb200_set_restrictions(self.h, len(r), r.ctypes.data))

    def set_stream(self, cuda_stream):
        """cuda_stream: integer handle, e.g. torch.cuda.current_stream().cuda_stream"""
        _chk(self.L.imdb200_set_stream(self.h, C.c_void_p(int(cuda_stream))))

    def set_press_calc(self, on=True):
        _chk(self.L.imdb200_set_press_calc(self.h, int(on)))

    def set_skin_skip(self, on=True):
        """on=False: walk every stored list entry like the reference (test hook; results are bit-identical)."""
        _chk(self.L.imdb200_set_skin_skip(self.h, int(on)))

    def set_berendsen(self, tauber, tot_kin_energy=0.0):
        """Berendsen variant of NVE (`ber` builds): tau_berendsen and the kinetic energy of the previous step."""
        _chk(self.L.imdb200_set_berendsen(self.h, float(tauber), float(tot_kin_energy)))

    def set_momenta(self, nummer, impuls):
        """Overwrite the momenta of the local atoms, matched by atom number (Andersen thermostat: maxwell() on the host)."""
        num = np.ascontiguousarray(nummer, np.int32); p = np.ascontiguousarray(impuls, np.float64)
        _chk(self.L.imdb200_set_momenta(self.h, len(num), num.ctypes.data, p.ctypes.data))

    def set_eta(self, eta):
        _chk(self.L.imdb200_set_eta(self.h, float(eta)))

    def invalidate_nbl(self):
        _chk(self.L.imdb200_invalidate_nblist(self.h))

    # --- step loop -----------------------------------------------------------------------
    def calc_forces(self, step=0):
        _chk(self.L.imdb200_calc_forces(self.h, int(step)))

    def move_atoms(self):
        _chk(self.L.imdb200_move_atoms(self.h))

    def check_nblist(self):
        _chk(self.L.imdb200_check_nblist(self.h))

    def step(self, n=1):
        _chk(self.L.imdb200_run(self.h, int(n)))

    run = step

    def lin_deform(self, dx, dy, dz, scale):
        v = [np.ascontiguousarray(x, np.float64) for x in (dx, dy, dz)]
        _chk(self.L.imdb200_lin_deform(self.h, v[0].ctypes.data, v[1].ctypes.data, v[2].ctypes.data, float(scale)))

    def deform_sample(self, deform_size, deform_shift, shear_def=None, deform_shear=None, deform_base=None):
        sh = np.ascontiguousarray(deform_shift, np.float64).reshape(-1, 3)
        n = len(sh)
        sd = np.zeros(n, np.int32) if shear_def is None else np.ascontiguousarray(shear_def, np.int32)
        ss = np.zeros((n, 3)) if deform_shear is None else np.ascontiguousarray(deform_shear, np.float64)
        bs = np.zeros((n, 3)) if deform_base is None else np.ascontiguousarray(deform_base, np.float64)
        _chk(self.L.imdb200_deform_sample(self.h, float(deform_size), sh.ctypes.data, sd.ctypes.data,
                                          ss.ctypes.data, bs.ctypes.data))

    # --- results -------------------------------------------------------------------------
    def raw_scalars(self):
        s = Scalars()
        _chk(self.L.imdb200_get_scalars(self.h, C.byref(s)))
        return s

    def scalars(self):
        s = self.raw_scalars()
        return dict(tot_pot_energy=s.tot_pot_energy, tot_kin_energy=s.tot_kin_energy, virial=s.virial,
                    vir_xx=0.0, vir_yy=0.0, vir_zz=0.0, vir_yz=0.0, vir_zx=0.0, vir_xy=0.0,
                    volume=s.volume, nactive=float(s.nactive), eta=s.eta, temperature=0.0, timestep=0.0,
                    max_displacement2=s.max_displacement2, nbl_len=int(s.nbl_len), is_short=int(s.is_short))

    @property
    def natoms(self):
        return int(self.L.imdb200_natoms_local(self.h))

    @property
    def have_valid_nbl(self):
        return int(self.raw_scalars().have_valid_nbl)

    @property
    def nbl_count(self):
        return int(self.raw_scalars().nbl_count)

    @property
    def cellsz(self):
        return float(self.raw_scalars().cellsz)

    def celldims(self):
        s = self.raw_scalars()
        return np.array(list(s.global_cell_dim), np.int32), np.array(list(s.cell_dim), np.int32)

    def set_npt_state(self, xi=0.0, Ekin_old=-1.0, pressure_ext=0.0):
        """NPT_iso hand-over: xi, twice the kinetic energy of the previous step (< 0: from the momenta), pressure_ext."""
        _chk(self.L.imdb200_set_npt_state(self.h, float(xi), float(Ekin_old), float(pressure_ext)))

    def set_npt_axial(self, xi, pressure_ext, d_pressure=(0.0, 0.0, 0.0), relax_dirs=(1, 1, 1), Ekin_old=-1.0, dyn_stress=None):
        """NPT_axial hand-over (globals xi, pressure_ext, relax_dirs, dyn_stress_x/y/z, Ekin_old of the reference)."""
        v = [np.ascontiguousarray(x, np.float64) for x in (xi, pressure_ext, d_pressure)]
        rd = np.ascontiguousarray(relax_dirs, np.int32)
        dy = None if dyn_stress is None else np.ascontiguousarray(dyn_stress, np.float64)
        _chk(self.L.imdb200_set_npt_axial(self.h, v[0].ctypes.data, v[1].ctypes.data, v[2].ctypes.data, rd.ctypes.data,
                                          float(Ekin_old), None if dy is None else dy.ctypes.data))

    def npt_axial(self):
        out = np.zeros(13)
        _chk(self.L.imdb200_get_npt_axial(self.h, out.ctypes.data))
        return dict(xi=out[0:3].copy(), stress=out[3:6].copy(), pressure_ext=out[6:9].copy(), dyn_stress=out[9:12].copy(),
                    Ekin_old=float(out[12]))

    def npt(self):
        out = np.zeros(4)
        _chk(self.L.imdb200_get_npt_state(self.h, out.ctypes.data))
        return dict(zip(("xi", "Ekin_old", "pressure", "pressure_ext"), out.tolist()))

    def box(self):
        out = np.zeros(9)
        _chk(self.L.imdb200_get_box(self.h, out.ctypes.data))
        return out.reshape(3, 3)

    def tot_presstens(self):
        return np.array(list(self.raw_scalars().tot_presstens))

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
        got = self.L.imdb200_get_atoms(self.h, *[d[k].ctypes.data for k in order])
        assert got == n
        if self.eeam:
            d["eam_p"] = np.zeros(n); d["dM"] = np.zeros(n)
            assert self.L.imdb200_get_eeam(self.h, d["eam_p"].ctypes.data, d["dM"].ctypes.data) == n
        if self.adp:
            d["adp_mu"] = np.zeros((n, 3)); d["adp_lambda"] = np.zeros((n, 6))
            assert self.L.imdb200_get_adp(self.h, d["adp_mu"].ctypes.data, d["adp_lambda"].ctypes.data) == n
        if sort:
            o = np.argsort(d["nummer"], kind="stable")
            d = {k: v[o] for k, v in d.items()}
        return d

    def nbl_pairs(self):
        """Full list as (nummer_i, nummer_j) + image shift of j; every pair appears in both directions."""
        cnt = self.L.imdb200_get_nblist(self.h, None, None, None, 0)
        if cnt < 0:
            raise IMDError("no valid neighbour list")
        pi = np.zeros(cnt, np.int32); pj = np.zeros(cnt, np.int32); sh = np.zeros((cnt, 3), np.int8)
        got = self.L.imdb200_get_nblist(self.h, pi.ctypes.data, pj.ctypes.data, sh.ctypes.data, cnt)
        assert got == cnt
        return np.stack([pi, pj], 1), sh

    def pair_int(self, which, col, r2):
        r2 = np.ascontiguousarray(np.atleast_1d(r2), np.float64)
        v = np.zeros_like(r2); g = np.zeros_like(r2)
        _chk(self.L.imdb200_pair_int(self.h, which, col, len(r2), r2.ctypes.data, v.ctypes.data, g.ctypes.data))
        return v, g

    def timers(self, reset=False):
        out = np.zeros(8)
        self.L.imdb200_get_timers(self.h, out.ctypes.data, int(reset))
        keys = ["rebuild_ms", "pass1_ms", "pass2_ms", "integrate_ms", "ghost_ms", "rebuilds", "steps", "_"]
        return dict(zip(keys, out.tolist()))
