"""Synthetic workload inputs in IMD's own file formats.

The reference ships no potential files or example inputs, so the
benchmark/parity workloads are generated here.  Everything is written in the formats the
reference reads:

* potential tables, format 2 (``#F 2 <ncols>`` / ``#E``; ``begin end step`` per column;
  then the columns one after the other) -- reader: src/imd_potential.c:394-462;
* potential tables, format 1 (one ``r2 V00 V01 ...`` line per sample, no header) --
  reader: src/imd_potential.c:297-376; this is what util/imd_mklj.c:50-66 writes;
* parameter files -- tag/value lines, src/imd_param.c:251-312.

The functional forms are smooth, Cu-like / Ni-Al-like / LJ-Ar-like toy models.  They are
workload generators, not physics: only the table *shape* (rows, columns, cut-offs) matters
for the benchmark, and parity is always measured against the reference run on the SAME files.
"""
from __future__ import annotations

import os
import numpy as np

# --- unit conventions of the survey's probe inputs ------------------
CU_A0 = 3.615          # Angstrom
CU_MASS = 0.0065850    # amu * 1.0364e-4  (IMD internal mass unit for eV/A/ps... see survey)
NI_MASS = 0.006083
AL_MASS = 0.002796
AR_A0 = 5.26
AR_MASS = 0.0041402


def _cut(r, rc, h):
    """Smooth cut-off psi(x) = x^4/(1+x^4), x=(r-rc)/h for r<rc, else 0."""
    x = np.minimum((r - rc) / h, 0.0)
    x4 = x ** 4
    return x4 / (1.0 + x4)


def eam_functions(kind: str = "cu"):
    """Return (phi(r, a, b), rho(r, a, b), F(rho, a), ntypes, r_cut) for a toy EAM model."""
    rc, h = 5.5, 0.6
    if kind == "cu":
        ntypes = 1
        D = [[0.12]]; al = [[1.45]]; r0 = [[2.62]]
        fe = [1.0]; beta = [4.2]; re = [2.556]
        A = [0.55]
    elif kind == "nial":
        ntypes = 2
        D = [[0.14, 0.16], [0.16, 0.10]]
        al = [[1.50, 1.40], [1.40, 1.25]]
        r0 = [[2.55, 2.60], [2.60, 2.80]]
        fe = [1.0, 0.8]; beta = [4.4, 3.6]; re = [2.49, 2.86]
        A = [0.60, 0.42]
    else:
        raise ValueError(kind)

    def phi(r, a, b):
        e = np.exp(-al[a][b] * (r - r0[a][b]))
        return D[a][b] * (e * e - 2.0 * e) * _cut(r, rc, h)

    def rho(r, a, b):
        # density received by an atom of type a from a neighbour of type b:
        # column a*ntypes+b of rho_h_tab (src/imd_forces_nbl.c:1176-1178)
        return fe[b] * np.exp(-beta[b] * (r / re[b] - 1.0)) * _cut(r, rc, h)

    def F(x, a):
        return -A[a] * np.sqrt(x + 0.05) + 2.0e-4 * x * x

    return phi, rho, F, ntypes, rc


def write_table2(path, begins, ends, steps, columns):
    """Write a format-2 table (src/imd_potential.c:394-462)."""
    ncols = len(columns)
    with open(path, "w") as f:
        f.write(f"#F 2 {ncols}\n#E\n")
        for b, e, s in zip(begins, ends, steps):
            f.write(f"{b:.16e} {e:.16e} {s:.16e}\n")
        f.write("\n")
        for col in columns:
            f.write("\n".join(f"{v:.16e}" for v in col))
            f.write("\n\n")


def write_table1(path, r2, columns, fmt="%.16e"):
    """Write a format-1 table (src/imd_potential.c:297-376): r2 V00 V01 ... per line."""
    arr = np.column_stack([r2] + list(columns))
    np.savetxt(path, arr, fmt=fmt)


def make_eam_tables(outdir, kind="cu", prefix=None, nr=2001, nrho=4001,
                    r2_begin=1.0, rho_end=40.0, per_column=False):
    """Write <prefix>_phi.pot, _rho.pot, _F.pot (format 2).  Returns dict of paths + r_cut.
    per_column: every column gets its own begin / end / step (phi_AB and phi_BA stay one function), the way hand-made IMD
    tables may look -- the case the kernels' general several-species path exists for."""
    os.makedirs(outdir, exist_ok=True)
    prefix = prefix or kind
    phi, rho, F, nt, rc = eam_functions(kind)
    if per_column:
        ncol = nt * nt
        paths = {
            "core_potential_file": os.path.join(outdir, f"{prefix}_phi.pot"),
            "atomic_e-density_file": os.path.join(outdir, f"{prefix}_rho.pot"),
            "embedding_energy_file": os.path.join(outdir, f"{prefix}_F.pot"),
        }
        def grid(b, e):
            st = (e - b) / (nr - 1)
            return b, e, st, np.sqrt(b + st * np.arange(nr))
        pg = [grid(r2_begin + 0.04 * ((c // nt) + (c % nt)), (rc - 0.07 * ((c // nt) + (c % nt))) ** 2) for c in range(ncol)]
        rg = [grid(r2_begin + 0.03 * (c % nt) + 0.01 * (c // nt), (rc - 0.11 * (c % nt) - 0.02 * (c // nt)) ** 2) for c in range(ncol)]
        write_table2(paths["core_potential_file"], [g[0] for g in pg], [g[1] for g in pg], [g[2] for g in pg],
                     [phi(pg[c][3], c // nt, c % nt) for c in range(ncol)])
        write_table2(paths["atomic_e-density_file"], [g[0] for g in rg], [g[1] for g in rg], [g[2] for g in rg],
                     [rho(rg[c][3], c // nt, c % nt) for c in range(ncol)])
        rstep = rho_end / (nrho - 1)
        x = rstep * np.arange(nrho)
        write_table2(paths["embedding_energy_file"], [0.0] * nt, [rho_end] * nt, [rstep] * nt, [F(x, a) for a in range(nt)])
        paths["r_cut"] = rc
        paths["ntypes"] = nt
        return paths
    r2_end = rc * rc
    step = (r2_end - r2_begin) / (nr - 1)
    r2 = r2_begin + step * np.arange(nr)
    r = np.sqrt(r2)
    ncol = nt * nt
    phic = [phi(r, c // nt, c % nt) for c in range(ncol)]
    rhoc = [rho(r, c // nt, c % nt) for c in range(ncol)]
    rstep = rho_end / (nrho - 1)
    x = rstep * np.arange(nrho)
    Fc = [F(x, a) for a in range(nt)]
    paths = {
        "core_potential_file": os.path.join(outdir, f"{prefix}_phi.pot"),
        "atomic_e-density_file": os.path.join(outdir, f"{prefix}_rho.pot"),
        "embedding_energy_file": os.path.join(outdir, f"{prefix}_F.pot"),
    }
    write_table2(paths["core_potential_file"], [r2_begin] * ncol, [r2_end] * ncol, [step] * ncol, phic)
    write_table2(paths["atomic_e-density_file"], [r2_begin] * ncol, [r2_end] * ncol, [step] * ncol, rhoc)
    write_table2(paths["embedding_energy_file"], [0.0] * nt, [rho_end] * nt, [rstep] * nt, Fc)
    paths["r_cut"] = rc
    paths["ntypes"] = nt
    return paths


def make_eeam_table(outdir, nt=1, prefix="eeam", npts=1201, p_end=30.0):
    """EEAM energy modification term M(p), p = sum_j rho_j(r_ij)^2 (`eeam_energy_file`, format 2, one column per
    type, not radial): a smooth synthetic function with non-trivial curvature."""
    os.makedirs(outdir, exist_ok=True)
    step = p_end / (npts - 1)
    x = step * np.arange(npts)
    cols = [(-0.08 - 0.02 * a) * np.sqrt(x + 0.5) + (6.0e-4 + 2.0e-4 * a) * x * x for a in range(nt)]
    path = os.path.join(outdir, f"{prefix}_M.pot")
    write_table2(path, [0.0] * nt, [p_end] * nt, [step] * nt, cols)
    return path


def make_adp_tables(outdir, nt=1, prefix="adp", nr=701, r2_begin=1.0, rc=5.5):
    """ADP dipole u(r) and quadrupole w(r) distortion functions (`adp_upotfile`, `adp_wpotfile`, format 2,
    ntypes^2 columns in r^2, radial): smooth synthetic functions that vanish at the cut-off."""
    os.makedirs(outdir, exist_ok=True)
    step = (rc * rc - r2_begin) / (nr - 1)
    r2 = r2_begin + step * np.arange(nr)
    r = np.sqrt(r2)
    ncol = nt * nt
    ucols = [(0.030 + 0.006 * ((c // nt) + (c % nt))) * np.exp(-0.9 * (r - 2.5)) * _cut(r, rc, 0.6) for c in range(ncol)]
    wcols = [(-0.012 + 0.003 * ((c // nt) + (c % nt))) * np.exp(-0.7 * (r - 2.5)) * _cut(r, rc, 0.6) for c in range(ncol)]
    pu, pw = os.path.join(outdir, f"{prefix}_u.pot"), os.path.join(outdir, f"{prefix}_w.pot")
    write_table2(pu, [r2_begin] * ncol, [rc * rc] * ncol, [step] * ncol, ucols)
    write_table2(pw, [r2_begin] * ncol, [rc * rc] * ncol, [step] * ncol, wcols)
    return pu, pw


def make_lj_table(outdir, name="lj_ar.pot", eps=0.0104, sigma=3.40, r_begin=2.0, r_cut=8.5,
                  nsteps=5000, ntypes=2):
    """Tabulated LJ pair potential in format 1, the way util/imd_mklj.c:50-66 lays it out
    (equidistant in r^2, one column per type pair, E*((s/r)^12 - 2 (s/r)^6) with s the
    position of the minimum).  Written at full double precision."""
    os.makedirs(outdir, exist_ok=True)
    rmin = sigma * 2.0 ** (1.0 / 6.0)
    step = (r_cut ** 2 - r_begin ** 2) / nsteps
    r2 = r_begin ** 2 + step * np.arange(nsteps + 1)
    r = np.sqrt(r2)
    s6 = (rmin / r) ** 6
    v = eps * (s6 * s6 - 2.0 * s6)
    path = os.path.join(outdir, name)
    write_table1(path, r2, [v] * (ntypes * ntypes))
    return {"potfile": path, "r_cut": r_cut, "ntypes": ntypes}


def write_param(path, **kw):
    """Write an IMD parameter file (tag value ... per line; src/imd_param.c:251-312).
    Sequence values are joined by blanks."""
    with open(path, "w") as f:
        for k, v in kw.items():
            if isinstance(v, (list, tuple, np.ndarray)):
                v = " ".join(_fmt(x) for x in v)
            else:
                v = _fmt(v)
            f.write(f"{k:24s} {v}\n")
    return path


def _fmt(x):
    if isinstance(x, (float, np.floating)):
        return repr(float(x))
    return str(x)


def cu_param(outdir, ncell=(8, 8, 8), *, name="cu", ensemble="nve", maxsteps=20, starttemp=0.05,
             seed=12345, timestep=0.001, nbl_margin=0.4, tables=None, coordname="_fcc", extra=None):
    """Parameter file of BASELINE config 2 (EAM Cu fcc, Verlet list + skin) at a given size."""
    os.makedirs(outdir, exist_ok=True)
    tables = tables or make_eam_tables(outdir, "cu")
    kw = dict(
        coordname=coordname, outfiles=os.path.join(outdir, name), ensemble=ensemble,
        maxsteps=maxsteps, startstep=0, timestep=timestep, ntypes=1, total_types=1,
        masses=CU_MASS, box_param=list(ncell), box_unit=CU_A0, pbc_dirs=[1, 1, 1],
        starttemp=starttemp, seed=seed, eng_int=0, checkpt_int=0,
        core_potential_file=tables["core_potential_file"],
        embedding_energy_file=tables["embedding_energy_file"],
        **{"atomic_e-density_file": tables["atomic_e-density_file"]},
        nbl_margin=nbl_margin, nbl_size=1.3,
    )
    if ensemble == "nvt":
        kw.update(endtemp=starttemp, tau_eta=0.1, eta=0.0)
    if extra:
        kw.update(extra)
    return write_param(os.path.join(outdir, name + ".param"), **kw)


def nial_param(outdir, ncell=(8, 8, 8), *, name="nial", ensemble="nvt", maxsteps=20, starttemp=0.05,
               seed=12345, timestep=0.001, nbl_margin=0.4, tables=None, extra=None):
    """Parameter file of BASELINE config 3 (binary EAM Ni-Al, B2 structure, two-species tables)."""
    os.makedirs(outdir, exist_ok=True)
    tables = tables or make_eam_tables(outdir, "nial")
    kw = dict(
        coordname="_b2", outfiles=os.path.join(outdir, name), ensemble=ensemble,
        maxsteps=maxsteps, startstep=0, timestep=timestep, ntypes=2, total_types=2,
        masses=[NI_MASS, AL_MASS], types=[0, 1], box_param=list(ncell), box_unit=2.88, pbc_dirs=[1, 1, 1],
        starttemp=starttemp, seed=seed, eng_int=0, checkpt_int=0,
        core_potential_file=tables["core_potential_file"],
        embedding_energy_file=tables["embedding_energy_file"],
        **{"atomic_e-density_file": tables["atomic_e-density_file"]},
        nbl_margin=nbl_margin, nbl_size=1.3,
    )
    if ensemble == "nvt":
        kw.update(endtemp=starttemp, tau_eta=0.1, eta=0.0)
    if extra:
        kw.update(extra)
    return write_param(os.path.join(outdir, name + ".param"), **kw)


def lj_param(outdir, ncell=(8, 8, 8), *, name="lj", maxsteps=20, starttemp=0.0043, seed=12345,
             timestep=0.002, nbl_margin=0.4, table=None, extra=None):
    """Parameter file of BASELINE config 1 (LJ Ar fcc, tabulated pair potential, NVE)."""
    os.makedirs(outdir, exist_ok=True)
    table = table or make_lj_table(outdir)
    kw = dict(
        coordname="_fcc", outfiles=os.path.join(outdir, name), ensemble="nve",
        maxsteps=maxsteps, startstep=0, timestep=timestep, ntypes=2, total_types=2,
        masses=[AR_MASS, AR_MASS], box_param=list(ncell), box_unit=AR_A0, pbc_dirs=[1, 1, 1],
        starttemp=starttemp, seed=seed, eng_int=0, checkpt_int=0,
        potfile=table["potfile"], nbl_margin=nbl_margin, nbl_size=1.3,
    )
    if extra:
        kw.update(extra)
    return write_param(os.path.join(outdir, name + ".param"), **kw)


def fcc_lattice(ncell, a0, offset=0.25):
    """Positions of an fcc crystal of ncell unit cells (like IMD's _fcc generator, src/imd_generate.c:306-484,
    but in plain lexicographic order).  Returns (n,3) float64 and the box matrix."""
    nx, ny, nz = ncell
    base = np.array([[0, 0, 0], [.5, .5, 0], [.5, 0, .5], [0, .5, .5]]) + offset
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cells = np.stack([ix, iy, iz], -1).reshape(-1, 1, 3).astype(np.float64)
    ort = ((cells + base[None]) * a0).reshape(-1, 3)
    box = np.diag([nx * a0, ny * a0, nz * a0]).astype(np.float64)
    return ort, box


def maxwell_momenta(n, mass, temperature, seed=1):
    """Gaussian momenta with sigma = sqrt(T m) and zero total momentum (what maxwell() produces,
    src/imd_maxwell.c:80-295, from numpy's generator instead of drand48)."""
    rng = np.random.default_rng(seed)
    p = rng.standard_normal((n, 3)) * np.sqrt(temperature * np.asarray(mass).reshape(-1, 1))
    p -= p.mean(axis=0)
    return p


def write_config(path, nummer, vsorte, masse, ort, impuls, box):
    """Atom configuration in IMD's ASCII format (read_atoms, src/imd_io_3d.c:44-): number, (virtual) type, mass,
    position, velocity; box vectors in the header (use with `box_from_header 1`)."""
    box = np.asarray(box, np.float64).reshape(3, 3)
    with open(path, "w") as f:
        f.write("#F A 1 1 1 3 3 0\n#C number type mass x y z vx vy vz\n")
        for tag, b in zip("XYZ", box):
            f.write("#%s %.17e %.17e %.17e\n" % (tag, b[0], b[1], b[2]))
        f.write("#E\n")
        vel = np.asarray(impuls) / np.asarray(masse)[:, None]
        for i in range(len(nummer)):
            f.write("%d %d %.17e %.17e %.17e %.17e %.17e %.17e %.17e\n" % (
                nummer[i], vsorte[i], masse[i], ort[i, 0], ort[i, 1], ort[i, 2], vel[i, 0], vel[i, 1], vel[i, 2]))
    return path
