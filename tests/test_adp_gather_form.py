"""CPU: the FULL-LIST GATHER formulation of the ADP terms that imd_b200/csrc/forces_adp.cu implements, written out
in numpy and held against the oracle (itself pinned to the reference's `adp` build by the fixtures).

The reference walks a half list and updates both atoms of a pair (mu_j -= u d, lambda_j += w d(x)d, KRAFT(j) -= f;
src/imd_forces_nbl.c:613-631, 1217-1305).  The CUDA engine stores both directions of every pair and lets every atom
gather its own sums: mu_i = sum_j u d_ij, lambda_i = sum_j w d_ij(x)d_ij, F_i = sum_j f(i,j), virial = -1/2 sum d.f,
per-atom stress -1/2 d(x)f.  This test checks those identities -- signs, table columns (col = it*nt+jt in the first pass,
col1 = jt*nt+it in the force pass), factors of 1/2 -- on both ADP fixtures.  It does not run CUDA code."""
import tempfile

import numpy as np
import pytest

from tests import common
from oracle import oracle as orc


def _gather_form(name):
    g = common.load_golden(name)
    tmp = tempfile.mkdtemp()
    paths = common.write_tables(g, tmp)
    nt = int(g["ntypes"])
    # oracle without and with ADP on the start state
    def mk(adp):
        kw = dict(pair=paths["pair"], embed=paths["embed"], rho=paths["rho"])
        if adp: kw.update(adp_u=paths["adp_u"], adp_w=paths["adp_w"])
        s = orc.OracleIMD(nt, g["box"], **kw)
        s.set_integrator(str(g["ensemble"]), float(g["timestep"]), float(g["temperature"]), float(g["eta0"]), float(g["isq_tau_eta"]))
        s.set_atoms(g["start:nummer"], g["start:sorte"], g["start:masse"], g["start:ort"], g["start:impuls"], vsorte=g["start:vsorte"])
        s.set_press_calc(True)
        s.calc_forces(0)
        return s
    s0, s1 = mk(False), mk(True)
    a0, a1 = s0.atoms(), s1.atoms()
    f_adp = a1["kraft"] - a0["kraft"]
    e_adp = a1["poteng"] - a0["poteng"]
    vir_adp = s1.scalars()["virial"] - s0.scalars()["virial"]
    st_adp = a1["presstens"] - a0["presstens"]
    # full list from the fixture's half list (nummer_i, nummer_j, shift of j)
    rows = g["nbl"].astype(np.int64)
    num = a1["nummer"]; idx = {int(v): k for k, v in enumerate(num)}
    x = a1["ort"]; typ = a1["sorte"]; box = g["box"]
    I = np.array([idx[int(r[0])] for r in rows]); J = np.array([idx[int(r[1])] for r in rows]); S = rows[:, 2:5].astype(float)
    # both directions
    ii = np.concatenate([I, J]); jj = np.concatenate([J, I]); ss = np.concatenate([S, -S])
    d = x[jj] + ss @ box - x[ii]
    r2 = (d * d).sum(1)
    # table lookups through the oracle (which: 4 = u, 5 = w)
    n = len(x)
    mu = np.zeros((n, 3)); la = np.zeros((n, 6))
    col = typ[ii] * nt + typ[jj]
    def look(which, cols, r2v):
        v = np.zeros(len(r2v)); gr = np.zeros(len(r2v))
        for c in np.unique(cols):
            m = cols == c
            vv, gg = s1.pair_int(which, int(c), r2v[m]); v[m] = vv; gr[m] = gg
        return v, gr
    u, du = look(4, col, r2); w, dw = look(5, col, r2)
    # ends: use the tables' end via value going to zero -- take r2 < 30.25
    inr = r2 < 30.25
    u, du, w, dw = u * inr, du * inr, w * inr, dw * inr
    np.add.at(mu, ii, u[:, None] * d)
    dd = np.stack([d[:, 0] * d[:, 0], d[:, 1] * d[:, 1], d[:, 2] * d[:, 2], d[:, 1] * d[:, 2], d[:, 2] * d[:, 0], d[:, 0] * d[:, 1]], 1)
    np.add.at(la, ii, w[:, None] * dd)
    err = dict(mu=np.abs(mu - a1["adp_mu"]).max() / np.abs(a1["adp_mu"]).max(),
               la=np.abs(la - a1["adp_lambda"]).max() / np.abs(a1["adp_lambda"]).max())
    tr = la[:, :3].sum(1) / 3
    e = 0.5 * (((la[:, :3] - tr[:, None]) ** 2).sum(1) + 2 * (la[:, 3:] ** 2).sum(1) + (mu ** 2).sum(1))
    err["energy"] = np.abs(e - e_adp).max() / np.abs(e_adp).max()
    # pass 2 in gather form, col1 = jt*nt+it
    col1 = typ[jj] * nt + typ[ii]
    u1, du1 = look(4, col1, r2); w1, dw1 = look(5, col1, r2)
    u1, du1, w1, dw1 = u1 * inr, du1 * inr, w1 * inr, dw1 * inr
    dm = mu[ii] - mu[jj]
    tmp = (dm * d).sum(1) * du1
    F = dm * u1[:, None] + tmp[:, None] * d
    L = la[ii] + la[jj]
    v = np.stack([L[:, 0] * d[:, 0] + L[:, 5] * d[:, 1] + L[:, 4] * d[:, 2],
                  L[:, 5] * d[:, 0] + L[:, 1] * d[:, 1] + L[:, 3] * d[:, 2],
                  L[:, 4] * d[:, 0] + L[:, 3] * d[:, 1] + L[:, 2] * d[:, 2]], 1)
    nu = L[:, :3].sum(1) / 3
    f1 = 2 * w1
    f2 = ((v * d).sum(1) - nu * r2) * dw1 - nu * f1
    F += f1[:, None] * v + f2[:, None] * d
    Fi = np.zeros((n, 3)); np.add.at(Fi, ii, F)
    err["force"] = np.abs(Fi - f_adp).max() / np.abs(f_adp).max()
    vir = -0.5 * (d * F).sum()
    err["virial"] = abs(vir - vir_adp) / abs(s1.scalars()["virial"])
    st = np.zeros((n, 6))
    dF = np.stack([d[:, 0] * F[:, 0], d[:, 1] * F[:, 1], d[:, 2] * F[:, 2], d[:, 1] * F[:, 2], d[:, 2] * F[:, 0], d[:, 0] * F[:, 1]], 1)
    np.add.at(st, ii, -0.5 * dF)
    err["stress"] = np.abs(st - st_adp).max() / np.abs(st_adp).max()
    return err



@pytest.mark.parametrize("name", ["cu_adp", "nial_adp"])
def test_adp_gather_form_equals_half_list_form(name):
    err = _gather_form(name)
    print(name, {k: f"{v:.1e}" for k, v in err.items()})
    assert err["mu"] < 1e-13 and err["la"] < 1e-13
    assert err["energy"] < 1e-11 and err["force"] < 1e-12 and err["stress"] < 1e-12
    assert err["virial"] < 1e-12
