// tables.cu -- potential tables: pot_table_t (src/types.h:416-428) -> derived device tables.
#include "internal.cuh"

// cellsz = max end of the radial tables (src/imd_potential.c:364, 406).  The cell size, the cell grid, the halo plan and
// the list cut-off (r2 < cellsz) all derive from it: when it changes they are derived again at the next
// geom_make_box (init_cells adds the margin once to a cellsz of 0, src/imd_geom_3d.c:122-126).
void tables_set_cellsz0(imdb200_sim *s, double cz)
{
  if (cz == s->cellsz0 && s->geom.cellsz != 0.0) return;
  s->cellsz0 = cz;
  s->geom.cellsz = 0.0;
  for (int d = 0; d < 3; d++) { s->min_height[d] = 0.0; s->max_height[d] = 0.0; }   // forces init_cells
  s->have_valid_nbl = 0;
}
#include <stdlib.h>
#include <string.h>
#include <vector>

static int fill_meta(TabMeta &m, const imdb200_pot_table *pt)
{
  if (pt->ncols > IMDB_MAXCOL) return imdb_fail(IMDB200_ERR_ARG, "potential table with %d columns (max %d)", pt->ncols, IMDB_MAXCOL);
  m.ncols = pt->ncols;
  m.nrows = pt->maxsteps;
  for (int c = 0; c < pt->ncols; c++) { m.begin[c] = pt->begin[c]; m.end[c] = pt->end[c]; m.invstep[c] = pt->invstep[c]; }
  return 0;
}

// ---- host-side preparation -------------------------------------------------------------------------------
// A private copy of one pot_table_t with the pad rows (and the spline's second derivatives) recomputed for
// the interpolation in use: init_threepoint / init_fourpoint / init_spline, src/imd_potential.c:1171-1272.
struct HostTab {
  const imdb200_pot_table *pt;
  std::vector<double> y, y2;     // [(maxsteps+2)][ncols]
  int mode;
};

static void prepare(HostTab &h, const imdb200_pot_table *pt, int mode, int radial)
{
  const int nc = pt->ncols;
  const size_t rows = (size_t) pt->maxsteps + 2;
  h.pt = pt; h.mode = mode;
  h.y.assign(pt->table, pt->table + rows * nc);
  h.y2.assign(mode == IMDB200_INTERP_SPLINE ? rows * nc : 0, 0.0);
  for (int col = 0; col < nc; col++) {
    double *y = h.y.data() + col;
    const long n = pt->len[col];
    if (mode == IMDB200_INTERP_4POINT) {
      if (n < 4) continue;
      y[n * nc]       =  4 * y[(n - 1) * nc] -  6 * y[(n - 2) * nc] +  4 * y[(n - 3) * nc] -     y[(n - 4) * nc];
      y[(n + 1) * nc] = 10 * y[(n - 1) * nc] - 20 * y[(n - 2) * nc] + 15 * y[(n - 3) * nc] - 4 * y[(n - 4) * nc];
    } else if (mode == IMDB200_INTERP_SPLINE) {
      if (n < 3) continue;
      double *y2 = h.y2.data() + col;
      const double step = pt->step[col];
      std::vector<double> u(rows, 0.0);
      y2[0] = u[0] = 0;                                   // natural spline at the left end
      for (long i = 1; i < n - 1; i++) {
        const double p = 0.5 * y2[(i - 1) * nc] + 2.0;
        y2[i * nc] = -0.5 / p;
        u[i] = (y[(i + 1) * nc] - 2 * y[i * nc] + y[(i - 1) * nc]) / step;
        u[i] = (6.0 * u[i] / (2 * step) - 0.5 * u[i - 1]) / p;
      }
      double qn = 0.0, un = 0.0;                          // radial functions: zero slope at the right end
      if (radial) { qn = 0.5; un = (3.0 / step) * (y[(n - 2) * nc] - y[(n - 1) * nc]) / step; }
      y2[(n - 1) * nc] = (un - qn * u[n - 2]) / (qn * y2[(n - 2) * nc] + 1.0);
      for (long k = n - 2; k >= 0; k--) y2[k * nc] = y2[k * nc] * y2[(k + 1) * nc] + u[k];
      y[n * nc] = 2 * y[(n - 1) * nc] - y[(n - 2) * nc] + step * step * y2[(n - 1) * nc];
      y2[n * nc] = 2 * y2[(n - 1) * nc] - y2[(n - 2) * nc];
    } else {
      if (n < 3) continue;
      y[n * nc]       = 3 * y[(n - 1) * nc] - 3 * y[(n - 2) * nc] + y[(n - 3) * nc];
      y[(n + 1) * nc] = 6 * y[(n - 1) * nc] - 8 * y[(n - 2) * nc] + 3 * y[(n - 3) * nc];
    }
  }
}

// Polynomial coefficients of interval k = (int)((r2-begin)*invstep), chi = (r2-begin)*invstep - k in [0,1):
//   val = c0 + chi*(c1 + chi*(c2 + chi*c3)),  grad = 2*istep*(c1 + chi*(2*c2 + 3*c3*chi))
// 3-point: the dv/d2v expressions of PAIR_INT2 (src/potaccess.h:345-349), c3 = 0.
// 4-point: the Lagrange cubic of PAIR_INT3 (:385-405) through samples k-1..k+2 collected by powers of chi; the
//          reference uses k = MAX(k,1), i.e. interval 0 evaluates the cubic of interval 1 at chi-1 -- that
//          polynomial is re-expanded around chi here so that the device indexes every mode the same way.
// spline : PAIR_INT_SP (:438-456) with a = 1-b collected by powers of b.
static void coef(const HostTab &h, int k, int col, double c[4])
{
  const imdb200_pot_table *pt = h.pt;
  const int nc = pt->ncols;
  if (h.mode == IMDB200_INTERP_4POINT) {
    const int kk = k < 1 ? 1 : k;
    const double *t = h.y.data() + (size_t) (kk - 1) * nc + col;
    const double p0 = t[0], p1 = t[nc], p2 = t[2 * nc], p3 = t[3 * nc];
    c[0] = p1;
    c[1] = -p0 / 3.0 - 0.5 * p1 + p2 - p3 / 6.0;
    c[2] = 0.5 * p0 - p1 + 0.5 * p2;
    c[3] = (p3 - p0) / 6.0 + 0.5 * (p1 - p2);
    if (k < 1) {                                            // P(chi - 1)
      const double a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3];
      c[0] = a0 - a1 + a2 - a3;
      c[1] = a1 - 2 * a2 + 3 * a3;
      c[2] = a2 - 3 * a3;
      c[3] = a3;
    }
  } else if (h.mode == IMDB200_INTERP_SPLINE) {
    const double *t = h.y.data() + (size_t) k * nc + col, *t2 = h.y2.data() + (size_t) k * nc + col;
    const double p1 = t[0], p2 = t[nc], d21 = t2[0], d22 = t2[nc];
    const double s6 = pt->step[col] * pt->step[col] / 6.0;
    c[0] = p1;
    c[1] = (p2 - p1) - s6 * (2 * d21 + d22);
    c[2] = 3 * s6 * d21;
    c[3] = s6 * (d22 - d21);
  } else {
    const double *t = h.y.data() + (size_t) k * nc + col;
    const double p0 = t[0], p1 = t[nc], p2 = t[2 * nc];
    const double dv = p1 - p0, d2v = p2 - 2 * p1 + p0;
    c[0] = p0;
    c[1] = dv - 0.5 * d2v;
    c[2] = 0.5 * d2v;
    c[3] = 0.0;
  }
}

template <typename T> static int upload(imdb200_sim *s, int slot, const std::vector<T> &h, const T **dev)
{
  void *p = nullptr;
  CUDA_TRY(cudaMalloc(&p, h.size() * sizeof(T)));
  CUDA_TRY(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  s->tab_mem[slot] = p;
  *dev = (const T *) p;
  return 0;
}

void tables_free(imdb200_sim *s)
{
  for (int i = 0; i < 12; i++) { if (s->tab_mem[i]) cudaFree(s->tab_mem[i]); s->tab_mem[i] = nullptr; }
  s->have_tabs = 0;
}

// (c0,c1), c2 and (cubic modes) (c2,c3) arrays of one table
static void split_coefs(const HostTab &h, std::vector<double2> &ab, std::vector<double> &c, std::vector<double2> &cd)
{
  const imdb200_pot_table *pt = h.pt;
  double c4[4];
  ab.assign((size_t) pt->maxsteps * pt->ncols, make_double2(0.0, 0.0));
  cd.assign((size_t) pt->maxsteps * pt->ncols, make_double2(0.0, 0.0));
  c.assign((((size_t) pt->maxsteps * pt->ncols + 1) / 2) * 2, 0.0);   // even length: staged 16 bytes at a time
  for (int k = 0; k < pt->maxsteps; k++)
    for (int col = 0; col < pt->ncols; col++) {
      coef(h, k, col, c4);
      ab[(size_t) k * pt->ncols + col] = make_double2(c4[0], c4[1]);
      cd[(size_t) k * pt->ncols + col] = make_double2(c4[2], c4[3]);
      c[(size_t) k * pt->ncols + col] = c4[2];
    }
}

int tables_upload(imdb200_sim *s, const imdb200_pot_table *pair, const imdb200_pot_table *embed,
                  const imdb200_pot_table *rho)
{
  if (!pair) return imdb_fail(IMDB200_ERR_ARG, "pair potential table is required");
  if ((embed == nullptr) != (rho == nullptr)) return imdb_fail(IMDB200_ERR_ARG, "EAM needs both embed and rho tables");
  const int nt = s->cfg.ntypes, mode = s->cfg.interpolation;
  if (mode < IMDB200_INTERP_3POINT || mode > IMDB200_INTERP_SPLINE) return imdb_fail(IMDB200_ERR_ARG, "unknown interpolation %d", mode);
  const bool cubic = mode != IMDB200_INTERP_3POINT;
  if (pair->ncols != nt * nt) return imdb_fail(IMDB200_ERR_ARG, "pair table has %d columns, need %d", pair->ncols, nt * nt);
  if (rho && (rho->ncols != nt * nt || embed->ncols != nt)) return imdb_fail(IMDB200_ERR_ARG, "EAM table column count mismatch");
  tables_free(s);
  DevTables &T = s->tabs;
  memset(&T, 0, sizeof(T));
  T.ntypes = nt;
  T.cubic = cubic;
  T.have_eam = rho != nullptr;
  TRY(fill_meta(T.pair, pair));
  HostTab hp, he, hr;
  prepare(hp, pair, mode, 1);
  std::vector<double2> ab, cd; std::vector<double> c;
  split_coefs(hp, ab, c, cd);
  TRY(upload(s, 0, ab, &T.pairAB));
  if (cubic) TRY(upload(s, 1, cd, &T.pairCD)); else TRY(upload(s, 1, c, &T.pairC));
  const size_t per1 = cubic ? 32 : 24;
  size_t bytes1 = (size_t) pair->maxsteps * pair->ncols * per1 + 8, bytes2 = 0;
  // cellsz = max end of the radial tables (src/imd_potential.c:364, 406)
  double cz = 0.0;
  std::vector<double2> hh;                                       // (h1,h2) of rho, [nrows][ncols]
  for (int col = 0; col < pair->ncols; col++) cz = cz > pair->end[col] ? cz : pair->end[col];
  if (rho) {
    for (int col = 0; col < rho->ncols; col++) cz = cz > rho->end[col] ? cz : rho->end[col];
    TRY(fill_meta(T.embed, embed));
    TRY(fill_meta(T.rho, rho));
    prepare(he, embed, mode, 0);
    prepare(hr, rho, mode, 1);
    {
      // per (k, type): c0 c1 c2 c3 | g1 g2 g3 -  with F' = g1 + chi*(g2 + chi*g3)
      double c4[4];
      std::vector<double> h((size_t) embed->maxsteps * embed->ncols * 8, 0.0);
      for (int k = 0; k < embed->maxsteps; k++)
        for (int col = 0; col < embed->ncols; col++) {
          coef(he, k, col, c4);
          double *o = &h[((size_t) k * embed->ncols + col) * 8];
          const double is2 = 2 * embed->invstep[col];
          o[0] = c4[0]; o[1] = c4[1]; o[2] = c4[2]; o[3] = c4[3];
          o[4] = is2 * c4[1]; o[5] = 2 * is2 * c4[2]; o[6] = 3 * is2 * c4[3];
        }
      TRY(upload(s, 2, h, &T.embedVG));
    }
    split_coefs(hr, ab, c, cd);
    TRY(upload(s, 3, ab, &T.rhoAB));
    if (cubic) TRY(upload(s, 4, cd, &T.rhoCD)); else TRY(upload(s, 4, c, &T.rhoC));
    hh.assign(ab.size(), make_double2(0.0, 0.0));
    std::vector<double> h3((ab.size() + 1) / 2 * 2, 0.0);
    for (int k = 0; k < rho->maxsteps; k++)
      for (int col = 0; col < rho->ncols; col++) {
        const size_t e = (size_t) k * rho->ncols + col;
        // rho'/2 = istep*(c1 + 2*chi*c2 + 3*chi^2*c3): exact scalings of the products the reference forms
        hh[e] = make_double2(rho->invstep[col] * ab[e].y, 2 * rho->invstep[col] * cd[e].x);
        h3[e] = 3 * rho->invstep[col] * cd[e].y;
      }
    TRY(upload(s, 5, hh, &T.rhoH));
    if (cubic) TRY(upload(s, 7, h3, &T.rhoH3));
    bytes1 += (size_t) rho->maxsteps * rho->ncols * per1 + 8;
    bytes2 = (size_t) rho->maxsteps * rho->ncols * (cubic ? 24 : 16) + 8;
    // Shared grid: phi and rho use the same begin/invstep in every column, so one (k,chi)
    // serves both lookups of a pair in pass 1.  Inside `r2 <= end` / `r2 < end` the MIN(r2,end)
    // clamp of PAIR_INT2 is inactive, so differing ends do not matter.
    T.shared_grid = 1;
    for (int col = 0; col < pair->ncols; col++)
      if (pair->begin[col] != rho->begin[col] || pair->invstep[col] != rho->invstep[col]) T.shared_grid = 0;
  }
  if (rho && nt == 1 && T.shared_grid && !cubic) {
    // (the cubic modes need 4+4 coefficients = four 16-byte loads with or without fusing: they use the split arrays)
    const int nr = pair->maxsteps > rho->maxsteps ? pair->maxsteps : rho->maxsteps;
    std::vector<double2> f((size_t) nr * 3, make_double2(0.0, 0.0));
    double c4[4];
    for (int k = 0; k < nr; k++) {
      double pc[3] = {0, 0, 0}, rc[3] = {0, 0, 0};
      if (k < pair->maxsteps) { coef(hp, k, 0, c4); pc[0] = c4[0]; pc[1] = c4[1]; pc[2] = c4[2]; }
      if (k < rho->maxsteps) { coef(hr, k, 0, c4); rc[0] = c4[0]; rc[1] = c4[1]; rc[2] = c4[2]; }
      f[3 * (size_t) k] = make_double2(pc[0], pc[1]);
      f[3 * (size_t) k + 1] = make_double2(pc[2], rc[2]);
      f[3 * (size_t) k + 2] = make_double2(rc[0], rc[1]);
    }
    TRY(upload(s, 6, f, &T.fused));
    T.fused_rows = nr;
    bytes1 = (size_t) nr * 48;
    // raw samples (phi_k, rho_k), pad rows included: a third of the shared memory, eight more FP64 operations per pair
    std::vector<double2> fr((size_t) nr + 2, make_double2(0.0, 0.0));
    for (int k = 0; k < nr + 2; k++) {
      if (k < pair->maxsteps + 2) fr[k].x = hp.y[(size_t) k];
      if (k < rho->maxsteps + 2) fr[k].y = hr.y[(size_t) k];
    }
    TRY(upload(s, 11, fr, &T.fraw));
    T.smem1_raw = (int) (fr.size() * 16);
  }
  if (nt > 1 && !cubic) {
    // raw samples of the distinct columns (see DevTables::rawP): two columns are the same function when their headers
    // and all their samples (pad rows included) agree bit for bit
    int rep[IMDB_MAXCOL];
    auto distinct = [&rep](const HostTab &h, signed char *umap, std::vector<double> &out) {
      const imdb200_pot_table *pt = h.pt;
      const int nc = pt->ncols; const size_t rows = (size_t) pt->maxsteps + 2;
      int nu = 0;
      for (int c = 0; c < nc; c++) {
        int u = -1;
        for (int q = 0; q < nu && u < 0; q++) {
          const int d = rep[q];
          bool same = pt->begin[c] == pt->begin[d] && pt->end[c] == pt->end[d] && pt->invstep[c] == pt->invstep[d];
          for (size_t r = 0; r < rows && same; r++) same = h.y[r * nc + c] == h.y[r * nc + d];
          if (same) u = q;
        }
        if (u < 0) { u = nu; rep[nu++] = c; }
        umap[c] = (signed char) u;
      }
      out.assign(rows * nu, 0.0);
      for (size_t r = 0; r < rows; r++) for (int q = 0; q < nu; q++) out[r * nu + q] = h.y[r * nc + rep[q]];
      return nu;
    };
    std::vector<double> rp, rr;
    size_t b2m = 0;
    T.nuP = distinct(hp, T.umapP, rp);
    if (rp.size() % 2) rp.push_back(0.0);                 // staged 16 bytes at a time
    TRY(upload(s, 11, rp, &T.rawP));
    size_t rawbytes = rp.size() * 8;
    if (rho) {
      T.nuR = distinct(hr, T.umapR, rr);
      if (rr.size() % 2) rr.push_back(0.0);
      // both raw blocks in one allocation slot would need a 13th slot: the rho block rides behind the fused slot (unused here)
      TRY(upload(s, 6, rr, &T.rawR));
      rawbytes += rr.size() * 8;
      // pass 2: (h1,h2) of the distinct rho columns (rep[] still holds their representatives); slot 7 is the cubic modes' h3
      std::vector<double2> hd((size_t) rho->maxsteps * T.nuR);
      for (int k = 0; k < rho->maxsteps; k++)
        for (int q = 0; q < T.nuR; q++) hd[(size_t) k * T.nuR + q] = hh[(size_t) k * rho->ncols + rep[q]];
      TRY(upload(s, 7, hd, &T.rhoHd));
      b2m = hd.size() * 16;
    }
    // One header for all columns of a table (the usual case: alloy EAM files carry one r grid): begin, end and invstep
    // are scalars in the kernels, and inside the cut-off the MIN(r2,end) clamp of PAIR_INT2 / DERIV_FUNC is inactive.
    // Tables with per-column headers take the general path (coefficient tables in HBM / L1).
    T.multi_uniform = 1;
    for (int col = 0; col < pair->ncols; col++) {
      if (pair->begin[col] != pair->begin[0] || pair->end[col] != pair->end[0] || pair->invstep[col] != pair->invstep[0]) T.multi_uniform = 0;
      if (rho && (rho->begin[col] != rho->begin[0] || rho->end[col] != rho->end[0] || rho->invstep[col] != rho->invstep[0])) T.multi_uniform = 0;
    }
    T.smem2m = (T.multi_uniform && b2m && b2m <= 160 * 1024) ? (int) b2m : 0;
    if (T.multi_uniform && rawbytes <= 160 * 1024) { T.raw_ok = 1; bytes1 = rawbytes + 16; }
  }
  // Tables are staged in shared memory when they leave at least ~100 KB of the 228 KB SM array to L1
  // (the position gathers live there); otherwise they stay in HBM and are served by L1/L2.
  const size_t limit = 128 * 1024;
  // several species, quadratic: shared memory holds the raw layout or nothing (the kernels' RAW path assumes multi_uniform)
  if (nt > 1 && !cubic) T.smem1 = T.raw_ok ? (int) bytes1 : 0;
  else T.smem1 = bytes1 <= limit ? (int) bytes1 : 0;
  T.smem2 = (bytes2 && bytes2 <= limit) ? (int) bytes2 : 0;
  tables_set_cellsz0(s, cz);
  s->have_tabs = 1;
  return 0;
}

// EEAM: the energy modification term M(p) (emod_pot, src/imd_potential.c:82-85), one column per type, not radial.
// Same derived layout as the embedding table: c0 c1 c2 c3 | g1 g2 g3 -.
int tables_upload_emod(imdb200_sim *s, const imdb200_pot_table *emod)
{
  if (!s->have_tabs || !s->tabs.have_eam) return imdb_fail(IMDB200_ERR_ARG, "set the EAM tables before the EEAM table");
  DevTables &T = s->tabs;
  if (s->tab_mem[8]) { cudaFree(s->tab_mem[8]); s->tab_mem[8] = nullptr; }
  T.have_eeam = 0; T.emodVG = nullptr;
  if (!emod) return 0;
  if (emod->ncols != T.ntypes) return imdb_fail(IMDB200_ERR_ARG, "EEAM table has %d columns, need %d", emod->ncols, T.ntypes);
  TRY(fill_meta(T.emod, emod));
  HostTab hm;
  prepare(hm, emod, s->cfg.interpolation, 0);
  double c4[4];
  std::vector<double> h((size_t) emod->maxsteps * emod->ncols * 8, 0.0);
  for (int k = 0; k < emod->maxsteps; k++)
    for (int col = 0; col < emod->ncols; col++) {
      coef(hm, k, col, c4);
      double *o = &h[((size_t) k * emod->ncols + col) * 8];
      const double is2 = 2 * emod->invstep[col];
      o[0] = c4[0]; o[1] = c4[1]; o[2] = c4[2]; o[3] = c4[3];
      o[4] = is2 * c4[1]; o[5] = 2 * is2 * c4[2]; o[6] = 3 * is2 * c4[3];
    }
  TRY(upload(s, 8, h, &T.emodVG));
  T.have_eeam = 1;
  return 0;
}

// ADP: dipole u(r) and quadrupole w(r) distortion functions (adp_upot, adp_wpot: ntypes^2 columns in r^2, radial,
// src/imd_potential.c:87-92), as (c0,c1,c2,c3) per interval and column.
static int upload_adp_one(imdb200_sim *s, int slot, const imdb200_pot_table *pt, TabMeta &meta, const double4 **dev)
{
  TRY(fill_meta(meta, pt));
  HostTab h;
  prepare(h, pt, s->cfg.interpolation, 1);
  std::vector<double4> k((size_t) pt->maxsteps * pt->ncols);
  double c4[4];
  for (int r = 0; r < pt->maxsteps; r++)
    for (int col = 0; col < pt->ncols; col++) { coef(h, r, col, c4); k[(size_t) r * pt->ncols + col] = make_double4(c4[0], c4[1], c4[2], c4[3]); }
  return upload(s, slot, k, dev);
}

int tables_upload_adp(imdb200_sim *s, const imdb200_pot_table *u, const imdb200_pot_table *w)
{
  if (!s->have_tabs || !s->tabs.have_eam) return imdb_fail(IMDB200_ERR_ARG, "set the EAM tables before the ADP tables");
  DevTables &T = s->tabs;
  for (int slot = 9; slot <= 10; slot++) if (s->tab_mem[slot]) { cudaFree(s->tab_mem[slot]); s->tab_mem[slot] = nullptr; }
  T.have_adp = 0; T.adpuK = T.adpwK = nullptr;
  if (!u && !w) return 0;
  if (!u || !w) return imdb_fail(IMDB200_ERR_ARG, "ADP needs both the u and the w table");
  const int nt = T.ntypes;
  if (u->ncols != nt * nt || w->ncols != nt * nt) return imdb_fail(IMDB200_ERR_ARG, "ADP tables need %d columns", nt * nt);
  TRY(upload_adp_one(s, 9, u, T.adpu, &T.adpuK));
  TRY(upload_adp_one(s, 10, w, T.adpw, &T.adpwK));
  // radial tables take part in cellsz = max end (src/imd_potential.c:406)
  double cz = s->cellsz0;
  for (int col = 0; col < nt * nt; col++) {
    if (u->end[col] > cz) cz = u->end[col];
    if (w->end[col] > cz) cz = w->end[col];
  }
  tables_set_cellsz0(s, cz);
  T.have_adp = 1;
  return 0;
}

// ---- test hook: PAIR_INT through the device lookup code -------------------------------------------
__global__ void k_pair_int(DevTables T, int which, int col, long n, const double *r2, double *pot, double *grad)
{
  long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (i >= n) return;
  const TabMeta &m = which == TAB_PAIR ? T.pair : (which == TAB_EMBED ? T.embed : T.rho);
  int k, sh = 0; double chi;
  tab_index(r2[i], m.begin[col], m.end[col], m.invstep[col], k, chi, sh);
  const size_t e = (size_t) k * m.ncols + col;
  if (which == TAB_EMBED) {
    const double *v = T.embedVG + e * 8;
    pot[i] = fma(chi, fma(chi, fma(chi, v[3], v[2]), v[1]), v[0]);
    grad[i] = fma(chi, fma(chi, v[6], v[5]), v[4]);
  } else if (T.cubic) {
    const double2 ab = which == TAB_PAIR ? T.pairAB[e] : T.rhoAB[e];
    const double2 cd = which == TAB_PAIR ? T.pairCD[e] : T.rhoCD[e];
    pot[i] = tab_val3(ab, cd, chi);
    grad[i] = which == TAB_PAIR ? tab_grad3(ab, cd, chi, 2.0 * m.invstep[col])
                                : 2.0 * fma(chi, fma(chi, T.rhoH3[e], T.rhoH[e].y), T.rhoH[e].x);
  } else {
    const double2 ab = which == TAB_PAIR ? T.pairAB[e] : T.rhoAB[e];
    const double c2 = which == TAB_PAIR ? T.pairC[e] : T.rhoC[e];
    pot[i] = tab_val(ab, c2, chi);
    // phi' as pass 1 forms it, rho' as pass 2 forms it
    grad[i] = which == TAB_PAIR ? tab_grad(ab, c2, chi, 2.0 * m.invstep[col])
                                : 2.0 * fma(chi, T.rhoH[e].y, T.rhoH[e].x);
  }
}

int tables_pair_int(imdb200_sim *s, int which, int col, long n, const double *r2, double *pot, double *grad)
{
  if (!s->have_tabs) return imdb_fail(IMDB200_ERR_ARG, "no potential tables loaded");
  if (which != TAB_PAIR && !s->tabs.have_eam) return imdb_fail(IMDB200_ERR_ARG, "no EAM tables loaded");
  double *d = nullptr;
  CUDA_TRY(cudaMalloc(&d, 3 * n * sizeof(double)));
  CUDA_TRY(cudaMemcpy(d, r2, n * sizeof(double), cudaMemcpyHostToDevice));
  k_pair_int<<<cdiv(n, 256), 256, 0, s->stream>>>(s->tabs, which, col, n, d, d + n, d + 2 * n);
  LAUNCH_CHECK();
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  CUDA_TRY(cudaMemcpy(pot, d + n, n * sizeof(double), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(grad, d + 2 * n, n * sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(d);
  return 0;
}
