// tables.cu -- potential tables: pot_table_t (src/types.h:416-428) -> derived device tables.
#include "internal.cuh"
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

// c0 c1 c2 g1 g2 of interval k, column col -- same dv/d2v expressions as PAIR_INT2
// (src/potaccess.h:345-349), evaluated here once in double instead of once per pair.
static inline void coef(const imdb200_pot_table *pt, int k, int col, double out[5])
{
  const int nc = pt->ncols;
  const double *t = pt->table + (size_t) k * nc + col;
  const double p0 = t[0], p1 = t[nc], p2 = t[2 * nc];
  const double dv = p1 - p0, d2v = p2 - 2 * p1 + p0, istep = pt->invstep[col];
  out[0] = p0;
  out[1] = dv - 0.5 * d2v;
  out[2] = 0.5 * d2v;
  out[3] = 2 * istep * out[1];
  out[4] = 4 * istep * out[2];
}

static int upload(imdb200_sim *s, int slot, const std::vector<double> &h, const double **dev)
{
  void *p = nullptr;
  CUDA_TRY(cudaMalloc(&p, h.size() * sizeof(double)));
  CUDA_TRY(cudaMemcpy(p, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
  s->tab_mem[slot] = p;
  *dev = (const double *) p;
  return 0;
}

void tables_free(imdb200_sim *s)
{
  for (int i = 0; i < 8; i++) { if (s->tab_mem[i]) cudaFree(s->tab_mem[i]); s->tab_mem[i] = nullptr; }
  s->have_tabs = 0;
}

int tables_upload(imdb200_sim *s, const imdb200_pot_table *pair, const imdb200_pot_table *embed,
                  const imdb200_pot_table *rho)
{
  if (!pair) return imdb_fail(IMDB200_ERR_ARG, "pair potential table is required");
  if ((embed == nullptr) != (rho == nullptr)) return imdb_fail(IMDB200_ERR_ARG, "EAM needs both embed and rho tables");
  const int nt = s->cfg.ntypes;
  if (pair->ncols != nt * nt) return imdb_fail(IMDB200_ERR_ARG, "pair table has %d columns, need %d", pair->ncols, nt * nt);
  if (rho && (rho->ncols != nt * nt || embed->ncols != nt)) return imdb_fail(IMDB200_ERR_ARG, "EAM table column count mismatch");
  tables_free(s);
  DevTables &T = s->tabs;
  memset(&T, 0, sizeof(T));
  T.ntypes = nt;
  T.have_eam = rho != nullptr;
  TRY(fill_meta(T.pair, pair));
  double c5[5];
  {
    std::vector<double> h((size_t) pair->maxsteps * pair->ncols * 6, 0.0);
    for (int k = 0; k < pair->maxsteps; k++)
      for (int c = 0; c < pair->ncols; c++) { coef(pair, k, c, c5); memcpy(&h[((size_t) k * pair->ncols + c) * 6], c5, 5 * sizeof(double)); }
    TRY(upload(s, 0, h, &T.pairVG));
  }
  // cellsz = max end of the radial tables (src/imd_potential.c:364, 406)
  double cz = 0.0;
  for (int c = 0; c < pair->ncols; c++) cz = cz > pair->end[c] ? cz : pair->end[c];
  if (rho) {
    for (int c = 0; c < rho->ncols; c++) cz = cz > rho->end[c] ? cz : rho->end[c];
    TRY(fill_meta(T.embed, embed));
    TRY(fill_meta(T.rho, rho));
    {
      std::vector<double> h((size_t) embed->maxsteps * embed->ncols * 6, 0.0);
      for (int k = 0; k < embed->maxsteps; k++)
        for (int c = 0; c < embed->ncols; c++) { coef(embed, k, c, c5); memcpy(&h[((size_t) k * embed->ncols + c) * 6], c5, 5 * sizeof(double)); }
      TRY(upload(s, 1, h, &T.embedVG));
    }
    {
      std::vector<double> hv((size_t) rho->maxsteps * rho->ncols * 4, 0.0), hg((size_t) rho->maxsteps * rho->ncols * 2, 0.0);
      for (int k = 0; k < rho->maxsteps; k++)
        for (int c = 0; c < rho->ncols; c++) {
          coef(rho, k, c, c5);
          memcpy(&hv[((size_t) k * rho->ncols + c) * 4], c5, 3 * sizeof(double));
          memcpy(&hg[((size_t) k * rho->ncols + c) * 2], c5 + 3, 2 * sizeof(double));
        }
      TRY(upload(s, 2, hv, &T.rhoV));
      TRY(upload(s, 3, hg, &T.rhoG));
    }
    // Shared grid: phi and rho use the same begin/invstep in every column, so one (k,chi)
    // serves both lookups of a pair in pass 1.  Inside `r2 <= end` / `r2 < end` the MIN(r2,end)
    // clamp of PAIR_INT2 is inactive, so differing ends do not matter.
    int fused = 1;
    for (int c = 0; c < pair->ncols; c++)
      if (pair->begin[c] != rho->begin[c] || pair->invstep[c] != rho->invstep[c]) fused = 0;
    if (fused) {
      const int nr = pair->maxsteps > rho->maxsteps ? pair->maxsteps : rho->maxsteps;
      std::vector<double> h((size_t) nr * pair->ncols * 8, 0.0);
      for (int k = 0; k < nr; k++)
        for (int c = 0; c < pair->ncols; c++) {
          double *o = &h[((size_t) k * pair->ncols + c) * 8];
          if (k < pair->maxsteps) { coef(pair, k, c, c5); memcpy(o, c5, 5 * sizeof(double)); }
          if (k < rho->maxsteps) { coef(rho, k, c, c5); memcpy(o + 5, c5, 3 * sizeof(double)); }
        }
      TRY(upload(s, 4, h, &T.fused1));
      T.fused = 1;
    }
  }
  s->cellsz0 = cz;
  s->have_tabs = 1;
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
  double c0, c1, c2, g1, g2;
  if (which == TAB_RHO) {
    const double *v = T.rhoV + ((size_t) k * m.ncols + col) * 4, *g = T.rhoG + ((size_t) k * m.ncols + col) * 2;
    c0 = v[0]; c1 = v[1]; c2 = v[2]; g1 = g[0]; g2 = g[1];
  } else {
    const double *v = (which == TAB_PAIR ? T.pairVG : T.embedVG) + ((size_t) k * m.ncols + col) * 6;
    c0 = v[0]; c1 = v[1]; c2 = v[2]; g1 = v[3]; g2 = v[4];
  }
  pot[i] = fma(chi, fma(chi, c2, c1), c0);
  grad[i] = fma(chi, g2, g1);
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
