/* oracle/imd_oracle.c -- TEST INFRASTRUCTURE ONLY (see imd_oracle.h).
 *
 * Order-faithful CPU restatement of IMD's NBL/EAM2 force-and-integrate path for
 * cpu_dim = 1 1 1 (serial build, buffer cells filled by in-process copies).  Each
 * function cites the reference file:line it follows.  Written from the reference's
 * behaviour, with our own data structures (flat atom arrays + per-cell index lists
 * instead of per-cell SoA blocks); arithmetic expressions keep the reference's
 * operation order so that results agree to the last bit wherever the summation
 * order is also the same.
 *
 * Compile with -ffp-contract=off (oracle/Makefile) so that no FMA is formed.
 */
#include "imd_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define SQR(a) ((a) * (a))
#define MAXV(a, b) ((a) > (b) ? (a) : (b))
#define MINV(a, b) ((a) < (b) ? (a) : (b))
#define PSTEP 50 /* src/config.h:251 */

typedef struct { double x, y, z; } vec3;

typedef struct { /* pot_table_t, src/types.h:416-428 */
  double *begin, *end, *step, *invstep;
  int *len, ncols, maxsteps;
  double *table;
  double *table2;      /* second derivatives, SPLINE only (pot_table_t.table2, src/types.h:416-428) */
  int loaded;
} ptab;

typedef struct { int n, cap; int *idx; } cellist;

struct orc_sim {
  int ntypes;
  vec3 box_x, box_y, box_z, tbox_x, tbox_y, tbox_z, height, min_height, max_height;
  double volume;
  int pbc[3];
  double nbl_margin, cellsz;
  int margin_added;
  ptab tab[6];         /* pair_pot, embed_pot, rho_h_tab, emod_pot (EEAM), adp_upot, adp_wpot (ADP) */
  int default_fmt;
  int interp;          /* ORC_INTERP_*: which PAIR_INT the build selects, src/potaccess.h:24-36 */
  /* atoms: [0,n) real, [n, n+ng) buffer-cell copies */
  long n, ng, cap;
  int *nummer, *sorte, *vsorte;
  double *masse, *ort, *impuls, *kraft, *poteng, *rho, *dF, *presstens, *nblpos;
  double *eam_p, *dM;  /* EEAM: p_i = sum rho_j^2 and M'(p_i) (EAM_P, EAM_DM, src/makros.h:96-99) */
  double *mu, *la;     /* ADP: dipole mu[3] and quadrupole lambda[6] = xx yy zz yz zx xy per atom (ADP_MU, ADP_LAMBDA) */
  long gstage[3]; /* end index of the z-, y-, x-stage buffer atoms */
  long *gsrc; /* source atom of each buffer atom (may itself be a buffer atom) */
  signed char *gshift; /* accumulated image shift (box units), for reporting only */
  /* cells */
  int gdim[3], cdim[3], nallcells, ncells;
  cellist *cells;
  int *cnp, *cnq; /* cell_nbrs_t: np, nq[14]  (src/types.h:379-382) */
  /* neighbour list (src/imd_forces_nbl.c:47) */
  long *tl; int *tb; long tb_cap; int *cl_off; int *cl_num;
  int have_valid_nbl, nbl_count;
  /* integrator */
  int ensemble; double timestep, temperature, eta, isq_tau_eta;
  /* NPT_iso (src/imd_integrate.c:1472-1729): barostat friction xi, twice the kinetic energy of the previous step,
     external pressure and its per-step increment, 1/tau_xi^2, the pressure the last step used */
  double xi, Ekin_old, pressure_ext, d_pressure, isq_tau_xi, pressure;
  /* NPT_axial (src/imd_integrate.c:1747-1959): one barostat per axis; Ekin_old and isq_tau_xi are shared with NPT_iso */
  double xi3[3], stress3[3], pext3[3], dpext3[3], dyn3[3]; int relax_dirs[3];
  double tauber;       /* > 0: Berendsen variant of NVE (`ber` builds, src/imd_integrate.c:44-53, 341-350) */
  int nvtypes; double *restr;
  /* results */
  double tot_pot_energy, tot_kin_energy, virial;
  double vir[3];       /* vir_xx, vir_yy, vir_zz: what P_AXIAL builds accumulate instead of the scalar virial (src/imd_forces_nbl.c:548-556) */
  long nactive;
  int is_short;
};

/* ------------------------------------------------------------------------------------ */
/* potential tables                                                                      */
/* ------------------------------------------------------------------------------------ */

/* init_threepoint, src/imd_potential.c:1256-1272 */
static void init_threepoint(ptab *pt)
{
  int col, nc = pt->ncols;
  for (col = 0; col < nc; col++) {
    double *y = pt->table + col;
    int n = pt->len[col];
    y[n * nc]       = 3 * y[(n - 1) * nc] - 3 * y[(n - 2) * nc] + y[(n - 3) * nc];
    y[(n + 1) * nc] = 6 * y[(n - 1) * nc] - 8 * y[(n - 2) * nc] + 3 * y[(n - 3) * nc];
  }
}

/* init_fourpoint, src/imd_potential.c:1171-1189 */
static void init_fourpoint(ptab *pt)
{
  int col, nc = pt->ncols;
  for (col = 0; col < nc; col++) {
    double *y = pt->table + col;
    int n = pt->len[col];
    y[n * nc]       =  4 * y[(n - 1) * nc] -  6 * y[(n - 2) * nc] +  4 * y[(n - 3) * nc] -     y[(n - 4) * nc];
    y[(n + 1) * nc] = 10 * y[(n - 1) * nc] - 20 * y[(n - 2) * nc] + 15 * y[(n - 3) * nc] - 4 * y[(n - 4) * nc];
  }
}

/* init_spline, src/imd_potential.c:1197-1246 */
static void init_spline(ptab *pt, int radial)
{
  int ncols = pt->ncols, size = pt->maxsteps + 2, col, n, i, k;
  double p, qn, un, step, *u, *y, *y2;
  pt->table2 = (double *) calloc((size_t) ncols * size, sizeof(double));
  u = (double *) calloc((size_t) size, sizeof(double));
  for (col = 0; col < ncols; col++) {
    y2 = pt->table2 + col;
    y = pt->table + col;
    n = pt->len[col];
    step = pt->step[col];
    y2[0] = u[0] = 0;
    for (i = 1; i < n - 1; i++) {
      p = 0.5 * y2[(i - 1) * ncols] + 2.0;
      y2[i * ncols] = -0.5 / p;
      u[i] = (y[(i + 1) * ncols] - 2 * y[i * ncols] + y[(i - 1) * ncols]) / step;
      u[i] = (6.0 * u[i] / (2 * step) - 0.5 * u[i - 1]) / p;
    }
    if (radial) {
      qn = 0.5;
      un = (3.0 / step) * (y[(n - 2) * ncols] - y[(n - 1) * ncols]) / step;
    } else {
      qn = un = 0.0;
    }
    y2[(n - 1) * ncols] = (un - qn * u[n - 2]) / (qn * y2[(n - 2) * ncols] + 1.0);
    for (k = n - 2; k >= 0; k--) y2[k * ncols] = y2[k * ncols] * y2[(k + 1) * ncols] + u[k];
    y[n * ncols] = 2 * y[(n - 1) * ncols] - y[(n - 2) * ncols] + step * step * y2[(n - 1) * ncols];
    y2[n * ncols] = 2 * y2[(n - 1) * ncols] - y2[(n - 2) * ncols];
  }
  free(u);
}

/* read_pot_table1, src/imd_potential.c:297-376 */
static int read_table1(orc_sim *s, ptab *pt, FILE *f, int radial)
{
  int ncols = pt->ncols, npot = 0, i, k;
  double val, r2 = 0, r2_start = 0, r2_step, delta;
  pt->maxsteps = PSTEP;
  pt->table = (double *) malloc(sizeof(double) * ncols * pt->maxsteps);
  while (!feof(f)) {
    if (((npot % PSTEP) == 0) && (npot > 0)) {
      pt->maxsteps += PSTEP;
      pt->table = (double *) realloc(pt->table, sizeof(double) * ncols * pt->maxsteps);
    }
    if (1 != fscanf(f, "%lf", &r2)) break;
    if (npot == 0) r2_start = r2;
    for (i = 0; i < ncols; ++i) {
      if (1 != fscanf(f, "%lf", &val)) return -1;
      pt->table[npot * ncols + i] = val;
      if (val != 0.0) { pt->end[i] = r2; pt->len[i] = npot + 1; }
    }
    ++npot;
  }
  r2_step = (r2 - r2_start) / (npot - 1);
  for (i = 0; i < ncols; ++i) {
    pt->begin[i] = r2_start;
    pt->step[i] = r2_step;
    pt->invstep[i] = 1.0 / r2_step;
    delta = pt->table[(npot - 1) * ncols + i];
    if (radial) {
      if (delta != 0.0)
        for (k = 0; k < npot; ++k) pt->table[k * ncols + i] -= delta;
      s->cellsz = MAXV(s->cellsz, pt->end[i]);
    }
  }
  pt->table = (double *) realloc(pt->table, sizeof(double) * ncols * (pt->maxsteps + 2));
  return 0;
}

/* read_pot_table2, src/imd_potential.c:394-462 */
static int read_table2(orc_sim *s, ptab *pt, FILE *f, int radial)
{
  int ncols = pt->ncols, i, k;
  double val, numstep, delta;
  for (i = 0; i < ncols; i++) {
    if (3 != fscanf(f, "%lf %lf %lf", &pt->begin[i], &pt->end[i], &pt->step[i])) return -1;
    if (radial) s->cellsz = MAXV(s->cellsz, pt->end[i]);
    pt->invstep[i] = 1.0 / pt->step[i];
    numstep = 1 + (pt->end[i] - pt->begin[i]) / pt->step[i];
    pt->len[i] = (int) (numstep + 0.49);
    pt->maxsteps = MAXV(pt->maxsteps, pt->len[i]);
  }
  pt->table = (double *) calloc((size_t) ncols * (pt->maxsteps + 2), sizeof(double));
  for (i = 0; i < ncols; i++)
    for (k = 0; k < pt->len[i]; k++) {
      if (1 != fscanf(f, "%lf", &val)) return -1;
      pt->table[k * ncols + i] = val;
    }
  if (radial)
    for (i = 0; i < ncols; i++) {
      delta = pt->table[(pt->len[i] - 1) * ncols + i];
      if (delta != 0.0)
        for (k = 0; k < pt->len[i]; k++) pt->table[k * ncols + i] -= delta;
    }
  return 0;
}

/* read_pot_table, src/imd_potential.c:161-282 */
int orc_read_table(orc_sim *s, int which, const char *path)
{
  ptab *pt = &s->tab[which];
  int radial = (which != ORC_EMBED && which != ORC_EMOD);   /* adp_upot, adp_wpot are radial (src/imd_potential.c:87-92) */
  int ncols = (which == ORC_EMBED || which == ORC_EMOD) ? s->ntypes : s->ntypes * s->ntypes;
  int have_header = 0, have_format = 0, end_header = 0, format, size = ncols, i, rc;
  char buffer[1024];
  FILE *f = fopen(path, "r");
  if (!f) return -1;
  format = (which == ORC_PAIR) ? s->default_fmt : 2; /* DEFAULT_POTFILE_TYPE, src/config.h:57-63 */
  do {
    if (!fgets(buffer, 1024, f)) { fclose(f); return -2; }
    if (buffer[0] == '#') {
      have_header = 1;
      end_header = (buffer[1] == 'E');
      if (buffer[1] == 'F') {
        if (2 != sscanf(buffer + 2, "%d%d", &format, &size)) { fclose(f); return -3; }
        if (size != ncols) { fclose(f); return -4; }
        if (format != 1 && format != 2) { fclose(f); return -5; }
        have_format = 1;
      }
    } else if (have_header) { fclose(f); return -6; }
    else end_header = 1;
  } while (!end_header);
  if (have_header && !have_format) { fclose(f); return -7; }
  if (!have_header) rewind(f);

  pt->maxsteps = 0; pt->ncols = ncols;
  pt->begin = (double *) calloc(ncols, sizeof(double));
  pt->end = (double *) calloc(ncols, sizeof(double));
  pt->step = (double *) calloc(ncols, sizeof(double));
  pt->invstep = (double *) calloc(ncols, sizeof(double));
  pt->len = (int *) calloc(ncols, sizeof(int));
  for (i = 0; i < ncols; ++i) { pt->end[i] = 0.0; pt->len[i] = 0; }
  rc = (format == 1) ? read_table1(s, pt, f, radial) : read_table2(s, pt, f, radial);
  fclose(f);
  if (rc) return -8;
  /* src/imd_potential.c:270-277 */
  if (s->interp == ORC_INTERP_4POINT) init_fourpoint(pt);
  else if (s->interp == ORC_INTERP_SPLINE) init_spline(pt, radial);
  else init_threepoint(pt);
  pt->loaded = 1;
  return 0;
}

/* PAIR_INT2, src/potaccess.h:323-354 (VAL_FUNC2 :465-495 and DERIV_FUNC2 :591-621 are the
   same computation returning only one of the two results) */
static inline void pair_int2(const ptab *pt, int col, int inc, double r2, double *pot, double *grad, int *is_short)
{
  double r2a, istep, chi, p0, p1, p2, dv, d2v;
  const double *ptr;
  int k;
  r2a = MINV(r2, pt->end[col]);
  r2a = r2a - pt->begin[col];
  if (r2a < 0) { r2a = 0; *is_short = 1; }
  istep = pt->invstep[col];
  r2a = r2a * istep;
  k = (int) (r2a);
  chi = r2a - k;
  ptr = pt->table + (size_t) k * inc + col;
  p0 = *ptr; ptr += inc;
  p1 = *ptr; ptr += inc;
  p2 = *ptr;
  dv = p1 - p0;
  d2v = p2 - 2 * p1 + p0;
  *pot = p0 + chi * dv + 0.5 * chi * (chi - 1) * d2v;
  *grad = 2 * istep * (dv + (chi - 0.5) * d2v);
}

/* PAIR_INT3, src/potaccess.h:365-407: cubic through the four samples k-1 .. k+2 */
static inline void pair_int3(const ptab *pt, int col, int inc, double r2, double *pot, double *grad, int *is_short)
{
  double r2a, istep, chi, p0, p1, p2, p3;
  double fac0, fac1, fac2, fac3, dfac0, dfac1, dfac2, dfac3;
  const double *ptr;
  int k;
  r2a = MINV(r2, pt->end[col]);
  r2a = r2a - pt->begin[col];
  if (r2a < 0) { r2a = 0; *is_short = 1; }
  istep = pt->invstep[col];
  r2a = r2a * istep;
  k = (int) (r2a);
  if (k < 1) k = 1;
  chi = r2a - k;
  fac0 = -(1.0 / 6.0) * chi * (chi - 1.0) * (chi - 2.0);
  fac1 = 0.5 * (chi * chi - 1.0) * (chi - 2.0);
  fac2 = -0.5 * chi * (chi + 1.0) * (chi - 2.0);
  fac3 = (1.0 / 6.0) * chi * (chi * chi - 1.0);
  dfac0 = -(1.0 / 6.0) * ((3.0 * chi - 6.0) * chi + 2.0);
  dfac1 = 0.5 * ((3.0 * chi - 4.0) * chi - 1.0);
  dfac2 = -0.5 * ((3.0 * chi - 2.0) * chi - 2.0);
  dfac3 = 1.0 / 6.0 * (3.0 * chi * chi - 1.0);
  ptr = pt->table + (size_t) (k - 1) * inc + col;
  p0 = *ptr; ptr += inc;
  p1 = *ptr; ptr += inc;
  p2 = *ptr; ptr += inc;
  p3 = *ptr;
  *pot = fac0 * p0 + fac1 * p1 + fac2 * p2 + fac3 * p3;
  *grad = 2 * istep * (dfac0 * p0 + dfac1 * p1 + dfac2 * p2 + dfac3 * p3);
}

/* PAIR_INT_SP, src/potaccess.h:418-457: cubic spline on (table, table2) */
static inline void pair_int_sp(const ptab *pt, int col, int inc, double r2, double *pot, double *grad, int *is_short)
{
  double r2a, a, b, a2, b2, istep, step, st6, p1, p2, d21, d22;
  int k;
  r2a = MINV(r2, pt->end[col]);
  r2a = r2a - pt->begin[col];
  if (r2a < 0) { r2a = 0; *is_short = 1; }
  istep = pt->invstep[col];
  step = pt->step[col];
  r2a = r2a * istep;
  k = (int) (r2a);
  b = r2a - k;
  a = 1.0 - b;
  k = k * inc + col;
  p1 = pt->table[k];
  d21 = pt->table2[k];
  k += inc;
  p2 = pt->table[k];
  d22 = pt->table2[k];
  a2 = a * a - 1;
  b2 = b * b - 1;
  st6 = step / 6;
  *pot = a * p1 + b * p2 + (a * a2 * d21 + b * b2 * d22) * st6 * step;
  *grad = 2 * ((p2 - p1) * istep + ((3 * b2 + 2) * d22 - (3 * a2 + 2) * d21) * st6);
}

/* PAIR_INT as the build selects it (src/potaccess.h:24-36) */
static inline void pair_int(const orc_sim *s, const ptab *pt, int col, int inc, double r2, double *pot, double *grad,
                            int *is_short)
{
  if (s->interp == ORC_INTERP_4POINT) pair_int3(pt, col, inc, r2, pot, grad, is_short);
  else if (s->interp == ORC_INTERP_SPLINE) pair_int_sp(pt, col, inc, r2, pot, grad, is_short);
  else pair_int2(pt, col, inc, r2, pot, grad, is_short);
}

void orc_pair_int(const orc_sim *s, int which, int col, double r2, double *pot, double *grad, int *is_short)
{
  int dummy = 0;
  pair_int(s, &s->tab[which], col, s->tab[which].ncols, r2, pot, grad, is_short ? is_short : &dummy);
}

void orc_set_interpolation(orc_sim *s, int mode) { s->interp = mode; }

int orc_table_info(const orc_sim *s, int which, int col, double *begin, double *end, double *step, int *len)
{
  const ptab *pt = &s->tab[which];
  if (!pt->loaded || col >= pt->ncols) return -1;
  *begin = pt->begin[col]; *end = pt->end[col]; *step = pt->step[col]; *len = pt->len[col];
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* box and cells                                                                         */
/* ------------------------------------------------------------------------------------ */

static vec3 vec_prod(vec3 u, vec3 v)
{
  vec3 w;
  w.x = u.y * v.z - u.z * v.y;
  w.y = u.z * v.x - u.x * v.z;
  w.z = u.x * v.y - u.y * v.x;
  return w;
}
#define SPROD(a, b) (((a).x * (b).x) + ((a).y * (b).y) + ((a).z * (b).z))

static void free_cells(orc_sim *s)
{
  int c;
  if (s->cells) {
    for (c = 0; c < s->nallcells; c++) free(s->cells[c].idx);
    free(s->cells);
  }
  free(s->cnp); free(s->cnq);
  s->cells = NULL; s->cnp = NULL; s->cnq = NULL;
}

static inline int cidx(const orc_sim *s, int i, int j, int k)
{
  return (i * s->cdim[1] + j) * s->cdim[2] + k; /* PTR_3D_V, src/makros.h:438-439 */
}

/* make_cell_lists (NBL version, AR half stencil), src/imd_geom_3d.c:823-953 */
static void make_cell_lists(orc_sim *s)
{
  int i, j, k, l, m, n, r, t, u, nn, qq, c = 0;
  int ipbc[3];
  s->nallcells = s->cdim[0] * s->cdim[1] * s->cdim[2];
  s->ncells = (s->cdim[0] - 2) * (s->cdim[1] - 2) * (s->cdim[2] - 2);
  s->cnp = (int *) malloc(sizeof(int) * s->ncells);
  s->cnq = (int *) malloc(sizeof(int) * s->ncells * 14);
  for (i = 1; i < s->cdim[0] - 1; ++i)
    for (j = 1; j < s->cdim[1] - 1; ++j)
      for (k = 1; k < s->cdim[2] - 1; ++k) {
        int *nq = s->cnq + 14 * c;
        s->cnp[c] = cidx(s, i, j, k);
        nq[0] = s->cnp[c];
        nn = 1;
        for (l = 0; l <= 1; ++l)
          for (m = -l; m <= 1; ++m)
            for (n = (l == 0 ? -m : -l); n <= 1; ++n) {
              r = i + l; t = j + m; u = k + n;
              qq = cidx(s, r, t, u);
              if (qq == s->cnp[c]) continue;
              r += -1; t += -1; u += -1; /* my_coord = 0 */
              ipbc[0] = 0; if (r < 0) ipbc[0]--; else if (r > s->gdim[0] - 1) ipbc[0]++;
              ipbc[1] = 0; if (t < 0) ipbc[1]--; else if (t > s->gdim[1] - 1) ipbc[1]++;
              ipbc[2] = 0; if (u < 0) ipbc[2]--; else if (u > s->gdim[2] - 1) ipbc[2]++;
              if (((s->pbc[0] == 1) || (ipbc[0] == 0)) && ((s->pbc[1] == 1) || (ipbc[1] == 0)) &&
                  ((s->pbc[2] == 1) || (ipbc[2] == 0)))
                nq[nn] = qq;
              else
                nq[nn] = -1;
              nn++;
            }
        c++;
      }
}

/* init_cells, src/imd_geom_3d.c:113-248 (cpu_dim = 1 1 1, no NPT tolerance) */
static void init_cells(orc_sim *s)
{
  vec3 cell_scale;
  int d;
  if (!s->margin_added) { /* :122-126 */
    s->cellsz = SQR(sqrt((double) s->cellsz) + s->nbl_margin);
    s->margin_added = 1;
  }
  cell_scale.x = sqrt(1.0 * s->cellsz / s->height.x);
  cell_scale.y = sqrt(1.0 * s->cellsz / s->height.y);
  cell_scale.z = sqrt(1.0 * s->cellsz / s->height.z);
  s->gdim[0] = (int) (1.0 / cell_scale.x);
  s->gdim[1] = (int) (1.0 / cell_scale.y);
  s->gdim[2] = (int) (1.0 / cell_scale.z);
  for (d = 0; d < 3; d++)
    if (s->gdim[d] < 1) { fprintf(stderr, "oracle: global_cell_dim too small\n"); exit(2); }
  s->min_height.x = s->cellsz * SQR(s->gdim[0]);
  s->min_height.y = s->cellsz * SQR(s->gdim[1]);
  s->min_height.z = s->cellsz * SQR(s->gdim[2]);
  s->max_height.x = s->cellsz * SQR(s->gdim[0] + 1);
  s->max_height.y = s->cellsz * SQR(s->gdim[1] + 1);
  s->max_height.z = s->cellsz * SQR(s->gdim[2] + 1);
  free_cells(s);
  for (d = 0; d < 3; d++) s->cdim[d] = s->gdim[d] + 2;
  make_cell_lists(s);
  s->cells = (cellist *) calloc(s->nallcells, sizeof(cellist));
  s->have_valid_nbl = 0;
}

/* make_box, src/imd_geom_3d.c:52-104 */
static void make_box(orc_sim *s)
{
  s->tbox_x = vec_prod(s->box_y, s->box_z);
  s->tbox_y = vec_prod(s->box_z, s->box_x);
  s->tbox_z = vec_prod(s->box_x, s->box_y);
  s->volume = SPROD(s->box_x, s->tbox_x);
  s->tbox_x.x /= s->volume; s->tbox_x.y /= s->volume; s->tbox_x.z /= s->volume;
  s->tbox_y.x /= s->volume; s->tbox_y.y /= s->volume; s->tbox_y.z /= s->volume;
  s->tbox_z.x /= s->volume; s->tbox_z.y /= s->volume; s->tbox_z.z /= s->volume;
  s->height.x = 1.0 / SPROD(s->tbox_x, s->tbox_x);
  s->height.y = 1.0 / SPROD(s->tbox_y, s->tbox_y);
  s->height.z = 1.0 / SPROD(s->tbox_z, s->tbox_z);
  if ((s->height.x < s->min_height.x) || (s->height.x > s->max_height.x) ||
      (s->height.y < s->min_height.y) || (s->height.y > s->max_height.y) ||
      (s->height.z < s->min_height.z) || (s->height.z > s->max_height.z))
    init_cells(s);
  if (0 > s->volume) s->volume = -s->volume;
}

/* cell_coord, src/imd_geom_3d.c:1054-1074 */
static void cell_coord(const orc_sim *s, double x, double y, double z, int c[3])
{
  c[0] = (int) (s->gdim[0] * (x * s->tbox_x.x + y * s->tbox_x.y + z * s->tbox_x.z));
  c[1] = (int) (s->gdim[1] * (x * s->tbox_y.x + y * s->tbox_y.y + z * s->tbox_y.z));
  c[2] = (int) (s->gdim[2] * (x * s->tbox_z.x + y * s->tbox_z.y + z * s->tbox_z.z));
  if (c[0] >= s->gdim[0]) c[0] = s->gdim[0] - 1; else if (c[0] < 0) c[0] = 0;
  if (c[1] >= s->gdim[1]) c[1] = s->gdim[1] - 1; else if (c[1] < 0) c[1] = 0;
  if (c[2] >= s->gdim[2]) c[2] = s->gdim[2] - 1; else if (c[2] < 0) c[2] = 0;
}

/* do_boundaries, src/imd_main_3d.c:1972-2059 */
static void do_boundaries(orc_sim *s)
{
  long l; double i; double *o;
  if (s->pbc[0] == 1)
    for (l = 0; l < s->n; ++l) {
      o = s->ort + 3 * l;
      i = -floor(o[0] * s->tbox_x.x + o[1] * s->tbox_x.y + o[2] * s->tbox_x.z);
      o[0] += i * s->box_x.x; o[1] += i * s->box_x.y; o[2] += i * s->box_x.z;
    }
  if (s->pbc[1] == 1)
    for (l = 0; l < s->n; ++l) {
      o = s->ort + 3 * l;
      i = -floor(o[0] * s->tbox_y.x + o[1] * s->tbox_y.y + o[2] * s->tbox_y.z);
      o[0] += i * s->box_y.x; o[1] += i * s->box_y.y; o[2] += i * s->box_y.z;
    }
  if (s->pbc[2] == 1)
    for (l = 0; l < s->n; ++l) {
      o = s->ort + 3 * l;
      i = -floor(o[0] * s->tbox_z.x + o[1] * s->tbox_z.y + o[2] * s->tbox_z.z);
      o[0] += i * s->box_z.x; o[1] += i * s->box_z.y; o[2] += i * s->box_z.z;
    }
}

static void cell_push(cellist *c, int a)
{
  if (c->n == c->cap) { c->cap = c->cap ? 2 * c->cap : 16; c->idx = (int *) realloc(c->idx, sizeof(int) * c->cap); }
  c->idx[c->n++] = a;
}

/* fix_cells, src/imd_fix_cells_3d.c:36-201: wrap, then put every atom into the cell that
   cell_coord/local_cell_coord (src/imd_geom_mpi_3d.c:119-128) assign.  The order of atoms
   inside a cell is history dependent in the reference and irrelevant to results
   (SURVEY.md section 9 item 1); here it is the storage order. */
static void fix_cells(orc_sim *s)
{
  long a; int c, cc[3];
  do_boundaries(s);
  for (c = 0; c < s->nallcells; c++) s->cells[c].n = 0;
  for (a = 0; a < s->n; a++) {
    cell_coord(s, s->ort[3 * a], s->ort[3 * a + 1], s->ort[3 * a + 2], cc);
    cell_push(&s->cells[cidx(s, cc[0] + 1, cc[1] + 1, cc[2] + 1)], (int) a);
  }
  s->ng = 0;
  s->have_valid_nbl = 0; /* :196-199 */
}

static void grow_atoms(orc_sim *s, long need)
{
  if (need <= s->cap) return;
  s->cap = need + need / 4 + 64;
#define GROW(p, T, m) p = (T *) realloc(p, sizeof(T) * (m) * s->cap)
  GROW(s->nummer, int, 1); GROW(s->sorte, int, 1); GROW(s->vsorte, int, 1);
  GROW(s->masse, double, 1); GROW(s->ort, double, 3); GROW(s->impuls, double, 3);
  GROW(s->kraft, double, 3); GROW(s->poteng, double, 1); GROW(s->rho, double, 1);
  GROW(s->dF, double, 1); GROW(s->presstens, double, 6); GROW(s->nblpos, double, 3);
  GROW(s->eam_p, double, 1); GROW(s->dM, double, 1); GROW(s->mu, double, 3); GROW(s->la, double, 6);
  GROW(s->gsrc, long, 1); GROW(s->gshift, signed char, 3);
#undef GROW
}

/* copy_cell, src/imd_comm_force_3d.c:726-778: positions (+shift) and types only */
static void copy_cell(orc_sim *s, int k, int l, int m, int r, int t, int u, vec3 v, int ax, int sgn, int first)
{
  cellist *from = &s->cells[cidx(s, k, l, m)], *to = &s->cells[cidx(s, r, t, u)];
  int i;
  if (first) {
    to->n = 0;
    for (i = 0; i < from->n; i++) {
      long g = s->n + s->ng, src = from->idx[i];
      grow_atoms(s, g + 1);
      s->ng++;
      s->gsrc[g] = src;
      s->nummer[g] = s->nummer[src];
      if (src >= s->n) memcpy(s->gshift + 3 * g, s->gshift + 3 * src, 3);
      else memset(s->gshift + 3 * g, 0, 3);
      if (sgn) s->gshift[3 * g + ax] = (signed char) sgn;
      cell_push(to, (int) g);
    }
  }
  for (i = 0; i < to->n; i++) {
    long g = to->idx[i], src = from->idx[i];
    s->ort[3 * g] = s->ort[3 * src] + v.x;
    s->ort[3 * g + 1] = s->ort[3 * src + 1] + v.y;
    s->ort[3 * g + 2] = s->ort[3 * src + 2] + v.z;
    s->sorte[g] = s->sorte[src];
  }
}

/* send_cells(copy_cell,...), src/imd_comm_force_3d.c:222-396, cpu_dim = 1 1 1, AR mode.
   `first` = buffer atoms are (re)created (after fix_cells); otherwise only refreshed. */
static void send_cells_pos(orc_sim *s, int first)
{
  int i, j; const int *cd = s->cdim;
  vec3 z0 = {0, 0, 0}, uvec = z0, dvec = z0, nvec = z0, svec = z0, evec = z0;
  if (s->pbc[0] == 1) evec = s->box_x;
  if (s->pbc[1] == 1) { nvec = s->box_y; svec.x = -s->box_y.x; svec.y = -s->box_y.y; svec.z = -s->box_y.z; }
  if (s->pbc[2] == 1) { uvec = s->box_z; dvec.x = -s->box_z.x; dvec.y = -s->box_z.y; dvec.z = -s->box_z.z; }
  /* Non-periodic directions: the reference still copies (with zero shift) but never
     references those buffer cells (nq = -1); we leave them empty. */
  if (s->pbc[2] == 1)
    for (i = 1; i < cd[0] - 1; ++i)
      for (j = 1; j < cd[1] - 1; ++j) {
        copy_cell(s, i, j, 1, i, j, cd[2] - 1, uvec, 2, +1, first);
        copy_cell(s, i, j, cd[2] - 2, i, j, 0, dvec, 2, -1, first);
      }
  if (first) s->gstage[0] = s->n + s->ng;
  if (s->pbc[1] == 1)
    for (i = 1; i < cd[0] - 1; ++i)
      for (j = 0; j < cd[2]; ++j) {
        copy_cell(s, i, 1, j, i, cd[1] - 1, j, nvec, 1, +1, first);
        copy_cell(s, i, cd[1] - 2, j, i, 0, j, svec, 1, -1, first);
      }
  if (first) s->gstage[1] = s->n + s->ng;
  if (s->pbc[0] == 1)
    for (i = 0; i < cd[1]; ++i)
      for (j = 0; j < cd[2]; ++j)
        copy_cell(s, 1, i, j, cd[0] - 1, i, j, evec, 0, +1, first);
  if (first) s->gstage[2] = s->n + s->ng;
}

/* send_cells(copy_dF,...): src/imd_comm_force_3d.c:1031-1060 -- same sweep, payload eam_dF.
   Buffer atoms were created in sweep order (z, then y, then x stage), so walking them in
   index order reproduces the staged copies. */
static void send_cells_dF(orc_sim *s)
{
  long g;
  for (g = s->n; g < s->n + s->ng; g++) {
    long t = s->gsrc[g]; int d;
    s->dF[g] = s->dF[t]; s->dM[g] = s->dM[t];            /* + EAM_DM :1044-1046 */
    for (d = 0; d < 3; d++) s->mu[3 * g + d] = s->mu[3 * t + d];   /* + ADP_MU, ADP_LAMBDA :1047-1057 */
    for (d = 0; d < 6; d++) s->la[6 * g + d] = s->la[6 * t + d];
  }
}

/* send_forces(add_rho,...) / send_forces(add_forces,...): src/imd_comm_force_3d.c:569-714,
   897-932, 1068-1096.  Reverse sweep x -> y -> z.  Buffer atoms were created stage by
   stage (gstage[]), and inside a stage in the reference's loop order, so a forward walk
   over each stage segment, stages taken in reverse, reproduces the reference's order of
   accumulation. */
static void send_forces_back(orc_sim *s, int what, int do_press)
{
  int st, d; long g, lo, hi;
  for (st = 2; st >= 0; st--) {
    lo = (st == 0) ? s->n : s->gstage[st - 1];
    hi = s->gstage[st];
    for (g = lo; g < hi; g++) {
      long t = s->gsrc[g];
      if (what == 0) { /* add_rho */
        s->rho[t] += s->rho[g];
        s->eam_p[t] += s->eam_p[g];              /* EEAM :1081-1083 */
        for (d = 0; d < 3; d++) s->mu[3 * t + d] += s->mu[3 * g + d];   /* ADP :1084-1094 */
        for (d = 0; d < 6; d++) s->la[6 * t + d] += s->la[6 * g + d];
      } else { /* add_forces */
        s->kraft[3 * t] += s->kraft[3 * g];
        s->kraft[3 * t + 1] += s->kraft[3 * g + 1];
        s->kraft[3 * t + 2] += s->kraft[3 * g + 2];
        s->poteng[t] += s->poteng[g];
        if (do_press) for (d = 0; d < 6; d++) s->presstens[6 * t + d] += s->presstens[6 * g + d];
      }
    }
  }
}

/* ------------------------------------------------------------------------------------ */
/* neighbour list                                                                        */
/* ------------------------------------------------------------------------------------ */

/* make_nblist, src/imd_forces_nbl.c:136-273 */
static void make_nblist(orc_sim *s)
{
  long n, tn, at, k; int c, i, m;
  long ntot = s->n + s->ng;
  /* reference positions :142-154 */
  for (k = 0; k < s->n; k++) {
    s->nblpos[3 * k] = s->ort[3 * k]; s->nblpos[3 * k + 1] = s->ort[3 * k + 1]; s->nblpos[3 * k + 2] = s->ort[3 * k + 2];
  }
  /* cl_off / cl_num :163-169, 211-216 */
  s->cl_off = (int *) realloc(s->cl_off, sizeof(int) * s->nallcells);
  s->cl_num = (int *) realloc(s->cl_num, sizeof(int) * (ntot + 1));
  at = 0;
  for (c = 0; c < s->nallcells; c++) { s->cl_off[c] = (int) at; at += s->cells[c].n; }
  n = 0;
  for (c = 0; c < s->nallcells; c++) for (i = 0; i < s->cells[c].n; i++) s->cl_num[n++] = c;
  s->tl = (long *) realloc(s->tl, sizeof(long) * (s->n + 2));
  /* :218-269 */
  n = 0; tn = 0; s->tl[0] = 0;
  for (c = 0; c < s->ncells; c++) {
    int c1 = s->cnp[c];
    cellist *p = &s->cells[c1];
    for (i = 0; i < p->n; i++) {
      const double *o1 = s->ort + 3 * (long) p->idx[i];
      double d1x = o1[0], d1y = o1[1], d1z = o1[2];
      for (m = 0; m < 14; m++) {
        int c2 = s->cnq[14 * c + m], jstart, j; cellist *q;
        if (c2 < 0) continue;
        jstart = (c2 == c1) ? i + 1 : 0;
        q = &s->cells[c2];
        for (j = jstart; j < q->n; j++) {
          const double *o2 = s->ort + 3 * (long) q->idx[j];
          double dx = o2[0] - d1x, dy = o2[1] - d1y, dz = o2[2] - d1z;
          double r2 = ((dx * dx) + (dy * dy)) + (dz * dz); /* SPROD3D, src/makros.h:409 */
          if (r2 < s->cellsz) {
            if (tn >= s->tb_cap) { s->tb_cap = s->tb_cap ? 2 * s->tb_cap : 1 << 16; s->tb = (int *) realloc(s->tb, sizeof(int) * s->tb_cap); }
            s->tb[tn++] = s->cl_off[c2] + j;
          }
        }
      }
      s->tl[++n] = tn;
    }
  }
  s->have_valid_nbl = 1;
  s->nbl_count++;
}

/* ------------------------------------------------------------------------------------ */
/* forces                                                                                */
/* ------------------------------------------------------------------------------------ */

/* calc_forces, src/imd_forces_nbl.c:281-1999 (PAIR and EAM2 branches, no P_AXIAL) */
void orc_calc_forces(orc_sim *s, int do_press_calc)
{
  const ptab *pair_pot = &s->tab[ORC_PAIR], *embed_pot = &s->tab[ORC_EMBED], *rho_h_tab = &s->tab[ORC_RHO];
  const ptab *emod_pot = &s->tab[ORC_EMOD], *adp_upot = &s->tab[ORC_ADP_U], *adp_wpot = &s->tab[ORC_ADP_W];
  const int nt = s->ntypes, inc = nt * nt, eam = rho_h_tab->loaded, eeam = eam && emod_pot->loaded, adp = eam && adp_upot->loaded && adp_wpot->loaded;
  long n, ntot, a; int c, i; long m;
  int is_short = 0, idummy = 0;

  if (0 == s->have_valid_nbl) { fix_cells(s); send_cells_pos(s, 1); make_nblist(s); } /* :304-317 */
  else send_cells_pos(s, 0);                                                          /* :314 */
  ntot = s->n + s->ng;

  s->tot_pot_energy = 0.0; s->virial = 0.0; s->vir[0] = s->vir[1] = s->vir[2] = 0.0; /* :319-331 */
  for (a = 0; a < ntot; a++) {              /* :333-401 */
    s->kraft[3 * a] = s->kraft[3 * a + 1] = s->kraft[3 * a + 2] = 0.0;
    for (i = 0; i < 6; i++) s->presstens[6 * a + i] = 0.0;
    s->poteng[a] = 0.0; s->rho[a] = 0.0; s->eam_p[a] = 0.0;
    for (i = 0; i < 3; i++) s->mu[3 * a + i] = 0.0;
    for (i = 0; i < 6; i++) s->la[6 * a + i] = 0.0;
  }

  /* atom index of list slot: slot = cl_off[c] + j */
#define SLOT2ATOM(slot) (s->cells[s->cl_num[slot]].idx[(slot) - s->cl_off[s->cl_num[slot]]])

  /* pair interactions :422-981 */
  n = 0;
  for (c = 0; c < s->ncells; c++) {
    cellist *p = &s->cells[s->cnp[c]];
    for (i = 0; i < p->n; i++) {
      long ia = p->idx[i];
      double pp[6] = {0, 0, 0, 0, 0, 0};
      double d1x = s->ort[3 * ia], d1y = s->ort[3 * ia + 1], d1z = s->ort[3 * ia + 2];
      double ffx = 0.0, ffy = 0.0, ffz = 0.0, ee = 0.0, eam_r = 0.0, eam_p = 0.0;
      double mu[3] = {0, 0, 0}, la[6] = {0, 0, 0, 0, 0, 0};
      int it = s->sorte[ia];
      for (m = s->tl[n]; m < s->tl[n + 1]; m++) {
        long ja = SLOT2ATOM(s->tb[m]);
        double dx = s->ort[3 * ja] - d1x, dy = s->ort[3 * ja + 1] - d1y, dz = s->ort[3 * ja + 2] - d1z;
        double r2 = ((dx * dx) + (dy * dy)) + (dz * dz);
        double pot, grad, rho_h = 0.0, fx, fy, fz;
        int jt = s->sorte[ja], col = it * nt + jt, col2 = jt * nt + it;
        if (r2 <= pair_pot->end[col]) { /* :493 */
          pair_int(s, pair_pot, col, inc, r2, &pot, &grad, &is_short);
          s->tot_pot_energy += pot;
          fx = dx * grad; fy = dy * grad; fz = dz * grad;
          s->kraft[3 * ja] -= fx; s->kraft[3 * ja + 1] -= fy; s->kraft[3 * ja + 2] -= fz;
          ffx += fx; ffy += fy; ffz += fz;
          pot *= 0.5;
          ee += pot; s->poteng[ja] += pot;
          s->virial -= r2 * grad;
          s->vir[0] -= dx * fx; s->vir[1] -= dy * fy; s->vir[2] -= dz * fz;   /* P_AXIAL :548-553 */
          if (do_press_calc) { /* :558-581 */
            fx *= 0.5; fy *= 0.5; fz *= 0.5;
            pp[0] -= dx * fx; s->presstens[6 * ja] -= dx * fx;         /* xx */
            pp[1] -= dy * fy; s->presstens[6 * ja + 1] -= dy * fy;     /* yy */
            pp[5] -= dx * fy; s->presstens[6 * ja + 5] -= dx * fy;     /* xy */
            pp[2] -= dz * fz; s->presstens[6 * ja + 2] -= dz * fz;     /* zz */
            pp[3] -= dy * fz; s->presstens[6 * ja + 3] -= dy * fz;     /* yz */
            pp[4] -= dz * fx; s->presstens[6 * ja + 4] -= dz * fx;     /* zx */
          }
        }
        if (eam) { /* :586-611 */
          double dummy;
          if (r2 < rho_h_tab->end[col]) {
            pair_int(s, rho_h_tab, col, inc, r2, &rho_h, &dummy, &is_short);
            eam_r += rho_h;
            if (eeam) eam_p += rho_h * rho_h;            /* :591-593 */
          }
          if (it == jt) {
            if (r2 < rho_h_tab->end[col]) { s->rho[ja] += rho_h; if (eeam) s->eam_p[ja] += rho_h * rho_h; }
          } else {
            if (r2 < rho_h_tab->end[col2]) {
              pair_int(s, rho_h_tab, col2, inc, r2, &rho_h, &dummy, &is_short);
              s->rho[ja] += rho_h;
              if (eeam) s->eam_p[ja] += rho_h * rho_h;   /* :606-608 */
            }
          }
        }
        if (adp) { /* :613-631 */
          double tmp, dummy;
          if (r2 < adp_upot->end[col]) {
            pair_int(s, adp_upot, col, inc, r2, &pot, &dummy, &is_short);
            tmp = pot * dx; mu[0] += tmp; s->mu[3 * ja] -= tmp;
            tmp = pot * dy; mu[1] += tmp; s->mu[3 * ja + 1] -= tmp;
            tmp = pot * dz; mu[2] += tmp; s->mu[3 * ja + 2] -= tmp;
          }
          if (r2 < adp_wpot->end[col]) {
            pair_int(s, adp_wpot, col, inc, r2, &pot, &dummy, &is_short);
            tmp = pot * dx * dx; la[0] += tmp; s->la[6 * ja] += tmp;         /* xx */
            tmp = pot * dy * dy; la[1] += tmp; s->la[6 * ja + 1] += tmp;     /* yy */
            tmp = pot * dz * dz; la[2] += tmp; s->la[6 * ja + 2] += tmp;     /* zz */
            tmp = pot * dy * dz; la[3] += tmp; s->la[6 * ja + 3] += tmp;     /* yz */
            tmp = pot * dz * dx; la[4] += tmp; s->la[6 * ja + 4] += tmp;     /* zx */
            tmp = pot * dx * dy; la[5] += tmp; s->la[6 * ja + 5] += tmp;     /* xy */
          }
        }
      }
      if (adp) { int d;                                  /* :919-929 */
        for (d = 0; d < 3; d++) s->mu[3 * ia + d] += mu[d];
        for (d = 0; d < 6; d++) s->la[6 * ia + d] += la[d]; }
      s->kraft[3 * ia] += ffx; s->kraft[3 * ia + 1] += ffy; s->kraft[3 * ia + 2] += ffz; /* :907-918 */
      s->poteng[ia] += ee;
      if (eam) s->rho[ia] += eam_r;
      if (eeam) s->eam_p[ia] += eam_p;                 /* :915-917 */
      if (do_press_calc) { int d; for (d = 0; d < 6; d++) s->presstens[6 * ia + d] += pp[d]; }
      n++;
    }
  }

  if (eam) {
    send_forces_back(s, 0, 0); /* :1076 */
    /* embedding energy :1079-1095 */
    for (c = 0; c < s->ncells; c++) {
      cellist *p = &s->cells[s->cnp[c]];
      for (i = 0; i < p->n; i++) {
        long ia = p->idx[i]; double pot;
        pair_int(s, embed_pot, s->sorte[ia], nt, s->rho[ia], &pot, &s->dF[ia], &idummy);
        s->poteng[ia] += pot;
        s->tot_pot_energy += pot;
        if (eeam) {                                      /* :1090-1095 */
          pair_int(s, emod_pot, s->sorte[ia], nt, s->eam_p[ia], &pot, &s->dM[ia], &idummy);
          s->poteng[ia] += pot;
          s->tot_pot_energy += pot;
        }
        if (adp) {                                       /* :1096-1110 */
          const double *L = s->la + 6 * ia, *M = s->mu + 3 * ia; double tr, tmp;
          tr = (L[0] + L[1] + L[2]) / 3.0;
          tmp = L[0] - tr; pot = tmp * tmp;
          tmp = L[1] - tr; pot += tmp * tmp;
          tmp = L[2] - tr; pot += tmp * tmp;
          tmp = L[3]; pot += (tmp * tmp) * 2.0;
          tmp = L[4]; pot += (tmp * tmp) * 2.0;
          tmp = L[5]; pot += (tmp * tmp) * 2.0;
          tmp = M[0]; pot += tmp * tmp;
          tmp = M[1]; pot += tmp * tmp;
          tmp = M[2]; pot += tmp * tmp;
          pot *= 0.5;
          s->poteng[ia] += pot;
          s->tot_pot_energy += pot;
        }
      }
    }
    send_cells_dF(s); /* :1115 */
    /* EAM force pass :1117-1322 */
    n = 0;
    for (c = 0; c < s->ncells; c++) {
      cellist *p = &s->cells[s->cnp[c]];
      for (i = 0; i < p->n; i++) {
        long ia = p->idx[i];
        double pp[6] = {0, 0, 0, 0, 0, 0};
        double d1x = s->ort[3 * ia], d1y = s->ort[3 * ia + 1], d1z = s->ort[3 * ia + 2];
        double ffx = 0.0, ffy = 0.0, ffz = 0.0;
        int it = s->sorte[ia];
        for (m = s->tl[n]; m < s->tl[n + 1]; m++) {
          long ja = SLOT2ATOM(s->tb[m]);
          double dx = s->ort[3 * ja] - d1x, dy = s->ort[3 * ja + 1] - d1y, dz = s->ort[3 * ja + 2] - d1z;
          double r2 = ((dx * dx) + (dy * dy)) + (dz * dz);
          int jt = s->sorte[ja], col1 = jt * nt + it, col2 = it * nt + jt;
          double fx = 0.0, fy = 0.0, fz = 0.0; int have_force = 0;
          if ((r2 < rho_h_tab->end[col1]) || (r2 < rho_h_tab->end[col2])) { /* :1172 */
            double rho_i = 0.0, rho_j = 0.0, rho_i_strich, rho_j_strich, grad;
            pair_int(s, rho_h_tab, col1, inc, r2, &rho_i, &rho_i_strich, &is_short);
            if (col1 == col2) { rho_j_strich = rho_i_strich; rho_j = rho_i; }
            else pair_int(s, rho_h_tab, col2, inc, r2, &rho_j, &rho_j_strich, &is_short);
            grad = 0.5 * (s->dF[ia] * rho_j_strich + s->dF[ja] * rho_i_strich); /* :1203 */
            if (eeam)                                    /* :1204-1208 */
              grad += (s->dM[ia] * rho_j * rho_j_strich + s->dM[ja] * rho_i * rho_i_strich);
            fx = dx * grad; fy = dy * grad; fz = dz * grad;
            have_force = 1;
          }
          if (adp) {
            if (r2 < adp_upot->end[col1]) {              /* dipole distortion :1217-1229 */
              double pot, grad, tmp, m0, m1, m2;
              pair_int(s, adp_upot, col1, inc, r2, &pot, &grad, &is_short);
              m0 = s->mu[3 * ia] - s->mu[3 * ja]; m1 = s->mu[3 * ia + 1] - s->mu[3 * ja + 1]; m2 = s->mu[3 * ia + 2] - s->mu[3 * ja + 2];
              tmp = (((m0 * dx) + (m1 * dy)) + (m2 * dz)) * grad;
              fx += m0 * pot + tmp * dx; fy += m1 * pot + tmp * dy; fz += m2 * pot + tmp * dz;
              have_force = 1;
            }
            if (r2 < adp_wpot->end[col1]) {              /* quadrupole distortion :1231-1254 */
              double pot, grad, nu, f1, f2, L[6], vx, vy, vz; int d;
              pair_int(s, adp_wpot, col1, inc, r2, &pot, &grad, &is_short);
              for (d = 0; d < 6; d++) L[d] = s->la[6 * ia + d] + s->la[6 * ja + d];
              vx = L[0] * dx + L[5] * dy + L[4] * dz;
              vy = L[5] * dx + L[1] * dy + L[3] * dz;
              vz = L[4] * dx + L[3] * dy + L[2] * dz;
              nu = (L[0] + L[1] + L[2]) / 3.0;
              f1 = 2.0 * pot;
              f2 = ((((vx * dx) + (vy * dy)) + (vz * dz)) - nu * r2) * grad - nu * f1;
              fx += f1 * vx + f2 * dx; fy += f1 * vy + f2 * dy; fz += f1 * vz + f2 * dz;
              have_force = 1;
            }
          }
          if (have_force) {                              /* :1267-1305 */
            s->kraft[3 * ja] -= fx; s->kraft[3 * ja + 1] -= fy; s->kraft[3 * ja + 2] -= fz;
            ffx += fx; ffy += fy; ffz += fz;
            s->virial -= ((dx * fx) + (dy * fy)) + (dz * fz); /* :1280 */
            s->vir[0] -= dx * fx; s->vir[1] -= dy * fy; s->vir[2] -= dz * fz;   /* P_AXIAL :1275-1279 */
            if (do_press_calc) { /* :1283-1304 */
              fx *= 0.5; fy *= 0.5; fz *= 0.5;
              pp[0] -= dx * fx; pp[1] -= dy * fy; pp[2] -= dz * fz;
              pp[3] -= dy * fz; pp[4] -= dz * fx; pp[5] -= dx * fy;
              s->presstens[6 * ja] -= dx * fx; s->presstens[6 * ja + 1] -= dy * fy;
              s->presstens[6 * ja + 2] -= dz * fz; s->presstens[6 * ja + 3] -= dy * fz;
              s->presstens[6 * ja + 4] -= dz * fx; s->presstens[6 * ja + 5] -= dx * fy;
            }
          }
        }
        s->kraft[3 * ia] += ffx; s->kraft[3 * ia + 1] += ffy; s->kraft[3 * ia + 2] += ffz;
        if (do_press_calc) { int d; for (d = 0; d < 6; d++) s->presstens[6 * ia + d] += pp[d]; }
        n++;
      }
    }
  }
  send_forces_back(s, 1, do_press_calc); /* :1997 */
  if (is_short) s->is_short = 1;
#undef SLOT2ATOM
}

/* ------------------------------------------------------------------------------------ */
/* integrators and list check                                                            */
/* ------------------------------------------------------------------------------------ */

/* move_atoms_nve (src/imd_integrate.c:32-497) and move_atoms_nvt (:891-1147); atoms are
   visited in cell-traversal order like the reference so that the energy sums round alike */
static void calc_dyn_pressure(orc_sim *s);
static void move_atoms_npt_iso(orc_sim *s, int do_press_calc);
static void move_atoms_npt_axial(orc_sim *s, int do_press_calc);

void orc_move_atoms(orc_sim *s, int do_press_calc)
{
  int c, i; const double dt = s->timestep;
  double E_kin_1 = 0.0, E_kin_2 = 0.0, reibung = 0, eins_d_reib = 0;
  if (s->ensemble == ORC_NPT_ISO) {
    if (s->Ekin_old < 0.0) { calc_dyn_pressure(s); if (s->isq_tau_xi == 0.0) s->xi = 0.0; }   /* :1493-1496 */
    move_atoms_npt_iso(s, do_press_calc);
    return;
  }
  if (s->ensemble == ORC_NPT_AXIAL) {
    if (s->Ekin_old < 0.0) {                                                          /* steps == steps_min :1756-1771 */
      int d;
      calc_dyn_pressure(s);
      for (d = 0; d < 3; d++) { if (s->isq_tau_xi == 0.0) s->xi3[d] = 0.0; s->xi3[d] *= s->relax_dirs[d]; }
    }
    move_atoms_npt_axial(s, do_press_calc);
    return;
  }
  double cc = 1.0;
  if (s->ensemble == ORC_NVE && s->tauber > 0.0) {   /* Ju Li's Berendsen thermostat, from the PREVIOUS step's Ekin :44-52 */
    cc = 1. - dt / s->tauber * ((2.0 * s->tot_kin_energy / s->nactive + 8.6174101569719990e-06) / (s->temperature + 8.6174101569719990e-06) - 1.);
    if (cc < 0.5) cc = 0.5;
    else if (cc > 2.0) cc = 2.0;
    cc = sqrt(cc);
  }
  if (s->ensemble == ORC_NVE) s->tot_kin_energy = 0.0;
  else {
    reibung = 1.0 - s->eta * dt / 2.0;               /* :907 */
    eins_d_reib = 1.0 / (1.0 + s->eta * dt / 2.0);   /* :908 */
  }
  for (c = 0; c < s->ncells; c++) {
    cellist *p = &s->cells[s->cnp[c]];
    for (i = 0; i < p->n; i++) {
      long a = p->idx[i];
      double *P = s->impuls + 3 * a, *F = s->kraft + 3 * a, *X = s->ort + 3 * a;
      const double *R = s->restr + 3 * s->vsorte[a];
      double m = s->masse[a], tmp;
      if (s->ensemble == ORC_NVE) {
        double k1 = (P[0] * P[0] + P[1] * P[1]) + P[2] * P[2], k2;
        F[0] *= R[0]; F[1] *= R[1]; F[2] *= R[2];                       /* :192-197 */
        P[0] += dt * F[0]; P[1] += dt * F[1]; P[2] += dt * F[2];        /* :213-217 */
        k2 = (P[0] * P[0] + P[1] * P[1]) + P[2] * P[2];
        s->tot_kin_energy += (k1 + k2) / (4 * m);                       /* :329-335 */
        if (s->tauber > 0.0) { P[0] *= cc; P[1] *= cc; P[2] *= cc; }    /* :341-350 */
      } else {
        E_kin_1 += ((P[0] * P[0] + P[1] * P[1]) + P[2] * P[2]) / m;    /* :951 */
        F[0] *= R[0]; F[1] *= R[1]; F[2] *= R[2];
        P[0] = (P[0] * reibung + dt * F[0]) * eins_d_reib * R[0];       /* :1020-1027 */
        P[1] = (P[1] * reibung + dt * F[1]) * eins_d_reib * R[1];
        P[2] = (P[2] * reibung + dt * F[2]) * eins_d_reib * R[2];
        E_kin_2 += ((P[0] * P[0] + P[1] * P[1]) + P[2] * P[2]) / m;
      }
      tmp = dt / m;                                                      /* :353-358 */
      X[0] += tmp * P[0]; X[1] += tmp * P[1]; X[2] += tmp * P[2];
      if (do_press_calc) {                                               /* :410-433 */
        double *S = s->presstens + 6 * a;
        S[0] += P[0] * P[0] / m; S[1] += P[1] * P[1] / m; S[2] += P[2] * P[2] / m;
        S[3] += P[1] * P[2] / m; S[4] += P[2] * P[0] / m; S[5] += P[0] * P[1] / m;
      }
    }
  }
  if (s->ensemble == ORC_NVT) {
    double ttt;
    s->tot_kin_energy = (E_kin_1 + E_kin_2) / 4.0;                      /* :1103 */
    ttt = s->nactive * s->temperature;
    s->eta += dt * (E_kin_2 / ttt - 1.0) * s->isq_tau_eta;              /* :1140-1141 */
  }
}

/* calc_dyn_pressure, src/imd_integrate.c:1403-1465: twice the kinetic energy from the current momenta */
static void calc_dyn_pressure(orc_sim *s)
{
  int c, i; double sx = 0.0, sy = 0.0, sz = 0.0;
  for (c = 0; c < s->ncells; c++) {
    cellist *p = &s->cells[s->cnp[c]];
    for (i = 0; i < p->n; i++) {
      long a = p->idx[i]; const double *P = s->impuls + 3 * a; double tmp = 1.0 / s->masse[a];
      sx += P[0] * P[0] * tmp; sy += P[1] * P[1] * tmp; sz += P[2] * P[2] * tmp;
    }
  }
  s->dyn3[0] = sx; s->dyn3[1] = sy; s->dyn3[2] = sz;
  s->Ekin_old = sx + sy; s->Ekin_old += sz;
}

/* move_atoms_npt_iso, src/imd_integrate.c:1472-1729 (no UNIAX, no restrictions in the reference either) */
static void move_atoms_npt_iso(orc_sim *s, int do_press_calc)
{
  int c, i; const double dt = s->timestep;
  double Ekin_new = 0.0, pfric, pifric, rfric, rifric, xi_old, ttt;
  s->pressure = (s->Ekin_old + s->virial) / (3 * s->volume);                         /* :1505 */
  xi_old = s->xi;
  s->xi += dt * (s->pressure - s->pressure_ext) * s->volume * s->isq_tau_xi / s->nactive;   /* :1509 */
  pfric = 1.0 - (xi_old + s->eta) * dt / 2.0;                                        /* :1512-1515 */
  pifric = 1.0 / (1.0 + (s->xi + s->eta) * dt / 2.0);
  rfric = 1.0 + (s->xi) * dt / 2.0;
  rifric = 1.0 / (1.0 - (s->xi) * dt / 2.0);
  for (c = 0; c < s->ncells; c++) {
    cellist *p = &s->cells[s->cnp[c]];
    for (i = 0; i < p->n; i++) {
      long a = p->idx[i];
      double *P = s->impuls + 3 * a, *F = s->kraft + 3 * a, *X = s->ort + 3 * a, m = s->masse[a], tmp;
      P[0] = (pfric * P[0] + dt * F[0]) * pifric;                                    /* :1572-1576 */
      P[1] = (pfric * P[1] + dt * F[1]) * pifric;
      P[2] = (pfric * P[2] + dt * F[2]) * pifric;
      Ekin_new += ((P[0] * P[0] + P[1] * P[1]) + P[2] * P[2]) / m;                   /* :1591 */
      tmp = dt / m;
      X[0] = (rfric * X[0] + P[0] * tmp) * rifric;                                   /* :1603-1607 */
      X[1] = (rfric * X[1] + P[1] * tmp) * rifric;
      X[2] = (rfric * X[2] + P[2] * tmp) * rifric;
      if (do_press_calc) {                                                           /* :1641-1652 */
        double *S = s->presstens + 6 * a;
        S[0] += P[0] * P[0] / m; S[1] += P[1] * P[1] / m; S[2] += P[2] * P[2] / m;
        S[3] += P[1] * P[2] / m; S[4] += P[2] * P[0] / m; S[5] += P[0] * P[1] / m;
      }
    }
  }
  s->tot_kin_energy = (s->Ekin_old + Ekin_new) / 4.0;                                /* :1691 */
  ttt = s->nactive * s->temperature;
  s->eta += dt * (Ekin_new / ttt - 1.0) * s->isq_tau_eta;                            /* :1695-1696 */
  s->Ekin_old = Ekin_new;
  ttt = (1.0 + s->xi * dt / 2.0) / (1.0 - s->xi * dt / 2.0);                         /* :1704-1717 */
  s->box_x.x *= ttt; s->box_x.y *= ttt; s->box_y.x *= ttt; s->box_y.y *= ttt;
  s->box_x.z *= ttt; s->box_y.z *= ttt; s->box_z.x *= ttt; s->box_z.y *= ttt; s->box_z.z *= ttt;
  make_box(s);
  s->pressure_ext += s->d_pressure;                                                  /* :1727 */
}

/* move_atoms_npt_axial, src/imd_integrate.c:1747-1959 */
static void move_atoms_npt_axial(orc_sim *s, int do_press_calc)
{
  int c, i, d; const double dt = s->timestep;
  double Ekin_new = 0.0, pfric[3], pifric[3], rfric[3], rifric[3], xi_old[3], ttt, tvec[3], dyn[3] = {0.0, 0.0, 0.0};
  for (d = 0; d < 3; d++) s->stress3[d] = (s->dyn3[d] + s->vir[d]) / s->volume;      /* :1775-1779 */
  ttt = dt * s->volume * s->isq_tau_xi / s->nactive;                                 /* :1782 */
  for (d = 0; d < 3; d++) {
    xi_old[d] = s->xi3[d]; s->xi3[d] += ttt * (s->stress3[d] - s->pext3[d]) * s->relax_dirs[d];   /* :1783-1787 */
    pfric[d]  =        1.0 - (xi_old[d] + s->eta) * dt / 2.0;                        /* :1790-1803 */
    pifric[d] = 1.0 / (1.0 + (s->xi3[d] + s->eta) * dt / 2.0);
    rfric[d]  =        1.0 + (s->xi3[d]         ) * dt / 2.0;
    rifric[d] = 1.0 / (1.0 - (s->xi3[d]         ) * dt / 2.0);
  }
  for (c = 0; c < s->ncells; c++) {
    cellist *p = &s->cells[s->cnp[c]];
    for (i = 0; i < p->n; i++) {
      long a = p->idx[i];
      double *P = s->impuls + 3 * a, *F = s->kraft + 3 * a, *X = s->ort + 3 * a, tmp = 1.0 / s->masse[a];
      const double *R = s->restr + 3 * s->vsorte[a];
      if (do_press_calc) {                                                           /* :1834-1845: before the kick */
        double *S = s->presstens + 6 * a;
        S[0] += P[0] * P[0] * tmp; S[1] += P[1] * P[1] * tmp; S[2] += P[2] * P[2] * tmp;
        S[3] += P[1] * P[2] * tmp; S[4] += P[2] * P[0] * tmp; S[5] += P[0] * P[1] * tmp;
      }
      for (d = 0; d < 3; d++) {
        P[d] = (pfric[d] * P[d] + dt * F[d]) * pifric[d];                            /* :1848-1855 */
        P[d] *= R[d];                                                                /* :1859-1864 */
      }
      dyn[0] += P[0] * P[0] * tmp; dyn[1] += P[1] * P[1] * tmp; dyn[2] += P[2] * P[2] * tmp;   /* :1868-1872 */
      Ekin_new += (((P[0] * P[0]) + (P[1] * P[1])) + (P[2] * P[2])) * tmp;           /* :1875 */
      tmp *= dt;
      for (d = 0; d < 3; d++) X[d] = (rfric[d] * X[d] + P[d] * tmp) * rifric[d];     /* :1878-1883 */
    }
  }
  for (d = 0; d < 3; d++) s->dyn3[d] = dyn[d];
  s->tot_kin_energy = (s->Ekin_old + Ekin_new) / 4.0;                                /* :1917 */
  ttt = s->nactive * s->temperature;
  s->eta += dt * (Ekin_new / ttt - 1.0) * s->isq_tau_eta;
  s->Ekin_old = Ekin_new;
  for (d = 0; d < 3; d++) tvec[d] = (1.0 + s->xi3[d] * dt / 2.0) / (1.0 - s->xi3[d] * dt / 2.0);   /* :1923-1937 */
  s->box_x.x *= tvec[0]; s->box_x.y *= tvec[0]; s->box_y.x *= tvec[1]; s->box_y.y *= tvec[1];
  s->box_x.z *= tvec[0]; s->box_y.z *= tvec[1]; s->box_z.x *= tvec[2]; s->box_z.y *= tvec[2]; s->box_z.z *= tvec[2];
  make_box(s);
  for (d = 0; d < 3; d++) s->pext3[d] += s->dpext3[d];                               /* :1955-1959 */
}

/* NPT_axial hand-over; Ekin_old < 0: the next move_atoms starts like steps == steps_min (calc_dyn_pressure, xi *= relax_dirs) */
void orc_set_npt_axial(orc_sim *s, const double *xi3, const double *pext3, const double *dpext3, const int *relax_dirs,
                       double Ekin_old, const double *dyn3, double isq_tau_xi)
{
  int d;
  for (d = 0; d < 3; d++) { s->xi3[d] = xi3[d]; s->pext3[d] = pext3[d]; s->dpext3[d] = dpext3[d]; s->relax_dirs[d] = relax_dirs[d];
                            s->dyn3[d] = dyn3 ? dyn3[d] : 0.0; }
  s->Ekin_old = Ekin_old; s->isq_tau_xi = isq_tau_xi;
}

/* out13 = xi[3], stress[3] of the last step, pressure_ext[3], dyn_stress[3], Ekin_old */
void orc_get_npt_axial(const orc_sim *s, double *out13)
{
  int d;
  for (d = 0; d < 3; d++) { out13[d] = s->xi3[d]; out13[3 + d] = s->stress3[d]; out13[6 + d] = s->pext3[d]; out13[9 + d] = s->dyn3[d]; }
  out13[12] = s->Ekin_old;
}

/* NPT_iso state: after a restart-like hand-over all of it comes from the caller; a fresh run calls it with
   xi = 0 and Ekin_old < 0, which makes the first step compute calc_dyn_pressure() like steps == steps_min does */
/* Berendsen variant of NVE; tot_kin_energy is the value the previous move_atoms left (0 at the start of a run) */
void orc_set_berendsen(orc_sim *s, double tauber, double tot_kin_energy)
{
  s->tauber = tauber; s->tot_kin_energy = tot_kin_energy;
}

void orc_set_npt(orc_sim *s, double xi, double Ekin_old, double pressure_ext, double d_pressure, double isq_tau_xi)
{
  s->xi = xi; s->Ekin_old = Ekin_old; s->pressure_ext = pressure_ext; s->d_pressure = d_pressure; s->isq_tau_xi = isq_tau_xi;
}

void orc_get_npt(const orc_sim *s, double out[4])
{
  out[0] = s->xi; out[1] = s->Ekin_old; out[2] = s->pressure; out[3] = s->pressure_ext;
}

/* check_nblist, src/imd_forces_nbl.c:2007-2037 */
void orc_check_nblist(orc_sim *s)
{
  long k; double max1 = 0.0;
  for (k = 0; k < s->n; k++) {
    double dx = s->ort[3 * k] - s->nblpos[3 * k], dy = s->ort[3 * k + 1] - s->nblpos[3 * k + 1],
           dz = s->ort[3 * k + 2] - s->nblpos[3 * k + 2];
    double r2 = ((dx * dx) + (dy * dy)) + (dz * dz);
    if (r2 > max1) max1 = r2;
  }
  if (max1 > SQR(0.5 * s->nbl_margin)) s->have_valid_nbl = 0;
}

/* main_loop body, src/imd_main_3d.c:405, 559, 768-772 */
void orc_step(orc_sim *s, int nsteps)
{
  int k;
  for (k = 0; k < nsteps; k++) { orc_calc_forces(s, 0); orc_move_atoms(s, 0); orc_check_nblist(s); }
}

/* lin_deform, src/imd_deform.c:35-119 */
void orc_lin_deform(orc_sim *s, const double dx[3], const double dy[3], const double dz[3], double scale)
{
  long a; double t[3];
  vec3 vx = {dx[0], dx[1], dx[2]}, vy = {dy[0], dy[1], dy[2]}, vz = {dz[0], dz[1], dz[2]};
  vec3 *b[3]; int k;
  for (a = 0; a < s->n; a++) {
    double *o = s->ort + 3 * a;
    t[0] = dx[0] * o[0] + dx[1] * o[1] + dx[2] * o[2];
    t[1] = dy[0] * o[0] + dy[1] * o[1] + dy[2] * o[2];
    t[2] = dz[0] * o[0] + dz[1] * o[1] + dz[2] * o[2];
    o[0] += scale * t[0]; o[1] += scale * t[1]; o[2] += scale * t[2];
  }
  b[0] = &s->box_x; b[1] = &s->box_y; b[2] = &s->box_z;
  for (k = 0; k < 3; k++) {
    t[0] = scale * SPROD(vx, *b[k]); t[1] = scale * SPROD(vy, *b[k]); t[2] = scale * SPROD(vz, *b[k]);
    b[k]->x += t[0]; b[k]->y += t[1]; b[k]->z += t[2];
  }
  make_box(s);
}

/* deform_sample, src/imd_deform.c:232-269: per virtual type a shift, optionally scaled by a shear profile */
void orc_deform_sample(orc_sim *s, double deform_size, const double *deform_shift, const int *shear_def,
                       const double *deform_shear, const double *deform_base)
{
  long a;
  for (a = 0; a < s->n; a++) {
    double *o = s->ort + 3 * a, shear;
    int sort = s->vsorte[a];
    if (shear_def && shear_def[sort] == 1) {
      vec3 ort, sh = {deform_shear[3 * sort], deform_shear[3 * sort + 1], deform_shear[3 * sort + 2]};
      ort.x = o[0] - deform_base[3 * sort];
      ort.y = o[1] - deform_base[3 * sort + 1];
      ort.z = o[2] - deform_base[3 * sort + 2];
      shear = SPROD(sh, ort);
    } else shear = 1.0;
    o[0] += shear * deform_size * deform_shift[3 * sort];
    o[1] += shear * deform_size * deform_shift[3 * sort + 1];
    o[2] += shear * deform_size * deform_shift[3 * sort + 2];
  }
}

/* nactive = number of degrees of freedom that move: 3 per atom of a real type, the restriction components of
   its virtual type otherwise (read_atoms, src/imd_io_3d.c:469-481; generate_atoms, src/imd_generate.c:450-451) */
static void count_nactive(orc_sim *s)
{
  long a, n = 0;
  for (a = 0; a < s->n; a++) {
    int v = s->vsorte[a];
    if (v < s->ntypes || !s->restr || v >= s->nvtypes) n += 3;
    else n += (long) s->restr[3 * v] + (long) s->restr[3 * v + 1] + (long) s->restr[3 * v + 2];
  }
  s->nactive = n;
}

/* ------------------------------------------------------------------------------------ */
/* construction / accessors                                                              */
/* ------------------------------------------------------------------------------------ */

orc_sim *orc_create(int ntypes, const double box[9], const int pbc[3], double nbl_margin)
{
  orc_sim *s = (orc_sim *) calloc(1, sizeof(orc_sim));
  s->ntypes = ntypes;
  s->box_x.x = box[0]; s->box_x.y = box[1]; s->box_x.z = box[2];
  s->box_y.x = box[3]; s->box_y.y = box[4]; s->box_y.z = box[5];
  s->box_z.x = box[6]; s->box_z.y = box[7]; s->box_z.z = box[8];
  s->pbc[0] = pbc[0]; s->pbc[1] = pbc[1]; s->pbc[2] = pbc[2];
  s->nbl_margin = nbl_margin;
  s->default_fmt = 1;
  s->ensemble = ORC_NVE; s->timestep = 0.0;
  s->nvtypes = ntypes;
  s->restr = (double *) malloc(sizeof(double) * 3 * ntypes);
  { int i; for (i = 0; i < 3 * ntypes; i++) s->restr[i] = 1.0; }
  return s;
}

void orc_destroy(orc_sim *s)
{
  int w;
  if (!s) return;
  for (w = 0; w < 6; w++) {
    free(s->tab[w].begin); free(s->tab[w].end); free(s->tab[w].step); free(s->tab[w].invstep);
    free(s->tab[w].len); free(s->tab[w].table); free(s->tab[w].table2);
  }
  free_cells(s);
  free(s->nummer); free(s->sorte); free(s->vsorte); free(s->masse); free(s->ort); free(s->impuls);
  free(s->kraft); free(s->poteng); free(s->rho); free(s->dF); free(s->presstens); free(s->nblpos);
  free(s->gsrc); free(s->gshift); free(s->tl); free(s->tb); free(s->cl_off); free(s->cl_num); free(s->restr);
  free(s);
}

void orc_set_atoms(orc_sim *s, long n, const int *nummer, const int *sorte, const int *vsorte,
                   const double *masse, const double *ort, const double *impuls)
{
  long a;
  grow_atoms(s, n + 1);
  s->n = n; s->ng = 0;
  for (a = 0; a < n; a++) {
    s->nummer[a] = nummer[a]; s->sorte[a] = sorte[a]; s->vsorte[a] = vsorte ? vsorte[a] : sorte[a];
    s->masse[a] = masse[a];
    memcpy(s->ort + 3 * a, ort + 3 * a, 3 * sizeof(double));
    if (impuls) memcpy(s->impuls + 3 * a, impuls + 3 * a, 3 * sizeof(double));
    else memset(s->impuls + 3 * a, 0, 3 * sizeof(double));
    s->kraft[3 * a] = s->kraft[3 * a + 1] = s->kraft[3 * a + 2] = 0.0;
    s->poteng[a] = s->rho[a] = s->dF[a] = 0.0;
  }
  count_nactive(s);
  if (!s->cells) make_box(s); /* first make_box -> init_cells (min/max_height start at 0) */
  s->have_valid_nbl = 0;
}

void orc_set_restrictions(orc_sim *s, int nvtypes, const double *restr3)
{
  s->nvtypes = nvtypes;
  s->restr = (double *) realloc(s->restr, sizeof(double) * 3 * nvtypes);
  memcpy(s->restr, restr3, sizeof(double) * 3 * nvtypes);
  count_nactive(s);
}

void orc_set_integrator(orc_sim *s, int ensemble, double timestep, double temperature, double eta, double isq_tau_eta)
{
  s->ensemble = ensemble; s->timestep = timestep; s->temperature = temperature;
  s->eta = eta; s->isq_tau_eta = isq_tau_eta;
}

void orc_set_box(orc_sim *s, const double box[9])
{
  s->box_x.x = box[0]; s->box_x.y = box[1]; s->box_x.z = box[2];
  s->box_y.x = box[3]; s->box_y.y = box[4]; s->box_y.z = box[5];
  s->box_z.x = box[6]; s->box_z.y = box[7]; s->box_z.z = box[8];
  make_box(s);
}

long orc_natoms(const orc_sim *s) { return s->n; }
int orc_have_valid_nbl(const orc_sim *s) { return s->have_valid_nbl; }
int orc_nbl_count(const orc_sim *s) { return s->nbl_count; }
double orc_cellsz(const orc_sim *s) { return s->cellsz; }
void orc_get_celldims(const orc_sim *s, int out6[6])
{
  int d; for (d = 0; d < 3; d++) { out6[d] = s->gdim[d]; out6[3 + d] = s->cdim[d]; }
}
void orc_get_scalars(const orc_sim *s, double out[14])
{
  int i; for (i = 0; i < 14; i++) out[i] = 0.0;
  out[0] = s->tot_pot_energy; out[1] = s->tot_kin_energy; out[2] = s->virial;
  out[3] = s->vir[0]; out[4] = s->vir[1]; out[5] = s->vir[2];
  out[9] = s->volume; out[10] = (double) s->nactive; out[11] = s->eta;
  out[12] = s->temperature; out[13] = s->timestep;
}
void orc_get_box(const orc_sim *s, double o[9])
{
  o[0] = s->box_x.x; o[1] = s->box_x.y; o[2] = s->box_x.z;
  o[3] = s->box_y.x; o[4] = s->box_y.y; o[5] = s->box_y.z;
  o[6] = s->box_z.x; o[7] = s->box_z.y; o[8] = s->box_z.z;
}

long orc_get_adp(const orc_sim *s, double *mu3, double *la6)
{
  if (mu3) memcpy(mu3, s->mu, sizeof(double) * 3 * s->n);
  if (la6) memcpy(la6, s->la, sizeof(double) * 6 * s->n);
  return s->n;
}

long orc_get_eeam(const orc_sim *s, double *eam_p, double *dM)
{
  long n;
  for (n = 0; n < s->n; n++) { if (eam_p) eam_p[n] = s->eam_p[n]; if (dM) dM[n] = s->dM[n]; }
  return s->n;
}

long orc_get_atoms(const orc_sim *s, int *nummer, int *sorte, int *vsorte, double *masse, double *ort,
                   double *impuls, double *kraft, double *poteng, double *rho, double *dF,
                   double *presstens, double *nblpos)
{
  long n; 
  for (n = 0; n < s->n; n++) {
    long a = n; /* storage order; callers sort by nummer */
    if (nummer) nummer[n] = s->nummer[a];
    if (sorte) sorte[n] = s->sorte[a];
    if (vsorte) vsorte[n] = s->vsorte[a];
    if (masse) masse[n] = s->masse[a];
    if (ort) memcpy(ort + 3 * n, s->ort + 3 * a, 24);
    if (impuls) memcpy(impuls + 3 * n, s->impuls + 3 * a, 24);
    if (kraft) memcpy(kraft + 3 * n, s->kraft + 3 * a, 24);
    if (poteng) poteng[n] = s->poteng[a];
    if (rho) rho[n] = s->rho[a];
    if (dF) dF[n] = s->dF[a];
    if (presstens) memcpy(presstens + 6 * n, s->presstens + 6 * a, 48);
    if (nblpos) memcpy(nblpos + 3 * n, s->nblpos + 3 * a, 24);
  }
  return s->n;
}

long orc_get_nbl_pairs(const orc_sim *s, int *pi, int *pj, signed char *shift, long cap)
{
  long n = 0, cnt = 0, m; int c, i;
  if (!s->have_valid_nbl || !s->tl) return -1;
  for (c = 0; c < s->ncells; c++) {
    const cellist *p = &s->cells[s->cnp[c]];
    for (i = 0; i < p->n; i++, n++)
      for (m = s->tl[n]; m < s->tl[n + 1]; m++) {
        int slot = s->tb[m], cc = s->cl_num[slot];
        long ja = s->cells[cc].idx[slot - s->cl_off[cc]];
        if (cnt < cap) {
          pi[cnt] = s->nummer[p->idx[i]];
          pj[cnt] = s->nummer[ja];
          if (shift) {
            if (ja >= s->n) memcpy(shift + 3 * cnt, s->gshift + 3 * ja, 3);
            else memset(shift + 3 * cnt, 0, 3);
          }
        }
        cnt++;
      }
  }
  return cnt;
}

/* calc_tot_presstens, src/imd_main_3d.c:2069-2130 */
void orc_tot_presstens(const orc_sim *s, double out6[6])
{
  long a; int d;
  for (d = 0; d < 6; d++) out6[d] = 0.0;
  for (a = 0; a < s->n; a++) for (d = 0; d < 6; d++) out6[d] += s->presstens[6 * a + d];
}
