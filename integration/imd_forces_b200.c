/* imd_forces_b200.c -- the reference-side binding of libimd_b200.so.
 *
 * Drop this file into IMD's src/ and select it as the force engine exactly like the Cell-BE engine is
 * selected (FORCESOURCES = imd_forces_cbe.c, src/Makefile:1141-1143): it takes the place of
 * imd_forces_nbl.c and re-defines the symbols IMD's step loop calls --
 *     calc_forces, make_nblist, check_nblist, estimate_nblist_size, deallocate_nblist
 * (src/prototypes.h:173, 195-198) -- on top of the C ABI in include/imd_b200.h.  The integrator is a
 * function pointer (src/globals.h:1444) set from the `ensemble` keyword; calc_forces re-points it to the
 * wrapper below on its first call, so no reference source file is edited.
 *
 * Everything else stays IMD's: main(), the parameter file, setup_potentials(), generate_atoms() /
 * read_atoms(), maxwell(), the .eng / checkpoint writers.  IMD's per-cell arrays remain the canonical host
 * copy: forces, energies, positions and momenta are written back into them (matched by NUMMER) every
 * IMD_B200_SYNC steps (environment variable, default 1 = every step, 0 = only when a writer is due), so
 * that the unmodified writers see current data.
 *
 * Builds with homdef / deform in the target name: IMD's lin_deform and deform_sample (src/imd_deform.c:35-119, 232-269)
 * edit the host cells, which the device never sees.  They are not in the force-engine file, so they cannot be replaced
 * by a second definition; the binding is linked with  -Wl,--wrap=lin_deform,--wrap=deform_sample  and the wrappers
 * below run IMD's own routine (box, host copy) and then the device's (imdb200_lin_deform / imdb200_deform_sample).
 * stress builds: per-atom PRESSTENS is downloaded whenever do_press_calc is set, so that calc_tot_presstens
 * (src/imd_main_3d.c:2069-2130) and the Press_xx.. columns of the .eng file (src/imd_io.c:2474-2480) work unchanged.
 * npt builds: ensembles npt_iso and npt_axial run on the device (xi, pressure_ext, Ekin_old, the per-axis stress and the
 * breathing box are mirrored back).
 *
 * MPI builds (imd_mpi_*, one rank per GPU): cpu_dim / my_coord are IMD's own (setup_mpi_topology,
 * src/imd_geom_mpi_3d.c:32-90), rank 0 creates the ncclUniqueId and MPI_Bcast carries it, every rank uploads the atoms
 * IMD distributed to it.  Halo exchange, atom migration and the reductions then run inside the library; because atoms
 * change ranks there, the host cells are refilled from what the device reports at every download (b200_rebin) instead
 * of being matched by NUMMER.
 *
 * Build (oracle/Makefile, target _ref/imd_b200_dropin):
 *   gcc -DNVE -DNVT -DEAM2 -DNBL ... <IMD core sources> integration/imd_forces_b200.c \
 *       -Iinclude -Limd_b200 -limd_b200 -Wl,-rpath,'$ORIGIN/../../imd_b200' -lm
 */
#include "imd.h"
#include "imd_b200.h"

int *tl = NULL, *tb = NULL, *cl_off = NULL, *cl_num = NULL;   /* kept for symbol compatibility */

static imdb200_sim *b200 = NULL;
static void (*imd_move_atoms)(void) = NULL;   /* what the ensemble keyword selected */
static long   b200_n = 0;
static int   *b_num = NULL, *b_sorte = NULL, *b_vsorte = NULL, *b_cell = NULL, *b_slot = NULL, *num2idx = NULL;
static double *b_masse = NULL, *b_ort = NULL, *b_impuls = NULL, *b_kraft = NULL, *b_poteng = NULL, *b_rho = NULL;
#ifdef STRESS_TENS
static double *b_press = NULL;
#endif
static int    num_max = 0, sync_int = 1;

static void b200_fatal(const char *msg) { error((char *) msg); }   /* IMD's abort convention, src/imd_misc.c:78 */

static void b200_check(int rc) { if (rc) error((char *) imdb200_last_error()); }

static void *b200_realloc(void *p, size_t bytes)
{
  void *q = realloc(p, bytes ? bytes : 1);
  if (q == NULL) error("imd_b200: out of host memory");
  return q;
}

/* is one of IMD's writers due at this step?  (src/imd_main_3d.c:690-694) */
static int output_due(int step)
{
  if (step >= steps_max) return 1;
  if (checkpt_int > 0 && step % checkpt_int == 0) return 1;
#ifdef FORCE
  if (force_int > 0 && step % force_int == 0) return 1;
#endif
#ifdef FNORM   /* deform / relaxation builds print sqrt(fnorm/nactive) and sqrt(f_max2) in every .eng line (src/imd_io.c:2413-2414) */
  if (eng_int > 0 && step % eng_int == 0) return 1;
#endif
  if (dist_int > 0 && step % dist_int == 0) return 1;
  if (pic_int > 0 && step % pic_int == 0) return 1;
  return 0;
}

static int sync_due(int step) { return (sync_int > 0 && step % sync_int == 0) || output_due(step); }

/* gather IMD's per-cell arrays into flat ones and hand them to the GPU */
static void b200_upload(void)
{
  long n = 0; int k, i;
  for (k = 0; k < NCELLS; k++) n += CELLPTR(k)->n;
  if (n != b200_n) {
    b200_n = n;
    b_num = b200_realloc(b_num, n * sizeof(int)); b_sorte = b200_realloc(b_sorte, n * sizeof(int));
    b_vsorte = b200_realloc(b_vsorte, n * sizeof(int)); b_cell = b200_realloc(b_cell, n * sizeof(int));
    b_slot = b200_realloc(b_slot, n * sizeof(int));
    b_masse = b200_realloc(b_masse, n * sizeof(double)); b_ort = b200_realloc(b_ort, 3 * n * sizeof(double));
    b_impuls = b200_realloc(b_impuls, 3 * n * sizeof(double)); b_kraft = b200_realloc(b_kraft, 3 * n * sizeof(double));
    b_poteng = b200_realloc(b_poteng, n * sizeof(double)); b_rho = b200_realloc(b_rho, n * sizeof(double));
#ifdef STRESS_TENS
    b_press = b200_realloc(b_press, 6 * n * sizeof(double));
#endif
  }
  n = 0; num_max = 0;
  for (k = 0; k < NCELLS; k++) {
    cell *p = CELLPTR(k);
    for (i = 0; i < p->n; i++, n++) {
      b_num[n] = NUMMER(p,i); b_sorte[n] = SORTE(p,i); b_vsorte[n] = VSORTE(p,i); b_masse[n] = MASSE(p,i);
      b_ort[3*n] = ORT(p,i,X); b_ort[3*n+1] = ORT(p,i,Y); b_ort[3*n+2] = ORT(p,i,Z);
      b_impuls[3*n] = IMPULS(p,i,X); b_impuls[3*n+1] = IMPULS(p,i,Y); b_impuls[3*n+2] = IMPULS(p,i,Z);
      b_cell[n] = k; b_slot[n] = i;
      if (b_num[n] > num_max) num_max = b_num[n];
    }
  }
  num2idx = b200_realloc(num2idx, ((size_t) num_max + 1) * sizeof(int));
  for (n = 0; n < b200_n; n++) num2idx[b_num[n]] = (int) n;
  b200_check(imdb200_set_atoms(b200, b200_n, b_num, b_sorte, b_vsorte, b_masse, b_ort, b_impuls));
}

#ifdef MPI
/* MPI builds: atoms migrate between GPUs inside the library, so the set of atoms of this rank changes.  The host cells
   are emptied and refilled from the device's arrays; every atom goes into the cell IMD's own cell_coord /
   local_cell_coord (src/imd_geom_3d.c:1054-1074, src/imd_geom_mpi_3d.c:119-128) assign it, clamped into the real cells
   (positions are only wrapped at list builds, an atom may sit a skin's width outside its domain). */
static int *r_num = NULL, *r_sorte = NULL, *r_vsorte = NULL;
static double *r_masse = NULL, *r_ort = NULL, *r_impuls = NULL, *r_kraft = NULL, *r_poteng = NULL, *r_rho = NULL, *r_press = NULL;
static long r_cap = 0;

static void b200_rebin(int press)
{
  long n = imdb200_natoms_local(b200), a, got; int k;
  if (n > r_cap) {
    r_cap = n + n / 8 + 64;
    r_num = b200_realloc(r_num, r_cap * sizeof(int)); r_sorte = b200_realloc(r_sorte, r_cap * sizeof(int));
    r_vsorte = b200_realloc(r_vsorte, r_cap * sizeof(int)); r_masse = b200_realloc(r_masse, r_cap * sizeof(double));
    r_ort = b200_realloc(r_ort, 3 * r_cap * sizeof(double)); r_impuls = b200_realloc(r_impuls, 3 * r_cap * sizeof(double));
    r_kraft = b200_realloc(r_kraft, 3 * r_cap * sizeof(double)); r_poteng = b200_realloc(r_poteng, r_cap * sizeof(double));
    r_rho = b200_realloc(r_rho, r_cap * sizeof(double)); r_press = b200_realloc(r_press, 6 * r_cap * sizeof(double));
  }
  got = imdb200_get_atoms(b200, r_num, r_sorte, r_vsorte, r_masse, r_ort, r_impuls, r_kraft, r_poteng, r_rho, NULL,
                          press ? r_press : NULL, NULL);
  if (got != n) error(got < 0 ? (char *) imdb200_last_error() : "imd_b200: atom count changed during the download");
  for (k = 0; k < nallcells; k++) (cell_array + k)->n = 0;
  for (a = 0; a < n; a++) {
    ivektor c = local_cell_coord(cell_coord(r_ort[3*a], r_ort[3*a+1], r_ort[3*a+2]));
    cell *p; int i;
    c.x = c.x < 1 ? 1 : (c.x > cell_dim.x - 2 ? cell_dim.x - 2 : c.x);
    c.y = c.y < 1 ? 1 : (c.y > cell_dim.y - 2 ? cell_dim.y - 2 : c.y);
    c.z = c.z < 1 ? 1 : (c.z > cell_dim.z - 2 ? cell_dim.z - 2 : c.z);
    p = PTR_VV(cell_array, c, cell_dim);
    if (p->n >= p->n_max) alloc_cell(p, p->n_max + incrsz);
    i = p->n++;
    NUMMER(p,i) = r_num[a]; SORTE(p,i) = r_sorte[a]; VSORTE(p,i) = r_vsorte[a]; MASSE(p,i) = r_masse[a];
    ORT(p,i,X) = r_ort[3*a]; ORT(p,i,Y) = r_ort[3*a+1]; ORT(p,i,Z) = r_ort[3*a+2];
    IMPULS(p,i,X) = r_impuls[3*a]; IMPULS(p,i,Y) = r_impuls[3*a+1]; IMPULS(p,i,Z) = r_impuls[3*a+2];
    KRAFT(p,i,X) = r_kraft[3*a]; KRAFT(p,i,Y) = r_kraft[3*a+1]; KRAFT(p,i,Z) = r_kraft[3*a+2];
    POTENG(p,i) = r_poteng[a];
#ifdef EAM2
    EAM_RHO(p,i) = r_rho[a];
#endif
#ifdef STRESS_TENS
    if (press) {
      PRESSTENS(p,i,xx) = r_press[6*a]; PRESSTENS(p,i,yy) = r_press[6*a+1]; PRESSTENS(p,i,zz) = r_press[6*a+2];
      PRESSTENS(p,i,yz) = r_press[6*a+3]; PRESSTENS(p,i,zx) = r_press[6*a+4]; PRESSTENS(p,i,xy) = r_press[6*a+5];
    }
#endif
  }
}
#endif

/* write device results back into IMD's cells; the device order is cell-sorted, atoms are matched by NUMMER */
static void b200_download(int forces, int state, int press)
{
  long n, got; 
  static int *d_num = NULL; static long d_cap = 0;
  double *pt = NULL;
#ifdef MPI
  (void) n; (void) got; (void) d_num; (void) d_cap; (void) pt; (void) forces; (void) state;
  b200_rebin(press);
  return;
#endif
  if (d_cap < b200_n) { d_cap = b200_n; d_num = b200_realloc(d_num, d_cap * sizeof(int)); }
#ifdef STRESS_TENS
  if (press) pt = b_press;
#endif
  got = imdb200_get_atoms(b200, d_num, NULL, NULL, NULL, state ? b_ort : NULL, state ? b_impuls : NULL,
                          forces ? b_kraft : NULL, forces ? b_poteng : NULL,
#ifdef EAM2
                          forces ? b_rho : NULL,
#else
                          NULL,
#endif
                          NULL, pt, NULL);
  if (got < 0) error((char *) imdb200_last_error());
  if (got != b200_n) error("imd_b200: atom count changed on the device");
  for (n = 0; n < got; n++) {
    int a = num2idx[d_num[n]];
    cell *p = CELLPTR(b_cell[a]); int i = b_slot[a];
    if (state) {
      ORT(p,i,X) = b_ort[3*n]; ORT(p,i,Y) = b_ort[3*n+1]; ORT(p,i,Z) = b_ort[3*n+2];
      IMPULS(p,i,X) = b_impuls[3*n]; IMPULS(p,i,Y) = b_impuls[3*n+1]; IMPULS(p,i,Z) = b_impuls[3*n+2];
    }
    if (forces) {
      KRAFT(p,i,X) = b_kraft[3*n]; KRAFT(p,i,Y) = b_kraft[3*n+1]; KRAFT(p,i,Z) = b_kraft[3*n+2];
      POTENG(p,i) = b_poteng[n];
#ifdef EAM2
      EAM_RHO(p,i) = b_rho[n];
#endif
    }
#ifdef STRESS_TENS
    if (pt) {      /* device order: xx yy zz yz zx xy (include/imd_b200.h) */
      PRESSTENS(p,i,xx) = pt[6*n]; PRESSTENS(p,i,yy) = pt[6*n+1]; PRESSTENS(p,i,zz) = pt[6*n+2];
      PRESSTENS(p,i,yz) = pt[6*n+3]; PRESSTENS(p,i,zx) = pt[6*n+4]; PRESSTENS(p,i,xy) = pt[6*n+5];
    }
#endif
  }
}

static void b200_scalars(int forces, int kinetic)
{
  imdb200_scalars sc;
  b200_check(imdb200_get_scalars(b200, &sc));
  if (forces) {
    tot_pot_energy = sc.tot_pot_energy; virial = sc.virial;
#ifdef P_AXIAL   /* these builds accumulate vir_xx/yy/zz instead and leave `virial` at 0 (src/imd_forces_nbl.c:548-556); the
                    per-axis sums stay in the library, IMD reads them back as stress_x/y/z */
    virial = 0.0;
#endif
  }
  if (kinetic) { tot_kin_energy = sc.tot_kin_energy; eta = sc.eta; }
  have_valid_nbl = sc.have_valid_nbl;
  nbl_count = sc.nbl_count;
  last_nbl_len = (int) (sc.nbl_len / 2);        /* the reference counts each pair once */
}

/* Where every atom sits in IMD's cells (b_cell, b_slot, by upload index): make_box() re-plans the host cells once the box
   has changed enough (init_cells moves the atoms into a new cell array, src/imd_geom_3d.c:353-398), and the map the
   downloads write through has to follow. */
static void b200_remap(void)
{
#ifndef MPI
  int k, i;
  for (k = 0; k < NCELLS; k++) {
    cell *p = CELLPTR(k);
    for (i = 0; i < p->n; i++) { const int a = num2idx[NUMMER(p,i)]; b_cell[a] = k; b_slot[a] = i; }
  }
#endif
}

/* the device's box -> IMD's box_x/y/z (+ make_box for tbox, volume, heights) */
static void b200_mirror_box(void)
{
  double b[9];
  const void *cells_before = (const void *) cell_array;
  const ivektor dim_before = cell_dim;
  b200_check(imdb200_get_box(b200, b));
  box_x.x = b[0]; box_x.y = b[1]; box_x.z = b[2];
  box_y.x = b[3]; box_y.y = b[4]; box_y.z = b[5];
  box_z.x = b[6]; box_z.y = b[7]; box_z.z = b[8];
  make_box();
  if ((const void *) cell_array != cells_before || cell_dim.x != dim_before.x || cell_dim.y != dim_before.y ||
      cell_dim.z != dim_before.z) b200_remap();
}

#ifdef NPT_iso
/* what move_atoms_npt_iso leaves in IMD's globals (src/imd_integrate.c:1509, 1691-1727) */
static void b200_mirror_npt(void)
{
  double st[4];
  b200_check(imdb200_get_npt_state(b200, st));
  xi.x = st[0]; Ekin_old = st[1]; pressure = st[2]; pressure_ext.x = st[3];
  b200_mirror_box();
}
#endif

#ifdef NPT_axial
/* what move_atoms_npt_axial leaves in IMD's globals (src/imd_integrate.c:1775-1787, 1917-1959) */
static void b200_mirror_npt_axial(void)
{
  double st[13];
  b200_check(imdb200_get_npt_axial(b200, st));
  xi.x = st[0]; xi.y = st[1]; xi.z = st[2];
  stress_x = st[3]; stress_y = st[4]; stress_z = st[5];
  pressure_ext.x = st[6]; pressure_ext.y = st[7]; pressure_ext.z = st[8];
  dyn_stress_x = st[9]; dyn_stress_y = st[10]; dyn_stress_z = st[11];
  Ekin_old = st[12];
  b200_mirror_box();
}
#endif

/* (*move_atoms)() = move_atoms_nve / move_atoms_nvt (src/imd_integrate.c:32, 891) */
static void b200_move_atoms(void)
{
  int press = 0;
#ifdef FNORM
  /* fnorm = sum F^2 and f_max2 = largest squared force component, of the restricted forces, as move_atoms_nve/nvt form
     them (src/imd_integrate.c:192-205, 1005-1013); from the forces calc_forces has just written into the cells */
  if (sync_due(steps)) {
    int k, i; real f2 = 0.0, m2 = 0.0;
    for (k = 0; k < NCELLS; k++) {
      cell *p = CELLPTR(k);
      for (i = 0; i < p->n; i++) {
        vektor *r = restrictions + VSORTE(p,i);
        real fx = KRAFT(p,i,X) * r->x, fy = KRAFT(p,i,Y) * r->y, fz = KRAFT(p,i,Z) * r->z;
        f2 += fx * fx + fy * fy + fz * fz;
        m2 = MAX(m2, MAX(fx * fx, MAX(fy * fy, fz * fz)));
      }
    }
    fnorm = f2; f_max2 = m2;
  }
#endif
  b200_check(imdb200_move_atoms(b200));
  b200_scalars(0, 1);
#ifdef NPT_iso
  if (ensemble == ENS_NPT_ISO) b200_mirror_npt();
#endif
#ifdef NPT_axial
  if (ensemble == ENS_NPT_AXIAL) b200_mirror_npt_axial();
#endif
#ifdef STRESS_TENS
  press = do_press_calc;                         /* kinetic part added by move_atoms: the tensor is complete now */
#endif
  if (sync_due(steps) || press) b200_download(0, sync_due(steps), press);   /* writers run after move_atoms of the same step (src/imd_main_3d.c:690-694) */
#ifdef AND
  { /* Andersen thermostat: every tempintv-th call of move_atoms_nve ends in maxwell(temperature) (src/imd_integrate.c:491-495);
       the host keeps maxwell, the new momenta go to the device by atom number */
    static int and_count = 0;
    ++and_count;
    if (ensemble == ENS_NVE && tempintv != 0 && 0 == and_count % tempintv) {
      long n = 0; int k, i;
      if (!sync_due(steps)) b200_download(0, 1, 0);
      maxwell(temperature);
      for (k = 0; k < NCELLS; k++) {
        cell *p = CELLPTR(k);
        for (i = 0; i < p->n; i++, n++) {
          b_num[n] = NUMMER(p,i);
          b_impuls[3*n] = IMPULS(p,i,X); b_impuls[3*n+1] = IMPULS(p,i,Y); b_impuls[3*n+2] = IMPULS(p,i,Z);
        }
      }
      b200_check(imdb200_set_momenta(b200, n, b_num, b_impuls));
      for (n = 0; n < b200_n; n++) num2idx[b_num[n]] = (int) n;    /* b_num was re-filled in today's cell order */
      b200_remap();
    }
  }
#endif
}

#ifdef IMD_B200_BACKTRACE   /* debugging aid of this binding (not part of IMD): where did a SIGSEGV come from */
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void b200_segv(int sig)
{
  void *bt[48]; int n = backtrace(bt, 48);
  backtrace_symbols_fd(bt, n, 2);
  _exit(128 + sig);
}
#endif

static void b200_init(void)
{
  imdb200_config cfg;
  char *e = getenv("IMD_B200_SYNC");
  if (e) sync_int = atoi(e);
#ifdef AND
  sync_int = 1;                                  /* the host cells have to follow every step, see calc_forces */
#endif
#ifdef IMD_B200_BACKTRACE
  signal(SIGSEGV, b200_segv);
#endif
  imdb200_set_error_handler(b200_fatal);
  imdb200_default_config(&cfg);
  cfg.ntypes = ntypes; cfg.total_types = vtypes;
  cfg.box_x[0] = box_x.x; cfg.box_x[1] = box_x.y; cfg.box_x[2] = box_x.z;
  cfg.box_y[0] = box_y.x; cfg.box_y[1] = box_y.y; cfg.box_y[2] = box_y.z;
  cfg.box_z[0] = box_z.x; cfg.box_z[1] = box_z.y; cfg.box_z[2] = box_z.z;
  cfg.pbc_dirs[0] = pbc_dirs.x; cfg.pbc_dirs[1] = pbc_dirs.y; cfg.pbc_dirs[2] = pbc_dirs.z;
  cfg.nbl_margin = nbl_margin; cfg.nbl_size = nbl_size; cfg.timestep = timestep;
  cfg.ensemble = (ensemble == ENS_NVT) ? IMDB200_ENS_NVT : IMDB200_ENS_NVE;
#ifdef NPT_iso
  if (ensemble == ENS_NPT_ISO) {                 /* src/imd_integrate.c:1493-1496, 1720-1727 */
    if (use_curr_pressure) error("imd_b200: use_curr_pressure is not supported with ensemble npt_iso");
    cfg.ensemble = IMDB200_ENS_NPT_ISO;
    cfg.xi = xi.x; cfg.isq_tau_xi = isq_tau_xi; cfg.pressure_ext = pressure_ext.x;
    cfg.d_pressure = (pressure_end.x - pressure_ext.x) / (steps_max - steps_min);
  } else
#endif
#ifdef NPT_axial
  if (ensemble == ENS_NPT_AXIAL) {               /* the per-axis state goes in after imdb200_create, see below */
    if (use_curr_pressure) error("imd_b200: use_curr_pressure is not supported with ensemble npt_axial");
    cfg.ensemble = IMDB200_ENS_NPT_AXIAL;
    cfg.isq_tau_xi = isq_tau_xi;
  } else
#endif
  if (ensemble != ENS_NVE && ensemble != ENS_NVT) error("imd_b200 supports the ensembles nve, nvt, npt_iso and npt_axial");
  cfg.temperature = temperature; cfg.eta = eta; cfg.isq_tau_eta = isq_tau_eta;
  /* `4point` / `spline` make targets (src/Makefile:1694-1701, src/potaccess.h:24-36) */
#if defined(FOURPOINT)
  cfg.interpolation = IMDB200_INTERP_4POINT;
#elif defined(SPLINE)
  cfg.interpolation = IMDB200_INTERP_SPLINE;
#endif
#ifdef MPI
  { /* one rank per GPU: the process grid is IMD's, the device is the rank's share of this node's GPUs */
    int ndev = imdb200_device_count();
    if (ndev < 1) error("imd_b200: no CUDA device");
    cfg.cpu_dim[0] = cpu_dim.x; cfg.cpu_dim[1] = cpu_dim.y; cfg.cpu_dim[2] = cpu_dim.z;
    cfg.my_coord[0] = my_coord.x; cfg.my_coord[1] = my_coord.y; cfg.my_coord[2] = my_coord.z;
    cfg.device = myid % ndev;
  }
#endif
  b200_check(imdb200_create(&cfg, &b200));
#ifdef NPT_axial
  if (ensemble == ENS_NPT_AXIAL) {               /* src/imd_integrate.c:1756-1771, 1941-1959 */
    double x3[3], p3[3], d3[3]; int r3[3];
    x3[0] = xi.x; x3[1] = xi.y; x3[2] = xi.z;
    p3[0] = pressure_ext.x; p3[1] = pressure_ext.y; p3[2] = pressure_ext.z;
    d3[0] = (pressure_end.x - pressure_ext.x) / (steps_max - steps_min);
    d3[1] = (pressure_end.y - pressure_ext.y) / (steps_max - steps_min);
    d3[2] = (pressure_end.z - pressure_ext.z) / (steps_max - steps_min);
    r3[0] = relax_dirs.x; r3[1] = relax_dirs.y; r3[2] = relax_dirs.z;
    b200_check(imdb200_set_npt_axial(b200, x3, p3, d3, r3, -1.0, NULL));
  }
#endif
#ifdef MPI
  if (num_cpus > 1) {
    char id[128];
    if (myid == 0) b200_check(imdb200_comm_unique_id(id));
    MPI_Bcast(id, 128, MPI_CHAR, 0, MPI_COMM_WORLD);
    b200_check(imdb200_comm_init(b200, id, myid, num_cpus));
  }
#endif
  /* pot_table_t (src/types.h:416-428) and imdb200_pot_table have the same layout */
#ifdef EAM2
  b200_check(imdb200_set_potentials(b200, (imdb200_pot_table *) &pair_pot, (imdb200_pot_table *) &embed_pot,
                                    (imdb200_pot_table *) &rho_h_tab));
#ifdef EEAM   /* `eeam` make targets: energy modification term M(p), src/imd_potential.c:82-85 */
  b200_check(imdb200_set_eeam_table(b200, (imdb200_pot_table *) &emod_pot));
#endif
#ifdef ADP    /* `adp` make targets: u(r), w(r), src/imd_potential.c:87-92 */
  b200_check(imdb200_set_adp_tables(b200, (imdb200_pot_table *) &adp_upot, (imdb200_pot_table *) &adp_wpot));
#endif
#else
  b200_check(imdb200_set_potentials(b200, (imdb200_pot_table *) &pair_pot, NULL, NULL));
#endif
  b200_check(imdb200_set_restrictions(b200, vtypes, (double *) restrictions));
  b200_upload();
  imd_move_atoms = move_atoms;
  move_atoms = b200_move_atoms;
}

/* void calc_forces(int steps)  (src/imd_forces_nbl.c:281-1999) */
void calc_forces(int steps)
{
#ifdef AND
  /* Andersen builds re-draw the momenta with IMD's own maxwell() (src/imd_integrate.c:491-495), which walks the HOST cells in
     order and takes one drand48 triple per atom: the host cells have to evolve exactly as in the reference, so IMD's own
     fix_cells() runs on the (downloaded) host positions wherever the reference runs it (src/imd_forces_nbl.c:304-311). */
  if (0 == have_valid_nbl) { fix_cells(); if (b200 != NULL) b200_remap(); }
#endif
  if (b200 == NULL) b200_init();
  else if (move_atoms != b200_move_atoms) {      /* a new simulation phase re-selected the integrator */
    imd_move_atoms = move_atoms; move_atoms = b200_move_atoms;
  }
#ifdef NVT
  if (ensemble == ENS_NVT) b200_check(imdb200_set_temperature(b200, temperature));   /* increment_temperature() */
#endif
#ifdef STRESS_TENS
  b200_check(imdb200_set_press_calc(b200, do_press_calc));   /* src/imd_main_3d.c:184-190 */
#endif
  b200_check(imdb200_calc_forces(b200, steps));
  b200_scalars(1, 0);
  nfc++;
  /* positions too: lin_deform / deform_sample of this step moved them on the device after the last download, and the
     .force dump is written between calc_forces and move_atoms (src/imd_main_3d.c:426-431) */
  if (sync_due(steps)) b200_download(1, 1, 0);
}

#ifdef HOMDEF
/* void lin_deform(vektor dx, vektor dy, vektor dz, real scale)  (src/imd_deform.c:35-119), linked with --wrap */
void __real_lin_deform(vektor dx, vektor dy, vektor dz, real scale);
void __wrap_lin_deform(vektor dx, vektor dy, vektor dz, real scale)
{
  double ax[3], ay[3], az[3];
  if (b200 == NULL) { __real_lin_deform(dx, dy, dz, scale); return; }    /* before the first calc_forces: host state only */
  ax[0] = dx.x; ax[1] = dx.y; ax[2] = dx.z; ay[0] = dy.x; ay[1] = dy.y; ay[2] = dy.z; az[0] = dz.x; az[1] = dz.y; az[2] = dz.z;
  b200_check(imdb200_lin_deform(b200, ax, ay, az, scale));
  b200_mirror_box();                                                     /* IMD's box follows the device's, bit for bit */
  b200_scalars(0, 0);
}
#endif

#ifdef HOMDEF
/* void relax_pressure(void)  (src/imd_deform.c:127-219), linked with --wrap.  It ends in a call of lin_deform inside its own
   translation unit, which the linker cannot redirect (--wrap only catches references across object files), so the unmodified
   function would deform IMD's host copy and leave the device behind.  Its few lines are restated here on IMD's own globals:
   the deviation of the stress tensor (calc_tot_presstens on the per-atom tensor the engine downloads while relax_rate > 0,
   src/imd_main_3d.c:183-194) from presstens_ext, turned into a strain by the bulk / shear modulus estimates and applied
   through the binding's lin_deform. */
void __real_relax_pressure(void);
void __wrap_relax_pressure(void)
{
  vektor ex = {0.0, 0.0, 0.0}, ey = {0.0, 0.0, 0.0}, ez = {0.0, 0.0, 0.0};
  if (b200 == NULL) { __real_relax_pressure(); return; }
#ifdef STRESS_TENS
  {
    real sxx, syy, szz, syz, szx, sxy, mean;
    calc_tot_presstens();
    sxx = tot_presstens.xx / volume - presstens_ext.xx;  syy = tot_presstens.yy / volume - presstens_ext.yy;
    szz = tot_presstens.zz / volume - presstens_ext.zz;  syz = tot_presstens.yz / volume - presstens_ext.yz;
    szx = tot_presstens.zx / volume - presstens_ext.zx;  sxy = tot_presstens.xy / volume - presstens_ext.xy;
    mean = (sxx * relax_dirs.x + syy * relax_dirs.y + szz * relax_dirs.z) / (relax_dirs.x + relax_dirs.y + relax_dirs.z);   /* :152 */
    if (relax_mode == RELAX_FULL || relax_mode == RELAX_AXIAL) {            /* :157-163 */
      ex.x = mean / bulk_module + (sxx - mean) / shear_module;
      ey.y = mean / bulk_module + (syy - mean) / shear_module;
      ez.z = mean / bulk_module + (szz - mean) / shear_module;
    } else ex.x = ey.y = ez.z = mean / bulk_module;                          /* :164-170 */
    if (relax_mode == RELAX_FULL) {                                          /* :171-177 */
      ex.y = ey.x = sxy / shear_module;
      ey.z = ez.y = syz / shear_module;
      ez.x = ex.z = szx / shear_module;
    }
  }
#else
  {                                                                          /* scalar pressure, relaxed towards zero :179-204 */
    real temp = 2.0 * tot_kin_energy / nactive;
    pressure = temp / (volume / natoms) + virial / (DIM * volume);
    ex.x = ey.y = ez.z = pressure / bulk_module;
  }
#endif
  if (relax_mode == RELAX_AXIAL) { ex.x *= relax_dirs.x; ey.y *= relax_dirs.y; ez.z *= relax_dirs.z; }   /* :206-212 */
  __wrap_lin_deform(ex, ey, ez, relax_rate);
}
#endif

#ifdef DEFORM
/* void deform_sample(void)  (src/imd_deform.c:232-269), linked with --wrap; main_loop calls check_nblist right after it */
void __real_deform_sample(void);
void __wrap_deform_sample(void)
{
  if (b200 == NULL) { __real_deform_sample(); return; }
  b200_check(imdb200_deform_sample(b200, deform_size, (double *) deform_shift, shear_def, (double *) deform_shear,
                                   (double *) deform_base));
}
#endif

/* void check_nblist(void)  (src/imd_forces_nbl.c:2007-2037) */
void check_nblist(void)
{
  if (b200 == NULL) { have_valid_nbl = 0; return; }
  b200_check(imdb200_check_nblist(b200));
  b200_scalars(0, 0);
}

/* void make_nblist(void)  (src/imd_forces_nbl.c:136-273) */
void make_nblist(void)
{
  if (b200 == NULL) b200_init();
  b200_check(imdb200_make_nblist(b200));
  b200_scalars(0, 0);
}

/* int estimate_nblist_size(void)  (src/imd_forces_nbl.c:74-128): the device sizes its own table */
int estimate_nblist_size(void) { return last_nbl_len; }

/* void deallocate_nblist(void)  (src/imd_forces_nbl.c:56-66) */
void deallocate_nblist(void) { have_valid_nbl = 0; if (b200) imdb200_invalidate_nblist(b200); }
