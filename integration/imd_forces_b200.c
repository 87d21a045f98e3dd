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
static int    num_max = 0, sync_int = 1;

static void b200_fatal(const char *msg) { error((char *) msg); }   /* IMD's abort convention, src/imd_misc.c:78 */

static void b200_check(int rc) { if (rc) error((char *) imdb200_last_error()); }

/* is one of IMD's writers due at this step?  (src/imd_main_3d.c:690-694) */
static int output_due(int step)
{
  if (step >= steps_max) return 1;
  if (checkpt_int > 0 && step % checkpt_int == 0) return 1;
#ifdef FORCE
  if (force_int > 0 && step % force_int == 0) return 1;
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
    b_num = realloc(b_num, n * sizeof(int)); b_sorte = realloc(b_sorte, n * sizeof(int));
    b_vsorte = realloc(b_vsorte, n * sizeof(int)); b_cell = realloc(b_cell, n * sizeof(int));
    b_slot = realloc(b_slot, n * sizeof(int));
    b_masse = realloc(b_masse, n * sizeof(double)); b_ort = realloc(b_ort, 3 * n * sizeof(double));
    b_impuls = realloc(b_impuls, 3 * n * sizeof(double)); b_kraft = realloc(b_kraft, 3 * n * sizeof(double));
    b_poteng = realloc(b_poteng, n * sizeof(double)); b_rho = realloc(b_rho, n * sizeof(double));
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
  num2idx = realloc(num2idx, (num_max + 1) * sizeof(int));
  for (n = 0; n < b200_n; n++) num2idx[b_num[n]] = (int) n;
  b200_check(imdb200_set_atoms(b200, b200_n, b_num, b_sorte, b_vsorte, b_masse, b_ort, b_impuls));
}

/* write device results back into IMD's cells; the device order is cell-sorted, atoms are matched by NUMMER */
static void b200_download(int forces, int state)
{
  long n, got; 
  static int *d_num = NULL; static long d_cap = 0;
  if (d_cap < b200_n) { d_cap = b200_n; d_num = realloc(d_num, d_cap * sizeof(int)); }
  got = imdb200_get_atoms(b200, d_num, NULL, NULL, NULL, state ? b_ort : NULL, state ? b_impuls : NULL,
                          forces ? b_kraft : NULL, forces ? b_poteng : NULL,
#ifdef EAM2
                          forces ? b_rho : NULL,
#else
                          NULL,
#endif
                          NULL, NULL, NULL);
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
  }
}

static void b200_scalars(int forces, int kinetic)
{
  imdb200_scalars sc;
  b200_check(imdb200_get_scalars(b200, &sc));
  if (forces) { tot_pot_energy = sc.tot_pot_energy; virial = sc.virial; }
  if (kinetic) { tot_kin_energy = sc.tot_kin_energy; eta = sc.eta; }
  have_valid_nbl = sc.have_valid_nbl;
  nbl_count = sc.nbl_count;
  last_nbl_len = (int) (sc.nbl_len / 2);        /* the reference counts each pair once */
}

/* (*move_atoms)() = move_atoms_nve / move_atoms_nvt (src/imd_integrate.c:32, 891) */
static void b200_move_atoms(void)
{
  b200_check(imdb200_move_atoms(b200));
  b200_scalars(0, 1);
  if (sync_due(steps)) b200_download(0, 1);      /* writers run after move_atoms of the same step (src/imd_main_3d.c:690-694) */
}

static void b200_init(void)
{
  imdb200_config cfg;
  char *e = getenv("IMD_B200_SYNC");
  if (e) sync_int = atoi(e);
  imdb200_set_error_handler(b200_fatal);
  imdb200_default_config(&cfg);
  cfg.ntypes = ntypes; cfg.total_types = vtypes;
  cfg.box_x[0] = box_x.x; cfg.box_x[1] = box_x.y; cfg.box_x[2] = box_x.z;
  cfg.box_y[0] = box_y.x; cfg.box_y[1] = box_y.y; cfg.box_y[2] = box_y.z;
  cfg.box_z[0] = box_z.x; cfg.box_z[1] = box_z.y; cfg.box_z[2] = box_z.z;
  cfg.pbc_dirs[0] = pbc_dirs.x; cfg.pbc_dirs[1] = pbc_dirs.y; cfg.pbc_dirs[2] = pbc_dirs.z;
  cfg.nbl_margin = nbl_margin; cfg.nbl_size = nbl_size; cfg.timestep = timestep;
  cfg.ensemble = (ensemble == ENS_NVT) ? IMDB200_ENS_NVT : IMDB200_ENS_NVE;
  if (ensemble != ENS_NVE && ensemble != ENS_NVT) error("imd_b200 supports ensemble nve and nvt");
  cfg.temperature = temperature; cfg.eta = eta; cfg.isq_tau_eta = isq_tau_eta;
  /* `4point` / `spline` make targets (src/Makefile:1694-1701, src/potaccess.h:24-36) */
#if defined(FOURPOINT)
  cfg.interpolation = IMDB200_INTERP_4POINT;
#elif defined(SPLINE)
  cfg.interpolation = IMDB200_INTERP_SPLINE;
#endif
  b200_check(imdb200_create(&cfg, &b200));
  /* pot_table_t (src/types.h:416-428) and imdb200_pot_table have the same layout */
#ifdef EAM2
  b200_check(imdb200_set_potentials(b200, (imdb200_pot_table *) &pair_pot, (imdb200_pot_table *) &embed_pot,
                                    (imdb200_pot_table *) &rho_h_tab));
#ifdef EEAM   /* `eeam` make targets: energy modification term M(p), src/imd_potential.c:82-85 */
  b200_check(imdb200_set_eeam_table(b200, (imdb200_pot_table *) &emod_pot));
#endif
#ifdef ADP    /* `adp` make targets: u(r), w(r), src/imd_potential.c:87-92 (device side not yet run on a GPU) */
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
  if (b200 == NULL) b200_init();
  else if (move_atoms != b200_move_atoms) {      /* a new simulation phase re-selected the integrator */
    imd_move_atoms = move_atoms; move_atoms = b200_move_atoms;
  }
#ifdef NVT
  if (ensemble == ENS_NVT) b200_check(imdb200_set_temperature(b200, temperature));   /* increment_temperature() */
#endif
  b200_check(imdb200_calc_forces(b200, steps));
  b200_scalars(1, 0);
  nfc++;
  if (sync_due(steps)) b200_download(1, 0);
}

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
