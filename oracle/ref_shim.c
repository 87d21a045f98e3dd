/* oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Our own glue, compiled TOGETHER WITH the unmodified reference sources (see
 * oracle/Makefile) into oracle/_ref/libimdref_*.so.  It lets a ctypes driver
 * (oracle/ref_driver.py) run the reference's step loop one call at a time and
 * read its global state at full binary precision -- the reference itself only
 * offers text output.
 *
 * Nothing here re-implements reference arithmetic: every function either calls a
 * reference symbol or copies reference memory out.
 *
 *   ref_setup()        = prologue of main()           (src/imd.c:73-135)
 *                      + prologue of main_loop()      (src/imd_main_3d.c:90-94)
 *   ref_calc_forces()  = calc_forces(steps)           (src/imd_forces_nbl.c:281)
 *   ref_move_atoms()   = (*move_atoms)()              (src/globals.h:1444)
 *   ref_check_nblist() = check_nblist()               (src/imd_forces_nbl.c:2007)
 *   ref_get_nbl_pairs(): tl/tb/cl_off/cl_num          (src/imd_forces_nbl.c:47)
 */
#include "imd.h"
#include "potaccess.h"

extern int *tl, *tb, *cl_off, *cl_num;

int ref_setup(const char *paramfile, int restart)
{
  strcpy(progname, "imd_ref");
  strcpy(paramfilename, paramfile);
  imdrestart = restart;
  imd_init_timer(&time_total,     0, NULL, NULL);
  imd_init_timer(&time_setup,     0, NULL, NULL);
  imd_init_timer(&time_main,      0, NULL, NULL);
  imd_init_timer(&time_output,    0, NULL, NULL);
  imd_init_timer(&time_input,     0, NULL, NULL);
  imd_init_timer(&time_integrate, 0, NULL, NULL);
  imd_init_timer(&time_forces,    0, NULL, NULL);
  read_parameters(paramfilename, 1);
  setup_potentials();
  if ('_' == infilename[0]) generate_atoms(infilename);
  else                      read_atoms(infilename);
  if (0 == imdrestart) {
    if (do_maxwell) maxwell(temperature);
    do_maxwell = 0;
  }
  steps = steps_min;
  return 0;
}

void ref_calc_forces(int step)  { calc_forces(step); }
void ref_move_atoms(void)       { move_atoms(); }
void ref_check_nblist(void)     { check_nblist(); }
void ref_set_press_calc(int on) {
#ifdef STRESS_TENS
  do_press_calc = on;
#endif
}
void ref_calc_tot_presstens(double *out6)
{
#ifdef STRESS_TENS
  calc_tot_presstens();
  out6[0] = tot_presstens.xx; out6[1] = tot_presstens.yy; out6[2] = tot_presstens.zz;
  out6[3] = tot_presstens.yz; out6[4] = tot_presstens.zx; out6[5] = tot_presstens.xy;
#endif
}
#ifdef HOMDEF
void ref_lin_deform(void) { lin_deform(lindef_x, lindef_y, lindef_z, lindef_size); }
#endif
#ifdef DEFORM
void ref_deform_sample(void) { deform_sample(); }
#endif

/* one step of main_loop (src/imd_main_3d.c:155-870) restricted to the hot path */
void ref_step(void)
{
  calc_forces(steps);
  move_atoms();
  check_nblist();
  steps++;
}

long   ref_natoms(void)         { return natoms; }
int    ref_nbl_count(void)      { return nbl_count; }
int    ref_have_valid_nbl(void) { return have_valid_nbl; }
double ref_cellsz(void)         { return cellsz; }
void   ref_get_scalars(double *out)
{
  out[0] = tot_pot_energy; out[1] = tot_kin_energy; out[2] = virial;
  out[3] = vir_xx; out[4] = vir_yy; out[5] = vir_zz;
  out[6] = vir_yz; out[7] = vir_zx; out[8] = vir_xy;
  out[9] = volume; out[10] = (double) nactive; out[11] = eta;
  out[12] = temperature; out[13] = timestep;
}
void ref_set_eta(double e) { eta = e; }
void ref_get_box(double *out9)
{
  out9[0] = box_x.x; out9[1] = box_x.y; out9[2] = box_x.z;
  out9[3] = box_y.x; out9[4] = box_y.y; out9[5] = box_y.z;
  out9[6] = box_z.x; out9[7] = box_z.y; out9[8] = box_z.z;
}
void ref_get_celldims(int *out6)
{
  out6[0] = global_cell_dim.x; out6[1] = global_cell_dim.y; out6[2] = global_cell_dim.z;
  out6[3] = cell_dim.x; out6[4] = cell_dim.y; out6[5] = cell_dim.z;
}

/* copy the per-atom state of all real atoms, in the reference's cell-traversal order */
long ref_get_atoms(int *nummer, int *sorte, int *vsorte, double *masse,
                   double *ort, double *impuls, double *kraft, double *poteng,
                   double *rho, double *dF, double *presstens, double *nblpos)
{
  long n = 0; int k, i;
  for (k = 0; k < NCELLS; k++) {
    cell *p = CELLPTR(k);
    for (i = 0; i < p->n; i++, n++) {
      if (nummer) nummer[n] = NUMMER(p,i);
      if (sorte)  sorte[n]  = SORTE(p,i);
      if (vsorte) vsorte[n] = VSORTE(p,i);
      if (masse)  masse[n]  = MASSE(p,i);
      if (ort)    { ort[3*n] = ORT(p,i,X); ort[3*n+1] = ORT(p,i,Y); ort[3*n+2] = ORT(p,i,Z); }
      if (impuls) { impuls[3*n] = IMPULS(p,i,X); impuls[3*n+1] = IMPULS(p,i,Y); impuls[3*n+2] = IMPULS(p,i,Z); }
      if (kraft)  { kraft[3*n] = KRAFT(p,i,X); kraft[3*n+1] = KRAFT(p,i,Y); kraft[3*n+2] = KRAFT(p,i,Z); }
      if (poteng) poteng[n] = POTENG(p,i);
#ifdef EAM2
      if (rho)    rho[n] = EAM_RHO(p,i);
      if (dF)     dF[n]  = EAM_DF(p,i);
#endif
#ifdef STRESS_TENS
      if (presstens) {
        presstens[6*n  ] = PRESSTENS(p,i,xx); presstens[6*n+1] = PRESSTENS(p,i,yy);
        presstens[6*n+2] = PRESSTENS(p,i,zz); presstens[6*n+3] = PRESSTENS(p,i,yz);
        presstens[6*n+4] = PRESSTENS(p,i,zx); presstens[6*n+5] = PRESSTENS(p,i,xy);
      }
#endif
      if (nblpos && p->nbl_pos) {
        nblpos[3*n] = NBL_POS(p,i,X); nblpos[3*n+1] = NBL_POS(p,i,Y); nblpos[3*n+2] = NBL_POS(p,i,Z);
      }
    }
  }
  return n;
}

/* ADP per-atom fields (src/imd_forces_nbl.c:613-631): mu[3], lambda[6] = xx yy zz yz zx xy */
long ref_get_adp(double *mu3, double *la6)
{
  long n = 0;
#ifdef ADP
  int k, i;
  for (k = 0; k < NCELLS; k++) {
    cell *p = CELLPTR(k);
    for (i = 0; i < p->n; i++, n++) {
      if (mu3) { mu3[3*n] = ADP_MU(p,i,X); mu3[3*n+1] = ADP_MU(p,i,Y); mu3[3*n+2] = ADP_MU(p,i,Z); }
      if (la6) { la6[6*n] = ADP_LAMBDA(p,i,xx); la6[6*n+1] = ADP_LAMBDA(p,i,yy); la6[6*n+2] = ADP_LAMBDA(p,i,zz);
                 la6[6*n+3] = ADP_LAMBDA(p,i,yz); la6[6*n+4] = ADP_LAMBDA(p,i,zx); la6[6*n+5] = ADP_LAMBDA(p,i,xy); }
    }
  }
#endif
  return n;
}

/* NPT_iso state (src/globals.h:407, 569-574): xi.x, Ekin_old, pressure, pressure_ext.x, isq_tau_xi */
void ref_get_npt(double *out5)
{
#ifdef NPT
  out5[0] = xi.x; out5[1] = Ekin_old; out5[2] = pressure; out5[3] = pressure_ext.x; out5[4] = isq_tau_xi;
#else
  out5[0] = out5[1] = out5[2] = out5[3] = out5[4] = 0.0;
#endif
}

/* NPT_axial state (src/globals.h:565-574, 627): xi, stress_x/y/z of the last move_atoms, pressure_ext, dyn_stress_x/y/z
   (the kinetic part the next move_atoms starts from), Ekin_old, relax_dirs */
void ref_get_npt_axial(double *out16)
{
#ifdef NPT_axial
  out16[0] = xi.x; out16[1] = xi.y; out16[2] = xi.z;
  out16[3] = stress_x; out16[4] = stress_y; out16[5] = stress_z;
  out16[6] = pressure_ext.x; out16[7] = pressure_ext.y; out16[8] = pressure_ext.z;
  out16[9] = dyn_stress_x; out16[10] = dyn_stress_y; out16[11] = dyn_stress_z;
  out16[12] = Ekin_old;
  out16[13] = relax_dirs.x; out16[14] = relax_dirs.y; out16[15] = relax_dirs.z;
#else
  int i; for (i = 0; i < 16; i++) out16[i] = 0.0;
#endif
}

/* EEAM per-atom fields (src/imd_forces_nbl.c:591-610, 1090-1095), same order as ref_get_atoms */
long ref_get_eeam(double *eam_p, double *eam_dM)
{
  long n = 0;
#ifdef EEAM
  int k, i;
  for (k = 0; k < NCELLS; k++) {
    cell *p = CELLPTR(k);
    for (i = 0; i < p->n; i++, n++) { if (eam_p) eam_p[n] = EAM_P(p,i); if (eam_dM) eam_dM[n] = EAM_DM(p,i); }
  }
#endif
  return n;
}

/* overwrite positions / momenta of the real atoms (same traversal order as ref_get_atoms);
   used to put reference and candidate on bit-identical states */
void ref_set_atoms(const double *ort, const double *impuls)
{
  long n = 0; int k, i;
  for (k = 0; k < NCELLS; k++) {
    cell *p = CELLPTR(k);
    for (i = 0; i < p->n; i++, n++) {
      if (ort)    { ORT(p,i,X) = ort[3*n]; ORT(p,i,Y) = ort[3*n+1]; ORT(p,i,Z) = ort[3*n+2]; }
      if (impuls) { IMPULS(p,i,X) = impuls[3*n]; IMPULS(p,i,Y) = impuls[3*n+1]; IMPULS(p,i,Z) = impuls[3*n+2]; }
    }
  }
  have_valid_nbl = 0;
}

/* Verlet list as (NUMMER_i, NUMMER_j, image shift of j in box units) triples.
   Ghost cells carry no NUMMER (copy_cell, src/imd_comm_force_3d.c:726-778, copies
   positions and types only), but copies preserve the order inside a cell, so for
   cpu_dim = 1 1 1 the ghost (cell c, slot j) is the image of slot j of the real
   cell obtained by wrapping c's coordinates. */
long ref_get_nbl_pairs(int *pi, int *pj, signed char *shift, long cap)
{
  long n = 0, cnt = 0; int k, i, m;
  if (!have_valid_nbl || tl == NULL) return -1;
  for (k = 0; k < ncells; k++) {
    cell *p = cell_array + cnbrs[k].np;
    for (i = 0; i < p->n; i++, n++) {
      for (m = tl[n]; m < tl[n+1]; m++) {
        int c  = cl_num[tb[m]];
        int j  = tb[m] - cl_off[c];
        int cx = c / (cell_dim.y * cell_dim.z);
        int cy = (c / cell_dim.z) % cell_dim.y;
        int cz = c % cell_dim.z;
        int sx = 0, sy = 0, sz = 0, src;
        if (cx == 0) { cx = cell_dim.x - 2; sx = -1; } else if (cx == cell_dim.x - 1) { cx = 1; sx = 1; }
        if (cy == 0) { cy = cell_dim.y - 2; sy = -1; } else if (cy == cell_dim.y - 1) { cy = 1; sy = 1; }
        if (cz == 0) { cz = cell_dim.z - 2; sz = -1; } else if (cz == cell_dim.z - 1) { cz = 1; sz = 1; }
        src = (cx * cell_dim.y + cy) * cell_dim.z + cz;
        if (cnt < cap) {
          pi[cnt] = NUMMER(p,i);
          pj[cnt] = NUMMER(cell_array + src, j);
          if (shift) { shift[3*cnt] = sx; shift[3*cnt+1] = sy; shift[3*cnt+2] = sz; }
        }
        cnt++;
      }
    }
  }
  return cnt;
}

/* potential table access through the reference's own macros (src/potaccess.h:24-36) */
void ref_pair_int(int which, int col, double r2, double *pot, double *grad)
{
  int is_short = 0; real p = 0, g = 0;
  pot_table_t *pt = NULL; int inc = ntypes * ntypes;
  if (which == 0) pt = &pair_pot;
#ifdef EAM2
  if (which == 1) { pt = &embed_pot; inc = ntypes; }
  if (which == 2) pt = &rho_h_tab;
#endif
#ifdef EEAM
  if (which == 3) { pt = &emod_pot; inc = ntypes; }
#endif
  if (!pt) return;
  PAIR_INT(p, g, *pt, col, inc, r2, is_short);
  *pot = p; *grad = g;
}
