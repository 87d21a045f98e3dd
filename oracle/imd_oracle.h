/* oracle/imd_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, single thread, order-faithful) of the reference's
 * force-and-integrate hot path.  It is the checker for the CUDA path; nothing in
 * imd_b200/ may include, link or call it.  Parity of this oracle itself is pinned
 * against the unmodified reference compiled into oracle/_ref (tests/test_oracle_vs_ref.py)
 * and against the committed fixtures in tests/golden/ (generated from that reference by
 * tools/make_golden.py) -- the reference tree ships no golden vectors of its own.
 */
#ifndef IMD_ORACLE_H
#define IMD_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_sim orc_sim;

/* which table: 0 = pair_pot (radial), 1 = embed_pot (not radial), 2 = rho_h_tab (radial) */
enum { ORC_PAIR = 0, ORC_EMBED = 1, ORC_RHO = 2, ORC_EMOD = 3, ORC_ADP_U = 4, ORC_ADP_W = 5 };  /* 3: emod_pot of EEAM builds (not radial) */
enum { ORC_NVE = 0, ORC_NVT = 1, ORC_NPT_ISO = 2, ORC_NPT_AXIAL = 3 };
/* table interpolation a reference build selects at compile time (src/potaccess.h:24-36, src/Makefile:1694-1701) */
enum { ORC_INTERP_3POINT = 0, ORC_INTERP_4POINT = 1, ORC_INTERP_SPLINE = 2 };

orc_sim *orc_create(int ntypes, const double box[9], const int pbc[3], double nbl_margin);
void     orc_destroy(orc_sim *s);

/* PAIR_INT2 (default), PAIR_INT3 (`4point` builds) or PAIR_INT_SP (`spline` builds); call before orc_read_table:
 * the pad rows / second-derivative table depend on it (init_threepoint / init_fourpoint / init_spline,
 * src/imd_potential.c:1171-1272) */
void orc_set_interpolation(orc_sim *s, int mode);
/* read_pot_table (src/imd_potential.c:161-282); returns 0 on success */
int  orc_read_table(orc_sim *s, int which, const char *path);
/* PAIR_INT2 / VAL_FUNC2 / DERIV_FUNC2 (src/potaccess.h:323-354, 465-495, 591-621) */
void orc_pair_int(const orc_sim *s, int which, int col, double r2, double *pot, double *grad, int *is_short);
int  orc_table_info(const orc_sim *s, int which, int col, double *begin, double *end, double *step, int *len);

void orc_set_atoms(orc_sim *s, long n, const int *nummer, const int *sorte, const int *vsorte,
                   const double *masse, const double *ort, const double *impuls);
void orc_set_restrictions(orc_sim *s, int nvtypes, const double *restr3);
void orc_set_integrator(orc_sim *s, int ensemble, double timestep, double temperature,
                        double eta, double isq_tau_eta);
/* NPT_iso (move_atoms_npt_iso, src/imd_integrate.c:1472-1729): xi, Ekin_old (< 0: compute it from the momenta at
 * the first step, like steps == steps_min), pressure_ext, its per-step increment, 1/tau_xi^2 */
void orc_set_npt(orc_sim *s, double xi, double Ekin_old, double pressure_ext, double d_pressure, double isq_tau_xi);
void orc_get_npt(const orc_sim *s, double out[4]);
/* NPT_axial (move_atoms_npt_axial, src/imd_integrate.c:1747-1959): per-axis xi, pressure_ext, its per-step increment, relax_dirs,
 * Ekin_old (< 0: steps == steps_min), dyn_stress_x/y/z the previous step left; out13 = xi, stress, pressure_ext, dyn_stress, Ekin_old */
void orc_set_npt_axial(orc_sim *s, const double *xi3, const double *pext3, const double *dpext3, const int *relax_dirs,
                       double Ekin_old, const double *dyn3, double isq_tau_xi);
void orc_get_npt_axial(const orc_sim *s, double *out13);
/* `ber` builds: Berendsen scaling of the momenta inside move_atoms_nve (src/imd_integrate.c:44-53, 341-350) */
void orc_set_berendsen(orc_sim *s, double tauber, double tot_kin_energy);      /* xi, Ekin_old, pressure of the last step, pressure_ext */
void orc_set_box(orc_sim *s, const double box[9]);        /* make_box, src/imd_geom_3d.c:52-104 */

void orc_calc_forces(orc_sim *s, int do_press_calc);      /* src/imd_forces_nbl.c:281-1999 */
void orc_move_atoms(orc_sim *s, int do_press_calc);       /* src/imd_integrate.c:32-497, 891-1147 */
void orc_check_nblist(orc_sim *s);                        /* src/imd_forces_nbl.c:2007-2037 */
void orc_step(orc_sim *s, int nsteps);
void orc_lin_deform(orc_sim *s, const double dx[3], const double dy[3], const double dz[3], double scale);
/* deform_sample (src/imd_deform.c:232-269); arrays indexed by virtual type, 3 doubles each */
void orc_deform_sample(orc_sim *s, double deform_size, const double *deform_shift, const int *shear_def,
                       const double *deform_shear, const double *deform_base);

long   orc_natoms(const orc_sim *s);
int    orc_have_valid_nbl(const orc_sim *s);
int    orc_nbl_count(const orc_sim *s);
double orc_cellsz(const orc_sim *s);
void   orc_get_celldims(const orc_sim *s, int out6[6]);
/* tot_pot_energy, tot_kin_energy, virial, vir_xx,yy,zz,yz,zx,xy, volume, nactive, eta, temperature, timestep */
void   orc_get_scalars(const orc_sim *s, double out[14]);
void   orc_get_box(const orc_sim *s, double out9[9]);
/* real atoms in current cell-traversal order; any pointer may be NULL */
long   orc_get_atoms(const orc_sim *s, int *nummer, int *sorte, int *vsorte, double *masse, double *ort,
                     double *impuls, double *kraft, double *poteng, double *rho, double *dF,
                     double *presstens, double *nblpos);
/* half Verlet list as (NUMMER_i, NUMMER_j, shift of j in box units); returns the pair count */
long   orc_get_nbl_pairs(const orc_sim *s, int *pi, int *pj, signed char *shift, long cap);
void   orc_tot_presstens(const orc_sim *s, double out6[6]);
/* EEAM: p_i = sum_j rho_j(r_ij)^2 and M'(p_i), storage order of orc_get_atoms */
long   orc_get_eeam(const orc_sim *s, double *eam_p, double *dM);
/* ADP: dipole mu[3] and quadrupole lambda[6] (xx yy zz yz zx xy) per atom, storage order of orc_get_atoms */
long   orc_get_adp(const orc_sim *s, double *mu3, double *la6);

#ifdef __cplusplus
}
#endif
#endif
