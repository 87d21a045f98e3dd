/* imd_b200.h -- C ABI of the B200-native force-and-integrate engine.
 *
 * This is the drop-in boundary for IMD's NVE/NVT step loop.  IMD has
 * no run-time plugin ABI: its "operator interface" is link-time substitution of
 *     calc_forces / make_nblist / check_nblist / fix_cells / move_atoms / lin_deform ...
 * selected by the Makefile (precedent: FORCESOURCES = imd_forces_cbe.c, src/Makefile:1141-1143;
 * the same surface is listed for scripting in src/imd.i:37-71).  Each entry point below
 * names the reference symbol it replaces.  All arguments are plain C: pointers, sizes and
 * doubles, host memory unless stated otherwise.  There is no CPU fallback anywhere: every
 * entry point needs a CUDA device of compute capability 10.0 (B200, sm_100a) and returns
 * IMDB200_ERR_CUDA (or calls the error handler) without one.
 *
 * Error convention: functions return 0 on success, a negative IMDB200_ERR_* otherwise, and
 * leave a message retrievable by imdb200_last_error().  A host that wants IMD's convention
 * (error() -> imderror() prints and exits, src/imd_misc.c:78-98, src/makros.h:23) installs
 * its handler with imdb200_set_error_handler(); it is then called before the function returns.
 *
 * Threading: one host thread (or process) per simulation handle / GPU, like one MPI rank per
 * domain in the reference.  Handles are independent.
 *
 * Numerical contract (BASELINE.json north_star; tests/common.py holds the bars):
 *   bit-exact   the neighbour SET (every pair is tested with the reference's operands and rounding:
 *               r2 = (dx*dx + dy*dy) + dz*dz without FMA, image positions added stage by stage), the cell grid,
 *               the rebuild decisions of check_nblist and therefore nbl_count
 *   <= 1e-10    forces, per-atom energies, rho, F', stress, Epot, virial, Ekin, eta -- per component, relative, with an
 *               absolute floor of 1e-12 of the largest component.  Inside the force loops two operations deliberately
 *               depart from the reference's rounding: r2 is formed with FMAs, and the table index (r2 - begin)*invstep
 *               is one FMA (tab_index_fast, csrc/internal.cuh).  Both move k/chi only when the argument lies within an
 *               ulp of a table node, where the interpolant is continuous: values change by ~1e-16 relative.  Sums run in
 *               a different order than the reference's half list (every atom gathers its own full list): ~1e-13.
 *   run to run  results are bit-reproducible (no floating-point atomics on any result; fixed reduction orders).
 */
#ifndef IMD_B200_H
#define IMD_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define IMDB200_VERSION 1

enum {
  IMDB200_OK = 0,
  IMDB200_ERR_ARG = -1,     /* bad argument / call order                                  */
  IMDB200_ERR_CUDA = -2,    /* CUDA runtime error, or no sm_100 device                    */
  IMDB200_ERR_NBL = -3,     /* "neighbor table full" (src/imd_forces_nbl.c:265-267)       */
  IMDB200_ERR_CELLS = -4,   /* "global_cell_dim too small" (src/imd_geom_3d.c:163-177)    */
  IMDB200_ERR_COMM = -5,    /* NCCL / halo exchange failure                               */
  IMDB200_ERR_IO = -6,      /* file errors of the host-side readers                       */
  IMDB200_ERR_EXPLODE = -7  /* "system seems to explode!" (src/imd_geom_3d.c:96)          */
};

enum { IMDB200_ENS_NVE = 0, IMDB200_ENS_NVT = 1, IMDB200_ENS_NPT_ISO = 2, IMDB200_ENS_NPT_AXIAL = 3 }; /* ensemble keyword, src/imd_param.c:377-444 */

/* Table interpolation.  The reference fixes it at compile time (src/potaccess.h:24-36; make targets with
 * `4point` or `spline` in their name, src/Makefile:1694-1701); here it is a run-time field of the config.
 *   3POINT  PAIR_INT2   quadratic through samples k..k+2            (src/potaccess.h:323-354), the default
 *   4POINT  PAIR_INT3   cubic through samples k-1..k+2, k >= 1      (src/potaccess.h:365-407)
 *   SPLINE  PAIR_INT_SP cubic spline on table + second derivatives  (src/potaccess.h:418-457) */
enum { IMDB200_INTERP_3POINT = 0, IMDB200_INTERP_4POINT = 1, IMDB200_INTERP_SPLINE = 2 };

typedef struct imdb200_sim imdb200_sim;

/* Mirrors pot_table_t (src/types.h:416-428): ncols columns interleaved row by row,
 * table[k*ncols+col], (maxsteps+2) rows allocated.  The pad rows behind len[col] (and, for splines, the
 * second-derivative table) depend on the interpolation: imdb200_set_potentials recomputes them from the
 * samples the way init_threepoint / init_fourpoint / init_spline do (src/imd_potential.c:1171-1272), so a
 * table padded by any of them is accepted. */
typedef struct {
  double *begin, *end, *step, *invstep;
  int *len;
  int ncols, maxsteps;
  double *table;
} imdb200_pot_table;

/* The run-time parameters of the hot path; names follow IMD's parameter-file keys
 * (src/imd_param.c) and globals (src/globals.h). */
typedef struct {
  int ntypes;              /* ntypes                                                       */
  int total_types;         /* total_types (virtual types; >= ntypes)                       */
  double box_x[3], box_y[3], box_z[3]; /* box_x / box_y / box_z                            */
  int pbc_dirs[3];         /* pbc_dirs                                                     */
  int cpu_dim[3];          /* cpu_dim: process (GPU) grid                                  */
  int my_coord[3];         /* this rank's coordinates in the grid (MPI_Cart_coords)        */
  double nbl_margin;       /* nbl_margin (default 0.4, src/globals.h:419)                  */
  double nbl_size;         /* nbl_size   (default 1.1)                                     */
  double timestep;         /* timestep                                                     */
  int ensemble;            /* IMDB200_ENS_*                                                */
  double temperature;      /* starttemp (NVT target)                                       */
  double eta;              /* eta                                                          */
  double isq_tau_eta;      /* 1/tau_eta^2                                                  */
  int device;              /* CUDA device ordinal, -1: current device                      */
  int lanes_per_atom;      /* 0 = choose; else 1,2,4,8,16,32 lanes cooperate on one atom   */
  int interpolation;       /* IMDB200_INTERP_*; what a `4point` / `spline` build of IMD selects */
  /* NPT_iso (ensemble npt_iso, src/imd_integrate.c:1472-1729) */
  double xi;               /* xi.x: barostat friction at the start                         */
  double isq_tau_xi;       /* 1/tau_xi^2                                                   */
  double pressure_ext;     /* pressure_start                                               */
  double d_pressure;       /* (pressure_end - pressure_start) / (steps_max - steps_min)    */
} imdb200_config;

void        imdb200_default_config(imdb200_config *cfg);
const char *imdb200_last_error(void);
void        imdb200_set_error_handler(void (*handler)(const char *msg));
/* number of kernels this library has launched in this process (bench.py: gpu_launches) */
long long   imdb200_kernel_launches(void);
/* number of sm_100 devices this process sees (MPI builds: device = rank % count; no reference counterpart) */
int         imdb200_device_count(void);

/* ---- setup ------------------------------------------------------------------------------ */
/* replaces: globals filled by read_parameters() + make_box() + init_cells()
 * (src/imd_param.c:3880, src/imd_geom_3d.c:52-104, 113-426) */
int  imdb200_create(const imdb200_config *cfg, imdb200_sim **out);
void imdb200_destroy(imdb200_sim *sim);

/* replaces: the device-side view of pair_pot, embed_pot, rho_h_tab after setup_potentials()
 * (src/imd_potential.c:43-130).  embed and rho may both be NULL: pair interactions only
 * (imd_nve_nbl build).  cellsz (max table end, in r^2) is taken from the tables the way
 * read_pot_table does (src/imd_potential.c:406). */
int  imdb200_set_potentials(imdb200_sim *sim, const imdb200_pot_table *pair,
                            const imdb200_pot_table *embed, const imdb200_pot_table *rho);

/* EEAM builds (make targets with `eeam`, src/Makefile:1161-1163): the energy modification term M(p_i),
 * p_i = sum_j rho_j(r_ij)^2 -- emod_pot, read from `eeam_energy_file` with ntypes columns, not radial
 * (src/imd_potential.c:82-85).  Call after imdb200_set_potentials; NULL switches the term off again.  Replaces the
 * `#ifdef EEAM` branches of calc_forces (src/imd_forces_nbl.c:591-610, 1090-1095, 1181-1208). */
int  imdb200_set_eeam_table(imdb200_sim *sim, const imdb200_pot_table *emod);
/* EAM_P and EAM_DM of the owned atoms, in the order of imdb200_get_atoms; returns the count */
long imdb200_get_eeam(imdb200_sim *sim, double *eam_p, double *eam_dM);

/* ADP builds (make targets with `adp`): dipole and quadrupole distortion functions u(r), w(r) -- adp_upot and
 * adp_wpot, read from `adp_upotfile` / `adp_wpotfile` with ntypes^2 columns in r^2 (src/imd_potential.c:87-92).
 * Call after imdb200_set_potentials; NULL, NULL switches the terms off.  Replaces the `#ifdef ADP` branches of
 * calc_forces (src/imd_forces_nbl.c:613-631, 919-929, 1096-1110, 1217-1255).  Green on B200 against the reference's `adp`
 * build (tests/test_gpu_parity.py::test_cuda_adp_matches_reference_fixture). */
int  imdb200_set_adp_tables(imdb200_sim *sim, const imdb200_pot_table *u, const imdb200_pot_table *w);
/* ADP_MU [n][3] and ADP_LAMBDA [n][6] = xx yy zz yz zx xy of the owned atoms, in the order of imdb200_get_atoms */
long imdb200_get_adp(imdb200_sim *sim, double *mu3, double *lambda6);

/* restrictions per virtual type (3 doubles each); default all 1 (src/imd_param.c:2053-2066) */
int  imdb200_set_restrictions(imdb200_sim *sim, int total_types, const double *restrictions);

/* replaces: the per-cell atom arrays filled by generate_atoms()/read_atoms()
 * (src/imd_generate.c, src/imd_io_3d.c:44-).  Arrays are in any order; ort/impuls are
 * n x 3.  Only atoms inside this rank's domain are kept when cpu_dim != 1 1 1.
 * vsorte may be NULL (= sorte), impuls may be NULL (= 0). */
int  imdb200_set_atoms(imdb200_sim *sim, long n, const int *nummer, const int *sorte,
                       const int *vsorte, const double *masse, const double *ort,
                       const double *impuls);

/* run all kernels of this handle on a caller-owned CUDA stream (cudaStream_t passed as void*), e.g.
 * torch.cuda.current_stream().cuda_stream, so that the caller's events bracket our launches */
int  imdb200_set_stream(imdb200_sim *sim, void *cuda_stream);

/* multi-GPU (replaces: MPI_Init + setup_mpi_topology, src/imd_mpi_util.c:48-63, src/imd_geom_mpi_3d.c:32-90):
 * one process per GPU and domain.  Rank 0 obtains a ncclUniqueId and the host distributes it (MPI_Bcast,
 * torch.distributed, a file ...); every rank then joins with its x-major rank over cpu_dim,
 * rank = (my_coord.x*cpu_dim.y + my_coord.y)*cpu_dim.z + my_coord.z.  Call between imdb200_set_potentials and
 * imdb200_set_atoms; every later call is collective (all ranks make the same calls in the same order). */
int  imdb200_comm_unique_id(void *id128);                 /* rank 0: fills 128 bytes        */
int  imdb200_comm_init(imdb200_sim *sim, const void *id128, int rank, int nranks);
/* replaces: send_forces(add_forces, pack_forces, unpack_forces) (src/imd_comm_force_3d.c:569-714, 897-1020) for a
 * caller-owned DEVICE field laid out field[c*stride + i], i < natoms_local + nghost_local, c < ncomp <= 8: what the
 * images accumulated is added to their owners, across GPUs where the owner lives on another one. */
int  imdb200_send_forces(imdb200_sim *sim, double *dev_field, int ncomp, long stride);
long imdb200_nghost_local(imdb200_sim *sim);

/* ---- the step loop ---------------------------------------------------------------------- */
/* replaces: void calc_forces(int steps)  (src/imd_forces_nbl.c:281-1999; prototypes.h:173) */
int  imdb200_calc_forces(imdb200_sim *sim, int steps);
/* replaces: (*move_atoms)() = move_atoms_nve / move_atoms_nvt (src/imd_integrate.c:32, 891) */
int  imdb200_move_atoms(imdb200_sim *sim);
/* replaces: void check_nblist(void) (src/imd_forces_nbl.c:2007-2037) */
int  imdb200_check_nblist(imdb200_sim *sim);
/* replaces: void fix_cells(void) (src/imd_fix_cells_3d.c:36-201) + do_boundaries */
int  imdb200_fix_cells(imdb200_sim *sim);
/* replaces: void make_nblist(void) (src/imd_forces_nbl.c:136-273) */
int  imdb200_make_nblist(imdb200_sim *sim);
/* nsteps iterations of { calc_forces; move_atoms; check_nblist } kept on the device
 * (main_loop body, src/imd_main_3d.c:405, 559, 768-772) */
int  imdb200_run(imdb200_sim *sim, int nsteps);
/* do_press_calc (src/imd_main_3d.c:183-195): accumulate the per-atom stress tensor */
int  imdb200_set_press_calc(imdb200_sim *sim, int on);
int  imdb200_invalidate_nblist(imdb200_sim *sim);         /* have_valid_nbl = 0              */
/* The list of every atom is grouped by build distance; by default a force call leaves out the groups that
 * cannot have come inside the cut-off given the largest displacement since the build (results are bit-identical
 * either way -- the entries left out would fail the r2 test of src/imd_forces_nbl.c:493, 588, 1172).
 * on = 0 walks every stored entry like the reference does (test hook). */
int  imdb200_set_skin_skip(imdb200_sim *sim, int on);
int  imdb200_set_eta(imdb200_sim *sim, double eta);
/* Overwrite the momenta of this rank's atoms, matched by atom number; positions, forces and the neighbour list stay as
 * they are.  replaces: host code that re-draws IMPULS between two steps -- maxwell(temperature) every `tempintv` steps in
 * `and` builds (Andersen thermostat, src/imd_integrate.c:491-495), which the binding keeps on the host (src/imd_maxwell.c
 * draws from drand48 in cell order).  n must be the local atom count; atoms this rank does not own are an error. */
int  imdb200_set_momenta(imdb200_sim *sim, long n, const int *nummer, const double *impuls);
int  imdb200_set_temperature(imdb200_sim *sim, double temperature);
/* replaces: the BER branch of move_atoms_nve in `ber` builds (src/imd_integrate.c:44-53, 341-350) and the parameter
   tau_berendsen (global tauber).  tauber > 0 switches the Berendsen scaling of the momenta on (ensemble nve; the target is
   imdb200_set_temperature's); tot_kin_energy is what the previous move_atoms left (IMD's global of that name), it drives
   the first scale factor.  tauber <= 0 switches it off. */
int  imdb200_set_berendsen(imdb200_sim *sim, double tauber, double tot_kin_energy);

/* replaces: lin_deform(dx,dy,dz,scale) (src/imd_deform.c:35-119) incl. make_box() */
int  imdb200_lin_deform(imdb200_sim *sim, const double dx[3], const double dy[3],
                        const double dz[3], double scale);
/* replaces: deform_sample() (src/imd_deform.c:232-269); arrays per virtual type */
int  imdb200_deform_sample(imdb200_sim *sim, double deform_size, const double *deform_shift,
                           const int *shear_def, const double *deform_shear,
                           const double *deform_base);

/* ---- results ---------------------------------------------------------------------------- */
/* globals after a step, summed over ranks like the MPI_Allreduce sites
 * (src/imd_forces_nbl.c:1975-1994, src/imd_integrate.c:437-466, 1104-1130) */
typedef struct {
  double tot_pot_energy, tot_kin_energy, virial;
  double volume;
  double eta;
  double max_displacement2;  /* check_nblist's max |x - nbl_pos|^2                          */
  double tot_presstens[6];   /* xx yy zz yz zx xy (calc_tot_presstens, imd_main_3d.c:2069)   */
  long long natoms, nactive;
  long long nbl_len;         /* stored (full-list) neighbour entries on this rank            */
  int have_valid_nbl, nbl_count, is_short;
  int global_cell_dim[3], cell_dim[3];
  double cellsz;
} imdb200_scalars;
int  imdb200_get_scalars(imdb200_sim *sim, imdb200_scalars *out);
/* NPT_iso hand-over between runs (the reference keeps xi, Ekin_old and pressure_ext in globals, src/globals.h:407,
 * 571-574): Ekin_old < 0 makes the next move_atoms compute it from the momenta (calc_dyn_pressure, what the
 * reference does at steps == steps_min).  out4 = xi, Ekin_old, pressure used by the last step, pressure_ext. */
int  imdb200_set_npt_state(imdb200_sim *sim, double xi, double Ekin_old, double pressure_ext);
int  imdb200_get_npt_state(imdb200_sim *sim, double out4[4]);
/* NPT_axial (ensemble npt_axial, move_atoms_npt_axial, src/imd_integrate.c:1747-1959): Nose-Hoover thermostat + one
 * barostat per box axis, driven by stress_x/y/z = (dyn_stress + vir)/volume of that axis.  The reference's `npt_axial`
 * builds define P_AXIAL and accumulate vir_xx/yy/zz in calc_forces (src/imd_forces_nbl.c:548-556); here the ensemble
 * always runs the per-atom-stress instances of the force kernels and sums their diagonal.  isq_tau_xi, temperature, eta
 * and isq_tau_eta come from imdb200_config.  xi3 = xi.x/y/z, pressure_ext3 = pressure_start, d_pressure3 =
 * (pressure_end - pressure_start)/(steps_max - steps_min), relax_dirs3 = relax_dirs (src/globals.h:627);
 * Ekin_old < 0 makes the next move_atoms start like steps == steps_min (calc_dyn_pressure, xi *= relax_dirs), else
 * dyn_stress3 = dyn_stress_x/y/z left by the previous step.  out13 = xi[3], stress_x/y/z of the last step, pressure_ext[3],
 * dyn_stress[3], Ekin_old. */
int  imdb200_set_npt_axial(imdb200_sim *sim, const double xi3[3], const double pressure_ext3[3], const double d_pressure3[3],
                           const int relax_dirs3[3], double Ekin_old, const double dyn_stress3[3]);
int  imdb200_get_npt_axial(imdb200_sim *sim, double out13[13]);
/* replaces: reading the globals box_x, box_y, box_z (src/globals.h) after lin_deform has changed them;
 * out9 = box_x, box_y, box_z */
int  imdb200_get_box(imdb200_sim *sim, double out9[9]);

/* copy the owned atoms back to the host (any pointer may be NULL); returns the count.
 * Order: the device's cell-sorted order; identify atoms by nummer (NUMMER, types.h:189). */
long imdb200_get_atoms(imdb200_sim *sim, int *nummer, int *sorte, int *vsorte, double *masse,
                       double *ort, double *impuls, double *kraft, double *poteng,
                       double *eam_rho, double *eam_dF, double *presstens, double *nbl_pos);
long imdb200_natoms_local(imdb200_sim *sim);
/* test hook: the neighbour list as (nummer_i, nummer_j, image shift of j) triples; the list is
 * stored in full (both directions).  Returns the number of entries (call with cap = 0 first). */
long imdb200_get_nblist(imdb200_sim *sim, int *nummer_i, int *nummer_j, signed char *shift3, long cap);
/* interpolate through the device tables with the kernels' own lookup code (test hook for
 * PAIR_INT2, src/potaccess.h:323-354); which: 0 pair, 1 embed, 2 rho */
int  imdb200_pair_int(imdb200_sim *sim, int which, int col, long n, const double *r2,
                      double *pot, double *grad);
/* device time (ms) spent in each phase since the last reset: rebuild, pass1, pass2, integrate, comm */
int  imdb200_get_timers(imdb200_sim *sim, double out_ms[8], int reset);

/* ---- host-side helpers kept from IMD's setup layer (plain C, no CUDA) -------------------- */
/* replaces: read_pot_table() + init_threepoint() (src/imd_potential.c:161-462, 1256-1272).
 * default_format: DEFAULT_POTFILE_TYPE (2 in EAM builds, 1 otherwise; src/config.h:57-63). */
int  imdb200_read_pot_table(imdb200_pot_table *pt, const char *filename, int ncols, int radial,
                            int ntypes, int default_format, double *cellsz);
void imdb200_free_pot_table(imdb200_pot_table *pt);

/* replaces: calc_cpu_dim() (src/imd_geom_mpi_3d.c:201-266): factorise num_cpus evenly, largest factor on the
 * axis with the largest requested cpu_dim */
void imdb200_calc_cpu_dim(int num_cpus, int cpu_dim[3]);
/* replaces: MPI_Cart_rank / MPI_Cart_coords on the cpugrid of setup_mpi_topology()
 * (src/imd_geom_mpi_3d.c:43-53): row-major, z fastest */
int  imdb200_cart_rank(const int coord[3], const int cpu_dim[3]);
void imdb200_cart_coords(int rank, const int cpu_dim[3], int coord[3]);
/* replaces: the neighbour ranks nbeast ... nbdwn of setup_mpi_topology() (src/imd_geom_mpi_3d.c:57-88) and the
 * periodic shift vectors of send_cells (src/imd_comm_force_3d.c:248-265).  Direction
 * d = (sx+1) + 3*(sy+1) + 9*(sz+1); peer[d] = -1 behind a free surface; code[d] = image shift, same encoding */
void imdb200_halo_peers(const int cpu_dim[3], const int my_coord[3], const int pbc_dirs[3], int peer[27],
                        int code[27]);
/* replaces: the message schedule of send_cells / send_forces (three sweeps of two messages,
 * src/imd_comm_force_3d.c:268-395, 587-714) by one message per neighbour rank: the order in which the receive
 * regions (recv_order, own-rank wraps included) and the send regions (send_order) of the 26 directions are laid out
 * so that everything exchanged with one peer is one contiguous, identically ordered slice on both sides */
void imdb200_halo_message_order(const int peer[27], int my_rank, int recv_order[26], int *n_recv,
                                int send_order[26], int *n_send);

#ifdef __cplusplus
}
#endif
#endif /* IMD_B200_H */
