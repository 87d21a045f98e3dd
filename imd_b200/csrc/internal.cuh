// internal.cuh -- shared declarations of the sm_100a engine behind include/imd_b200.h.
//
// Device data layout (all arrays live in HBM, one simulation handle per GPU):
//   pos   double4[n_own + n_ghost]  x,y,z + (vsorte<<32 | sorte) bit-cast into .w
//                                   owned atoms first, sorted by cell (x slowest, z fastest like
//                                   PTR_3D_V, src/makros.h:438), inside a cell by atom number;
//                                   ghost (buffer-cell) images after them, grouped by ghost cell
//   mom   double4[n_own]            px,py,pz + mass in .w          (IMPULS, MASSE)
//   frc   double4[n_own]            fx,fy,fz + per-atom Epot in .w (KRAFT, POTENG)
//   rho   double[n_own]             host electron density          (EAM_RHO)
//   dF    double[n_own + n_ghost]   2 F'(rho) of owners and images (EAM_DF)
//   nblpos double[3][n_pad]         reference positions of the skin check (NBL_POS)
//   nbl   int32[n_pad*L/32][max_nb/L][32]  FULL neighbour list in warp blocks: the 32 lanes of a warp
//                                   (L lanes per atom) read one contiguous 128-byte row per iteration
//                                   and a warp's whole list is one contiguous block; nnb int32[n_own]
// 32-byte atom records make every gather exactly one 32-byte sector.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/imd_b200.h"

#define IMDB_MAXCOL 16 // ntypes <= 4
#define IMDB_TWO52 4503599627370496.0

// ---- potential tables in derived (polynomial-coefficient) form ---------------------------------
// PAIR_INT2 (src/potaccess.h:323-354) evaluates on interval k, with chi in [0,1):
//    val  = p0 + chi*dv + 0.5*chi*(chi-1)*d2v,   grad = 2*istep*(dv + (chi-0.5)*d2v)
// with dv = p1-p0, d2v = p2-2*p1+p0.  We precompute per (k,col) on the host, in double:
//    c0 = p0, c1 = dv - 0.5*d2v, c2 = 0.5*d2v
// so that val = c0 + chi*(c1 + chi*c2) and grad = 2*istep*(c1 + 2*chi*c2): 2-4 DFMA instead of
// ~12 DP ops, and 24 bytes per lookup (one 16-byte + one 8-byte fetch) instead of three strided
// ones.  The force kernels are bound by the L1/shared-memory gather pipe (profiles/), so bytes per
// lookup are what counts: the gradient of phi is derived from (c1,c2) in registers, and pass 2
// fetches only a 16-byte (h1,h2) = istep*(c1, 2*c2) pair, grad/2 = h1 + chi*h2.
// k and chi are computed with the reference's operation sequence (see tab_index()).
enum { TAB_PAIR = 0, TAB_EMBED = 1, TAB_RHO = 2 };

struct TabMeta {          // per-column header of a pot_table_t
  double begin[IMDB_MAXCOL], end[IMDB_MAXCOL], invstep[IMDB_MAXCOL];
  int ncols, nrows;       // nrows = rows of the derived table (maxsteps)
};

struct DevTables {
  TabMeta pair, embed, rho;
  const double2 *pairAB;  // [nrows][ncols]  (c0,c1) of phi
  const double  *pairC;   // [nrows][ncols]  c2 of phi
  const double2 *rhoAB;   // [nrows][ncols]  (c0,c1) of rho           (pass 1)
  const double  *rhoC;    // [nrows][ncols]  c2 of rho
  const double2 *rhoH;    // [nrows][ncols]  (h1,h2): rho'/2 = h1+chi*h2 (pass 2; the 0.5 of :1203 folded in)
  const double  *embedVG; // [nrows][ntypes][8]  c0 c1 c2 c3 g1 g2 g3 -  value+grad of F (once per atom)
  // cubic interpolation (IMDB200_INTERP_4POINT / _SPLINE): val = c0+chi*(c1+chi*(c2+chi*c3)); pairC/rhoC unused
  const double2 *pairCD;  // [nrows][ncols]  (c2,c3) of phi
  const double2 *rhoCD;   // [nrows][ncols]  (c2,c3) of rho
  const double  *rhoH3;   // [nrows][ncols]  h3: rho'/2 = h1+chi*(h2+chi*h3)
  int cubic;
  // EEAM (extended EAM): energy modification term M(p), p = sum rho^2 -- same layout as embedVG
  TabMeta emod;
  const double  *emodVG;
  int have_eeam;
  // ADP: dipole u(r) and quadrupole w(r) distortion functions, (c0,c1,c2,c3) per interval and column (forces_adp.cu)
  TabMeta adpu, adpw;
  const double4 *adpuK, *adpwK;
  int have_adp;
  const double2 *fused;   // single species, phi and rho on one r^2 grid: [nrows][3] = (phi c0,c1) (phi c2, rho c2) (rho c0,c1),
                          // one 48-byte record per interval = three 16-byte loads per pair in pass 1
  int fused_rows;
  const double2 *fraw;    // the same single-species case as raw samples: [fused_rows+2] = (phi_k, rho_k), 16 bytes per row
  int smem1_raw;
  int have_eam, shared_grid, ntypes;
  int smem1, smem2;       // dynamic shared memory (bytes) to stage the pass-1 / pass-2 tables; 0 = leave in HBM/L1
  // Several species, quadratic interpolation: pass 1 stages the RAW samples of the DISTINCT columns (phi_AB == phi_BA,
  // rho depends on the neighbour's type only in standard EAM: 3 + 2 instead of 4 + 4 columns for two species), 8 bytes
  // per sample instead of 24 per interval -- 80 KB for the 2001-row Ni-Al tables, which do not fit in coefficient form.
  // The kernel forms dv, d2v from rows k, k+1, k+2 like PAIR_INT2 itself (src/potaccess.h:345-349).
  const double *rawP, *rawR;          // [nrows+2][nuP], [nrows+2][nuR]
  int nuP, nuR, raw_ok;               // distinct columns; raw_ok: smem1 holds the raw layout
  signed char umapP[IMDB_MAXCOL], umapR[IMDB_MAXCOL];
  // Several species, pass 2: (h1,h2) of the DISTINCT rho columns, [nrows][nuR], staged in shared memory (smem2m bytes).
  // multi_uniform: every column of the pair table has the same begin / end / invstep, and so has every column of the rho
  // table -- the condition for the shared-memory paths of both passes (raw_ok, smem2m), which keep the headers in registers
  const double2 *rhoHd;
  int multi_uniform, smem2m;
};

// ---- geometry ------------------------------------------------------------------------------------
struct Geom {
  double box[3][3];       // box_x, box_y, box_z
  double tbox[3][3];      // reciprocal vectors, SPROD(box_i,tbox_j) = delta_ij (make_box)
  int pbc[3];
  int gdim[3];            // global_cell_dim
  int cdim[3];            // local cell_dim incl. the two buffer layers
  int coff[3];            // my_coord * (cell_dim-2): global cell coordinate of local cell 1
  int nall;               // cdim.x*cdim.y*cdim.z
  double cellsz;          // (r_cut_max + nbl_margin)^2
};

// one buffer (ghost) cell of the local grid and where its content comes from
struct GhostCell { int dst, src; int code; int peer; };
// code = (sx+1) + 3*(sy+1) + 9*(sz+1), s in {-1,0,1}: image shift in box units; src/peer: the source
// cell on this rank (peer == own rank) or -1 when the content arrives from another rank

// One of the 26 directions d = (sx+1) + 3*(sy+1) + 9*(sz+1) of the halo.  Receive side: the buffer cells
// on face/edge/corner d are filled from rank `peer` (possibly this rank: periodic wrap on one GPU).
// Send side: the owned cells adjacent to face d go to `peer`, where they fill direction 26-d.
struct DirPlan {
  int peer;                       // -1: no neighbour (non-periodic edge)
  int ncells;                     // cells in the region (same count on both sides)
  int gcell_off;                  // first entry in gcells (receive cells), direction-major
  int scell_off;                  // first entry in scells (send cells), -1 if peer is self or none
  int recv_off, recv_cnt;         // ghost atoms [n_own+recv_off, +recv_cnt) after a rebuild
  int send_off, send_cnt;         // slice of send_idx / the send buffers
};

// Everything exchanged with one neighbour rank: the regions of all directions that lead to it are adjacent
// (comm_plan), so each exchange is one message per peer.
struct PeerPlan {
  int peer;
  int ncells, gcell_off, scell_off;   // slices of gcells / scells (fixed by the cell grid)
  int recv_off, recv_cnt;             // ghost atoms [n_own+recv_off, +recv_cnt) after a rebuild
  int send_off, send_cnt;             // slice of send_idx / the send buffers
};

struct imdb200_sim {
  imdb200_config cfg;
  Geom geom;
  double height[3], min_height[3], max_height[3], volume, volume_init;
  double cellsz0;                 // max table end (r^2) before the margin is added
  DevTables tabs;
  void *tab_mem[12];              // device allocations behind DevTables
  int have_tabs;
  cudaStream_t stream; int own_stream;
  // atoms
  long n_own, n_ghost, cap_atoms; // capacity of the per-atom arrays
  long long natoms_global;        // sum of n_own over all ranks
  int need_filter;                // set_atoms was given atoms of other domains too: drop them at the first binning
  double4 *pos, *pos_alt, *mom, *mom_alt, *frc;
  int *nummer, *nummer_alt;
  double *rho, *dF, *nblpos, *presstens; // presstens [6][cap] SoA
  double *eam_p, *dM;             // EEAM: p_i = sum rho^2 (owners) and M'(p_i) (owners and images)
  double *adp_mu, *adp_la; long adp_cap;   // ADP: mu [3][adp_cap], lambda [6][adp_cap] (xx yy zz yz zx xy), owners and images
  double4 *posdf;                 // single-species EAM: x,y,z + 2F'(rho) in .w, the pass-2 gather record
  int *cellid, *cellid_alt, *perm;
  void *xfer; size_t xfer_bytes;  // staging for set_atoms / get_atoms
  // cells; cell_count/cell_start have nall + NBIN_EXTRA entries: 27 bins for atoms leaving in direction d
  // and one for atoms of foreign domains dropped by the first binning
  int *cell_count, *cell_start, *cell_fill, *cell_code;
  GhostCell *gcells; int n_gcells;
  int *gcount, *gstart;
  int *gsrc;                      // per ghost atom: source atom on this rank, -1 = arrives from a peer
  int *ghost_num;                 // per ghost atom: NUMMER (test hook / diagnostics)
  double4 *ghost_raw;             // per ghost atom: unshifted position as the owner holds it
  int *scan_tmp;
  // halo plan (comm.cu)
  DirPlan dir[27];
  PeerPlan peers[26]; int n_peers;
  int *scells; int n_scells;      // owned cells to send, direction-major
  int *scount, *sstart;
  int *send_idx; long n_send, cap_send;
  double4 *sendbuf4; double *sendbuf1; int *sendbufi;
  int *h_starts;                  // pinned: gstart/sstart read back at a rebuild
  // neighbour list
  int *nbl, *nnb; long nbl_cap_rows; int max_nb, lanes; long n_pad;
  unsigned long long *nnbc;       // per atom: cumulative entry counts per skin class, 12 bits each (see NBL_CLASSES)
  float4 *posf;                   // single-precision copy of pos, pre-filter of the list build only
  int have_valid_nbl, nbl_count; long long nbl_len;
  int cell_words;                 // 32-bit candidate words per cell in the list build (cells of up to 32*cell_words atoms)
  int skin_skip;                  // 1: the force kernels skip list entries that cannot be inside the cut-off yet
  int skin_all;                   // box/positions changed outside move_atoms since the build: use every entry
  double disp2;                   // max squared displacement of the current positions since the build; <0 unknown
  // pos / posdf as linear textures for the TEX share of the gathers (forces.cu)
  cudaTextureObject_t tex_pos, tex_posdf; const void *tex_pos_ptr, *tex_posdf_ptr; size_t tex_pos_bytes, tex_posdf_bytes;
  int tex_ok;
  // restrictions / deformation tables per virtual type
  double *restr; int n_restr;
  // scalars
  double *d_scal;      // device scalar block of this rank, see SC_*
  double *d_glob;      // the same summed over ranks (== d_scal on one rank)
  double *d_all;       // [nranks][SC_COUNT] all-gather target
  double *h_scal;      // pinned mirror of d_glob
  double *d_partial; size_t partial_bytes;   // per-block partial sums
  int *d_flags, *h_flags;
  // device-side step control (api.cu, run_async): gate = 0 turns every kernel of a step into a no-op (a step that was
  // queued before the host knew that the previous one invalidated the list); disp2 = max squared displacement since the
  // list build as of the end of the previous step, from which the force kernels pick the skin classes to walk
  struct StepCtl *d_ctl;
  // Boundary-first processing order of the force kernels (multi-GPU): worder lists the warps (32 thread slots each)
  // whose atoms lie in the outermost layer of owned cells first, the interior ones behind them.  The boundary part
  // of a pass runs as its own launch; what it produces is stored into the neighbours' ghost regions over NVLink
  // (comm_p2p.cu) while the interior part runs, and the consumer waits for it only where it needs the images.
  int *worder, *wflag, *wscan; long n_warps, n_bwarps, worder_cap;
  int split_part;                    // which part the next force launch covers: 0 all, 1 boundary, 2 interior
  int pos_sent_early;                // the boundary positions of the next step are already on their way (pass 2 of this one)
  int maxd2_zeroed;                  // SC_MAXD2 was cleared by the reduction kernel behind pass 2 (zero_before_move)
  int zero_before_move;              // imdb200_run, unfused step: ask that reduction kernel to clear it
  int fuse_step;                     // the step being queued runs move_atoms in the tail of pass 2
  struct StepSlot *d_slot, *h_slot;   // per-step results, ring of 3 (device + pinned mirror)
  cudaEvent_t ev_slot[3];
  int press_calc, is_short, short_warned;
  long long nactive;   // sum of the restriction components over all atoms (3N by default)
  int nactive_dirty;   // atoms or restrictions changed: recount at the next rebuild / step
  // NPT_iso: barostat friction, twice the global kinetic energy after the last step (< 0: unknown), external pressure,
  // the pressure the last step used
  double npt_xi, npt_ekin_old, npt_pressure_ext, npt_pressure;
  // NPT_axial: per-axis xi, stress of the last step, pressure_ext and its increment, dyn_stress the last step left, relax_dirs
  double ax_xi[3], ax_stress[3], ax_pext[3], ax_dpext[3], ax_dyn[3]; int ax_relax[3];
  double eta;
  double tauber;       // > 0: Berendsen variant of NVE (imdb200_set_berendsen)
  // timers
  cudaEvent_t ev[16];
  double t_ms[8];
  // comm
  void *nccl_comm; int rank, nranks;
  // halo exchange over peer memory (comm_p2p.cu; off unless IMDB200_HALO_P2P=1)
  int p2p_on, p2p_step; void *p2p;   // p2p_step: this exchange belongs to a step without a list build
};

// The list of an atom is stored in NBL_CLASSES+1 groups by the pair distance r_b at build time:
// group 0: r_b <= rc (largest table cut-off), group q: rc+(q-1)w < r_b <= rc+q*w with w = nbl_margin/NBL_CLASSES.
// A pair of group q can only be inside the cut-off once 2*max|displacement| > (q-1)*w, so a force call walks the
// groups 0..cls only (cls from the last skin check); the entries it leaves out would all fail the r2 test.
#define NBL_CLASSES 4
// A list entry is the index of the neighbour (own atom or image, < 2^29, checked in launch_build) with the neighbour's TYPE in
// bits 29-30 (ntypes <= 4): the several-species force kernels need the type to pick the table column, and with it in the entry
// pass 2 gathers one 32-byte record (x, y, z, F') per neighbour instead of a position record plus F'.  Single species: plain index.
#define NBL_TSHIFT 29
#define NBL_JMASK 0x1fffffff
#define NBL_CBITS 12

#define NBIN_EXTRA 29   // 27 leave directions + 1 dropped + 1 spare (exclusive-scan total)

enum { SC_EPOT = 0, SC_VIRIAL, SC_EKIN, SC_EKIN2, SC_MAXD2, SC_ETA, SC_PXX, SC_PYY, SC_PZZ, SC_PYZ, SC_PZX,
       SC_PXY, SC_EKIN1, SC_SHORT, SC_DYNX, SC_DYNY, SC_DYNZ, SC_COUNT = 20 };   // SC_SHORT: is_short of any rank (max);
       // SC_DYN*: dyn_stress_x/y/z of NPT_axial; the last slot is scratch
enum { FL_SHORT = 0, FL_NBL_OVERFLOW, FL_MAXNB, FL_NGHOST, FL_LOST, FL_NSEND, FL_BADTYPE, FL_CELLFULL, FL_COUNT = 8 };

struct StepCtl { int gate; int pad; double disp2; };
struct StepSlot { int executed, valid; int flags[FL_COUNT]; double scal[SC_COUNT]; };

// ---- error handling --------------------------------------------------------------------------------
int imdb_fail(int code, const char *fmt, ...);
extern long long g_kernel_launches;
#define CUDA_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
  return imdb_fail(IMDB200_ERR_CUDA, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)
#define LAUNCH_CHECK() do { g_kernel_launches++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) \
  return imdb_fail(IMDB200_ERR_CUDA, "%s:%d: kernel launch: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); } while (0)
#define TRY(call) do { int r_ = (call); if (r_) return r_; } while (0)

static inline int cdiv(long a, long b) { return (int) ((a + b - 1) / b); }

// position of entry m of atom i in the warp-blocked neighbour list (L lanes per atom, R = max_nb/L rows)
__host__ __device__ inline size_t nbl_index(long i, int m, int L, int R)
{
  const long slot = i * L + (m % L);
  return (size_t) (slot >> 5) * ((size_t) R * 32) + (size_t) (m / L) * 32 + (size_t) (slot & 31);
}

// ---- cross-file entry points -------------------------------------------------------------------------
int tables_upload(imdb200_sim *s, const imdb200_pot_table *pair, const imdb200_pot_table *embed,
                  const imdb200_pot_table *rho);
void tables_free(imdb200_sim *s);
void tables_set_cellsz0(imdb200_sim *s, double cz);   // new interaction range: cell grid and list cut-off are re-derived
int tables_pair_int(imdb200_sim *s, int which, int col, long n, const double *r2, double *pot, double *grad);
int tables_upload_emod(imdb200_sim *s, const imdb200_pot_table *emod);   // EEAM energy modification term
int tables_upload_adp(imdb200_sim *s, const imdb200_pot_table *u, const imdb200_pot_table *w);   // ADP u(r), w(r)

int geom_make_box(imdb200_sim *s);           // make_box + init_cells when needed
int cells_ensure_capacity(imdb200_sim *s, long n_atoms_total);
int cells_rebuild(imdb200_sim *s);            // fix_cells + ghost images + make_nblist
int cells_count_nactive(imdb200_sim *s);     // nactive from the virtual types and restriction vectors (collective)
int scan_exclusive(imdb200_sim *s, const int *in, int *out, int n, int *total_dev);

int comm_plan(imdb200_sim *s);                // halo plan for the current cell grid (after init_cells)
void comm_free(imdb200_sim *s);
int comm_migrate(imdb200_sim *s, const int *h_counts, long n_stay, long *n_new);  // send_atoms (fix_cells)
int comm_setup_ghosts(imdb200_sim *s);        // per-cell counts, ghost ranges, send lists (at a rebuild)
int comm_ghost_pos(imdb200_sim *s);           // send_cells(copy_cell,pack_cell,unpack_cell)
int comm_ghost_dF(imdb200_sim *s);            // send_cells(copy_dF,pack_dF,unpack_dF)
int comm_ghost_field(imdb200_sim *s, double *field, int ncomp, long stride);   // owners -> images for field[c*stride + i], c < ncomp <= 8
int comm_ghost_dM(imdb200_sim *s);            // the EAM_DM part of copy_dF in EEAM builds (src/imd_comm_force_3d.c:1044-1046)
int comm_reverse_add(imdb200_sim *s, double *field, int ncomp, long stride);  // send_forces(add_*,...)
int comm_sync_scalars(imdb200_sim *s);        // the MPI_Allreduce sites
extern "C" int comm_p2p_enable(imdb200_sim *s);
void comm_p2p_free(imdb200_sim *s);
int comm_p2p_defer_free(imdb200_sim *s, void *ptr);   // 1: the peer-memory halo keeps ptr until its importers have unmapped it
int comm_p2p_setup(imdb200_sim *s, int (*allgather)(imdb200_sim *, const void *, void *, size_t));
int comm_p2p_ready(const imdb200_sim *s);
int comm_p2p_positions(imdb200_sim *s);
int comm_p2p_dF(imdb200_sim *s);
int comm_p2p_begin(imdb200_sim *s, int kind);     // kind 0 positions, 1 F': store my boundary values into the peers' ghost regions, signal
int comm_p2p_end(imdb200_sim *s, int kind);       // wait until every peer has done the same towards me
int comm_ghost_pos_finish(imdb200_sim *s);    // images from the raw copies (after comm_p2p_end(s, 0))
int comm_ghost_dF_finish(imdb200_sim *s);
int forces_split_possible(const imdb200_sim *s);
int cells_build_worder(imdb200_sim *s);
int comm_allgather_ll(imdb200_sim *s, long long mine, long long *all);

int forces_pass1(imdb200_sim *s);             // pair + rho + embedding
int forces_pass2(imdb200_sim *s, int fuse_move);   // EAM force pass; fuse_move: move_atoms + check_nblist in its tail
int forces_can_fuse_move(const imdb200_sim *s);
int forces_adp_pass1(imdb200_sim *s);         // ADP: mu, lambda and the ADP energy (after pass 1)
int forces_adp_halo(imdb200_sim *s);          // ADP: mu, lambda of the owners into the images
int forces_adp_pass2(imdb200_sim *s);         // ADP: dipole and quadrupole forces (after pass 2)
int adp_ensure_arrays(imdb200_sim *s);
int forces_textures(imdb200_sim *s);
void forces_free_textures(imdb200_sim *s);
int forces_pass1_quad(imdb200_sim *s);        // forces.cu is compiled once per interpolation order (quadratic / cubic)
int forces_pass2_quad(imdb200_sim *s, int fuse_move);
int forces_pass1_cubic(imdb200_sim *s);
int forces_pass2_cubic(imdb200_sim *s, int fuse_move);
int forces_pass1_quad_eeam(imdb200_sim *s);
int forces_pass2_quad_eeam(imdb200_sim *s, int fuse_move);
int forces_pass1_cubic_eeam(imdb200_sim *s);
int forces_pass2_cubic_eeam(imdb200_sim *s, int fuse_move);
int integrate_finish(imdb200_sim *s, int nblocks_move);   // reductions + Nose-Hoover update after the per-atom part
int integrate_move(imdb200_sim *s);           // move_atoms_nve/nvt + check_nblist fused
int integrate_move_npt(imdb200_sim *s);       // move_atoms_npt_iso + check_nblist fused; integrate_npt_after_fetch follows the scalar fetch
int integrate_npt_dyn_pressure(imdb200_sim *s);   // calc_dyn_pressure into SC_EKIN2 (local share)
int integrate_npt_after_fetch(imdb200_sim *s);
int integrate_axial_dyn_pressure(imdb200_sim *s);   // calc_dyn_pressure per axis into SC_DYN*, SC_EKIN2 (local share)
int integrate_axial_virial(imdb200_sim *s);         // vir_xx/yy/zz = sums of the per-atom stress after calc_forces, into SC_PXX..
int integrate_move_axial(imdb200_sim *s);           // move_atoms_npt_axial + check_nblist; integrate_axial_after_fetch follows the fetch
int integrate_axial_after_fetch(imdb200_sim *s);
int reduce_finish(imdb200_sim *s, int nblocks, int nvals, const int *slots, int accumulate_mask, int zero_maxd2 = 0);
int step_snapshot_disp2(imdb200_sim *s, int reset);   // StepCtl::disp2 <- SC_MAXD2 of the global block (reset: 0)

// ---- device helpers ------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ int sorte_of(double w) { return (int) (__double_as_longlong(w) & 0xffffffffLL); }
__device__ __forceinline__ int vsorte_of(double w) { return (int) (__double_as_longlong(w) >> 32); }
__device__ __forceinline__ double pack_types(int sorte, int vsorte)
{ return __longlong_as_double(((long long) vsorte << 32) | (unsigned int) sorte); }

// r2 = SPROD(d,d) with the reference's rounding: ((dx*dx)+(dy*dy))+(dz*dz), no FMA
// (src/makros.h:409, src/imd_forces_nbl.c:252-258).
__device__ __forceinline__ double r2_exact(double dx, double dy, double dz)
{ return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)); }

// Table index exactly as PAIR_INT2 computes it (src/potaccess.h:329-341):
//   r2a = MIN(r2,end) - begin; if (r2a<0) {r2a=0; is_short=1;}  r2a *= invstep; k=(int)r2a; chi=r2a-k
// (int) truncation of the non-negative r2a is done on the FP64 pipe: adding 2^52 with
// round-toward-zero leaves floor(r2a) in the low mantissa bits, and subtracting 2^52 again gives
// it back as a double -- both exact -- instead of the slow F2I/I2F conversion pipe.
__device__ __forceinline__ void tab_index(double r2, double begin, double end, double invstep,
                                          int &k, double &chi, int &is_short)
{
  double r2a = fmin(r2, end) - begin;
  if (r2a < 0.0) { r2a = 0.0; is_short = 1; }
  r2a = __dmul_rn(r2a, invstep);
  double tk = __dadd_rz(r2a, IMDB_TWO52);
  k = __double2loint(tk);
  chi = r2a - (tk - IMDB_TWO52);
}

// Same index for the force kernels: r2 is known to be inside the column's cut-off (the MIN clamp is
// inactive) and the subtraction and scaling are contracted into one DFMA.  k and chi can differ from
// tab_index() only when r2a*invstep is within an ulp of an integer; the interpolant is continuous
// there, so values move by ~1e-16 relative (parity bar 1e-10).
__device__ __forceinline__ void tab_index_fast(double r2, double nbegin_istep, double invstep, int &k, double &chi,
                                               int &is_short)
{
  double t = fma(r2, invstep, nbegin_istep);     // (r2 - begin) * invstep
  if (t < 0.0) { t = 0.0; is_short = 1; }
  const double tk = __dadd_rz(t, IMDB_TWO52);
  k = __double2loint(tk);
  chi = t - (tk - IMDB_TWO52);
}

// val = c0 + chi*(c1 + chi*c2); grad = G*(c1 + 2*chi*c2) with G = 2*invstep
__device__ __forceinline__ double tab_val(double2 ab, double c2, double chi) { return fma(chi, fma(chi, c2, ab.y), ab.x); }
__device__ __forceinline__ double tab_grad(double2 ab, double c2, double chi, double G) { return G * fma(chi + chi, c2, ab.y); }
// cubic modes: val = c0 + chi*(c1 + chi*(c2 + chi*c3)); grad = G*(c1 + chi*(2*c2 + 3*c3*chi))
__device__ __forceinline__ double tab_val3(double2 ab, double2 cd, double chi)
{ return fma(chi, fma(chi, fma(chi, cd.y, cd.x), ab.y), ab.x); }
__device__ __forceinline__ double tab_grad3(double2 ab, double2 cd, double chi, double G)
{ return G * fma(chi, fma(3.0 * chi, cd.y, cd.x + cd.x), ab.y); }

// copy_cell (src/imd_comm_force_3d.c:726-778) for all three sweeps at once.  The reference adds
// the box vectors stage by stage (up/down, then north/south, then east/west; :268-395), so the
// image position is ((x + sz*box_z) + sy*box_y) + sx*box_x, each add rounded.
__device__ __forceinline__ double4 image_pos(double4 p, int code, const Geom &g)
{
  int sx = code % 3 - 1, sy = (code / 3) % 3 - 1, sz = code / 9 - 1;
  if (sz) { double f = (double) sz; p.x = __dadd_rn(p.x, f * g.box[2][0]); p.y = __dadd_rn(p.y, f * g.box[2][1]); p.z = __dadd_rn(p.z, f * g.box[2][2]); }
  if (sy) { double f = (double) sy; p.x = __dadd_rn(p.x, f * g.box[1][0]); p.y = __dadd_rn(p.y, f * g.box[1][1]); p.z = __dadd_rn(p.z, f * g.box[1][2]); }
  if (sx) { double f = (double) sx; p.x = __dadd_rn(p.x, f * g.box[0][0]); p.y = __dadd_rn(p.y, f * g.box[0][1]); p.z = __dadd_rn(p.z, f * g.box[0][2]); }
  return p;
}

// a kernel of a queued step whose list turned out invalid does nothing (StepCtl::gate)
#define STEP_GATE(ctl) do { if ((ctl) != nullptr && (ctl)->gate == 0) return; } while (0)

// Highest list group a force call has to walk (NBL_CLASSES above).  Group q >= 1 holds pairs with build distance
// r_b > rc + (q-1) w; |r - r_b| <= 2 dmax, so they are out of reach while 2 dmax <= (q-1) w.  d2 = dmax^2 (< 0: unknown).
__device__ __forceinline__ int skin_class_of(double d2, double w)
{
  if (d2 == 0.0) return 0;
  if (!(d2 > 0.0)) return NBL_CLASSES;
  const double reach = 2.0 * sqrt(d2) * (1.0 + 1e-9) + 1e-12;
  const int c = (int) floor(reach / w) + 1;
  return c < NBL_CLASSES ? c : NBL_CLASSES;
}

__device__ __forceinline__ double2 ld2(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }

// one 32-byte atom record = one sector, fetched with a single 256-bit load (LDG.E.ENL2.256, sm_100)
__device__ __forceinline__ double4 ld_atom(const double4 *p)
{
  double4 v;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One atom of move_atoms_nve / move_atoms_nvt (src/imd_integrate.c:192-217, 328-358, 907-908, 1020-1027) and of
// check_nblist (src/imd_forces_nbl.c:2007-2037).  Shared by k_move_atoms and by the fused tail of k_pass2 so that
// imdb200_run and the separate calls produce bit-identical trajectories.  f is the (restricted) force, rx..rz the
// restriction vector of the atom's virtual type, red[0..1] the kinetic-energy partial sums, returns |x - nbl_pos|^2.
// Berendsen variant of NVE (`ber` builds, src/imd_integrate.c:44-53): scale factor of the momenta from the kinetic
// energy the PREVIOUS move_atoms left; tauber <= 0: plain NVE (factor 1).
__host__ __device__ inline double berendsen_cc(double tot_kin_energy, double nactive, double temperature, double dt, double tauber)
{
  if (!(tauber > 0.0)) return 1.0;
  const double kb = 8.6174101569719990e-06;
  double cc = 1. - dt / tauber * ((2.0 * tot_kin_energy / nactive + kb) / (temperature + kb) - 1.);
  if (cc < 0.5) cc = 0.5; else if (cc > 2.0) cc = 2.0;
  return sqrt(cc);
}

template <bool NVT>
__device__ __forceinline__ double integrate_atom(double4 &x, double4 &p, const double4 &f, double dt, double eta,
                                                 double rx, double ry, double rz, double nx, double ny, double nz,
                                                 double (&red)[2], double cc = 1.0)
{
  const double m = p.w;
  if (!NVT) {
    const double k1 = p.x * p.x + p.y * p.y + p.z * p.z;
    p.x += dt * f.x; p.y += dt * f.y; p.z += dt * f.z;              // :213-217
    const double k2 = p.x * p.x + p.y * p.y + p.z * p.z;
    red[0] = (k1 + k2) / (4 * m);                                   // :329-335
    red[1] = 0.0;
    p.x *= cc; p.y *= cc; p.z *= cc;                                // BER :341-350 (after the energy, before the move)
  } else {
    const double reibung = 1.0 - eta * dt / 2.0;                    // :907
    const double eins_d_reib = 1.0 / (1.0 + eta * dt / 2.0);        // :908
    red[0] = (p.x * p.x + p.y * p.y + p.z * p.z) / m;               // E_kin_1 :951
    p.x = (p.x * reibung + dt * f.x) * eins_d_reib * rx;            // :1020-1027
    p.y = (p.y * reibung + dt * f.y) * eins_d_reib * ry;
    p.z = (p.z * reibung + dt * f.z) * eins_d_reib * rz;
    red[1] = (p.x * p.x + p.y * p.y + p.z * p.z) / m;               // E_kin_2
  }
  const double tmp = dt / m;                                         // :353-358
  x.x += tmp * p.x; x.y += tmp * p.y; x.z += tmp * p.z;
  // check_nblist: same operands and rounding as the reference (exact -> identical rebuild steps)
  return r2_exact(x.x - nx, x.y - ny, x.z - nz);
}

// Block-wide sum of NV values; thread 0 of the block writes them to partial[blockIdx.x*NV + v].
template <int NV> __device__ __forceinline__ void block_sum_store(double (&v)[NV], double *partial)
{
  __shared__ double sm[32][NV];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; i++) v[i] = warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; i++) sm[w][i] = v[i];
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int i = 0; i < NV; i++) {
      double x = (lane < nw) ? sm[lane][i] : 0.0;
      x = warp_sum(x);
      if (lane == 0) partial[(size_t) blockIdx.x * NV + i] = x;
    }
  }
}
#endif
