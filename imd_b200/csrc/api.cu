// api.cu -- the C ABI of include/imd_b200.h: handle life cycle, host<->device transfers and the
// step loop of main_loop (src/imd_main_3d.c:155-870) kept resident on the device.
#include "internal.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>

int integrate_check_nblist(imdb200_sim *s);
int integrate_lin_deform(imdb200_sim *s, const double dx[3], const double dy[3], const double dz[3], double scale);
int integrate_deform_sample(imdb200_sim *s, int nvt, double size, const double *shift, const int *shear_def,
                            const double *shear, const double *base);

long long g_kernel_launches = 0;
static thread_local char g_err[1024] = "";
static void (*g_handler)(const char *) = nullptr;

int imdb_fail(int code, const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  if (g_handler) g_handler(g_err);
  return code;
}

// StepCtl::disp2 <- max squared displacement since the list build (the global SC_MAXD2), or 0 right after a build
__global__ void k_snapshot_disp2(StepCtl *ctl, const double *glob, int reset) { ctl->disp2 = reset ? 0.0 : glob[SC_MAXD2]; }

int step_snapshot_disp2(imdb200_sim *s, int reset)
{
  k_snapshot_disp2<<<1, 1, 0, s->stream>>>(s->d_ctl, s->d_glob, reset);
  LAUNCH_CHECK();
  return 0;
}

extern "C" {

void imdb200_destroy(imdb200_sim *s);
const char *imdb200_last_error(void) { return g_err; }
void imdb200_set_error_handler(void (*h)(const char *)) { g_handler = h; }
long long imdb200_kernel_launches(void) { return g_kernel_launches; }
int imdb200_device_count(void) { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0; }

void imdb200_default_config(imdb200_config *c)
{
  memset(c, 0, sizeof(*c));
  c->ntypes = 1; c->total_types = 1;
  c->pbc_dirs[0] = c->pbc_dirs[1] = c->pbc_dirs[2] = 1;
  c->cpu_dim[0] = c->cpu_dim[1] = c->cpu_dim[2] = 1;
  c->nbl_margin = 0.4;   /* src/globals.h:419 */
  c->nbl_size = 1.1;     /* src/globals.h:420 */
  c->ensemble = IMDB200_ENS_NVE;
  c->device = -1;
  c->lanes_per_atom = 0;
}

static int create_device_state(imdb200_sim *s)
{
  CUDA_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  s->own_stream = 1;
  CUDA_TRY(cudaMalloc(&s->d_scal, SC_COUNT * sizeof(double)));
  CUDA_TRY(cudaMemset(s->d_scal, 0, SC_COUNT * sizeof(double)));
  CUDA_TRY(cudaMemcpy(s->d_scal + SC_ETA, &s->eta, sizeof(double), cudaMemcpyHostToDevice));
  s->d_glob = s->d_scal;            // one rank: the local block is the global one
  CUDA_TRY(cudaMallocHost(&s->h_scal, SC_COUNT * sizeof(double)));
  memset(s->h_scal, 0, SC_COUNT * sizeof(double));
  CUDA_TRY(cudaMalloc(&s->d_flags, FL_COUNT * sizeof(int)));
  CUDA_TRY(cudaMemset(s->d_flags, 0, FL_COUNT * sizeof(int)));
  CUDA_TRY(cudaMallocHost(&s->h_flags, FL_COUNT * sizeof(int)));
  for (int i = 0; i < 16; i++) CUDA_TRY(cudaEventCreate(&s->ev[i]));
  const StepCtl ctl0 = {1, 0, -1.0};               // gate open, displacement unknown
  CUDA_TRY(cudaMalloc(&s->d_ctl, sizeof(StepCtl)));
  CUDA_TRY(cudaMemcpy(s->d_ctl, &ctl0, sizeof(ctl0), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(&s->d_slot, 3 * sizeof(StepSlot)));
  CUDA_TRY(cudaMallocHost(&s->h_slot, 3 * sizeof(StepSlot)));
  for (int i = 0; i < 3; i++) CUDA_TRY(cudaEventCreateWithFlags(&s->ev_slot[i], cudaEventDisableTiming));
  return 0;
}

int imdb200_create(const imdb200_config *cfg, imdb200_sim **out)
{
  if (!cfg || !out) return imdb_fail(IMDB200_ERR_ARG, "null argument");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return imdb_fail(IMDB200_ERR_CUDA, "no CUDA device available (%s); imd_b200 has no CPU path", cudaGetErrorString(e));
  if (cfg->device >= 0) CUDA_TRY(cudaSetDevice(cfg->device));
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return imdb_fail(IMDB200_ERR_CUDA, "device %d (%s) is sm_%d%d; this library is built for sm_100a only", dev, prop.name, prop.major, prop.minor);
  if (cfg->ntypes < 1 || cfg->ntypes * cfg->ntypes > IMDB_MAXCOL) return imdb_fail(IMDB200_ERR_ARG, "ntypes must be 1..4");
  imdb200_sim *s = (imdb200_sim *) calloc(1, sizeof(imdb200_sim));
  if (!s) return imdb_fail(IMDB200_ERR_ARG, "out of host memory");
  s->cfg = *cfg;
  s->cfg.device = dev;
  if (s->cfg.total_types < s->cfg.ntypes) s->cfg.total_types = s->cfg.ntypes;
  if (s->cfg.nbl_size < 1.0) s->cfg.nbl_size = 1.1;
  for (int d = 0; d < 3; d++) {
    s->geom.box[0][d] = cfg->box_x[d]; s->geom.box[1][d] = cfg->box_y[d]; s->geom.box[2][d] = cfg->box_z[d];
    s->geom.pbc[d] = cfg->pbc_dirs[d];
    if (cfg->cpu_dim[d] < 1) { free(s); return imdb_fail(IMDB200_ERR_ARG, "cpu_dim must be >= 1"); }
  }
  s->nranks = cfg->cpu_dim[0] * cfg->cpu_dim[1] * cfg->cpu_dim[2];
  s->rank = (cfg->my_coord[0] * cfg->cpu_dim[1] + cfg->my_coord[1]) * cfg->cpu_dim[2] + cfg->my_coord[2];
  s->eta = cfg->eta;
  s->skin_skip = 1; s->disp2 = -1.0;
  s->npt_xi = cfg->xi; s->npt_ekin_old = -1.0; s->npt_pressure_ext = cfg->pressure_ext;
  for (int d = 0; d < 3; d++) { s->ax_xi[d] = cfg->xi; s->ax_pext[d] = cfg->pressure_ext; s->ax_dpext[d] = cfg->d_pressure; s->ax_relax[d] = 1; }
  if (cfg->ensemble == IMDB200_ENS_NPT_AXIAL) s->press_calc = 1;   // see imdb200_set_press_calc
  const int rc = create_device_state(s);
  if (rc) { imdb200_destroy(s); return rc; }       // nothing of a half-built handle is left behind
  *out = s;
  return 0;
}

void imdb200_destroy(imdb200_sim *s)
{
  if (!s) return;
  cudaSetDevice(s->cfg.device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  tables_free(s);
  comm_free(s);
  forces_free_textures(s);
  void *ptrs[] = {s->posdf, s->pos, s->pos_alt, s->mom, s->mom_alt, s->frc, s->nummer, s->nummer_alt, s->rho, s->dF,
                  s->nblpos, s->presstens, s->cellid, s->cellid_alt, s->perm, s->cell_count, s->cell_start,
                  s->cell_fill, s->cell_code, s->gsrc, s->ghost_num, s->ghost_raw, s->scan_tmp, s->nbl, s->nnb,
                  s->restr, s->d_scal, s->d_partial, s->d_flags, s->xfer, s->nnbc, s->posf, s->eam_p, s->dM, s->adp_mu, s->adp_la};
  for (void *p : ptrs) if (p) cudaFree(p);
  if (s->h_scal) cudaFreeHost(s->h_scal);
  if (s->h_flags) cudaFreeHost(s->h_flags);
  if (s->d_ctl) cudaFree(s->d_ctl);
  if (s->d_slot) cudaFree(s->d_slot);
  if (s->h_slot) cudaFreeHost(s->h_slot);
  for (int i = 0; i < 3; i++) if (s->ev_slot[i]) cudaEventDestroy(s->ev_slot[i]);
  for (int i = 0; i < 16; i++) if (s->ev[i]) cudaEventDestroy(s->ev[i]);
  if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
  free(s);
}

int imdb200_set_potentials(imdb200_sim *s, const imdb200_pot_table *pair, const imdb200_pot_table *embed,
                           const imdb200_pot_table *rho)
{
  if (!s) return imdb_fail(IMDB200_ERR_ARG, "null handle");
  CUDA_TRY(cudaSetDevice(s->cfg.device));
  TRY(tables_upload(s, pair, embed, rho));
  return geom_make_box(s);
}

int imdb200_set_eeam_table(imdb200_sim *s, const imdb200_pot_table *emod)
{
  if (!s) return imdb_fail(IMDB200_ERR_ARG, "null handle");
  CUDA_TRY(cudaSetDevice(s->cfg.device));
  return tables_upload_emod(s, emod);
}

int imdb200_set_adp_tables(imdb200_sim *s, const imdb200_pot_table *u, const imdb200_pot_table *w)
{
  if (!s) return imdb_fail(IMDB200_ERR_ARG, "null handle");
  CUDA_TRY(cudaSetDevice(s->cfg.device));
  TRY(tables_upload_adp(s, u, w));
  TRY(adp_ensure_arrays(s));
  s->have_valid_nbl = 0;
  return geom_make_box(s);                 // cellsz may have grown
}

int imdb200_set_stream(imdb200_sim *s, void *stream)
{
  if (!s) return imdb_fail(IMDB200_ERR_ARG, "null handle");
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  if (s->own_stream) { cudaStreamDestroy(s->stream); s->own_stream = 0; }
  s->stream = (cudaStream_t) stream;
  return 0;
}

int imdb200_set_restrictions(imdb200_sim *s, int total_types, const double *r)
{
  if (!s || total_types < 1 || !r) return imdb_fail(IMDB200_ERR_ARG, "bad restrictions");
  CUDA_TRY(cudaSetDevice(s->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(s->stream));        // queued kernels may still read the old table
  if (s->restr) cudaFree(s->restr);
  s->restr = nullptr;
  CUDA_TRY(cudaMalloc(&s->restr, 3 * total_types * sizeof(double)));
  CUDA_TRY(cudaMemcpy(s->restr, r, 3 * total_types * sizeof(double), cudaMemcpyHostToDevice));
  int all1 = 1;
  for (int i = 0; i < 3 * total_types; i++) if (r[i] != 1.0) all1 = 0;
  s->n_restr = all1 ? 0 : total_types;
  s->nactive_dirty = 1;
  return 0;
}

static int ensure_partials(imdb200_sim *s)
{
  const size_t need = ((size_t) s->cap_atoms * 32 / 128 + 64) * 8 * sizeof(double);
  if (s->d_partial && need <= s->partial_bytes) return 0;      // no allocation on a repeated set_atoms
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  if (s->d_partial) cudaFree(s->d_partial);
  s->d_partial = nullptr; s->partial_bytes = 0;
  CUDA_TRY(cudaMalloc(&s->d_partial, need));
  s->partial_bytes = need;
  return 0;
}

// ---- host <-> device transfer of atom arrays: raw user arrays are copied as they are and (un)packed on
// the device, so the host side does no per-atom work and pinned user buffers are copied at full speed ------
static int ensure_xfer(imdb200_sim *s, size_t bytes)
{
  if (bytes <= s->xfer_bytes) return 0;
  if (s->xfer) cudaFree(s->xfer);
  s->xfer = nullptr; s->xfer_bytes = 0;
  CUDA_TRY(cudaMalloc(&s->xfer, bytes));
  s->xfer_bytes = bytes;
  return 0;
}

__global__ void k_pack_atoms(long n, const double *ort, const double *impuls, const double *masse, const int *sorte,
                             const int *vsorte, int ntypes, double4 *pos, double4 *mom, double4 *frc, int *flags)
{
  long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int so = sorte[i], vs = vsorte ? vsorte[i] : so;
  if (so < 0 || so >= ntypes) atomicExch(&flags[FL_BADTYPE], 1);
  pos[i] = make_double4(ort[3 * i], ort[3 * i + 1], ort[3 * i + 2], pack_types(so, vs));
  mom[i] = impuls ? make_double4(impuls[3 * i], impuls[3 * i + 1], impuls[3 * i + 2], masse[i])
                  : make_double4(0.0, 0.0, 0.0, masse[i]);
  frc[i] = make_double4(0.0, 0.0, 0.0, 0.0);
}

int imdb200_set_atoms(imdb200_sim *s, long n, const int *nummer, const int *sorte, const int *vsorte,
                      const double *masse, const double *ort, const double *impuls)
{
  if (!s || n <= 0 || !nummer || !sorte || !masse || !ort) return imdb_fail(IMDB200_ERR_ARG, "bad atom arrays");
  if (!s->have_tabs) return imdb_fail(IMDB200_ERR_ARG, "call imdb200_set_potentials before imdb200_set_atoms");
  CUDA_TRY(cudaSetDevice(s->cfg.device));
  if (s->nranks > 1 && !s->nccl_comm) return imdb_fail(IMDB200_ERR_COMM, "cpu_dim has %d ranks: call imdb200_comm_init before imdb200_set_atoms", s->nranks);
  cudaStream_t st = s->stream;
  s->n_own = 0;
  TRY(cells_ensure_capacity(s, n + n / 2 + 4096));
  TRY(ensure_partials(s));
  // staging layout: ort[3n] impuls[3n] masse[n] | sorte[n] vsorte[n]
  TRY(ensure_xfer(s, (size_t) n * (7 * sizeof(double) + 2 * sizeof(int))));
  double *d_ort = (double *) s->xfer, *d_imp = d_ort + 3 * n, *d_m = d_imp + 3 * n;
  int *d_so = (int *) (d_m + n), *d_vs = d_so + n;
  CUDA_TRY(cudaMemcpyAsync(d_ort, ort, 3 * n * sizeof(double), cudaMemcpyHostToDevice, st));
  if (impuls) CUDA_TRY(cudaMemcpyAsync(d_imp, impuls, 3 * n * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_m, masse, n * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_so, sorte, n * sizeof(int), cudaMemcpyHostToDevice, st));
  if (vsorte) CUDA_TRY(cudaMemcpyAsync(d_vs, vsorte, n * sizeof(int), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(s->nummer, nummer, n * sizeof(int), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(s->d_flags, 0, FL_COUNT * sizeof(int), st));
  k_pack_atoms<<<cdiv(n, 256), 256, 0, st>>>(n, d_ort, impuls ? d_imp : nullptr, d_m, d_so, vsorte ? d_vs : nullptr,
                                             s->cfg.ntypes, s->pos, s->mom, s->frc, s->d_flags);
  LAUNCH_CHECK();
  CUDA_TRY(cudaMemcpyAsync(s->h_flags, s->d_flags, FL_COUNT * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (s->h_flags[FL_BADTYPE]) return imdb_fail(IMDB200_ERR_ARG, "an atom has a type outside 0..ntypes-1");
  s->n_own = n;
  s->natoms_global = s->nranks > 1 ? 0 : n;     // multi-rank: counted after the first binning
  s->need_filter = s->nranks > 1;               // keep only the atoms of this rank's domain
  s->nactive = 3 * (long long) n;               /* recounted with the restriction vectors after the first binning */
  s->nactive_dirty = 1;
  s->n_ghost = 0;
  s->have_valid_nbl = 0;
  s->nbl_count = 0;
  if (s->lanes == 0) {
    int L = s->cfg.lanes_per_atom;
    if (L == 0) { L = 1; while (L < 32 && n * L < 400000) L *= 2; }
    if (L < 1 || L > 32 || (L & (L - 1))) return imdb_fail(IMDB200_ERR_ARG, "lanes_per_atom must be a power of two <= 32");
    s->lanes = L;
  }
  return 0;
}

// ---- the step loop ------------------------------------------------------------------------------------
// End of one queued step of imdb200_run: check_nblist's decision (src/imd_forces_nbl.c:2036) taken on the device.
// An executed step (gate open) publishes its scalars and closes the gate when the list has become invalid, which
// turns the kernels of the step queued behind it into no-ops; a gated step reports executed = 0.
__global__ void k_step_end(StepCtl *ctl, const double *glob, const int *flags, double lim2, StepSlot *out)
{
  const int t = threadIdx.x;
  const int open = ctl->gate;
  __syncwarp();
  if (t < SC_COUNT) out->scal[t] = glob[t];
  if (t < FL_COUNT) out->flags[t] = flags[t];
  if (t == 0) {
    const int valid = open && !(glob[SC_MAXD2] > lim2);
    out->executed = open;
    out->valid = valid;
    if (open) { ctl->disp2 = glob[SC_MAXD2]; ctl->gate = valid; }
  }
}
__global__ void k_open_gate(StepCtl *ctl) { ctl->gate = 1; }

static int fetch_scalars(imdb200_sim *s)
{
  TRY(comm_sync_scalars(s));     // MPI_Allreduce sites: every rank sees the sums over all domains
  CUDA_TRY(cudaMemcpyAsync(s->h_scal, s->d_glob, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaMemcpyAsync(s->h_flags, s->d_flags, FL_COUNT * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  if (s->h_flags[FL_SHORT]) s->is_short = 1;
  s->eta = s->h_scal[SC_ETA];
  return 0;
}

static int ready(imdb200_sim *s)
{
  if (!s) return imdb_fail(IMDB200_ERR_ARG, "null handle");
  if (!s->have_tabs || (s->n_own <= 0 && s->nranks == 1)) return imdb_fail(IMDB200_ERR_ARG, "potentials and atoms must be set first");
  CUDA_TRY(cudaSetDevice(s->cfg.device));
  if (s->nactive_dirty && s->have_valid_nbl) TRY(cells_count_nactive(s));   // restrictions changed between rebuilds
  return 0;
}

static int calc_forces_async(imdb200_sim *s)
{
  s->p2p_step = s->have_valid_nbl;                     // steps with a list build keep the NCCL exchange
  if (!s->have_valid_nbl) TRY(cells_rebuild(s));       // fix_cells, send_cells, make_nblist (:304-317)
  else TRY(comm_ghost_pos(s));                         // send_cells(copy_cell,...) (:314)
  TRY(forces_pass1(s));
  if (s->tabs.have_adp) TRY(forces_adp_pass1(s));
  if (s->tabs.have_eam) {
    TRY(comm_ghost_dF(s));                             // send_cells(copy_dF,...) (:1115)
    if (s->tabs.have_eeam) TRY(comm_ghost_dM(s));
    if (s->tabs.have_adp) TRY(forces_adp_halo(s));
    TRY(forces_pass2(s, 0));
    if (s->tabs.have_adp) TRY(forces_adp_pass2(s));
  }
  return 0;
}

int imdb200_calc_forces(imdb200_sim *s, int steps)
{
  (void) steps;
  TRY(ready(s));
  TRY(calc_forces_async(s));
  TRY(fetch_scalars(s));
  if (s->h_flags[FL_SHORT] && !s->short_warned && s->rank == 0) {   // once per handle, rank 0 speaks for all (the flag is global)
    fprintf(stderr, "Short distance, pair, step %d!\n", steps); /* :982 */
    s->short_warned = 1;
  }
  return 0;
}

static void apply_check(imdb200_sim *s)
{
  const double lim = 0.5 * s->cfg.nbl_margin;
  if (s->h_scal[SC_MAXD2] > lim * lim) s->have_valid_nbl = 0;   /* src/imd_forces_nbl.c:2036 */
}

// NPT_iso: the barostat needs the global virial of this step and Ekin_old on the host before the launch
static int move_npt(imdb200_sim *s)
{
  if (s->npt_ekin_old < 0.0) {                       // steps == steps_min in the reference (:1493-1496)
    TRY(integrate_npt_dyn_pressure(s));
    TRY(fetch_scalars(s));
    s->npt_ekin_old = s->h_scal[SC_EKIN2];
    if (s->cfg.isq_tau_xi == 0.0) s->npt_xi = 0.0;
  }
  TRY(integrate_move_npt(s));
  TRY(fetch_scalars(s));
  return integrate_npt_after_fetch(s);
}

// NPT_axial: vir_xx/yy/zz of this step (sums of the per-atom stress), then like move_npt
static int move_npt_axial(imdb200_sim *s)
{
  if (s->npt_ekin_old < 0.0) {                       // steps == steps_min in the reference (:1756-1771)
    TRY(integrate_axial_dyn_pressure(s));
    TRY(fetch_scalars(s));
    s->npt_ekin_old = s->h_scal[SC_EKIN2];
    s->ax_dyn[0] = s->h_scal[SC_DYNX]; s->ax_dyn[1] = s->h_scal[SC_DYNY]; s->ax_dyn[2] = s->h_scal[SC_DYNZ];
    for (int d = 0; d < 3; d++) { if (s->cfg.isq_tau_xi == 0.0) s->ax_xi[d] = 0.0; s->ax_xi[d] *= s->ax_relax[d]; }
  }
  TRY(integrate_axial_virial(s));
  TRY(fetch_scalars(s));
  TRY(integrate_move_axial(s));
  TRY(fetch_scalars(s));
  return integrate_axial_after_fetch(s);
}

int imdb200_set_npt_axial(imdb200_sim *s, const double xi3[3], const double pext3[3], const double dpext3[3],
                          const int relax3[3], double Ekin_old, const double dyn3[3])
{
  if (!s || !xi3 || !pext3) return imdb_fail(IMDB200_ERR_ARG, "null argument");
  for (int d = 0; d < 3; d++) {
    s->ax_xi[d] = xi3[d]; s->ax_pext[d] = pext3[d]; s->ax_dpext[d] = dpext3 ? dpext3[d] : 0.0;
    s->ax_relax[d] = relax3 ? relax3[d] : 1; s->ax_dyn[d] = dyn3 ? dyn3[d] : 0.0;
  }
  if (Ekin_old >= 0.0 && !dyn3) return imdb_fail(IMDB200_ERR_ARG, "npt_axial: Ekin_old without dyn_stress");
  s->npt_ekin_old = Ekin_old;
  return 0;
}

int imdb200_get_npt_axial(imdb200_sim *s, double out13[13])
{
  if (!s || !out13) return imdb_fail(IMDB200_ERR_ARG, "null argument");
  for (int d = 0; d < 3; d++) { out13[d] = s->ax_xi[d]; out13[3 + d] = s->ax_stress[d]; out13[6 + d] = s->ax_pext[d]; out13[9 + d] = s->ax_dyn[d]; }
  out13[12] = s->npt_ekin_old;
  return 0;
}

int imdb200_move_atoms(imdb200_sim *s)
{
  TRY(ready(s));
  if (s->nbl_count == 0) return imdb_fail(IMDB200_ERR_ARG, "move_atoms before the first calc_forces");
  if (s->cfg.ensemble == IMDB200_ENS_NPT_ISO) TRY(move_npt(s));
  else if (s->cfg.ensemble == IMDB200_ENS_NPT_AXIAL) TRY(move_npt_axial(s));
  else { TRY(integrate_move(s)); TRY(fetch_scalars(s)); }
  s->disp2 = s->h_scal[SC_MAXD2];
  return step_snapshot_disp2(s, 0);
}

int imdb200_set_npt_state(imdb200_sim *s, double xi, double Ekin_old, double pressure_ext)
{
  if (!s) return imdb_fail(IMDB200_ERR_ARG, "null handle");
  s->npt_xi = xi; s->npt_ekin_old = Ekin_old; s->npt_pressure_ext = pressure_ext;
  return 0;
}

int imdb200_get_npt_state(imdb200_sim *s, double out4[4])
{
  if (!s || !out4) return imdb_fail(IMDB200_ERR_ARG, "null argument");
  out4[0] = s->npt_xi; out4[1] = s->npt_ekin_old; out4[2] = s->npt_pressure; out4[3] = s->npt_pressure_ext;
  return 0;
}

int imdb200_check_nblist(imdb200_sim *s)
{
  TRY(ready(s));
  if (s->nbl_count == 0) { s->have_valid_nbl = 0; return 0; }
  TRY(integrate_check_nblist(s));
  TRY(fetch_scalars(s));
  s->disp2 = s->h_scal[SC_MAXD2];
  apply_check(s);
  return step_snapshot_disp2(s, 0);
}

int imdb200_fix_cells(imdb200_sim *s) { TRY(ready(s)); s->have_valid_nbl = 0; return cells_rebuild(s); }
int imdb200_make_nblist(imdb200_sim *s) { TRY(ready(s)); s->have_valid_nbl = 0; return cells_rebuild(s); }
int imdb200_invalidate_nblist(imdb200_sim *s) { if (!s) return IMDB200_ERR_ARG; s->have_valid_nbl = 0; return 0; }
int imdb200_set_skin_skip(imdb200_sim *s, int on) { if (!s) return IMDB200_ERR_ARG; s->skin_skip = on ? 1 : 0; return 0; }
// npt_axial needs the per-axis virial every step: its force kernels always run the per-atom-stress instances
int imdb200_set_press_calc(imdb200_sim *s, int on)
{ if (!s) return IMDB200_ERR_ARG; s->press_calc = (on || s->cfg.ensemble == IMDB200_ENS_NPT_AXIAL) ? 1 : 0; return 0; }

int imdb200_set_eta(imdb200_sim *s, double eta)
{
  if (!s) return IMDB200_ERR_ARG;
  s->eta = eta;
  CUDA_TRY(cudaSetDevice(s->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  CUDA_TRY(cudaMemcpy(s->d_scal + SC_ETA, &eta, sizeof(double), cudaMemcpyHostToDevice));
  return 0;
}
__global__ void k_set_momenta(double4 *mom, const double *buf, long n)
{
  const long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 p = mom[i];
  p.x = buf[3 * i]; p.y = buf[3 * i + 1]; p.z = buf[3 * i + 2];       // .w keeps the mass
  mom[i] = p;
}

int imdb200_set_momenta(imdb200_sim *s, long n, const int *nummer, const double *impuls)
{
  if (!s || !nummer || !impuls) return imdb_fail(IMDB200_ERR_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(s->cfg.device));
  if (n != s->n_own) return imdb_fail(IMDB200_ERR_ARG, "set_momenta: %ld atoms given, this rank holds %ld", n, s->n_own);
  if (n == 0) return 0;
  // the device order is cell-sorted: bring the caller's rows into it through the atom numbers
  std::vector<int> dnum(n);
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  CUDA_TRY(cudaMemcpy(dnum.data(), s->nummer, n * sizeof(int), cudaMemcpyDeviceToHost));
  int mx = 0;
  for (long a = 0; a < n; a++) { if (nummer[a] < 0) return imdb_fail(IMDB200_ERR_ARG, "negative atom number"); if (nummer[a] > mx) mx = nummer[a]; }
  std::vector<int> where((size_t) mx + 1, -1);
  for (long a = 0; a < n; a++) where[nummer[a]] = (int) a;
  std::vector<double> buf((size_t) 3 * n);
  for (long i = 0; i < n; i++) {
    const int a = (dnum[i] >= 0 && dnum[i] <= mx) ? where[dnum[i]] : -1;
    if (a < 0) return imdb_fail(IMDB200_ERR_ARG, "set_momenta: atom %d of this rank is not in the caller's list", dnum[i]);
    buf[3 * i] = impuls[3 * (size_t) a]; buf[3 * i + 1] = impuls[3 * (size_t) a + 1]; buf[3 * i + 2] = impuls[3 * (size_t) a + 2];
  }
  double *d_buf = nullptr;
  CUDA_TRY(cudaMalloc(&d_buf, buf.size() * sizeof(double)));
  cudaError_t e = cudaMemcpy(d_buf, buf.data(), buf.size() * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) { k_set_momenta<<<cdiv(n, 256), 256, 0, s->stream>>>(s->mom, d_buf, n); g_kernel_launches++; e = cudaStreamSynchronize(s->stream); }
  cudaFree(d_buf);
  if (e != cudaSuccess) return imdb_fail(IMDB200_ERR_CUDA, "set_momenta: %s", cudaGetErrorString(e));
  return 0;
}

int imdb200_set_berendsen(imdb200_sim *s, double tauber, double tot_kin_energy)
{
  if (!s) return IMDB200_ERR_ARG;
  if (s->cfg.ensemble != IMDB200_ENS_NVE && tauber > 0.0) return imdb_fail(IMDB200_ERR_ARG, "the Berendsen variant belongs to ensemble nve");
  CUDA_TRY(cudaSetDevice(s->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  s->tauber = tauber;
  s->h_scal[SC_EKIN] = tot_kin_energy;
  // the kinetic energy the previous move_atoms left, as every rank sees it (the GLOBAL value enters the scale factor)
  CUDA_TRY(cudaMemcpy(s->d_glob + SC_EKIN, &tot_kin_energy, sizeof(double), cudaMemcpyHostToDevice));
  if (s->d_glob != s->d_scal) {
    const double mine = s->rank == 0 ? tot_kin_energy : 0.0;      // the local blocks are summed over the ranks
    CUDA_TRY(cudaMemcpy(s->d_scal + SC_EKIN, &mine, sizeof(double), cudaMemcpyHostToDevice));
  }
  return 0;
}
int imdb200_set_temperature(imdb200_sim *s, double t) { if (!s) return IMDB200_ERR_ARG; s->cfg.temperature = t; return 0; }

// One step of imdb200_run, queued without waiting for anything: forces, move_atoms, check_nblist on the device,
// results into result slot `slot`.  `built`: the list was rebuilt just before (the ghost positions are current).
static int queue_step(imdb200_sim *s, int slot, bool built)
{
  cudaEvent_t *ev = s->ev + 5 * slot;
  s->p2p_step = !built;                        // the exchanges of a step with a list build stay on the NCCL path
  // single-species EAM: move_atoms + check_nblist ride in the tail of pass 2 (bit-identical, one kernel less)
  const int fuse = forces_can_fuse_move(s);
  // Overlapped halo (peer-memory exchange + boundary-first order): every pass runs its boundary warps first, their
  // results go straight into the neighbours' ghost regions, and the interior warps run while the data is in flight;
  // the wait sits in front of the next kernel that reads images.
  const int split = comm_p2p_ready(s) && forces_split_possible(s);
  if (!built) {
    cudaEventRecord(ev[0], s->stream);
    if (split && s->pos_sent_early) { TRY(comm_p2p_end(s, 0)); TRY(comm_ghost_pos_finish(s)); }   // sent during pass 2 of the last step
    else TRY(comm_ghost_pos(s));                                                   // send_cells(copy_cell,...) (:314)
  }
  s->pos_sent_early = 0;
  cudaEventRecord(ev[1], s->stream);
  s->fuse_step = fuse; s->zero_before_move = !fuse;
  if (split) {
    s->split_part = 1; TRY(forces_pass1(s));
    TRY(comm_p2p_begin(s, 1));                                                     // F' of the boundary atoms on its way ...
    s->split_part = 2; TRY(forces_pass1(s));                                       // ... while the interior is evaluated
    s->split_part = 0;
    cudaEventRecord(ev[2], s->stream);
    TRY(comm_p2p_end(s, 1)); TRY(comm_ghost_dF_finish(s));
    s->split_part = 1; TRY(forces_pass2(s, fuse));                                 // boundary atoms: forces, move_atoms
    TRY(comm_p2p_begin(s, 0)); s->pos_sent_early = 1;                              // their new positions for the next step
    s->split_part = 2; TRY(forces_pass2(s, fuse));
    s->split_part = 0;
  } else {
    TRY(forces_pass1(s));
    cudaEventRecord(ev[2], s->stream);
    if (s->tabs.have_eam) {
      TRY(comm_ghost_dF(s));                                                       // send_cells(copy_dF,...) (:1115)
      if (s->tabs.have_eeam) TRY(comm_ghost_dM(s));
      TRY(forces_pass2(s, fuse));
    } else s->maxd2_zeroed = 0;
  }
  cudaEventRecord(ev[3], s->stream);
  if (fuse) TRY(integrate_finish(s, 0)); else TRY(integrate_move(s));
  s->fuse_step = 0; s->zero_before_move = 0;
  cudaEventRecord(ev[4], s->stream);
  TRY(comm_sync_scalars(s));                  // the MPI_Allreduce sites: every rank sees the sums and the global maximum
  const double lim = 0.5 * s->cfg.nbl_margin;
  k_step_end<<<1, 32, 0, s->stream>>>(s->d_ctl, s->d_glob, s->d_flags, lim * lim, s->d_slot + slot); LAUNCH_CHECK();
  CUDA_TRY(cudaMemcpyAsync(s->h_slot + slot, s->d_slot + slot, sizeof(StepSlot), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaEventRecord(s->ev_slot[slot], s->stream));
  return 0;
}

// The host looks at a finished step: scalars, flags, timers, and whether the list survived it.
static int retire_step(imdb200_sim *s, int slot, bool built, double rebuild_ms, int *executed)
{
  CUDA_TRY(cudaEventSynchronize(s->ev_slot[slot]));
  const StepSlot &r = s->h_slot[slot];
  *executed = r.executed;
  if (!r.executed) return 0;                    // it was queued behind a step that invalidated the list: nothing happened
  memcpy(s->h_scal, r.scal, sizeof(r.scal));
  memcpy(s->h_flags, r.flags, sizeof(r.flags));
  if (s->h_flags[FL_SHORT]) s->is_short = 1;
  s->eta = s->h_scal[SC_ETA];
  s->disp2 = s->h_scal[SC_MAXD2];
  if (!r.valid) s->have_valid_nbl = 0;
  cudaEvent_t *ev = s->ev + 5 * slot;
  float ms;
  if (built) s->t_ms[0] += rebuild_ms; else { cudaEventElapsedTime(&ms, ev[0], ev[1]); s->t_ms[4] += ms; }
  cudaEventElapsedTime(&ms, ev[1], ev[2]); s->t_ms[1] += ms;
  cudaEventElapsedTime(&ms, ev[2], ev[3]); s->t_ms[2] += ms;
  cudaEventElapsedTime(&ms, ev[3], ev[4]); s->t_ms[3] += ms;
  s->t_ms[5] += built ? 1.0 : 0.0;
  s->t_ms[6] += 1.0;
  return 0;
}

// main_loop (src/imd_main_3d.c:155-870) for nve / nvt, resident on the device.  The host never waits for the step it
// has just queued: it queues the next one first and then looks at the previous one (scalars, rebuild decision).  The
// decision itself is taken on the device (k_step_end); when a step invalidates the list, the step already queued
// behind it finds the gate closed and does nothing, the host rebuilds the list and queues it again.
static int run_async(imdb200_sim *s, int nsteps)
{
  int done = 0, inflight = 0, head = 0, tail = 0;
  s->pos_sent_early = 0;                       // the caller may have changed positions or box since the last call
  bool built[3] = {false, false, false};
  double reb_ms[3] = {0.0, 0.0, 0.0};
  int rc = 0;
  while (done < nsteps && rc == 0) {
    bool fresh = false;
    double rms = 0.0;
    if (!s->have_valid_nbl) {                  // only reached with nothing in flight
      cudaEventRecord(s->ev[15], s->stream);
      if ((rc = cells_rebuild(s)) != 0) break; // fix_cells, send_cells, make_nblist (:304-317); synchronises
      k_open_gate<<<1, 1, 0, s->stream>>>(s->d_ctl); g_kernel_launches++;
      cudaEventRecord(s->ev[14], s->stream);
      cudaEventSynchronize(s->ev[14]);
      float ms; cudaEventElapsedTime(&ms, s->ev[15], s->ev[14]); rms = ms;
      fresh = true;
    }
    while (inflight < 2 && done + inflight < nsteps && rc == 0) {
      built[head] = fresh; reb_ms[head] = rms; fresh = false;
      rc = queue_step(s, head, built[head]);
      head = (head + 1) % 3; inflight++;
    }
    if (rc) break;
    int executed = 0;
    rc = retire_step(s, tail, built[tail], reb_ms[tail], &executed);
    tail = (tail + 1) % 3; inflight--;
    done += executed;
    if (rc == 0 && !s->have_valid_nbl)         // what is still queued will find the gate closed: drain it
      while (inflight > 0 && rc == 0) { rc = retire_step(s, tail, false, 0.0, &executed); tail = (tail + 1) % 3; inflight--; done += executed; }
  }
  // leave the gate open for the stand-alone calls (the host flag have_valid_nbl carries the decision from here on)
  cudaStreamSynchronize(s->stream);
  k_open_gate<<<1, 1, 0, s->stream>>>(s->d_ctl); g_kernel_launches++;
  return rc;
}

int imdb200_run(imdb200_sim *s, int nsteps)
{
  TRY(ready(s));
  const bool host_barostat = s->cfg.ensemble == IMDB200_ENS_NPT_ISO || s->cfg.ensemble == IMDB200_ENS_NPT_AXIAL;
  if (!host_barostat && !s->tabs.have_adp) {
    TRY(run_async(s, nsteps));
    if (s->is_short && !s->short_warned && s->rank == 0) { fprintf(stderr, "Short distance!\n"); s->short_warned = 1; }
    return 0;
  }
  // npt_iso (the host advances the barostat between the force and the move kernels) and ADP: one step at a time
  for (int k = 0; k < nsteps; k++) {
    const bool rebuild = !s->have_valid_nbl;
    cudaEventRecord(s->ev[0], s->stream);
    s->p2p_step = !rebuild;
    if (rebuild) TRY(cells_rebuild(s)); else TRY(comm_ghost_pos(s));
    cudaEventRecord(s->ev[1], s->stream);
    TRY(forces_pass1(s));
    cudaEventRecord(s->ev[2], s->stream);
    if (s->tabs.have_adp) TRY(forces_adp_pass1(s));
    if (s->tabs.have_eam) {
      TRY(comm_ghost_dF(s));
      if (s->tabs.have_eeam) TRY(comm_ghost_dM(s));
      if (s->tabs.have_adp) TRY(forces_adp_halo(s));
      TRY(forces_pass2(s, 0));
      if (s->tabs.have_adp) TRY(forces_adp_pass2(s));
    }
    cudaEventRecord(s->ev[3], s->stream);
    if (s->cfg.ensemble == IMDB200_ENS_NPT_ISO) { TRY(fetch_scalars(s)); TRY(move_npt(s)); }   // virial first, see move_npt
    else if (s->cfg.ensemble == IMDB200_ENS_NPT_AXIAL) TRY(move_npt_axial(s));
    else TRY(integrate_move(s));
    cudaEventRecord(s->ev[4], s->stream);
    if (!host_barostat) TRY(fetch_scalars(s));
    s->disp2 = s->h_scal[SC_MAXD2];
    TRY(step_snapshot_disp2(s, 0));
    apply_check(s);
    float ms;
    cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]); s->t_ms[rebuild ? 0 : 4] += ms;
    cudaEventElapsedTime(&ms, s->ev[1], s->ev[2]); s->t_ms[1] += ms;
    cudaEventElapsedTime(&ms, s->ev[2], s->ev[3]); s->t_ms[2] += ms;
    cudaEventElapsedTime(&ms, s->ev[3], s->ev[4]); s->t_ms[3] += ms;
    s->t_ms[5] += rebuild ? 1.0 : 0.0;
    s->t_ms[6] += 1.0;
  }
  if (s->is_short && !s->short_warned && s->rank == 0) { fprintf(stderr, "Short distance!\n"); s->short_warned = 1; }
  return 0;
}

int imdb200_get_timers(imdb200_sim *s, double out[8], int reset)
{
  if (!s) return IMDB200_ERR_ARG;
  memcpy(out, s->t_ms, sizeof(s->t_ms));
  if (reset) memset(s->t_ms, 0, sizeof(s->t_ms));
  return 0;
}

int imdb200_lin_deform(imdb200_sim *s, const double dx[3], const double dy[3], const double dz[3], double scale)
{
  TRY(ready(s));
  TRY(integrate_lin_deform(s, dx, dy, dz, scale));
  s->skin_all = 1;                 // images move with the box: the displacement bound no longer holds
  // box vectors: box += scale * (D box)  (src/imd_deform.c:71-106)
  Geom &g = s->geom;
  for (int b = 0; b < 3; b++) {
    double t0 = scale * (dx[0] * g.box[b][0] + dx[1] * g.box[b][1] + dx[2] * g.box[b][2]);
    double t1 = scale * (dy[0] * g.box[b][0] + dy[1] * g.box[b][1] + dy[2] * g.box[b][2]);
    double t2 = scale * (dz[0] * g.box[b][0] + dz[1] * g.box[b][1] + dz[2] * g.box[b][2]);
    g.box[b][0] += t0; g.box[b][1] += t1; g.box[b][2] += t2;
  }
  return geom_make_box(s);
}

int imdb200_deform_sample(imdb200_sim *s, double size, const double *shift, const int *shear_def,
                          const double *shear, const double *base)
{
  TRY(ready(s));
  s->skin_all = 1;
  return integrate_deform_sample(s, s->cfg.total_types, size, shift, shear_def, shear, base);
}

// ---- results ----------------------------------------------------------------------------------------------
int imdb200_get_scalars(imdb200_sim *s, imdb200_scalars *o)
{
  if (!s || !o) return imdb_fail(IMDB200_ERR_ARG, "null argument");
  memset(o, 0, sizeof(*o));
  // host-side only: the values fetched by the last calc_forces / move_atoms / check_nblist / run.  No device
  // access and no communication, so a single rank may ask (the step calls themselves are collective).
  o->tot_pot_energy = s->h_scal[SC_EPOT];
  o->tot_kin_energy = s->h_scal[SC_EKIN];
  o->virial = s->h_scal[SC_VIRIAL];
  o->volume = s->volume;
  o->eta = s->eta;
  o->max_displacement2 = s->h_scal[SC_MAXD2];
  for (int d = 0; d < 6; d++) o->tot_presstens[d] = s->h_scal[SC_PXX + d];
  o->natoms = s->natoms_global ? s->natoms_global : s->n_own; o->nactive = s->nactive;
  o->nbl_len = s->nbl_len;
  o->have_valid_nbl = s->have_valid_nbl; o->nbl_count = s->nbl_count; o->is_short = s->is_short;
  for (int d = 0; d < 3; d++) { o->global_cell_dim[d] = s->geom.gdim[d]; o->cell_dim[d] = s->geom.cdim[d]; }
  o->cellsz = s->geom.cellsz;
  return 0;
}

int imdb200_get_box(imdb200_sim *s, double out9[9])
{
  if (!s || !out9) return imdb_fail(IMDB200_ERR_ARG, "null argument");
  for (int b = 0; b < 3; b++) for (int d = 0; d < 3; d++) out9[3 * b + d] = s->geom.box[b][d];
  return 0;
}

__global__ void k_unpack_soa(const double *src, long stride, int ncomp, long n, double *out);

// the getters return the atom count, 0 when there is nothing to report and -1 after a CUDA error (imdb200_last_error)
#define GET_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
  imdb_fail(IMDB200_ERR_CUDA, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return -1; } } while (0)
#define GET_LAUNCHED() do { g_kernel_launches++; GET_TRY(cudaGetLastError()); } while (0)

long imdb200_get_adp(imdb200_sim *s, double *mu3, double *la6)
{
  if (!s || s->n_own <= 0 || !s->tabs.have_adp || !s->adp_mu) return 0;
  GET_TRY(cudaSetDevice(s->cfg.device));
  const long n = s->n_own;
  cudaStream_t st = s->stream;
  if (ensure_xfer(s, (size_t) n * 6 * sizeof(double))) return -1;
  double *b = (double *) s->xfer;
  const int nb = cdiv(n, 256);
  if (mu3) { k_unpack_soa<<<nb, 256, 0, st>>>(s->adp_mu, s->adp_cap, 3, n, b); GET_LAUNCHED();
             GET_TRY(cudaMemcpyAsync(mu3, b, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, st)); GET_TRY(cudaStreamSynchronize(st)); }
  if (la6) { k_unpack_soa<<<nb, 256, 0, st>>>(s->adp_la, s->adp_cap, 6, n, b); GET_LAUNCHED();
             GET_TRY(cudaMemcpyAsync(la6, b, 6 * n * sizeof(double), cudaMemcpyDeviceToHost, st)); }
  GET_TRY(cudaStreamSynchronize(st));
  return n;
}

long imdb200_get_eeam(imdb200_sim *s, double *eam_p, double *dM)
{
  if (!s || s->n_own <= 0 || !s->tabs.have_eeam) return 0;
  GET_TRY(cudaSetDevice(s->cfg.device));
  const long n = s->n_own;
  if (eam_p) GET_TRY(cudaMemcpyAsync(eam_p, s->eam_p, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (dM) GET_TRY(cudaMemcpyAsync(dM, s->dM, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  GET_TRY(cudaStreamSynchronize(s->stream));
  return n;
}

long imdb200_natoms_local(imdb200_sim *s) { return s ? s->n_own : 0; }
long imdb200_nghost_local(imdb200_sim *s) { return s ? s->n_ghost : 0; }

int imdb200_send_forces(imdb200_sim *s, double *dev_field, int ncomp, long stride)
{
  TRY(ready(s));
  if (!dev_field || stride < s->n_own + s->n_ghost) return imdb_fail(IMDB200_ERR_ARG, "field must hold natoms_local + nghost_local entries per component");
  if (!s->have_valid_nbl) return imdb_fail(IMDB200_ERR_ARG, "no buffer cells yet: call imdb200_calc_forces or imdb200_fix_cells first");
  TRY(comm_reverse_add(s, dev_field, ncomp, stride));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return 0;
}

// unpack kernels: device records -> the plain arrays of the C ABI
__global__ void k_unpack3(const double4 *src, long n, double *v3, double *w)
{
  long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 p = src[i];
  if (v3) { v3[3 * i] = p.x; v3[3 * i + 1] = p.y; v3[3 * i + 2] = p.z; }
  if (w) w[i] = p.w;
}
__global__ void k_unpack_types(const double4 *pos, long n, int *sorte, int *vsorte)
{
  long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double w = pos[i].w;
  sorte[i] = sorte_of(w); vsorte[i] = vsorte_of(w);
}
__global__ void k_unpack_soa(const double *src, long stride, int ncomp, long n, double *out)
{
  long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int c = 0; c < ncomp; c++) out[(size_t) ncomp * i + c] = src[c * stride + i];
}

long imdb200_get_atoms(imdb200_sim *s, int *nummer, int *sorte, int *vsorte, double *masse, double *ort,
                       double *impuls, double *kraft, double *poteng, double *rho, double *dF,
                       double *presstens, double *nblpos)
{
  if (!s || s->n_own <= 0) return 0;
  GET_TRY(cudaSetDevice(s->cfg.device));
  const long n = s->n_own;
  cudaStream_t st = s->stream;
  if (ensure_xfer(s, (size_t) n * 7 * sizeof(double))) return -1;
  // one staging area, re-used in stream order: up to 6 components + 1 scalar
  double *b_ort = (double *) s->xfer, *b_imp = b_ort, *b_frc = b_ort, *b_pt = b_ort, *b_np = b_ort, *b_m = b_ort + 6 * n, *b_e = b_m;
  int *bi = (int *) b_ort;
  const int nb = cdiv(n, 256);
#define D2H(dst, src, bytes) GET_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st))
  if (ort) { k_unpack3<<<nb, 256, 0, st>>>(s->pos, n, b_ort, nullptr); GET_LAUNCHED(); D2H(ort, b_ort, 3 * n * sizeof(double)); }
  if (sorte || vsorte) {
    k_unpack_types<<<nb, 256, 0, st>>>(s->pos, n, bi, bi + n); GET_LAUNCHED();
    if (sorte) D2H(sorte, bi, n * sizeof(int));
    if (vsorte) D2H(vsorte, bi + n, n * sizeof(int));
  }
  if (impuls || masse) {
    k_unpack3<<<nb, 256, 0, st>>>(s->mom, n, impuls ? b_imp : nullptr, masse ? b_m : nullptr); GET_LAUNCHED();
    if (impuls) D2H(impuls, b_imp, 3 * n * sizeof(double));
    if (masse) D2H(masse, b_m, n * sizeof(double));
  }
  if (kraft || poteng) {
    k_unpack3<<<nb, 256, 0, st>>>(s->frc, n, kraft ? b_frc : nullptr, poteng ? b_e : nullptr); GET_LAUNCHED();
    if (kraft) D2H(kraft, b_frc, 3 * n * sizeof(double));
    if (poteng) D2H(poteng, b_e, n * sizeof(double));
  }
  if (nummer) D2H(nummer, s->nummer, n * sizeof(int));
  if (rho) { if (s->tabs.have_eam) D2H(rho, s->rho, n * sizeof(double)); else memset(rho, 0, n * sizeof(double)); }
  if (dF) { if (s->tabs.have_eam) D2H(dF, s->dF, n * sizeof(double)); else memset(dF, 0, n * sizeof(double)); }
  if (presstens) { k_unpack_soa<<<nb, 256, 0, st>>>(s->presstens, s->cap_atoms, 6, n, b_pt); GET_LAUNCHED(); D2H(presstens, b_pt, 6 * n * sizeof(double)); }
  if (nblpos) { k_unpack_soa<<<nb, 256, 0, st>>>(s->nblpos, s->cap_atoms, 3, n, b_np); GET_LAUNCHED(); D2H(nblpos, b_np, 3 * n * sizeof(double)); }
#undef D2H
  GET_TRY(cudaStreamSynchronize(st));
  return n;
}

long imdb200_get_nblist(imdb200_sim *s, int *ni, int *nj, signed char *shift3, long cap)
{
  if (!s || !s->have_valid_nbl || !s->nbl) return -1;
  cudaSetDevice(s->cfg.device);
  cudaStreamSynchronize(s->stream);
  if (cap <= 0 || !ni) return (long) s->nbl_len;
  const long n = s->n_own, ntot = s->n_own + s->n_ghost;
  const int L = s->lanes;
  std::vector<int> nnb(n), num(n), gnum(s->n_ghost > 0 ? s->n_ghost : 1), cid(ntot), code(s->geom.nall + 1);
  std::vector<int> nbl((size_t) s->n_pad * s->max_nb);
  cudaMemcpy(nnb.data(), s->nnb, n * sizeof(int), cudaMemcpyDeviceToHost);
  cudaMemcpy(num.data(), s->nummer, n * sizeof(int), cudaMemcpyDeviceToHost);
  if (s->n_ghost) cudaMemcpy(gnum.data(), s->ghost_num, s->n_ghost * sizeof(int), cudaMemcpyDeviceToHost);
  cudaMemcpy(cid.data(), s->cellid, ntot * sizeof(int), cudaMemcpyDeviceToHost);
  cudaMemcpy(code.data(), s->cell_code, s->geom.nall * sizeof(int), cudaMemcpyDeviceToHost);
  cudaMemcpy(nbl.data(), s->nbl, nbl.size() * sizeof(int), cudaMemcpyDeviceToHost);
  long cnt = 0;
  for (long i = 0; i < n; i++)
    for (int m = 0; m < nnb[i]; m++) {
      int j = nbl[nbl_index(i, m, L, s->max_nb / L)] & NBL_JMASK;   // bits 29-30: the neighbour's type
      int sx = 0, sy = 0, sz = 0, jn;
      // image shift of j as this rank applies it (periodic wrap); 0 for images of a neighbour domain's atoms
      if (j >= n) { int c = code[cid[j]]; sx = c % 3 - 1; sy = (c / 3) % 3 - 1; sz = c / 9 - 1; jn = gnum[j - n]; }
      else jn = num[j];
      if (cnt < cap) {
        ni[cnt] = num[i]; nj[cnt] = jn;
        if (shift3) { shift3[3 * cnt] = (signed char) sx; shift3[3 * cnt + 1] = (signed char) sy; shift3[3 * cnt + 2] = (signed char) sz; }
      }
      cnt++;
    }
  return cnt;
}

int imdb200_pair_int(imdb200_sim *s, int which, int col, long n, const double *r2, double *pot, double *grad)
{
  if (!s) return imdb_fail(IMDB200_ERR_ARG, "null handle");
  CUDA_TRY(cudaSetDevice(s->cfg.device));
  return tables_pair_int(s, which, col, n, r2, pot, grad);
}

} // extern "C"
