// integrate.cu -- leap-frog NVE / Nose-Hoover NVT step fused with the skin check, and deformation.
//   move_atoms_nve   src/imd_integrate.c:32-497   (hot lines :192-217, 328-358, 410-433)
//   move_atoms_nvt   src/imd_integrate.c:891-1147 (hot lines :907-908, 951, 1020-1027, 1103, 1140-1141)
//   check_nblist     src/imd_forces_nbl.c:2007-2037
//   lin_deform / deform_sample   src/imd_deform.c:35-119, 232-269
#include "internal.cuh"

struct IArgs {
  double4 *pos, *mom, *frc;
  const double *nblpos; long nstride;
  const double *restr;
  double *presstens; long pstride;
  long n;
  double dt;
  const double *scal;
  double *partial;
  unsigned long long *maxd2;
  const StepCtl *ctl;
  const double *glob; double nactive, temperature, tauber;   // Berendsen: previous tot_kin_energy (global), target
};

#define IBLOCK 256

__device__ __forceinline__ double block_max(double v)
{
  __shared__ double smx[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (lane == 0) smx[w] = v;
  __syncthreads();
  v = (threadIdx.x < nw) ? smx[threadIdx.x] : 0.0;
  if (w == 0) for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <bool NVT, bool STRESS, bool RESTR>
__global__ void __launch_bounds__(IBLOCK) k_move_atoms(IArgs a)
{
  STEP_GATE(a.ctl);
  const long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  double red[2] = {0.0, 0.0};
  double d2 = 0.0;
  const double cc = NVT ? 1.0 : berendsen_cc(a.glob[SC_EKIN], a.nactive, a.temperature, a.dt, a.tauber);
  if (i < a.n) {
    double4 x = a.pos[i], p = a.mom[i], f = a.frc[i];
    const double m = p.w, dt = a.dt;
    double rx = 1.0, ry = 1.0, rz = 1.0;
    if (RESTR) {
      const double *r = a.restr + 3 * vsorte_of(x.w);
      rx = r[0]; ry = r[1]; rz = r[2];
      f.x *= rx; f.y *= ry; f.z *= rz;                                // :192-197
      a.frc[i] = f;
    }
    const double eta = NVT ? a.scal[SC_ETA] : 0.0;
    const double nx = a.nblpos[i], ny = a.nblpos[a.nstride + i], nz = a.nblpos[2 * a.nstride + i];
    d2 = integrate_atom<NVT>(x, p, f, dt, eta, rx, ry, rz, nx, ny, nz, red, cc);
    a.mom[i] = p;
    a.pos[i] = x;
    if (STRESS) {                                                      // :410-433
      double *s = a.presstens + i;
      s[0] += p.x * p.x / m; s[a.pstride] += p.y * p.y / m; s[2 * a.pstride] += p.z * p.z / m;
      s[3 * a.pstride] += p.y * p.z / m; s[4 * a.pstride] += p.z * p.x / m; s[5 * a.pstride] += p.x * p.y / m;
    }
  }
  d2 = block_max(d2);
  if (threadIdx.x == 0) atomicMax(a.maxd2, (unsigned long long) __double_as_longlong(d2));
  block_sum_store<2>(red, a.partial);
}

__global__ void k_check_nblist(const double4 *pos, const double *nblpos, long nstride, long n, unsigned long long *maxd2)
{
  const long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  double d2 = 0.0;
  if (i < n) {
    const double4 x = pos[i];
    d2 = r2_exact(x.x - nblpos[i], x.y - nblpos[nstride + i], x.z - nblpos[2 * nstride + i]);
  }
  d2 = block_max(d2);
  if (threadIdx.x == 0) atomicMax(maxd2, (unsigned long long) __double_as_longlong(d2));
}

// tot_kin_energy and the Nose-Hoover variable (src/imd_integrate.c:1103, 1140-1141)
// scal: this rank's block, glob: the block summed over ranks (the same pointer on one GPU).  The local
// share of tot_kin_energy is linear in E_kin_1/2, so it is formed here and summed with the rest; eta is
// advanced from the GLOBAL E_kin_2 and comes out identical on every rank.
__global__ void k_nvt_finish(double *scal, const double *glob, double dt, double nactive, double temperature,
                             double isq_tau_eta, const StepCtl *ctl)
{
  STEP_GATE(ctl);
  scal[SC_EKIN] = (scal[SC_EKIN1] + scal[SC_EKIN2]) / 4.0;
  const double e2 = glob[SC_EKIN2];
  const double ttt = nactive * temperature;
  scal[SC_ETA] += dt * (e2 / ttt - 1.0) * isq_tau_eta;
}

// presstens totals (calc_tot_presstens, src/imd_main_3d.c:2069-2130)
__global__ void k_sum_presstens(const double *presstens, long pstride, long n, double *partial)
{
  const long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  double v[6];
#pragma unroll
  for (int d = 0; d < 6; d++) v[d] = (i < n) ? presstens[d * pstride + i] : 0.0;
  block_sum_store<6>(v, partial);
}

int integrate_move(imdb200_sim *s)
{
  IArgs a;
  a.pos = s->pos; a.mom = s->mom; a.frc = s->frc;
  a.nblpos = s->nblpos; a.nstride = s->cap_atoms;
  a.restr = s->restr;
  a.presstens = s->presstens; a.pstride = s->cap_atoms;
  a.n = s->n_own; a.dt = s->cfg.timestep;
  a.scal = s->d_scal; a.partial = s->d_partial;
  a.maxd2 = (unsigned long long *) (s->d_scal + SC_MAXD2);
  a.ctl = s->d_ctl;
  a.glob = s->d_glob; a.nactive = (double) s->nactive; a.temperature = s->cfg.temperature; a.tauber = s->tauber;
  const int nb = cdiv(s->n_own, IBLOCK);
  // SC_MAXD2 starts from 0: cleared by the reduction kernel behind pass 2 inside imdb200_run (zero_before_move), by a
  // memset for the stand-alone call
  if (!s->maxd2_zeroed) CUDA_TRY(cudaMemsetAsync(s->d_scal + SC_MAXD2, 0, sizeof(double), s->stream));
  s->maxd2_zeroed = 0;
  const bool nvt = s->cfg.ensemble == IMDB200_ENS_NVT, st = s->press_calc != 0, re = s->n_restr > 0;
#define GO(N, S, R) k_move_atoms<N, S, R><<<nb, IBLOCK, 0, s->stream>>>(a)
  if (nvt) { if (st) { if (re) GO(true, true, true); else GO(true, true, false); }
             else    { if (re) GO(true, false, true); else GO(true, false, false); } }
  else     { if (st) { if (re) GO(false, true, true); else GO(false, true, false); }
             else    { if (re) GO(false, false, true); else GO(false, false, false); } }
#undef GO
  LAUNCH_CHECK();
  return integrate_finish(s, nb);
}

// what follows the per-atom part: kinetic-energy sums, Nose-Hoover update, stress totals
int integrate_finish(imdb200_sim *s, int nb)
{
  const bool nvt = s->cfg.ensemble == IMDB200_ENS_NVT || s->cfg.ensemble == IMDB200_ENS_NPT_ISO || s->cfg.ensemble == IMDB200_ENS_NPT_AXIAL,
             st = s->press_calc != 0;
  if (nvt) {
    if (nb > 0) { const int slots[2] = {SC_EKIN1, SC_EKIN2}; TRY(reduce_finish(s, nb, 2, slots, 0)); }
    TRY(comm_sync_scalars(s));    // MPI_Allreduce of E_kin_1/2 (src/imd_integrate.c:1104-1130)
    k_nvt_finish<<<1, 1, 0, s->stream>>>(s->d_scal, s->d_glob, s->cfg.timestep, (double) s->nactive,
                                         s->cfg.temperature, s->cfg.isq_tau_eta, s->d_ctl);
    LAUNCH_CHECK();
  } else if (nb > 0) {
    const int slots[2] = {SC_EKIN, SC_EKIN2};
    TRY(reduce_finish(s, nb, 2, slots, 0));
  }
  if (st) {
    const int nbp = cdiv(s->n_own, IBLOCK);
    k_sum_presstens<<<nbp, IBLOCK, 0, s->stream>>>(s->presstens, s->cap_atoms, s->n_own, s->d_partial); LAUNCH_CHECK();
    const int slots[6] = {SC_PXX, SC_PYY, SC_PZZ, SC_PYZ, SC_PZX, SC_PXY};
    TRY(reduce_finish(s, nbp, 6, slots, 0));
  }
  return 0;
}

// ---- NPT_iso: Nose-Hoover thermostat + isotropic barostat (move_atoms_npt_iso, src/imd_integrate.c:1472-1729) --------
// Per atom:  p = (pfric*p + dt*F)*pifric,  x = (rfric*x + p*dt/m)*rifric  (:1572-1576, 1603-1607); the four factors
// come from eta, xi_old and xi, which the host advances from the global pressure before the launch.  red[0] is twice
// the kinetic energy before the kick (the reference's Ekin_old), red[1] after it (Ekin_new), as in NVT.
struct NptArgs {
  double4 *pos, *mom; const double4 *frc;
  const double *nblpos; long nstride;
  double *presstens; long pstride;
  long n;
  double dt, pfric, pifric, rfric, rifric;
  double *partial;
  unsigned long long *maxd2;
};

template <bool STRESS>
__global__ void __launch_bounds__(IBLOCK) k_move_atoms_npt(NptArgs a)
{
  const long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  double red[2] = {0.0, 0.0};
  double d2 = 0.0;
  if (i < a.n) {
    double4 x = a.pos[i], p = a.mom[i];
    const double4 f = a.frc[i];
    const double m = p.w;
    red[0] = (p.x * p.x + p.y * p.y + p.z * p.z) / m;
    p.x = (a.pfric * p.x + a.dt * f.x) * a.pifric;
    p.y = (a.pfric * p.y + a.dt * f.y) * a.pifric;
    p.z = (a.pfric * p.z + a.dt * f.z) * a.pifric;
    red[1] = (p.x * p.x + p.y * p.y + p.z * p.z) / m;                  // :1591
    const double tmp = a.dt / m;
    x.x = (a.rfric * x.x + p.x * tmp) * a.rifric;
    x.y = (a.rfric * x.y + p.y * tmp) * a.rifric;
    x.z = (a.rfric * x.z + p.z * tmp) * a.rifric;
    a.mom[i] = p;
    a.pos[i] = x;
    d2 = r2_exact(x.x - a.nblpos[i], x.y - a.nblpos[a.nstride + i], x.z - a.nblpos[2 * a.nstride + i]);
    if (STRESS) {                                                      // :1641-1652
      double *s = a.presstens + i;
      s[0] += p.x * p.x / m; s[a.pstride] += p.y * p.y / m; s[2 * a.pstride] += p.z * p.z / m;
      s[3 * a.pstride] += p.y * p.z / m; s[4 * a.pstride] += p.z * p.x / m; s[5 * a.pstride] += p.x * p.y / m;
    }
  }
  d2 = block_max(d2);
  if (threadIdx.x == 0) atomicMax(a.maxd2, (unsigned long long) __double_as_longlong(d2));
  block_sum_store<2>(red, a.partial);
}

// calc_dyn_pressure (:1403-1465): twice the kinetic energy from the current momenta, local share into SC_EKIN2
__global__ void __launch_bounds__(IBLOCK) k_dyn_pressure(const double4 *mom, long n, double *partial)
{
  const long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  double red[1] = {0.0};
  if (i < n) { const double4 p = mom[i]; red[0] = (p.x * p.x + p.y * p.y + p.z * p.z) / p.w; }
  block_sum_store<1>(red, partial);
}

int integrate_npt_dyn_pressure(imdb200_sim *s)
{
  const int nb = cdiv(s->n_own, IBLOCK);
  if (nb > 0) {
    k_dyn_pressure<<<nb, IBLOCK, 0, s->stream>>>(s->mom, s->n_own, s->d_partial); LAUNCH_CHECK();
    const int slots[1] = {SC_EKIN2};
    TRY(reduce_finish(s, nb, 1, slots, 0));
  }
  return 0;
}

// The caller has fetched the scalars of calc_forces (global virial) and knows the global Ekin_old.
int integrate_move_npt(imdb200_sim *s)
{
  const double dt = s->cfg.timestep, vol = s->volume;
  s->npt_pressure = (s->npt_ekin_old + s->h_scal[SC_VIRIAL]) / (3 * vol);                         // :1505
  const double xi_old = s->npt_xi;
  s->npt_xi += dt * (s->npt_pressure - s->npt_pressure_ext) * vol * s->cfg.isq_tau_xi / (double) s->nactive;   // :1509
  NptArgs a;
  a.pos = s->pos; a.mom = s->mom; a.frc = s->frc;
  a.nblpos = s->nblpos; a.nstride = s->cap_atoms;
  a.presstens = s->presstens; a.pstride = s->cap_atoms;
  a.n = s->n_own; a.dt = dt;
  a.pfric = 1.0 - (xi_old + s->eta) * dt / 2.0;                                                    // :1512-1515
  a.pifric = 1.0 / (1.0 + (s->npt_xi + s->eta) * dt / 2.0);
  a.rfric = 1.0 + (s->npt_xi) * dt / 2.0;
  a.rifric = 1.0 / (1.0 - (s->npt_xi) * dt / 2.0);
  a.partial = s->d_partial;
  a.maxd2 = (unsigned long long *) (s->d_scal + SC_MAXD2);
  const int nb = cdiv(s->n_own, IBLOCK);
  CUDA_TRY(cudaMemsetAsync(s->d_scal + SC_MAXD2, 0, sizeof(double), s->stream));
  if (nb > 0) {
    if (s->press_calc) k_move_atoms_npt<true><<<nb, IBLOCK, 0, s->stream>>>(a);
    else k_move_atoms_npt<false><<<nb, IBLOCK, 0, s->stream>>>(a);
    LAUNCH_CHECK();
  }
  // E_kin sums, tot_kin_energy = (Ekin_old + Ekin_new)/4 and the eta update are NVT's (:1691-1696)
  return integrate_finish(s, nb);
}

// After the scalar fetch that follows integrate_move_npt: remember Ekin_new, let the box breathe (:1704-1718),
// advance the external pressure (:1727).
int integrate_npt_after_fetch(imdb200_sim *s)
{
  s->npt_ekin_old = s->h_scal[SC_EKIN2];
  const double dt = s->cfg.timestep;
  const double ttt = (1.0 + s->npt_xi * dt / 2.0) / (1.0 - s->npt_xi * dt / 2.0);
  if (ttt < 0) return imdb_fail(IMDB200_ERR_EXPLODE, "box size has become negative!");
  Geom &g = s->geom;
  for (int b = 0; b < 3; b++) for (int d = 0; d < 3; d++) g.box[b][d] *= ttt;
  s->skin_all = 1;                 // the images move with the box: the displacement bound of the skin classes is void
  s->npt_pressure_ext += s->cfg.d_pressure;
  return geom_make_box(s);
}

// ---- NPT_axial: Nose-Hoover thermostat + one barostat per box axis (move_atoms_npt_axial, src/imd_integrate.c:1747-1959) ----
// The per-axis virial is what P_AXIAL builds accumulate in calc_forces (vir_xx -= d.x*force.x, src/imd_forces_nbl.c:548-556,
// 1275-1279); here it is the sum over atoms of the per-atom stress tensor of the STRESS instances of the force kernels
// (-1/2 d (x) f per atom and pair end), which this ensemble therefore always runs.
struct AxArgs {
  double4 *pos, *mom; const double4 *frc;
  const double *nblpos; long nstride;
  const double *restr;
  double *presstens; long pstride; int stress;
  long n;
  double dt, pfric[3], pifric[3], rfric[3], rifric[3];
  double *partial;
  unsigned long long *maxd2;
};

template <bool RESTR>
__global__ void __launch_bounds__(IBLOCK) k_move_atoms_axial(AxArgs a)
{
  const long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  // twice the kinetic energy before the kick (= the reference's Ekin_old: the sum the previous step formed from the very
  // same momenta, or calc_dyn_pressure at the first step) and after it (Ekin_new), then dyn_stress_x/y/z
  double red[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  double d2 = 0.0;
  if (i < a.n) {
    double4 x = a.pos[i], p = a.mom[i];
    const double4 f = a.frc[i];
    double tmp = 1.0 / p.w;
    red[0] = (p.x * p.x + p.y * p.y + p.z * p.z) * tmp;
    if (a.stress) {                                                    // kinetic part from the momenta BEFORE the kick (:1834-1845)
      double *s = a.presstens + i;
      s[0] += p.x * p.x * tmp; s[a.pstride] += p.y * p.y * tmp; s[2 * a.pstride] += p.z * p.z * tmp;
      s[3 * a.pstride] += p.y * p.z * tmp; s[4 * a.pstride] += p.z * p.x * tmp; s[5 * a.pstride] += p.x * p.y * tmp;
    }
    p.x = (a.pfric[0] * p.x + a.dt * f.x) * a.pifric[0];              // :1848-1855
    p.y = (a.pfric[1] * p.y + a.dt * f.y) * a.pifric[1];
    p.z = (a.pfric[2] * p.z + a.dt * f.z) * a.pifric[2];
    if (RESTR) { const double *r = a.restr + 3 * vsorte_of(x.w); p.x *= r[0]; p.y *= r[1]; p.z *= r[2]; }   // :1859-1864
    red[2] = p.x * p.x * tmp; red[3] = p.y * p.y * tmp; red[4] = p.z * p.z * tmp;   // :1868-1872
    red[1] = (p.x * p.x + p.y * p.y + p.z * p.z) * tmp;                // :1875
    tmp *= a.dt;
    x.x = (a.rfric[0] * x.x + p.x * tmp) * a.rifric[0];                // :1878-1883
    x.y = (a.rfric[1] * x.y + p.y * tmp) * a.rifric[1];
    x.z = (a.rfric[2] * x.z + p.z * tmp) * a.rifric[2];
    a.mom[i] = p;
    a.pos[i] = x;
    d2 = r2_exact(x.x - a.nblpos[i], x.y - a.nblpos[a.nstride + i], x.z - a.nblpos[2 * a.nstride + i]);
  }
  d2 = block_max(d2);
  if (threadIdx.x == 0) atomicMax(a.maxd2, (unsigned long long) __double_as_longlong(d2));
  block_sum_store<5>(red, a.partial);
}

// calc_dyn_pressure (:1403-1465) per axis
__global__ void __launch_bounds__(IBLOCK) k_dyn_pressure_axial(const double4 *mom, long n, double *partial)
{
  const long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  double red[4] = {0.0, 0.0, 0.0, 0.0};
  if (i < n) { const double4 p = mom[i]; const double tmp = 1.0 / p.w;
               red[1] = p.x * p.x * tmp; red[2] = p.y * p.y * tmp; red[3] = p.z * p.z * tmp; red[0] = (red[1] + red[2]) + red[3]; }
  block_sum_store<4>(red, partial);
}

int integrate_axial_dyn_pressure(imdb200_sim *s)
{
  const int nb = cdiv(s->n_own, IBLOCK);
  if (nb > 0) {
    k_dyn_pressure_axial<<<nb, IBLOCK, 0, s->stream>>>(s->mom, s->n_own, s->d_partial); LAUNCH_CHECK();
    const int slots[4] = {SC_EKIN2, SC_DYNX, SC_DYNY, SC_DYNZ};
    TRY(reduce_finish(s, nb, 4, slots, 0));
  }
  return 0;
}

int integrate_axial_virial(imdb200_sim *s)
{
  const int nbp = cdiv(s->n_own, IBLOCK);
  if (nbp > 0) {
    k_sum_presstens<<<nbp, IBLOCK, 0, s->stream>>>(s->presstens, s->cap_atoms, s->n_own, s->d_partial); LAUNCH_CHECK();
    const int slots[6] = {SC_PXX, SC_PYY, SC_PZZ, SC_PYZ, SC_PZX, SC_PXY};
    TRY(reduce_finish(s, nbp, 6, slots, 0));
  }
  return 0;
}

// The caller has fetched the global vir_xx/yy/zz (SC_PXX..) and holds the global dyn_stress and Ekin_old of the last step.
int integrate_move_axial(imdb200_sim *s)
{
  const double dt = s->cfg.timestep, vol = s->volume;
  const double vir[3] = {s->h_scal[SC_PXX], s->h_scal[SC_PYY], s->h_scal[SC_PZZ]};
  AxArgs a;
  const double ttt = dt * vol * s->cfg.isq_tau_xi / (double) s->nactive;                           // :1782
  for (int d = 0; d < 3; d++) {
    s->ax_stress[d] = (s->ax_dyn[d] + vir[d]) / vol;                                               // :1775-1779
    const double xi_old = s->ax_xi[d];
    s->ax_xi[d] += ttt * (s->ax_stress[d] - s->ax_pext[d]) * s->ax_relax[d];                       // :1783-1787
    a.pfric[d]  =        1.0 - (xi_old      + s->eta) * dt / 2.0;                                  // :1790-1803
    a.pifric[d] = 1.0 / (1.0 + (s->ax_xi[d] + s->eta) * dt / 2.0);
    a.rfric[d]  =        1.0 + (s->ax_xi[d]         ) * dt / 2.0;
    a.rifric[d] = 1.0 / (1.0 - (s->ax_xi[d]         ) * dt / 2.0);
  }
  a.pos = s->pos; a.mom = s->mom; a.frc = s->frc;
  a.nblpos = s->nblpos; a.nstride = s->cap_atoms;
  a.restr = s->restr;
  a.presstens = s->presstens; a.pstride = s->cap_atoms; a.stress = 1;
  a.n = s->n_own; a.dt = dt;
  a.partial = s->d_partial;
  a.maxd2 = (unsigned long long *) (s->d_scal + SC_MAXD2);
  const int nb = cdiv(s->n_own, IBLOCK);
  CUDA_TRY(cudaMemsetAsync(s->d_scal + SC_MAXD2, 0, sizeof(double), s->stream));
  if (nb > 0) {
    if (s->n_restr > 0) k_move_atoms_axial<true><<<nb, IBLOCK, 0, s->stream>>>(a);
    else k_move_atoms_axial<false><<<nb, IBLOCK, 0, s->stream>>>(a);
    LAUNCH_CHECK();
    const int slots[5] = {SC_EKIN1, SC_EKIN2, SC_DYNX, SC_DYNY, SC_DYNZ};
    TRY(reduce_finish(s, nb, 5, slots, 0));
  }
  // tot_kin_energy = (Ekin_old + Ekin_new)/4, the eta update from the global Ekin_new (:1917-1920) and the stress totals
  // of the step (virial + kinetic part) are NVT's
  return integrate_finish(s, 0);
}

// After the scalar fetch: dyn_stress and Ekin_old for the next step, the box (:1923-1937), the pressure ramp (:1955-1959).
int integrate_axial_after_fetch(imdb200_sim *s)
{
  const double dt = s->cfg.timestep;
  s->npt_ekin_old = s->h_scal[SC_EKIN2];
  s->ax_dyn[0] = s->h_scal[SC_DYNX]; s->ax_dyn[1] = s->h_scal[SC_DYNY]; s->ax_dyn[2] = s->h_scal[SC_DYNZ];
  double tvec[3];
  for (int d = 0; d < 3; d++) {
    tvec[d] = (1.0 + s->ax_xi[d] * dt / 2.0) / (1.0 - s->ax_xi[d] * dt / 2.0);
    if (tvec[d] < 0) return imdb_fail(IMDB200_ERR_EXPLODE, "box size has become negative!");
  }
  Geom &g = s->geom;
  for (int b = 0; b < 3; b++) for (int d = 0; d < 3; d++) g.box[b][d] *= tvec[b];   // box_x *= tvec.x, box_y *= tvec.y, box_z *= tvec.z
  s->skin_all = 1;
  for (int d = 0; d < 3; d++) s->ax_pext[d] += s->ax_dpext[d];
  return geom_make_box(s);
}

int integrate_check_nblist(imdb200_sim *s)
{
  CUDA_TRY(cudaMemsetAsync(s->d_scal + SC_MAXD2, 0, sizeof(double), s->stream));
  k_check_nblist<<<cdiv(s->n_own, IBLOCK), IBLOCK, 0, s->stream>>>(s->pos, s->nblpos, s->cap_atoms, s->n_own,
                                                                    (unsigned long long *) (s->d_scal + SC_MAXD2));
  LAUNCH_CHECK();
  return 0;
}

// ---- deformation ----------------------------------------------------------------------------------------
struct Mat3 { double m[3][3]; };
__global__ void k_lin_deform(double4 *pos, long n, Mat3 D, double scale)
{
  const long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 x = pos[i];
  // tmport = D x ; ort += scale * tmport  (src/imd_deform.c:59-67), reference operation order
  const double t0 = __dadd_rn(__dadd_rn(__dmul_rn(D.m[0][0], x.x), __dmul_rn(D.m[0][1], x.y)), __dmul_rn(D.m[0][2], x.z));
  const double t1 = __dadd_rn(__dadd_rn(__dmul_rn(D.m[1][0], x.x), __dmul_rn(D.m[1][1], x.y)), __dmul_rn(D.m[1][2], x.z));
  const double t2 = __dadd_rn(__dadd_rn(__dmul_rn(D.m[2][0], x.x), __dmul_rn(D.m[2][1], x.y)), __dmul_rn(D.m[2][2], x.z));
  x.x = __dadd_rn(x.x, __dmul_rn(scale, t0));
  x.y = __dadd_rn(x.y, __dmul_rn(scale, t1));
  x.z = __dadd_rn(x.z, __dmul_rn(scale, t2));
  pos[i] = x;
}

int integrate_lin_deform(imdb200_sim *s, const double dx[3], const double dy[3], const double dz[3], double scale)
{
  Mat3 D;
  for (int d = 0; d < 3; d++) { D.m[0][d] = dx[d]; D.m[1][d] = dy[d]; D.m[2][d] = dz[d]; }
  k_lin_deform<<<cdiv(s->n_own, 256), 256, 0, s->stream>>>(s->pos, s->n_own, D, scale);
  LAUNCH_CHECK();
  return 0;
}

struct DefArgs { const double *shift, *shear, *base; const int *shear_def; double size; };
__global__ void k_deform_sample(double4 *pos, long n, DefArgs d)
{
  const long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 x = pos[i];
  const int vs = vsorte_of(x.w);
  double shear = 1.0;
  if (d.shear_def[vs] == 1) {                                          // src/imd_deform.c:248-255
    const double ox = x.x - d.base[3 * vs], oy = x.y - d.base[3 * vs + 1], oz = x.z - d.base[3 * vs + 2];
    shear = __dadd_rn(__dadd_rn(__dmul_rn(d.shear[3 * vs], ox), __dmul_rn(d.shear[3 * vs + 1], oy)), __dmul_rn(d.shear[3 * vs + 2], oz));
  }
  const double f = __dmul_rn(shear, d.size);                           // :260-264
  x.x = __dadd_rn(x.x, __dmul_rn(f, d.shift[3 * vs]));
  x.y = __dadd_rn(x.y, __dmul_rn(f, d.shift[3 * vs + 1]));
  x.z = __dadd_rn(x.z, __dmul_rn(f, d.shift[3 * vs + 2]));
  pos[i] = x;
}

int integrate_deform_sample(imdb200_sim *s, int nvt, double size, const double *shift, const int *shear_def,
                            const double *shear, const double *base)
{
  double *d = nullptr; int *di = nullptr;
  CUDA_TRY(cudaMalloc(&d, 9 * nvt * sizeof(double)));
  CUDA_TRY(cudaMalloc(&di, nvt * sizeof(int)));
  CUDA_TRY(cudaMemcpy(d, shift, 3 * nvt * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(d + 3 * nvt, shear, 3 * nvt * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(d + 6 * nvt, base, 3 * nvt * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(di, shear_def, nvt * sizeof(int), cudaMemcpyHostToDevice));
  DefArgs a; a.shift = d; a.shear = d + 3 * nvt; a.base = d + 6 * nvt; a.shear_def = di; a.size = size;
  k_deform_sample<<<cdiv(s->n_own, 256), 256, 0, s->stream>>>(s->pos, s->n_own, a);
  LAUNCH_CHECK();
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  cudaFree(d); cudaFree(di);
  return 0;
}
