// forces.cu -- tabulated pair + two-pass EAM forces over the Verlet list (calc_forces,
// src/imd_forces_nbl.c:281-1999; PAIR and EAM2 branches).
//
// The reference walks a HALF list and scatters -f, rho and Epot/2 into atom j (:513-545, 595-610,
// 1267-1281).  On a GPU that scatter would be ~8 FP64 atomics per pair; instead the list is FULL
// and every atom gathers: each pair is evaluated from both ends, no atomics, no write conflicts,
// and the per-atom results (force, Epot, rho, stress) are complete when the thread ends, so the
// embedding lookup (:1079-1095) is fused into the tail of pass 1.  Sums differ from the reference
// only in their order (parity tolerance 1e-10, noise floor 1e-13 -- SURVEY.md section 0).
//
// Thread mapping: L lanes cooperate on one atom (L = 1 for big systems, up to 32 for small ones);
// lane l takes list entries l, l+L, ... (the list rows are lane-interleaved, so a warp reads one
// contiguous 128-byte row segment per iteration) and the partial sums are combined with xor
// shuffles.  Per pair: one 32-byte gather of (x,y,z,type) -- plus 8 bytes of dF in pass 2.
#include "internal.cuh"

struct FArgs {
  const double4 *pos;
  double4 *frc;
  double *rho, *dF;
  const int *nbl, *nnb;
  long n_own, rowstride;
  double *presstens; long pstride;
  double *partial;
  int *flags;
};

#define FBLOCK 128

template <int L> __device__ __forceinline__ double lanes_sum(double v)
{
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double4 ld_pos(const double4 *p)
{
  const double2 *q = reinterpret_cast<const double2 *>(p);
  double2 a = __ldg(q), b = __ldg(q + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

// ----------------------------------------------------------------------------------------------------
// pass 1: pair potential + host electron density (src/imd_forces_nbl.c:422-981), then the embedding
// energy F(rho_i) and 2F'(rho_i) (:1079-1095)
// ----------------------------------------------------------------------------------------------------
template <int L, bool EAM, bool MULTI, bool FUSED, bool STRESS>
__global__ void __launch_bounds__(FBLOCK) k_pass1(FArgs a, DevTables T)
{
  const long gt = blockIdx.x * (long) blockDim.x + threadIdx.x;
  const long i = gt / L;
  const int sub = (int) (gt % L);
  const bool act = i < a.n_own;
  const int nt = T.ntypes;
  double fx = 0.0, fy = 0.0, fz = 0.0, ee = 0.0, rh = 0.0, vir = 0.0;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0, s5 = 0.0;
  int is_short = 0, it = 0;
  if (act) {
    const double4 xi = a.pos[i];
    if (MULTI) it = sorte_of(xi.w);
    const int nn = a.nnb[i];
    const int *row = a.nbl + i * L + sub;
    for (int m = sub; m < nn; m += L, row += a.rowstride) {
      const int j = __ldcs(row);
      const double4 xj = ld_pos(a.pos + j);
      const double dx = xj.x - xi.x, dy = xj.y - xi.y, dz = xj.z - xi.z;
      const double r2 = r2_exact(dx, dy, dz);
      const int col = MULTI ? it * nt + sorte_of(xj.w) : 0;
      if (FUSED) {
        const bool inp = r2 <= T.pair.end[col];            // :493
        const bool inr = r2 < T.rho.end[col];              // :588
        if (inp || inr) {
          // inside its own cut-off the MIN(r2,end) clamp of PAIR_INT2 is inactive
          int k; double chi;
          tab_index(r2, T.pair.begin[col], r2, T.pair.invstep[col], k, chi, is_short);
          const double *e = T.fused1 + ((size_t) k * T.pair.ncols + col) * 8;
          const double2 e0 = ld2(e), e1 = ld2(e + 2), e2 = ld2(e + 4), e3 = ld2(e + 6);
          if (inp) {
            const double pot = fma(chi, fma(chi, e1.x, e0.y), e0.x);
            const double grad = fma(chi, e2.x, e1.y);
            const double gx = dx * grad, gy = dy * grad, gz = dz * grad;
            fx += gx; fy += gy; fz += gz;
            ee += pot;
            vir = fma(r2, grad, vir);
            if (STRESS) { s0 = fma(dx, gx, s0); s1 = fma(dy, gy, s1); s2 = fma(dz, gz, s2);
                          s3 = fma(dy, gz, s3); s4 = fma(dz, gx, s4); s5 = fma(dx, gy, s5); }
          }
          if (inr) rh += fma(chi, fma(chi, e3.y, e3.x), e2.y);
        }
      } else {
        if (r2 <= T.pair.end[col]) {
          int k; double chi;
          tab_index(r2, T.pair.begin[col], T.pair.end[col], T.pair.invstep[col], k, chi, is_short);
          const double *e = T.pairVG + ((size_t) k * T.pair.ncols + col) * 6;
          const double2 e0 = ld2(e), e1 = ld2(e + 2), e2 = ld2(e + 4);
          const double pot = fma(chi, fma(chi, e1.x, e0.y), e0.x);
          const double grad = fma(chi, e2.x, e1.y);
          const double gx = dx * grad, gy = dy * grad, gz = dz * grad;
          fx += gx; fy += gy; fz += gz;
          ee += pot;
          vir = fma(r2, grad, vir);
          if (STRESS) { s0 = fma(dx, gx, s0); s1 = fma(dy, gy, s1); s2 = fma(dz, gz, s2);
                        s3 = fma(dy, gz, s3); s4 = fma(dz, gx, s4); s5 = fma(dx, gy, s5); }
        }
        if (EAM && r2 < T.rho.end[col]) {
          int k; double chi;
          tab_index(r2, T.rho.begin[col], T.rho.end[col], T.rho.invstep[col], k, chi, is_short);
          const double *e = T.rhoV + ((size_t) k * T.rho.ncols + col) * 4;
          const double2 e0 = ld2(e), e1 = ld2(e + 2);
          rh += fma(chi, fma(chi, e1.x, e0.y), e0.x);
        }
      }
    }
  }
  if (L > 1) {
    fx = lanes_sum<L>(fx); fy = lanes_sum<L>(fy); fz = lanes_sum<L>(fz);
    ee = lanes_sum<L>(ee); vir = lanes_sum<L>(vir);
    if (EAM) rh = lanes_sum<L>(rh);
    if (STRESS) { s0 = lanes_sum<L>(s0); s1 = lanes_sum<L>(s1); s2 = lanes_sum<L>(s2);
                  s3 = lanes_sum<L>(s3); s4 = lanes_sum<L>(s4); s5 = lanes_sum<L>(s5); }
  }
  double red[2] = {0.0, 0.0};
  if (act && sub == 0) {
    double epot = 0.5 * ee;                                // pot *= 0.5 on both atoms (:535-545)
    if (EAM) {
      int k, dummy = 0; double chi;                       // PAIR_INT(pot, EAM_DF, embed_pot, ...) :1086
      tab_index(rh, T.embed.begin[it], T.embed.end[it], T.embed.invstep[it], k, chi, dummy);
      const double *e = T.embedVG + ((size_t) k * T.embed.ncols + it) * 6;
      const double2 e0 = ld2(e), e1 = ld2(e + 2), e2 = ld2(e + 4);
      epot += fma(chi, fma(chi, e1.x, e0.y), e0.x);
      a.rho[i] = rh;
      a.dF[i] = fma(chi, e2.x, e1.y);
    }
    a.frc[i] = make_double4(fx, fy, fz, epot);
    if (STRESS) {                                          // -0.5 d (x) f per atom (:558-581)
      double *p = a.presstens + i;
      p[0] = -0.5 * s0; p[a.pstride] = -0.5 * s1; p[2 * a.pstride] = -0.5 * s2;
      p[3 * a.pstride] = -0.5 * s3; p[4 * a.pstride] = -0.5 * s4; p[5 * a.pstride] = -0.5 * s5;
    }
    red[0] = epot;
    red[1] = -0.5 * vir;                                   // virial -= r2*grad once per pair (:555)
  }
  if (is_short) atomicExch(&a.flags[FL_SHORT], 1);
  block_sum_store<2>(red, a.partial);
}

// ----------------------------------------------------------------------------------------------------
// pass 2: EAM forces (src/imd_forces_nbl.c:1117-1322)
// ----------------------------------------------------------------------------------------------------
template <int L, bool MULTI, bool STRESS>
__global__ void __launch_bounds__(FBLOCK) k_pass2(FArgs a, DevTables T)
{
  const long gt = blockIdx.x * (long) blockDim.x + threadIdx.x;
  const long i = gt / L;
  const int sub = (int) (gt % L);
  const bool act = i < a.n_own;
  const int nt = T.ntypes;
  double fx = 0.0, fy = 0.0, fz = 0.0, vir = 0.0;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0, s5 = 0.0;
  int is_short = 0;
  if (act) {
    const double4 xi = a.pos[i];
    const int it = MULTI ? sorte_of(xi.w) : 0;
    const double dFi = a.dF[i];
    const int nn = a.nnb[i];
    const int *row = a.nbl + i * L + sub;
    for (int m = sub; m < nn; m += L, row += a.rowstride) {
      const int j = __ldcs(row);
      const double4 xj = ld_pos(a.pos + j);
      const double dx = xj.x - xi.x, dy = xj.y - xi.y, dz = xj.z - xi.z;
      const double r2 = r2_exact(dx, dy, dz);
      double grad;
      if (!MULTI) {
        if (!(r2 < T.rho.end[0])) continue;                // :1172
        int k; double chi;
        tab_index(r2, T.rho.begin[0], r2, T.rho.invstep[0], k, chi, is_short);
        const double2 g = ld2(T.rhoG + (size_t) k * 2);
        grad = 0.5 * (dFi + __ldg(a.dF + j)) * fma(chi, g.y, g.x);   // :1203, col1 == col2
      } else {
        const int jt = sorte_of(xj.w);
        const int col1 = jt * nt + it, col2 = it * nt + jt;
        if (!((r2 < T.rho.end[col1]) || (r2 < T.rho.end[col2]))) continue;
        int k; double chi;
        // rho_i' from column col1, rho_j' from col2; both evaluated with the clamp, as DERIV_FUNC does
        tab_index(r2, T.rho.begin[col1], T.rho.end[col1], T.rho.invstep[col1], k, chi, is_short);
        double2 g = ld2(T.rhoG + ((size_t) k * T.rho.ncols + col1) * 2);
        const double rho_i_strich = fma(chi, g.y, g.x);
        double rho_j_strich = rho_i_strich;
        if (col1 != col2) {
          tab_index(r2, T.rho.begin[col2], T.rho.end[col2], T.rho.invstep[col2], k, chi, is_short);
          g = ld2(T.rhoG + ((size_t) k * T.rho.ncols + col2) * 2);
          rho_j_strich = fma(chi, g.y, g.x);
        }
        grad = 0.5 * (dFi * rho_j_strich + __ldg(a.dF + j) * rho_i_strich);
      }
      const double gx = dx * grad, gy = dy * grad, gz = dz * grad;
      fx += gx; fy += gy; fz += gz;
      vir = fma(r2, grad, vir);                            // SPROD(d,force) = r2*grad (:1280)
      if (STRESS) { s0 = fma(dx, gx, s0); s1 = fma(dy, gy, s1); s2 = fma(dz, gz, s2);
                    s3 = fma(dy, gz, s3); s4 = fma(dz, gx, s4); s5 = fma(dx, gy, s5); }
    }
  }
  if (L > 1) {
    fx = lanes_sum<L>(fx); fy = lanes_sum<L>(fy); fz = lanes_sum<L>(fz); vir = lanes_sum<L>(vir);
    if (STRESS) { s0 = lanes_sum<L>(s0); s1 = lanes_sum<L>(s1); s2 = lanes_sum<L>(s2);
                  s3 = lanes_sum<L>(s3); s4 = lanes_sum<L>(s4); s5 = lanes_sum<L>(s5); }
  }
  double red[1] = {0.0};
  if (act && sub == 0) {
    double4 f = a.frc[i];
    f.x += fx; f.y += fy; f.z += fz;
    a.frc[i] = f;
    if (STRESS) {
      double *p = a.presstens + i;
      p[0] -= 0.5 * s0; p[a.pstride] -= 0.5 * s1; p[2 * a.pstride] -= 0.5 * s2;
      p[3 * a.pstride] -= 0.5 * s3; p[4 * a.pstride] -= 0.5 * s4; p[5 * a.pstride] -= 0.5 * s5;
    }
    red[0] = -0.5 * vir;
  }
  if (is_short) atomicExch(&a.flags[FL_SHORT], 1);
  block_sum_store<1>(red, a.partial);
}

// ----------------------------------------------------------------------------------------------------
// deterministic second reduction stage: one block sums the per-block partials in a fixed order
// (replaces the MPI_Allreduce operand build-up of src/imd_forces_nbl.c:1975-1994 on one rank)
// ----------------------------------------------------------------------------------------------------
struct Slots { int s[8]; };
__global__ void k_reduce_partials(const double *partial, int nblocks, int nv, double *scal, Slots slots, int accumulate)
{
  __shared__ double sm[256];
  for (int v = 0; v < nv; v++) {
    double x = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 256) x += partial[(size_t) b * nv + v];
    sm[threadIdx.x] = x;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) { if (accumulate) scal[slots.s[v]] += sm[0]; else scal[slots.s[v]] = sm[0]; }
    __syncthreads();
  }
}

int reduce_finish(imdb200_sim *s, int nblocks, int nvals, const int *slots, int accumulate)
{
  Slots sl;
  for (int i = 0; i < 8; i++) sl.s[i] = i < nvals ? slots[i] : 0;
  k_reduce_partials<<<1, 256, 0, s->stream>>>(s->d_partial, nblocks, nvals, s->d_scal, sl, accumulate);
  LAUNCH_CHECK();
  return 0;
}

// ----------------------------------------------------------------------------------------------------
// launch wrappers: pick the template instance
// ----------------------------------------------------------------------------------------------------
static FArgs make_args(imdb200_sim *s)
{
  FArgs a;
  a.pos = s->pos; a.frc = s->frc; a.rho = s->rho; a.dF = s->dF; a.nbl = s->nbl; a.nnb = s->nnb;
  a.n_own = s->n_own; a.rowstride = s->n_pad * s->lanes;
  a.presstens = s->presstens; a.pstride = s->cap_atoms;
  a.partial = s->d_partial; a.flags = s->d_flags;
  return a;
}

template <int L, bool EAM, bool MULTI, bool FUSED>
static void launch1(imdb200_sim *s, const FArgs &a, int nb)
{
  if (s->press_calc) k_pass1<L, EAM, MULTI, FUSED, true><<<nb, FBLOCK, 0, s->stream>>>(a, s->tabs);
  else               k_pass1<L, EAM, MULTI, FUSED, false><<<nb, FBLOCK, 0, s->stream>>>(a, s->tabs);
}
template <int L> static void launch1_L(imdb200_sim *s, const FArgs &a, int nb)
{
  const bool multi = s->tabs.ntypes > 1;
  if (!s->tabs.have_eam) { if (multi) launch1<L, false, true, false>(s, a, nb); else launch1<L, false, false, false>(s, a, nb); }
  else if (s->tabs.fused) { if (multi) launch1<L, true, true, true>(s, a, nb); else launch1<L, true, false, true>(s, a, nb); }
  else { if (multi) launch1<L, true, true, false>(s, a, nb); else launch1<L, true, false, false>(s, a, nb); }
}
template <int L> static void launch2_L(imdb200_sim *s, const FArgs &a, int nb)
{
  const bool multi = s->tabs.ntypes > 1;
  if (multi) { if (s->press_calc) k_pass2<L, true, true><<<nb, FBLOCK, 0, s->stream>>>(a, s->tabs);
               else k_pass2<L, true, false><<<nb, FBLOCK, 0, s->stream>>>(a, s->tabs); }
  else { if (s->press_calc) k_pass2<L, false, true><<<nb, FBLOCK, 0, s->stream>>>(a, s->tabs);
         else k_pass2<L, false, false><<<nb, FBLOCK, 0, s->stream>>>(a, s->tabs); }
}

static int nblocks_for(imdb200_sim *s) { return cdiv(s->n_own * s->lanes, FBLOCK); }

int forces_pass1(imdb200_sim *s)
{
  FArgs a = make_args(s);
  const int nb = nblocks_for(s);
  switch (s->lanes) {
    case 1: launch1_L<1>(s, a, nb); break;
    case 2: launch1_L<2>(s, a, nb); break;
    case 4: launch1_L<4>(s, a, nb); break;
    case 8: launch1_L<8>(s, a, nb); break;
    case 16: launch1_L<16>(s, a, nb); break;
    case 32: launch1_L<32>(s, a, nb); break;
    default: return imdb_fail(IMDB200_ERR_ARG, "lanes_per_atom must be a power of two <= 32");
  }
  LAUNCH_CHECK();
  const int slots[2] = {SC_EPOT, SC_VIRIAL};
  return reduce_finish(s, nb, 2, slots, 0);
}

int forces_pass2(imdb200_sim *s)
{
  FArgs a = make_args(s);
  const int nb = nblocks_for(s);
  switch (s->lanes) {
    case 1: launch2_L<1>(s, a, nb); break;
    case 2: launch2_L<2>(s, a, nb); break;
    case 4: launch2_L<4>(s, a, nb); break;
    case 8: launch2_L<8>(s, a, nb); break;
    case 16: launch2_L<16>(s, a, nb); break;
    case 32: launch2_L<32>(s, a, nb); break;
    default: return imdb_fail(IMDB200_ERR_ARG, "lanes_per_atom must be a power of two <= 32");
  }
  LAUNCH_CHECK();
  const int slots[1] = {SC_VIRIAL};
  return reduce_finish(s, nb, 1, slots, 1);
}
