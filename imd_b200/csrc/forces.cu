// forces.cu -- tabulated pair + two-pass EAM forces over the Verlet list (calc_forces,
// src/imd_forces_nbl.c:281-1999; PAIR and EAM2 branches).
//
// The reference walks a HALF list and scatters -f, rho and Epot/2 into atom j (:513-545, 595-610,
// 1267-1281).  On a GPU that scatter would be ~8 FP64 atomics per pair (shared-memory FP64 atomics are
// CAS loops, global ones are bound by the L2 atomic units); instead the list is FULL and every atom
// gathers: each pair is evaluated from both ends, no atomics, no write conflicts, and the per-atom
// results (force, Epot, rho, stress) are complete when the thread ends, so the embedding lookup
// (:1079-1095) is fused into the tail of pass 1.  Sums differ from the reference only in their order
// (parity tolerance 1e-10, noise floor 1e-13 -- SURVEY.md section 0).
//
// What bounds these kernels (profiles/): the L1/shared-memory data pipe (128 B/clk/SM), i.e. the bytes
// gathered per pair, not HBM and not the FP64 pipe.  The layout is therefore built around bytes per pair:
//   * one 32-byte atom record per neighbour, fetched with ONE 256-bit load; in single-species EAM the
//     pass-2 record carries 2F'(rho_j) in .w so that pass 2 needs no second gather;
//   * potential tables staged in shared memory in coefficient form, 24 bytes per lookup in pass 1
//     (value; the gradient of phi is derived in registers) and 16 bytes in pass 2;
//   * persistent CTAs, one per SM, each walking a contiguous range of cell-sorted atoms so that the
//     position gathers hit in L1;
//   * the list is stored in warp blocks: a warp streams one contiguous block, 128 bytes per iteration.
// Thread mapping: L lanes cooperate on one atom (L = 1 for big systems, up to 32 for small ones); lane l
// takes list entries l, l+L, ... and the partial sums are combined with xor shuffles.
#include "internal.cuh"
#include <stdlib.h>
#include <string.h>

// entries per block of the software-pipelined list walk, and threads per CTA (one persistent CTA per SM)
#ifndef IMDB_DEPTH
#define IMDB_DEPTH 4
#endif
#ifndef IMDB_NT
#define IMDB_NT 640
#endif
#define FDEPTH IMDB_DEPTH
// pass 2 has its own block depth and CTA size
#ifndef IMDB_DEPTH2
#define IMDB_DEPTH2 4
#endif
#ifndef IMDB_NT2
#define IMDB_NT2 640
#endif
#define FDEPTH2 IMDB_DEPTH2
// Entries per block whose position gather goes through the TEX front end of the L1 (tex1Dfetch on a linear int4
// texture over the same records) instead of the LSU one (ld.global).  The L1 data stage has one pipe per front end;
// pass 1 keeps the LSU pipe busy with the table look-ups in shared memory, so all its gathers go through TEX
// (1.62 -> 1.33 ms at 4 M atoms), pass 2 has little table traffic and splits its gathers between the two pipes.
#ifndef IMDB_TEX1
#define IMDB_TEX1 4
#endif
#ifndef IMDB_TEX2
#define IMDB_TEX2 2
#endif
#ifndef IMDB_TEX1_SKIP
#define IMDB_TEX1_SKIP 4
#endif
// This file is compiled twice (Makefile): IMDB_CUBIC=0 holds the quadratic (PAIR_INT2) kernels and everything the
// two builds share, IMDB_CUBIC=1 the same kernels for the cubic table modes (PAIR_INT3 / PAIR_INT_SP, one more
// coefficient per lookup).  The kernels carry the flag as a template argument so that their symbols differ.
// IMDB_EEAM=1 adds the extended-EAM terms (EEAM builds of the reference, src/imd_forces_nbl.c:591-610, 1090-1095,
// 1181-1208): p_i = sum rho^2 in pass 1, a second embedding look-up M(p_i), and the dM terms in pass 2.
// Timing experiments only (tools/build_variant.sh; results are WRONG with either of them, never part of the product):
// IMDB_EXP_BCAST=C makes the C lanes of a group walk the list of the group's first atom, i.e. every gather
// instruction touches 32/C distinct records -- the cost of a gather with C-fold broadcast at unchanged list length;
// IMDB_EXP_KCONST pins the table interval to two rows -- look-ups without bank conflicts.
#ifndef IMDB_EXP_BCAST
#define IMDB_EXP_BCAST 1
#endif
#ifndef IMDB_EXP_SKIP
#define IMDB_EXP_SKIP 0          // experiment: every IMDB_EXP_SKIP-th neighbour is not gathered (what a gather mask would save)
#endif
// list rows are requested two blocks ahead instead of one (the rows stream from HBM: ~1 us under load)
#ifndef IMDB_PF2
#define IMDB_PF2 0
#endif
// double-buffered gathers: see the comment at GATHER1
#ifndef IMDB_DB
#define IMDB_DB 0
#endif
#ifndef IMDB_EXP_NOGATHER
#define IMDB_EXP_NOGATHER 0
#endif
// single-species pass 1 on raw table samples (32 KB of shared memory instead of 96 KB, eight more FP64 operations per pair)
#ifndef IMDB_RAW1
#define IMDB_RAW1 0
#endif
#ifdef IMDB_EXP_KCONST
#define EXP_K(k) ((k) = 100 + ((k) & 1))
#else
#define EXP_K(k) ((void) 0)
#endif
// branch-free block bodies (see FAST1 / FAST2 below): measured on B200 at 4 M atoms, pass 1 1.368 -> 1.330 ms,
// pass 2 1.054 -> 1.201 ms (its body is too short to gain from overlap, the masking costs more) -- on for pass 1 only
#ifndef IMDB_BRANCHFREE
#define IMDB_BRANCHFREE 1
#endif
#ifndef IMDB_BRANCHFREE2
#define IMDB_BRANCHFREE2 0
#endif
#ifndef IMDB_CUBIC
#define IMDB_CUBIC 0
#endif
#ifndef IMDB_EEAM
#define IMDB_EEAM 0
#endif
#if IMDB_CUBIC && IMDB_EEAM
#define IMPL(name) name##_cubic_eeam
#elif IMDB_CUBIC
#define IMPL(name) name##_cubic
#elif IMDB_EEAM
#define IMPL(name) name##_quad_eeam
#else
#define IMPL(name) name##_quad
#endif
#define IMDB_BASE_TU (!IMDB_CUBIC && !IMDB_EEAM)
static constexpr bool CUBIC = IMDB_CUBIC != 0;
static constexpr bool EEAMC = IMDB_EEAM != 0;

struct FArgs {
  const double4 *pos;
  double4 *posdf;
  double4 *frc;
  double *rho, *dF;
  double *eam_p, *dM;            // EEAM: p_i = sum rho^2 and M'(p_i) (EAM_P, EAM_DM)
  const int *nbl;
  cudaTextureObject_t tpos, tposdf;   // the same atom records as linear int4 textures (two texels per atom)
  int use_tex;                        // 0: the atom arrays exceed the 1-D linear texture limit, every gather on the LSU path
  const unsigned long long *nnbc;
  int cls_fixed;                 // highest skin class to walk, or -1: from ctl->disp2 (skin_class_of)
  double cls_w;                  // width of a skin class
  const StepCtl *ctl;
  int pblock0;                        // first block slot of this launch in the partial-sum array
  const int *worder; long w0, wn;     // warps [w0, w0+wn) of the processing order (nullptr: identity, see imdb200_sim::worder)
  long n_own;
  int rows;                      // max_nb / L
  double *presstens; long pstride;
  double *partial;
  int *flags;
  // fused move_atoms + check_nblist in the tail of pass 2 (single-species EAM: pass 2 gathers posdf, never pos)
  double4 *pos_rw, *mom;
  const double *nblpos; long nstride;
  const double *scal;
  unsigned long long *maxd2;
  double dt; int nvt;
  const double *glob; double nactive, temperature, tauber;   // Berendsen variant of NVE (integrate_atom)
};

__device__ __forceinline__ double4 ld_atom_tex(cudaTextureObject_t t, int j)
{
  const int4 a = tex1Dfetch<int4>(t, 2 * j), b = tex1Dfetch<int4>(t, 2 * j + 1);
  return make_double4(__hiloint2double(a.y, a.x), __hiloint2double(a.w, a.z), __hiloint2double(b.y, b.x), __hiloint2double(b.w, b.z));
}

// 16-byte shared-memory load by 32-bit shared-window address (no generic-address conversion in the inner loop)
__device__ __forceinline__ double2 lds2(unsigned addr)
{
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

template <int L> __device__ __forceinline__ double lanes_sum(double v)
{
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// cooperative copy global -> shared, 16 bytes per thread per iteration
__device__ __forceinline__ void stage(void *dst, const void *src, int bytes)
{
  const int4 *s = reinterpret_cast<const int4 *>(src);
  int4 *d = reinterpret_cast<int4 *>(dst);
  for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) d[i] = __ldg(s + i);
}

// ----------------------------------------------------------------------------------------------------
// pass 1: pair potential + host electron density (src/imd_forces_nbl.c:422-981), then the embedding
// energy F(rho_i) and 2F'(rho_i) (:1079-1095)
// ----------------------------------------------------------------------------------------------------
template <int NT, int L, bool EAM, bool MULTI, bool SHARED, bool STRESS, bool TSMEM, bool CUB, bool EE>
__global__ void __launch_bounds__(NT, 1) k_pass1(FArgs a, DevTables T)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  STEP_GATE(a.ctl);
  const int cls_shift = NBL_CBITS * (a.cls_fixed >= 0 ? a.cls_fixed : skin_class_of(a.ctl->disp2, a.cls_w));
  const double2 *pAB = T.pairAB, *rAB = T.rhoAB;
  const double *pC = T.pairC, *rC = T.rhoC;
  const double2 *pCD = T.pairCD, *rCD = T.rhoCD;           // cubic modes: (c2,c3)
  constexpr bool FUSED = EAM && !MULTI && SHARED && !CUB;  // one 48-byte record per interval, see DevTables::fused
  constexpr bool FAST1 = FUSED && TSMEM && !STRESS && !EE && IMDB_BRANCHFREE;   // the branch-free block body below
  constexpr bool RAW = MULTI && TSMEM && !CUB;             // several species: raw samples of the distinct columns in shared memory
  constexpr bool MU = MULTI && !RAW;                       // several species with per-column table headers (general path)
  const double *rawP = nullptr, *rawR = nullptr;
  const unsigned s_tab = (unsigned) __cvta_generic_to_shared(smem_raw);
  const unsigned k_max = (unsigned) (T.fused_rows - 1);
  const double2 *fT = T.fused;
  if (TSMEM && FUSED && IMDB_RAW1 && FAST1) {
    stage(smem_raw, T.fraw, (T.fused_rows + 2) * 16);
    __syncthreads();
  } else if (TSMEM && FUSED) {
    stage(smem_raw, T.fused, T.fused_rows * 48);
    __syncthreads();
    fT = reinterpret_cast<const double2 *>(smem_raw);
  } else if (TSMEM && CUB) {
    // [phi (c0,c1)] [rho (c0,c1)] [phi (c2,c3)] [rho (c2,c3)]
    const int np = T.pair.nrows * T.pair.ncols, nr = EAM ? T.rho.nrows * T.rho.ncols : 0;
    double2 *sAB = reinterpret_cast<double2 *>(smem_raw);
    stage(sAB, T.pairAB, np * 16);
    stage(sAB + np + nr, T.pairCD, np * 16);
    if (EAM) { stage(sAB + np, T.rhoAB, nr * 16); stage(sAB + 2 * np + nr, T.rhoCD, nr * 16); }
    __syncthreads();
    pAB = sAB; rAB = sAB + np; pCD = sAB + np + nr; rCD = sAB + 2 * np + nr;
  } else if (TSMEM && RAW) {
    // raw samples of the distinct columns: [phi (nrows+2) x nuP] [rho (nrows+2) x nuR], both padded to 16 bytes
    const int npd = ((T.pair.nrows + 2) * T.nuP + 1) & ~1, nrd = EAM ? ((T.rho.nrows + 2) * T.nuR + 1) & ~1 : 0;
    double *sP = reinterpret_cast<double *>(smem_raw);
    stage(sP, T.rawP, npd * 8);
    if (EAM) stage(sP + npd, T.rawR, nrd * 8);
    __syncthreads();
    rawP = sP; rawR = sP + npd;
  } else if (TSMEM) {
    // [phi (c0,c1)] [rho (c0,c1)] [phi c2] [rho c2]
    const int np = T.pair.nrows * T.pair.ncols, nr = EAM ? T.rho.nrows * T.rho.ncols : 0;
    const int npe = (np + 1) & ~1, nre = (nr + 1) & ~1;                // c2 arrays are padded to even length
    double2 *sAB = reinterpret_cast<double2 *>(smem_raw);
    double *sC = reinterpret_cast<double *>(sAB + np + nr);
    stage(sAB, T.pairAB, np * 16);
    stage(sC, T.pairC, npe * 8);
    if (EAM) { stage(sAB + np, T.rhoAB, nr * 16); stage(sC + npe, T.rhoC, nre * 8); }
    __syncthreads();
    pAB = sAB; rAB = sAB + np; pC = sC; rC = sC + npe;
  }
  const int nt = T.ntypes, nuP = T.nuP, nuR = T.nuR;
  // per-column constants of the single-species case (and of several species with one header per table) live in registers
  const double p_end0 = T.pair.end[0], p_is0 = T.pair.invstep[0], p_nb0 = -T.pair.begin[0] * T.pair.invstep[0];
  const double r_end0 = EAM ? T.rho.end[0] : 0.0, r_is0 = EAM ? T.rho.invstep[0] : 0.0,
               r_nb0 = EAM ? -T.rho.begin[0] * T.rho.invstep[0] : 0.0;
  // every CTA walks a contiguous run of the warp order: whole warps of 32 thread slots
  const long per = (a.wn + gridDim.x - 1) / gridDim.x;
  const long w_end = min(a.wn, (long) (blockIdx.x + 1) * per);
  double red[2] = {0.0, 0.0};
  int is_short = 0;
  for (long wv = (long) blockIdx.x * per + (threadIdx.x >> 5); wv < w_end; wv += NT / 32) {
    const long slot = (a.worder ? (long) a.worder[a.w0 + wv] : a.w0 + wv) * 32 + (threadIdx.x & 31);
    const long i = slot / L;
    const int sub = (int) (slot % L);
    const bool act = i < a.n_own;
    double fx = 0.0, fy = 0.0, fz = 0.0, ee = 0.0, rh = 0.0, vir = 0.0, ph = 0.0;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0, s5 = 0.0;
    int it = 0;
    double4 xi = make_double4(0.0, 0.0, 0.0, 0.0);
    unsigned upk = 0u, urk = 0u;                           // RAW: distinct-column index of (it, jt), four bits per jt
    if (act) {
      xi = a.pos[i];
      if (MULTI) it = sorte_of(xi.w);
      if (RAW) {
        for (int jt = 0; jt < nt; jt++) { upk |= (unsigned) T.umapP[it * nt + jt] << (4 * jt);
                                          if (EAM) urk |= (unsigned) T.umapR[it * nt + jt] << (4 * jt); }
      }
      const long lslot = slot & ~(long) (IMDB_EXP_BCAST - 1);
      const int nn = (int) ((a.nnbc[lslot / L] >> cls_shift) & ((1u << NBL_CBITS) - 1));
      const int *row = a.nbl + (size_t) (lslot >> 5) * ((size_t) a.rows * 32) + (lslot & 31);
      // Software pipeline over the list: the D entries of a block are in registers when the block starts (they
      // were loaded during the previous block), their D position gathers are issued together, the next block's
      // entries are requested, and only then the arithmetic of the block runs.  The list stream comes from HBM
      // (~1 us) and the gathers from L1/L2: without this every warp sits out both latencies once per entry.
      int jq[FDEPTH];
#pragma unroll
      for (int d = 0; d < FDEPTH; d++) jq[d] = (sub + d * L < nn) ? __ldcs(row + d * 32) : -1;
#if IMDB_PF2
      int jn[FDEPTH];                                     // the list rows of the block after the next one
#pragma unroll
      for (int d = 0; d < FDEPTH; d++) jn[d] = (sub + (FDEPTH + d) * L < nn) ? __ldcs(row + (FDEPTH + d) * 32) : -1;
#endif
#if IMDB_EXP_NOGATHER   // experiment: neighbour positions made up from the index, no memory access (arithmetic + table floor)
#define GATHER1(X) _Pragma("unroll") for (int d = 0; d < FDEPTH; d++) { const int j_ = jq[d]; \
          X[d] = j_ >= 0 ? make_double4(xi.x + 1.0 + (j_ & 7) * 0.3, xi.y + ((j_ >> 3) & 7) * 0.3, xi.z + ((j_ >> 6) & 7) * 0.3, xi.w) : xi; }
#else
#define GATHER1(X) _Pragma("unroll") for (int d = 0; d < FDEPTH; d++) { const int j_ = jq[d] >= 0 ? (MULTI ? jq[d] & NBL_JMASK : jq[d]) : (int) i; \
          X[d] = (d < IMDB_TEX1 && a.use_tex) ? ld_atom_tex(a.tpos, j_) : ld_atom(a.pos + j_); }
#endif
#if IMDB_DB
      // Double buffering: the gathers of block b+1 are in flight while block b is evaluated (its own were issued one
      // iteration earlier), so a warp overlaps the L2 round trip of its gathers with its own arithmetic.
      double4 xq[FDEPTH];
      GATHER1(xq)
#pragma unroll
      for (int d = 0; d < FDEPTH; d++) jq[d] = (sub + (FDEPTH + d) * L < nn) ? __ldcs(row + (FDEPTH + d) * 32) : -1;
#endif
      for (int m = sub; m < nn; m += FDEPTH * L, row += FDEPTH * 32) {
#if IMDB_DB
        double4 xn[FDEPTH];
        GATHER1(xn)
#pragma unroll
        for (int d = 0; d < FDEPTH; d++) jq[d] = (m + (2 * FDEPTH + d) * L < nn) ? __ldcs(row + (2 * FDEPTH + d) * 32) : -1;
#else
        double4 xq[FDEPTH];
        GATHER1(xq)
#if IMDB_PF2
#pragma unroll
        for (int d = 0; d < FDEPTH; d++) { jq[d] = jn[d]; jn[d] = (m + (2 * FDEPTH + d) * L < nn) ? __ldcs(row + (2 * FDEPTH + d) * 32) : -1; }
#else
#pragma unroll
        for (int d = 0; d < FDEPTH; d++) jq[d] = (m + (FDEPTH + d) * L < nn) ? __ldcs(row + (FDEPTH + d) * 32) : -1;
#endif
#endif
        if (FAST1) {
          // Benchmark path (single species, phi and rho on one grid, quadratic tables in shared memory): the D entries of
          // a block are evaluated WITHOUT branches -- entries beyond the list end or outside the cut-offs run through
          // the same arithmetic with a clamped table index and are masked out of the sums (adding +0.0 leaves every
          // sum bit-identical to skipping the entry).  Straight-line code lets the D independent dependency chains
          // (gather -> r2 -> index -> LDS -> polynomials) overlap inside one warp; with the branchy form every entry
          // waited out its own chain (5 warps per scheduler cannot hide that), see profiles/README.md round 2.
#pragma unroll
          for (int d = 0; d < FDEPTH; d++) {
            const bool valid = m + d * L < nn;
            const double4 xj = xq[d];
            const double dx = xj.x - xi.x, dy = xj.y - xi.y, dz = xj.z - xi.z;
            const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            const bool inp = valid && r2 <= p_end0;                                  // :493
            const bool inr = valid && r2 < r_end0;                                   // :588
            double t = fma(r2, p_is0, p_nb0);                                        // (r2 - begin) * invstep
            if (t < 0.0) { t = 0.0; if (inp || inr) is_short = 1; }
            const double tk = __dadd_rz(t, IMDB_TWO52);
            // entries beyond the cut-off index past the table: clamp the row (they are masked out below)
#if IMDB_RAW1
            const unsigned rec = s_tab + 16u * min((unsigned) __double2loint(tk), k_max);
            const double chi = t - (tk - IMDB_TWO52);
            const double2 q0 = lds2(rec), q1 = lds2(rec + 16), q2 = lds2(rec + 32);    // (phi, rho) of rows k, k+1, k+2
            const double pc2 = 0.5 * ((q2.x - 2 * q1.x) + q0.x), pc1 = (q1.x - q0.x) - pc2;
            const double rc2 = 0.5 * ((q2.y - 2 * q1.y) + q0.y), rc1 = (q1.y - q0.y) - rc2;
            double pot = fma(chi, fma(chi, pc2, pc1), q0.x), grad = (p_is0 + p_is0) * fma(chi + chi, pc2, pc1),
                   rv = fma(chi, fma(chi, rc2, rc1), q0.y);
#else
            const unsigned rec = s_tab + 48u * min((unsigned) __double2loint(tk), k_max);
            const double chi = t - (tk - IMDB_TWO52);
            const double2 a0 = lds2(rec), a1 = lds2(rec + 16), a2 = lds2(rec + 32);    // (phi c0,c1) (phi c2, rho c2) (rho c0,c1)
            double pot = tab_val(a0, a1.x, chi), grad = tab_grad(a0, a1.x, chi, p_is0 + p_is0), rv = tab_val(a2, a1.y, chi);
#endif
            pot = inp ? pot : 0.0; grad = inp ? grad : 0.0; rv = inr ? rv : 0.0;
            fx = fma(dx, grad, fx); fy = fma(dy, grad, fy); fz = fma(dz, grad, fz);
            ee += pot;
            vir = fma(r2, grad, vir);
            rh += rv;
          }
        } else
#pragma unroll
        for (int d = 0; d < FDEPTH; d++) {
        if (m + d * L >= nn) break;
        const double4 xj = xq[d];
        const double dx = xj.x - xi.x, dy = xj.y - xi.y, dz = xj.z - xi.z;
        const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
        const int jt = MULTI ? sorte_of(xj.w) : 0;
        const int col = MULTI ? it * nt + jt : 0;
        const bool inp = r2 <= (MU ? T.pair.end[col] : p_end0);                  // :493
        const bool inr = EAM && r2 < (MU ? T.rho.end[col] : r_end0);             // :588
        if (!(inp || inr)) continue;
        int k = 0, kr = 0; double chi = 0.0, chir = 0.0;
        const double pis = MU ? T.pair.invstep[col] : p_is0;
        if (inp || SHARED) { tab_index_fast(r2, MU ? -T.pair.begin[col] * pis : p_nb0, pis, k, chi, is_short); EXP_K(k); }
        if (EAM) {
          if (SHARED) { kr = k; chir = chi; }
          else if (inr) { const double ris = MU ? T.rho.invstep[col] : r_is0;
                          tab_index_fast(r2, MU ? -T.rho.begin[col] * ris : r_nb0, ris, kr, chir, is_short); }
        }
        double2 fmid = make_double2(0.0, 0.0);
        if (FUSED) fmid = fT[3 * k + 1];                     // (phi c2, rho c2)
        if (inp) {
          const int e = MULTI ? k * T.pair.ncols + col : k;
          double pot, grad;
          if (RAW) {
            // PAIR_INT2's own operands: p0, p1, p2 of rows k..k+2, dv = p1-p0, d2v = p2-2p1+p0 (src/potaccess.h:345-349);
            // (c0,c1,c2) = (p0, dv - d2v/2, d2v/2) are what the coefficient tables hold, formed here with the same operations
            const double *t = rawP + k * nuP + ((upk >> (4 * jt)) & 15u);
            const double p0 = t[0], p1 = t[nuP], p2 = t[2 * nuP];
            const double c2 = 0.5 * ((p2 - 2 * p1) + p0), c1 = (p1 - p0) - c2;
            pot = tab_val(make_double2(p0, c1), c2, chi); grad = tab_grad(make_double2(p0, c1), c2, chi, pis + pis);
          } else {
          const double2 ab = FUSED ? fT[3 * k] : pAB[e];
          if (CUB) { const double2 cd = pCD[e]; pot = tab_val3(ab, cd, chi); grad = tab_grad3(ab, cd, chi, pis + pis); }
          else { const double c2 = FUSED ? fmid.x : pC[e]; pot = tab_val(ab, c2, chi); grad = tab_grad(ab, c2, chi, pis + pis); }
          }
          fx = fma(dx, grad, fx); fy = fma(dy, grad, fy); fz = fma(dz, grad, fz);
          ee += pot;
          vir = fma(r2, grad, vir);
          if (STRESS) { const double gx = dx * grad, gy = dy * grad, gz = dz * grad;
                        s0 = fma(dx, gx, s0); s1 = fma(dy, gy, s1); s2 = fma(dz, gz, s2);
                        s3 = fma(dy, gz, s3); s4 = fma(dz, gx, s4); s5 = fma(dx, gy, s5); }
        }
        if (inr) {
          const int e = MULTI ? kr * T.rho.ncols + col : kr;
          double rv;
          if (RAW) {
            const double *t = rawR + kr * nuR + ((urk >> (4 * jt)) & 15u);
            const double p0 = t[0], p1 = t[nuR], p2 = t[2 * nuR];
            const double c2 = 0.5 * ((p2 - 2 * p1) + p0), c1 = (p1 - p0) - c2;
            rv = tab_val(make_double2(p0, c1), c2, chir);
          }
          else if (CUB) rv = tab_val3(rAB[e], rCD[e], chir);
          else rv = FUSED ? tab_val(fT[3 * kr + 2], fmid.y, chir) : tab_val(rAB[e], rC[e], chir);
          rh += rv;
          if (EE) ph = fma(rv, rv, ph);                      // eam_p += rho_h*rho_h (:591-593)
        }
        }
#if IMDB_DB
#pragma unroll
        for (int d = 0; d < FDEPTH; d++) xq[d] = xn[d];
#endif
      }
#undef GATHER1
    }
    if (L > 1) {
      fx = lanes_sum<L>(fx); fy = lanes_sum<L>(fy); fz = lanes_sum<L>(fz);
      ee = lanes_sum<L>(ee); vir = lanes_sum<L>(vir);
      if (EAM) rh = lanes_sum<L>(rh);
      if (EE) ph = lanes_sum<L>(ph);
      if (STRESS) { s0 = lanes_sum<L>(s0); s1 = lanes_sum<L>(s1); s2 = lanes_sum<L>(s2);
                    s3 = lanes_sum<L>(s3); s4 = lanes_sum<L>(s4); s5 = lanes_sum<L>(s5); }
    }
    if (act && sub == 0) {
      double epot = 0.5 * ee;                              // pot *= 0.5 on both atoms (:535-545)
      if (EAM) {
        int k, dummy = 0; double chi;                     // PAIR_INT(pot, EAM_DF, embed_pot, ...) :1086
        tab_index(rh, T.embed.begin[it], T.embed.end[it], T.embed.invstep[it], k, chi, dummy);
        const double *e = T.embedVG + ((size_t) k * T.embed.ncols + it) * 8;   // c0 c1 | c2 c3 | g1 g2 | g3 -
        const double2 e0 = ld2(e), e1 = ld2(e + 2), e2 = ld2(e + 4);
        double dF;
        if (CUB) {
          epot += tab_val3(e0, e1, chi);
          dF = fma(chi, fma(chi, __ldg(e + 6), e2.y), e2.x);
        } else {
          epot += fma(chi, fma(chi, e1.x, e0.y), e0.x);
          dF = fma(chi, e2.y, e2.x);
        }
        a.rho[i] = rh;
        a.dF[i] = dF;
        if (EE) {                                          // PAIR_INT(pot, EAM_DM, emod_pot, ..., EAM_P) :1091-1094
          tab_index(ph, T.emod.begin[it], T.emod.end[it], T.emod.invstep[it], k, chi, dummy);
          const double *m = T.emodVG + ((size_t) k * T.emod.ncols + it) * 8;
          const double2 m0 = ld2(m), m1 = ld2(m + 2), m2 = ld2(m + 4);
          epot += tab_val3(m0, CUB ? m1 : make_double2(m1.x, 0.0), chi);
          a.eam_p[i] = ph;
          a.dM[i] = fma(chi, CUB ? fma(chi, __ldg(m + 6), m2.y) : m2.y, m2.x);
        }
        a.posdf[i] = make_double4(xi.x, xi.y, xi.z, dF);     // the pass-2 gather record
      }
      a.frc[i] = make_double4(fx, fy, fz, epot);
      if (STRESS) {                                        // -0.5 d (x) f per atom (:558-581)
        double *p = a.presstens + i;
        p[0] = -0.5 * s0; p[a.pstride] = -0.5 * s1; p[2 * a.pstride] = -0.5 * s2;
        p[3 * a.pstride] = -0.5 * s3; p[4 * a.pstride] = -0.5 * s4; p[5 * a.pstride] = -0.5 * s5;
      }
      red[0] += epot;
      red[1] += -0.5 * vir;                                // virial -= r2*grad once per pair (:555)
    }
  }
  if (is_short) atomicExch(&a.flags[FL_SHORT], 1);
  block_sum_store<2>(red, a.partial + (size_t) a.pblock0 * 2);
}

// ----------------------------------------------------------------------------------------------------
// pass 2: EAM forces (src/imd_forces_nbl.c:1117-1322)
// ----------------------------------------------------------------------------------------------------
template <int NT, int L, bool MULTI, bool STRESS, bool TSMEM, bool FUSE, bool CUB, bool EE>
__global__ void __launch_bounds__(NT, 1) k_pass2(FArgs a, DevTables T)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  STEP_GATE(a.ctl);
  const int cls_shift = NBL_CBITS * (a.cls_fixed >= 0 ? a.cls_fixed : skin_class_of(a.ctl->disp2, a.cls_w));
  const double2 *rH = T.rhoH;
  const double *rH3 = T.rhoH3;                       // cubic modes: rho'/2 = h1 + chi*(h2 + chi*h3)
  constexpr bool FAST2 = !MULTI && !EE && !CUB && !STRESS && TSMEM && IMDB_BRANCHFREE2;
  // Several species with the tables in shared memory (the host passes TSMEM only when this applies: quadratic, no EEAM, every
  // rho column on one r^2 grid): (h1,h2) of the DISTINCT rho columns, see DevTables::rhoHd
  constexpr bool FASTM = MULTI && TSMEM && !CUB && !EE;
  const unsigned s_tab = (unsigned) __cvta_generic_to_shared(smem_raw);
  const unsigned k_max = (unsigned) (T.rho.nrows - 1);
  const int nuR = T.nuR;
  if (FASTM) {
    stage(smem_raw, T.rhoHd, T.rho.nrows * nuR * 16);
    __syncthreads();
    rH = reinterpret_cast<const double2 *>(smem_raw);
  } else if (TSMEM) {
    const int nr = T.rho.nrows * T.rho.ncols;
    stage(smem_raw, T.rhoH, nr * 16);
    if (CUB) stage(smem_raw + (size_t) nr * 16, T.rhoH3, ((nr + 1) & ~1) * 8);
    __syncthreads();
    rH = reinterpret_cast<const double2 *>(smem_raw);
    rH3 = reinterpret_cast<const double *>(smem_raw + (size_t) nr * 16);
  }
  const int nt = T.ntypes;
  const double r_end0 = T.rho.end[0], r_is0 = T.rho.invstep[0], r_nb0 = -T.rho.begin[0] * T.rho.invstep[0];
  const double4 *gat = a.posdf;                      // x,y,z,F' in one record; the neighbour's type rides in the list entry
  const long per = (a.wn + gridDim.x - 1) / gridDim.x;
  const long w_end = min(a.wn, (long) (blockIdx.x + 1) * per);
  double red[3] = {0.0, 0.0, 0.0};                  // virial, and with FUSE the two kinetic-energy sums
  double d2max = 0.0;
  // FUSE: SC_EKIN still holds the previous step's kinetic energy here (this step's is written by the reduction behind us)
  const double ber_cc = FUSE ? berendsen_cc(a.glob[SC_EKIN], a.nactive, a.temperature, a.dt, a.tauber) : 1.0;
  int is_short = 0;
  for (long wv = (long) blockIdx.x * per + (threadIdx.x >> 5); wv < w_end; wv += NT / 32) {
    const long slot = (a.worder ? (long) a.worder[a.w0 + wv] : a.w0 + wv) * 32 + (threadIdx.x & 31);
    const long i = slot / L;
    const int sub = (int) (slot % L);
    const bool act = i < a.n_own;
    double fx = 0.0, fy = 0.0, fz = 0.0, vir = 0.0;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0, s5 = 0.0;
    if (act) {
      const double4 xi = gat[i];
      const int it = MULTI ? sorte_of(reinterpret_cast<const double *>(a.pos + i)[3]) : 0;
      const double dFi = xi.w;
      unsigned uk1 = 0u, uk2 = 0u;                        // FASTM: distinct-column index of (jt, it) and (it, jt), four bits per jt
      if (FASTM) for (int jt = 0; jt < nt; jt++) { uk1 |= (unsigned) T.umapR[jt * nt + it] << (4 * jt);
                                                   uk2 |= (unsigned) T.umapR[it * nt + jt] << (4 * jt); }
      const double dMi = EE ? a.dM[i] : 0.0;
      const long lslot = slot & ~(long) (IMDB_EXP_BCAST - 1);
      const int nn = (int) ((a.nnbc[lslot / L] >> cls_shift) & ((1u << NBL_CBITS) - 1));
      const int *row = a.nbl + (size_t) (lslot >> 5) * ((size_t) a.rows * 32) + (lslot & 31);
      int jq[FDEPTH2];                                    // software pipeline as in pass 1
#pragma unroll
      for (int d = 0; d < FDEPTH2; d++) jq[d] = (sub + d * L < nn) ? __ldcs(row + d * 32) : -1;
#if IMDB_PF2
      int jn[FDEPTH2];
#pragma unroll
      for (int d = 0; d < FDEPTH2; d++) jn[d] = (sub + (FDEPTH2 + d) * L < nn) ? __ldcs(row + (FDEPTH2 + d) * 32) : -1;
#endif
      // the TEX share is issued first (longer latency) and consumed last (entries FDEPTH2-IMDB_TEX2 .. FDEPTH2-1)
#if IMDB_EXP_NOGATHER
#define GATHER2(X, JC) _Pragma("unroll") for (int d = 0; d < FDEPTH2; d++) { const int j_ = jq[d]; if (MULTI || EE) JC[d] = j_ >= 0 ? j_ : (int) i; \
          X[d] = j_ >= 0 ? make_double4(xi.x + 1.0 + (j_ & 7) * 0.3, xi.y + ((j_ >> 3) & 7) * 0.3, xi.z + ((j_ >> 6) & 7) * 0.3, xi.w) : xi; }
#else
#define GATHER2(X, JC) _Pragma("unroll") for (int dd = 0; dd < FDEPTH2; dd++) { const int d = FDEPTH2 - 1 - dd; \
          const int e_ = jq[d] >= 0 ? jq[d] : (int) i; if (MULTI || EE) JC[d] = e_; const int j = MULTI ? e_ & NBL_JMASK : e_; \
          X[d] = (d >= FDEPTH2 - IMDB_TEX2 && a.use_tex) ? ld_atom_tex(a.tposdf, j) : ld_atom(gat + j); }
#endif
#if IMDB_DB
      double4 xq[FDEPTH2];
      int jc[(MULTI || EE) ? FDEPTH2 : 1];
      GATHER2(xq, jc)
#pragma unroll
      for (int d = 0; d < FDEPTH2; d++) jq[d] = (sub + (FDEPTH2 + d) * L < nn) ? __ldcs(row + (FDEPTH2 + d) * 32) : -1;
#endif
      for (int m = sub; m < nn; m += FDEPTH2 * L, row += FDEPTH2 * 32) {
#if IMDB_DB
        double4 xn[FDEPTH2];
        int jcn[(MULTI || EE) ? FDEPTH2 : 1];
        GATHER2(xn, jcn)
#pragma unroll
        for (int d = 0; d < FDEPTH2; d++) jq[d] = (m + (2 * FDEPTH2 + d) * L < nn) ? __ldcs(row + (2 * FDEPTH2 + d) * 32) : -1;
#else
        double4 xq[FDEPTH2];
        int jc[(MULTI || EE) ? FDEPTH2 : 1];
        GATHER2(xq, jc)
#if IMDB_PF2
#pragma unroll
        for (int d = 0; d < FDEPTH2; d++) { jq[d] = jn[d]; jn[d] = (m + (2 * FDEPTH2 + d) * L < nn) ? __ldcs(row + (2 * FDEPTH2 + d) * 32) : -1; }
#else
#pragma unroll
        for (int d = 0; d < FDEPTH2; d++) jq[d] = (m + (FDEPTH2 + d) * L < nn) ? __ldcs(row + (FDEPTH2 + d) * 32) : -1;
#endif
#endif
        if (FAST2) {
          // branch-free block body, see pass 1
#pragma unroll
          for (int d = 0; d < FDEPTH2; d++) {
            const bool valid = m + d * L < nn;
            const double4 xj = xq[d];
            const double dx = xj.x - xi.x, dy = xj.y - xi.y, dz = xj.z - xi.z;
            const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            const bool in = valid && r2 < r_end0;                                    // :1172
            double t = fma(r2, r_is0, r_nb0);
            if (t < 0.0) { t = 0.0; if (in) is_short = 1; }
            const double tk = __dadd_rz(t, IMDB_TWO52);
            const double2 h = lds2(s_tab + 16u * min((unsigned) __double2loint(tk), k_max));
            const double chi = t - (tk - IMDB_TWO52);
            double grad = (dFi + xj.w) * fma(chi, h.y, h.x);                         // 0.5*(dF_i+dF_j)*rho' (:1203)
            grad = in ? grad : 0.0;
            fx = fma(dx, grad, fx); fy = fma(dy, grad, fy); fz = fma(dz, grad, fz);
            vir = fma(r2, grad, vir);
          }
        } else
#pragma unroll
        for (int d = 0; d < FDEPTH2; d++) {
        if (m + d * L >= nn) break;
        const double4 xj = xq[d];
        const int j = (MULTI || EE) ? (MULTI ? jc[d] & NBL_JMASK : jc[d]) : 0;
        const double dx = xj.x - xi.x, dy = xj.y - xi.y, dz = xj.z - xi.z;
        const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
        double grad;
        if (FASTM) {
          // one header for all rho columns (DevTables::multi_uniform): one cut-off, one (k, chi), no clamp inside it
          if (!(r2 < r_end0)) continue;                                                // :1172
          int k; double chi;
          tab_index_fast(r2, r_nb0, r_is0, k, chi, is_short);
          const int jt = jc[d] >> NBL_TSHIFT;
          const unsigned u1 = (uk1 >> (4 * jt)) & 15u, u2 = (uk2 >> (4 * jt)) & 15u;
          const double2 h1 = rH[k * nuR + u1];
          const double rho_i_strich = fma(chi, h1.y, h1.x);                            // column (jt, it)
          double rho_j_strich = rho_i_strich;
          if (u1 != u2) { const double2 h2 = rH[k * nuR + u2]; rho_j_strich = fma(chi, h2.y, h2.x); }   // column (it, jt)
          grad = dFi * rho_j_strich + xj.w * rho_i_strich;
        } else if (!MULTI) {
          if (!(r2 < r_end0)) continue;                    // :1172
          int k; double chi;
          tab_index_fast(r2, r_nb0, r_is0, k, chi, is_short);
          EXP_K(k);
          if (!EE) {
            const double2 h = rH[k];
            const double hs = CUB ? fma(chi, rH3[k], h.y) : h.y;
            grad = (dFi + xj.w) * fma(chi, hs, h.x);       // 0.5*(dF_i+dF_j)*rho' (:1203), col1 == col2
          } else {
            // EEAM needs rho(r) as well: value and rho'/2 from the pass-1 coefficient arrays
            const double2 ab = T.rhoAB[k];
            double rv, hv;
            if (CUB) { const double2 cd = T.rhoCD[k]; rv = tab_val3(ab, cd, chi); hv = r_is0 * fma(chi, fma(3.0 * chi, cd.y, cd.x + cd.x), ab.y); }
            else { const double c2 = T.rhoC[k]; rv = tab_val(ab, c2, chi); hv = r_is0 * fma(chi + chi, c2, ab.y); }
            // 0.5*(dF_i+dF_j)*rho' + (dM_i+dM_j)*rho*rho'  (:1203-1208), rho' = 2*hv
            grad = (dFi + xj.w) * hv + (dMi + __ldg(a.dM + j)) * (rv * (hv + hv));
          }
        } else {
          const int jt = jc[d] >> NBL_TSHIFT;
          const int col1 = jt * nt + it, col2 = it * nt + jt;
          if (!((r2 < T.rho.end[col1]) || (r2 < T.rho.end[col2]))) continue;
          int k; double chi;
          // rho_i' from column col1, rho_j' from col2; both evaluated with the clamp, as DERIV_FUNC does
          tab_index(r2, T.rho.begin[col1], T.rho.end[col1], T.rho.invstep[col1], k, chi, is_short);
          double2 h = rH[k * T.rho.ncols + col1];
          const double rho_i_strich = fma(chi, CUB ? fma(chi, rH3[k * T.rho.ncols + col1], h.y) : h.y, h.x);
          double rho_j_strich = rho_i_strich, rho_i = 0.0, rho_j = 0.0;
          if (EE) { const int e = k * T.rho.ncols + col1;
                    rho_i = CUB ? tab_val3(T.rhoAB[e], T.rhoCD[e], chi) : tab_val(T.rhoAB[e], T.rhoC[e], chi); rho_j = rho_i; }
          if (col1 != col2) {
            tab_index(r2, T.rho.begin[col2], T.rho.end[col2], T.rho.invstep[col2], k, chi, is_short);
            h = rH[k * T.rho.ncols + col2];
            rho_j_strich = fma(chi, CUB ? fma(chi, rH3[k * T.rho.ncols + col2], h.y) : h.y, h.x);
            if (EE) { const int e = k * T.rho.ncols + col2;
                      rho_j = CUB ? tab_val3(T.rhoAB[e], T.rhoCD[e], chi) : tab_val(T.rhoAB[e], T.rhoC[e], chi); }
          }
          grad = dFi * rho_j_strich + xj.w * rho_i_strich;
          // + dM_i*rho_j*rho_j' + dM_j*rho_i*rho_i' (:1204-1208); the "strich" values here are rho'/2
          if (EE) grad += 2.0 * (dMi * rho_j * rho_j_strich + __ldg(a.dM + j) * rho_i * rho_i_strich);
        }
        fx = fma(dx, grad, fx); fy = fma(dy, grad, fy); fz = fma(dz, grad, fz);
        vir = fma(r2, grad, vir);                          // SPROD(d,force) = r2*grad (:1280)
        if (STRESS) { const double gx = dx * grad, gy = dy * grad, gz = dz * grad;
                      s0 = fma(dx, gx, s0); s1 = fma(dy, gy, s1); s2 = fma(dz, gz, s2);
                      s3 = fma(dy, gz, s3); s4 = fma(dz, gx, s4); s5 = fma(dx, gy, s5); }
        }
#if IMDB_DB
#pragma unroll
        for (int d = 0; d < FDEPTH2; d++) { xq[d] = xn[d]; if (MULTI || EE) jc[d] = jcn[d]; }
#endif
      }
#undef GATHER2
    }
    if (L > 1) {
      fx = lanes_sum<L>(fx); fy = lanes_sum<L>(fy); fz = lanes_sum<L>(fz); vir = lanes_sum<L>(vir);
      if (STRESS) { s0 = lanes_sum<L>(s0); s1 = lanes_sum<L>(s1); s2 = lanes_sum<L>(s2);
                    s3 = lanes_sum<L>(s3); s4 = lanes_sum<L>(s4); s5 = lanes_sum<L>(s5); }
    }
    if (act && sub == 0) {
      double4 f = a.frc[i];
      f.x += fx; f.y += fy; f.z += fz;
      a.frc[i] = f;
      if (FUSE) {                                        // move_atoms + check_nblist of this atom, see integrate_atom()
        const double4 xo = gat[i];
        double4 x = xo, p = a.mom[i];
        double rk[2];
        const double nx = a.nblpos[i], ny = a.nblpos[a.nstride + i], nz = a.nblpos[2 * a.nstride + i];
        const double d2 = a.nvt ? integrate_atom<true>(x, p, f, a.dt, a.scal[SC_ETA], 1.0, 1.0, 1.0, nx, ny, nz, rk)
                                : integrate_atom<false>(x, p, f, a.dt, 0.0, 1.0, 1.0, 1.0, nx, ny, nz, rk, ber_cc);
        a.mom[i] = p;
        double *xw = reinterpret_cast<double *>(a.pos_rw + i);   // .w (the types) stays as it is
        *reinterpret_cast<double2 *>(xw) = make_double2(x.x, x.y);
        xw[2] = x.z;
        red[1] += rk[0]; red[2] += rk[1];
        d2max = fmax(d2max, d2);
      }
      if (STRESS) {
        double *p = a.presstens + i;
        p[0] -= 0.5 * s0; p[a.pstride] -= 0.5 * s1; p[2 * a.pstride] -= 0.5 * s2;
        p[3 * a.pstride] -= 0.5 * s3; p[4 * a.pstride] -= 0.5 * s4; p[5 * a.pstride] -= 0.5 * s5;
      }
      red[0] += -0.5 * vir;
    }
  }
  if (is_short) atomicExch(&a.flags[FL_SHORT], 1);
  if (FUSE) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d2max = fmax(d2max, __shfl_xor_sync(0xffffffffu, d2max, o));
    if ((threadIdx.x & 31) == 0) atomicMax(a.maxd2, (unsigned long long) __double_as_longlong(d2max));
    block_sum_store<3>(red, a.partial + (size_t) a.pblock0 * 3);
  } else {
    double r1[1] = {red[0]};
    block_sum_store<1>(r1, a.partial + (size_t) a.pblock0);
  }
}

#if IMDB_BASE_TU
// ----------------------------------------------------------------------------------------------------
// deterministic second reduction stage: one block sums the per-block partials in a fixed order
// (replaces the MPI_Allreduce operand build-up of src/imd_forces_nbl.c:1975-1994 on one rank)
// ----------------------------------------------------------------------------------------------------
struct Slots { int s[8]; };
__global__ void k_reduce_partials(const double *partial, int nblocks, int nv, double *scal, Slots slots, int accumulate_mask,
                                  const StepCtl *ctl, int zero_maxd2)
{
  __shared__ double sm[256];
  STEP_GATE(ctl);
  if (zero_maxd2 && threadIdx.x == 0) scal[SC_MAXD2] = 0.0;     // the kernel that follows accumulates the new maximum
  for (int v = 0; v < nv; v++) {
    double x = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 256) x += partial[(size_t) b * nv + v];
    sm[threadIdx.x] = x;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) { if ((accumulate_mask >> v) & 1) scal[slots.s[v]] += sm[0]; else scal[slots.s[v]] = sm[0]; }
    __syncthreads();
  }
}

int reduce_finish(imdb200_sim *s, int nblocks, int nvals, const int *slots, int accumulate_mask, int zero_maxd2)
{
  Slots sl;
  for (int i = 0; i < 8; i++) sl.s[i] = i < nvals ? slots[i] : 0;
  k_reduce_partials<<<1, 256, 0, s->stream>>>(s->d_partial, nblocks, nvals, s->d_scal, sl, accumulate_mask, s->d_ctl, zero_maxd2);
  LAUNCH_CHECK();
  return 0;
}

#endif  // IMDB_BASE_TU

// ----------------------------------------------------------------------------------------------------
// launch wrappers: pick the template instance
// ----------------------------------------------------------------------------------------------------
// Highest list group a force call has to walk: fixed by the host when the displacement bound is void (skin skipping
// off, box or positions changed outside move_atoms since the build), else the kernels derive it from StepCtl::disp2
// (skin_class_of, internal.cuh).
static int skin_class_fixed(const imdb200_sim *s)
{
  return (!s->skin_skip || s->skin_all) ? NBL_CLASSES : -1;
}

#if IMDB_BASE_TU
// The atom records as linear textures.  The L1 of an SM has two front ends, LSU (ld.global, shared memory) and TEX;
// the force passes saturate the LSU data pipe with table look-ups + gathers while TEX idles, so a fixed share of
// the gathers of every block goes through TEX instead (profiles/README.md, "two pipes").
static int tex_make(cudaTextureObject_t *t, const void **cur, const void *ptr, size_t bytes, size_t *cur_bytes)
{
  if (*cur == ptr && *cur_bytes == bytes) return 0;
  if (*cur) cudaDestroyTextureObject(*t);
  *cur = nullptr; *t = 0;
  cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
  rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = const_cast<void *>(ptr);
  rd.res.linear.desc = cudaCreateChannelDesc<int4>(); rd.res.linear.sizeInBytes = bytes;
  cudaTextureDesc td; memset(&td, 0, sizeof(td));
  td.readMode = cudaReadModeElementType;
  CUDA_TRY(cudaCreateTextureObject(t, &rd, &td, nullptr));
  *cur = ptr; *cur_bytes = bytes;
  return 0;
}

int forces_textures(imdb200_sim *s)
{
  static size_t max_texels = 0;
  if (!max_texels) { cudaDeviceProp p; CUDA_TRY(cudaGetDeviceProperties(&p, s->cfg.device)); max_texels = (size_t) p.maxTexture1DLinear; }
  const size_t bytes = (size_t) s->cap_atoms * sizeof(double4);
  s->tex_ok = bytes / 16 <= max_texels;          // beyond the 1-D linear texture limit every gather stays on the LSU path
  if (!s->tex_ok) return 0;
  TRY(tex_make(&s->tex_pos, &s->tex_pos_ptr, s->pos, bytes, &s->tex_pos_bytes));
  TRY(tex_make(&s->tex_posdf, &s->tex_posdf_ptr, s->posdf, bytes, &s->tex_posdf_bytes));
  return 0;
}

void forces_free_textures(imdb200_sim *s)
{
  if (s->tex_pos_ptr) cudaDestroyTextureObject(s->tex_pos);
  if (s->tex_posdf_ptr) cudaDestroyTextureObject(s->tex_posdf);
  s->tex_pos_ptr = s->tex_posdf_ptr = nullptr;
}
#endif

static int g_num_sms = 0;
// CTAs of a launch over nw warps: persistent, at most one per SM
static int part_blocks(imdb200_sim *s, long nw, int nt)
{
  if (!g_num_sms) { cudaDeviceProp p; cudaGetDeviceProperties(&p, s->cfg.device); g_num_sms = p.multiProcessorCount; }
  const long need = (nw * 32 + nt - 1) / nt;
  return (int) (need < g_num_sms ? (need > 0 ? need : 1) : g_num_sms);
}

static FArgs make_args(imdb200_sim *s, int nt)
{
  FArgs a;
  a.tpos = s->tex_pos; a.tposdf = s->tex_posdf; a.use_tex = s->tex_ok;
  a.pos = s->pos; a.posdf = s->posdf; a.frc = s->frc; a.rho = s->rho; a.dF = s->dF; a.eam_p = s->eam_p; a.dM = s->dM; a.nbl = s->nbl; a.nnbc = s->nnbc;
  a.cls_fixed = skin_class_fixed(s); a.cls_w = s->cfg.nbl_margin / NBL_CLASSES; a.ctl = s->d_ctl;
  const long nw = (s->n_own * s->lanes + 31) / 32;
  a.worder = nullptr; a.w0 = 0; a.wn = nw;
  if (s->split_part == 1) { a.worder = s->worder; a.w0 = 0; a.wn = s->n_bwarps; }
  else if (s->split_part == 2) { a.worder = s->worder; a.w0 = s->n_bwarps; a.wn = nw - s->n_bwarps; }
  // the per-block partial sums of the interior launch sit behind those of the boundary launch
  a.pblock0 = s->split_part == 2 ? part_blocks(s, s->n_bwarps, nt) : 0;
  a.n_own = s->n_own; a.rows = s->max_nb / s->lanes;
  a.presstens = s->presstens; a.pstride = s->cap_atoms;
  a.partial = s->d_partial; a.flags = s->d_flags;
  a.pos_rw = s->pos; a.mom = s->mom; a.nblpos = s->nblpos; a.nstride = s->cap_atoms; a.scal = s->d_scal;
  a.maxd2 = (unsigned long long *) (s->d_scal + SC_MAXD2);
  a.dt = s->cfg.timestep; a.nvt = s->cfg.ensemble == IMDB200_ENS_NVT;
  a.glob = s->d_glob; a.nactive = (double) s->nactive; a.temperature = s->cfg.temperature; a.tauber = s->tauber;
  return a;
}

static int grid_for(imdb200_sim *s, int nt)
{
  const long nw = (s->n_own * s->lanes + 31) / 32;
  if (s->split_part == 1) return part_blocks(s, s->n_bwarps, nt);
  if (s->split_part == 2) return part_blocks(s, nw - s->n_bwarps, nt);
  return part_blocks(s, nw, nt);
}
// blocks whose partial sums the reduction behind a pass has to add up (both parts of a split pass)
static int reduce_blocks(imdb200_sim *s, int nt)
{
  const long nw = (s->n_own * s->lanes + 31) / 32;
  if (s->split_part == 2) return part_blocks(s, s->n_bwarps, nt) + part_blocks(s, nw - s->n_bwarps, nt);
  return grid_for(s, nt);
}

template <typename K> static int launch_k(K kern, imdb200_sim *s, const FArgs &a, int nt, int smem)
{
  if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<grid_for(s, nt), nt, smem, s->stream>>>(a, s->tabs);
  LAUNCH_CHECK();
  return 0;
}

// the per-atom stress variant needs twelve more accumulator registers: half the threads, twice the registers
#define P1(L, EAM, MULTI, SHARED) \
  (s->press_calc ? (ts ? launch_k(k_pass1<512, L, EAM, MULTI, SHARED, true, true, CUBIC, EEAMC>, s, a, 512, sm) \
                       : launch_k(k_pass1<512, L, EAM, MULTI, SHARED, true, false, CUBIC, EEAMC>, s, a, 512, 0)) \
                 : (ts ? launch_k(k_pass1<IMDB_NT, L, EAM, MULTI, SHARED, false, true, CUBIC, EEAMC>, s, a, IMDB_NT, sm) \
                       : launch_k(k_pass1<IMDB_NT, L, EAM, MULTI, SHARED, false, false, CUBIC, EEAMC>, s, a, IMDB_NT, 0)))

template <int L> static int launch1_L(imdb200_sim *s, const FArgs &a)
{
  const bool multi = s->tabs.ntypes > 1, ts = s->tabs.smem1 > 0;
  // (the raw single-species layout is what the branch-free kernel instance stages: no stress, no EEAM terms)
  const bool raw1 = IMDB_RAW1 && IMDB_BRANCHFREE && !multi && s->tabs.smem1_raw > 0 && s->tabs.fused && !s->press_calc && !EEAMC && !CUBIC;
  const int sm = raw1 ? s->tabs.smem1_raw : s->tabs.smem1;
#if IMDB_EEAM
  if (!s->tabs.have_eam) return imdb_fail(IMDB200_ERR_ARG, "EEAM needs EAM tables");
#else
  if (!s->tabs.have_eam) return multi ? P1(L, false, true, false) : P1(L, false, false, false);
#endif
  if (s->tabs.shared_grid) return multi ? P1(L, true, true, true) : P1(L, true, false, true);
  return multi ? P1(L, true, true, false) : P1(L, true, false, false);
}

#define P2(L, MULTI, FUSE) \
  (s->press_calc ? (ts ? launch_k(k_pass2<512, L, MULTI, true, true, false, CUBIC, EEAMC>, s, a, 512, sm) \
                       : launch_k(k_pass2<512, L, MULTI, true, false, false, CUBIC, EEAMC>, s, a, 512, 0)) \
                 : (ts ? launch_k(k_pass2<IMDB_NT2, L, MULTI, false, true, FUSE, CUBIC, EEAMC>, s, a, IMDB_NT2, sm) \
                       : launch_k(k_pass2<IMDB_NT2, L, MULTI, false, false, FUSE, CUBIC, EEAMC>, s, a, IMDB_NT2, 0)))

template <int L> static int launch2_L(imdb200_sim *s, const FArgs &a, int fuse)
{
  const bool multi = s->tabs.ntypes > 1;
  // single-species EEAM reads rhoAB/rhoC; several species stage the distinct (h1,h2) columns (quadratic, no EEAM, one rho grid)
  const bool ts = multi ? (s->tabs.smem2m > 0 && !EEAMC && !CUBIC) : (s->tabs.smem2 > 0 && !EEAMC);
  const int sm = multi ? s->tabs.smem2m : s->tabs.smem2;
  if (multi) return fuse ? P2(L, true, true) : P2(L, true, false);
  return fuse ? P2(L, false, true) : P2(L, false, false);
}

#if IMDB_BASE_TU
// move_atoms can ride in the tail of pass 2 -- pass 2 gathers posdf, never pos -- when neither the per-atom stress nor
// restriction vectors are in play
int forces_can_fuse_move(const imdb200_sim *s)
{ return s->tabs.have_eam && !s->press_calc && s->n_restr == 0 &&
         s->cfg.ensemble != IMDB200_ENS_NPT_ISO && !s->tabs.have_adp; }

// the passes can run boundary part first / interior part second when there is a boundary-first order and the
// integrator rides in pass 2 (positions of the boundary atoms are final after its boundary part)
int forces_split_possible(const imdb200_sim *s)
{ return s->nranks > 1 && s->worder && s->n_bwarps > 0 && s->n_bwarps < s->n_warps && forces_can_fuse_move(s) && !s->tabs.have_eeam; }

// forces.cu is compiled four times: quadratic / cubic table interpolation, each without / with the EEAM terms
int forces_pass1(imdb200_sim *s)
{
  if (s->tabs.have_eeam) return s->tabs.cubic ? forces_pass1_cubic_eeam(s) : forces_pass1_quad_eeam(s);
  return s->tabs.cubic ? forces_pass1_cubic(s) : forces_pass1_quad(s);
}
int forces_pass2(imdb200_sim *s, int fuse)
{
  if (s->tabs.have_eeam) return s->tabs.cubic ? forces_pass2_cubic_eeam(s, fuse) : forces_pass2_quad_eeam(s, fuse);
  return s->tabs.cubic ? forces_pass2_cubic(s, fuse) : forces_pass2_quad(s, fuse);
}
#endif

int IMPL(forces_pass1)(imdb200_sim *s)
{
  TRY(forces_textures(s));
  FArgs a = make_args(s, s->press_calc ? 512 : IMDB_NT);
  switch (s->lanes) {
    case 1: TRY(launch1_L<1>(s, a)); break;
    case 2: TRY(launch1_L<2>(s, a)); break;
    case 4: TRY(launch1_L<4>(s, a)); break;
    case 8: TRY(launch1_L<8>(s, a)); break;
    case 16: TRY(launch1_L<16>(s, a)); break;
    case 32: TRY(launch1_L<32>(s, a)); break;
    default: return imdb_fail(IMDB200_ERR_ARG, "lanes_per_atom must be a power of two <= 32");
  }
  if (s->split_part == 1) return 0;                      // the interior launch follows; one reduction over both
  const int slots[2] = {SC_EPOT, SC_VIRIAL};
  // fuse_step (imdb200_run): SC_MAXD2 is cleared here, in stream order before pass 2, whose fused integrator
  // accumulates the new maximum (a kernel, not a memset, so that a gated step leaves it alone)
  return reduce_finish(s, reduce_blocks(s, s->press_calc ? 512 : IMDB_NT), 2, slots, 0, s->fuse_step);
}

int IMPL(forces_pass2)(imdb200_sim *s, int fuse)
{
  TRY(forces_textures(s));
  FArgs a = make_args(s, s->press_calc ? 512 : IMDB_NT2);
  if (fuse && !forces_can_fuse_move(s)) return imdb_fail(IMDB200_ERR_ARG, "fused move_atoms is not available in this configuration");
  switch (s->lanes) {
    case 1: TRY(launch2_L<1>(s, a, fuse)); break;
    case 2: TRY(launch2_L<2>(s, a, fuse)); break;
    case 4: TRY(launch2_L<4>(s, a, fuse)); break;
    case 8: TRY(launch2_L<8>(s, a, fuse)); break;
    case 16: TRY(launch2_L<16>(s, a, fuse)); break;
    case 32: TRY(launch2_L<32>(s, a, fuse)); break;
    default: return imdb_fail(IMDB200_ERR_ARG, "lanes_per_atom must be a power of two <= 32");
  }
  if (s->split_part == 1) return 0;
  const int nb = reduce_blocks(s, s->press_calc ? 512 : IMDB_NT2);
  if (fuse) {
    const bool nvt = s->cfg.ensemble == IMDB200_ENS_NVT;
    const int slots[3] = {SC_VIRIAL, nvt ? SC_EKIN1 : SC_EKIN, SC_EKIN2};
    return reduce_finish(s, nb, 3, slots, 1);            // the virial adds to pass 1's, the kinetic sums replace
  }
  const int slots[1] = {SC_VIRIAL};
  s->maxd2_zeroed = s->zero_before_move;
  return reduce_finish(s, nb, 1, slots, 1, s->zero_before_move);
}
