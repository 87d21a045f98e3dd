// cells.cu -- box/cell geometry, cell binning + sort (fix_cells), ghost images (send_cells) and the
// Verlet neighbour-list build (make_nblist).
//
// Reference behaviour restated here (not its data structures):
//   make_box / init_cells      src/imd_geom_3d.c:52-104, 113-248
//   cell_coord                 src/imd_geom_3d.c:1054-1074
//   do_boundaries              src/imd_main_3d.c:1972-2059
//   fix_cells                  src/imd_fix_cells_3d.c:36-201
//   send_cells + copy_cell     src/imd_comm_force_3d.c:222-396, 726-778
//   make_nblist                src/imd_forces_nbl.c:136-273
#include "internal.cuh"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <chrono>

// IMDB200_DEBUG_REBUILD=1: wall-clock time of the sections of a list build on stderr (host waits included)
static bool g_dbg_rebuild = getenv("IMDB200_DEBUG_REBUILD") != nullptr;
struct RebuildClock {
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void lap(imdb200_sim *s, const char *what) {
    if (!g_dbg_rebuild) return;
    cudaStreamSynchronize(s->stream);
    const auto n = std::chrono::steady_clock::now();
    fprintf(stderr, "[rebuild rank %d] %-28s %8.3f ms\n", s->rank, what, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};

// =====================================================================================================
// host: box and cell grid
// =====================================================================================================
static void cross(const double u[3], const double v[3], double w[3])
{
  w[0] = u[1] * v[2] - u[2] * v[1];
  w[1] = u[2] * v[0] - u[0] * v[2];
  w[2] = u[0] * v[1] - u[1] * v[0];
}
static double dot(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// init_cells (src/imd_geom_3d.c:113-248): cell grid from cellsz and the box heights.
static int init_cells(imdb200_sim *s)
{
  Geom &g = s->geom;
  if (!s->have_tabs) return imdb_fail(IMDB200_ERR_ARG, "set the potentials before the atoms (cellsz unknown)");
  if (g.cellsz == 0.0) { // margin is added once (:122-126)
    double r = sqrt(s->cellsz0) + s->cfg.nbl_margin;
    g.cellsz = r * r;
  }
  for (int d = 0; d < 3; d++) {
    double cell_scale = sqrt(1.0 * g.cellsz / s->height[d]);
    g.gdim[d] = (int) (1.0 / cell_scale);
    int cd = s->cfg.cpu_dim[d];
    if (g.gdim[d] % cd) g.gdim[d] = (g.gdim[d] / cd) * cd;
    if (g.gdim[d] < cd) return imdb_fail(IMDB200_ERR_CELLS, "global_cell_dim too small, need at least %d", cd);
    s->min_height[d] = g.cellsz * (double) g.gdim[d] * g.gdim[d];
    s->max_height[d] = g.cellsz * (double) (g.gdim[d] + cd) * (g.gdim[d] + cd);
    g.cdim[d] = g.gdim[d] / cd + 2;
    g.coff[d] = s->cfg.my_coord[d] * (g.cdim[d] - 2);
  }
  g.nall = g.cdim[0] * g.cdim[1] * g.cdim[2];
  if (s->cell_count) { cudaFree(s->cell_count); cudaFree(s->cell_start); cudaFree(s->cell_fill); cudaFree(s->cell_code); cudaFree(s->scan_tmp); }
  CUDA_TRY(cudaMalloc(&s->cell_code, (g.nall + NBIN_EXTRA) * sizeof(int)));
  CUDA_TRY(cudaMalloc(&s->cell_count, (g.nall + NBIN_EXTRA) * sizeof(int)));
  CUDA_TRY(cudaMalloc(&s->cell_start, (g.nall + NBIN_EXTRA) * sizeof(int)));
  CUDA_TRY(cudaMalloc(&s->cell_fill, (g.nall + NBIN_EXTRA) * sizeof(int)));
  CUDA_TRY(cudaMalloc(&s->scan_tmp, (g.nall / 1024 + 1024) * sizeof(int)));
  TRY(comm_plan(s));
  s->have_valid_nbl = 0;
  return 0;
}

// make_box (src/imd_geom_3d.c:52-104)
int geom_make_box(imdb200_sim *s)
{
  Geom &g = s->geom;
  cross(g.box[1], g.box[2], g.tbox[0]);
  cross(g.box[2], g.box[0], g.tbox[1]);
  cross(g.box[0], g.box[1], g.tbox[2]);
  s->volume = dot(g.box[0], g.tbox[0]);
  if (s->volume == 0.0) return imdb_fail(IMDB200_ERR_ARG, "Box Edges are parallel.");
  for (int i = 0; i < 3; i++) for (int d = 0; d < 3; d++) g.tbox[i][d] /= s->volume;
  int redo = 0;
  for (int d = 0; d < 3; d++) {
    s->height[d] = 1.0 / dot(g.tbox[d], g.tbox[d]);
    if (s->height[d] < s->min_height[d] || s->height[d] > s->max_height[d]) redo = 1;
  }
  if (redo) TRY(init_cells(s));
  if (s->volume < 0) s->volume = -s->volume;
  if (s->volume_init == 0.0) s->volume_init = s->volume;
  else if (s->volume > 8 * s->volume_init) return imdb_fail(IMDB200_ERR_EXPLODE, "system seems to explode!");
  return 0;
}

// =====================================================================================================
// per-atom array capacity
// =====================================================================================================
// exported: peers may have this array mapped over CUDA IPC (comm_p2p.cu) -- the old allocation is parked, not freed
template <typename T> static int regrow(T **p, long old_n, long new_cap, imdb200_sim *exported = nullptr)
{
  T *q = nullptr;
  CUDA_TRY(cudaMalloc(&q, new_cap * sizeof(T)));
  if (*p && old_n > 0) CUDA_TRY(cudaMemcpy(q, *p, old_n * sizeof(T), cudaMemcpyDeviceToDevice));
  if (*p && !(exported && comm_p2p_defer_free(exported, *p))) cudaFree(*p);
  *p = q;
  return 0;
}

int cells_ensure_capacity(imdb200_sim *s, long need)
{
  if (need <= s->cap_atoms) return 0;
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  long cap = need + need / 8 + 1024;
  long n = s->n_own;
  TRY(regrow(&s->pos, n, cap)); TRY(regrow(&s->pos_alt, 0, cap));
  TRY(regrow(&s->mom, n, cap)); TRY(regrow(&s->mom_alt, 0, cap));
  TRY(regrow(&s->frc, n, cap)); TRY(regrow(&s->posdf, 0, cap));
  TRY(regrow(&s->nummer, n, cap)); TRY(regrow(&s->nummer_alt, 0, cap));
  TRY(regrow(&s->rho, n, cap)); TRY(regrow(&s->dF, n, cap, s));
  TRY(regrow(&s->eam_p, 0, cap)); TRY(regrow(&s->dM, 0, cap));
  TRY(regrow(&s->cellid, n, cap)); TRY(regrow(&s->cellid_alt, 0, cap)); TRY(regrow(&s->perm, 0, cap));
  TRY(regrow(&s->gsrc, 0, cap)); TRY(regrow(&s->ghost_num, 0, cap)); TRY(regrow(&s->ghost_raw, 0, cap, s));
  TRY(regrow(&s->posf, 0, cap));
  // SoA blocks whose stride is the capacity: contents are rebuilt before use
  if (s->nblpos) cudaFree(s->nblpos);
  if (s->presstens) cudaFree(s->presstens);
  s->nblpos = nullptr; s->presstens = nullptr;
  CUDA_TRY(cudaMalloc(&s->nblpos, 3 * cap * sizeof(double)));
  CUDA_TRY(cudaMalloc(&s->presstens, 6 * cap * sizeof(double)));
  CUDA_TRY(cudaMemset(s->presstens, 0, 6 * cap * sizeof(double)));
  s->cap_atoms = cap;
  s->have_valid_nbl = 0;
  return adp_ensure_arrays(s);             // ADP: mu / lambda follow the capacity
}

// =====================================================================================================
// exclusive scan (int32), three small kernels
// =====================================================================================================
__global__ void k_scan_blocks(const int *in, int *out, int n, int *sums)
{
  __shared__ int sm[1024];
  const int base = blockIdx.x * 1024, t = threadIdx.x;
  int v = (base + t < n) ? in[base + t] : 0;
  sm[t] = v;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    int x = (t >= o) ? sm[t - o] : 0;
    __syncthreads();
    sm[t] += x;
    __syncthreads();
  }
  if (base + t < n) out[base + t] = sm[t] - v;
  if (t == 1023) sums[blockIdx.x] = sm[t];
}
__global__ void k_scan_sums(int *sums, int nb, int *total)
{
  __shared__ int sm[1024];
  __shared__ int carry;
  const int t = threadIdx.x;
  if (t == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    int v = (base + t < nb) ? sums[base + t] : 0;
    sm[t] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int x = (t >= o) ? sm[t - o] : 0;
      __syncthreads();
      sm[t] += x;
      __syncthreads();
    }
    if (base + t < nb) sums[base + t] = carry + sm[t] - v;
    __syncthreads();
    if (t == 1023) carry += sm[t];
    __syncthreads();
  }
  if (t == 0 && total) *total = carry;
}
__global__ void k_scan_add(int *out, int n, const int *sums)
{
  int i = blockIdx.x * 1024 + threadIdx.x;
  if (i < n) out[i] += sums[blockIdx.x];
}

int scan_exclusive(imdb200_sim *s, const int *in, int *out, int n, int *total_dev)
{
  if (n <= 0) { if (total_dev) CUDA_TRY(cudaMemsetAsync(total_dev, 0, sizeof(int), s->stream)); return 0; }
  int nb = cdiv(n, 1024);
  k_scan_blocks<<<nb, 1024, 0, s->stream>>>(in, out, n, s->scan_tmp); LAUNCH_CHECK();
  k_scan_sums<<<1, 1024, 0, s->stream>>>(s->scan_tmp, nb, total_dev); LAUNCH_CHECK();
  k_scan_add<<<nb, 1024, 0, s->stream>>>(out, n, s->scan_tmp); LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// fix_cells: wrap, bin, sort
// =====================================================================================================
__device__ __forceinline__ double sprod_exact(double x, double y, double z, const double t[3])
{ return __dadd_rn(__dadd_rn(__dmul_rn(x, t[0]), __dmul_rn(y, t[1])), __dmul_rn(z, t[2])); }

// do_boundaries (src/imd_main_3d.c:1972-2059) + cell_coord (src/imd_geom_3d.c:1054-1074) +
// local_cell_coord (src/imd_geom_mpi_3d.c:119-128), same operation order, no FMA.
// An atom whose cell belongs to another rank goes into one of the extra bins behind the cell grid:
// nall + d for "leaves in direction d" (fix_cells, src/imd_fix_cells_3d.c:100-175: it is sent to that
// neighbour), or nall + 27 when set_atoms handed us atoms of other domains (they are dropped).
struct BinArgs { int cpu_dim[3], my_coord[3]; int filter; };
__global__ void k_wrap_bin(double4 *pos, long n, Geom g, BinArgs b, int *cellid, int *cell_count, int *flags)
{
  long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 p = pos[i];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    if (g.pbc[d] == 1) {
      double f = -floor(sprod_exact(p.x, p.y, p.z, g.tbox[d]));
      p.x = __dadd_rn(p.x, __dmul_rn(f, g.box[d][0]));
      p.y = __dadd_rn(p.y, __dmul_rn(f, g.box[d][1]));
      p.z = __dadd_rn(p.z, __dmul_rn(f, g.box[d][2]));
    }
  }
  pos[i] = p;
  int c[3], dir = 13;
  bool away = false, lost = false;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    int v = __double2int_rz(__dmul_rn((double) g.gdim[d], sprod_exact(p.x, p.y, p.z, g.tbox[d])));
    if (v >= g.gdim[d]) v = g.gdim[d] - 1; else if (v < 0) v = 0;
    const int per = g.cdim[d] - 2, owner = v / per, me = b.my_coord[d], np = b.cpu_dim[d];
    if (owner != me) {
      away = true;
      const int stride = d == 0 ? 1 : (d == 1 ? 3 : 9);
      // the side follows from the coordinates; the periodic wrap only joins the two ends where the axis is
      // periodic (with cpu_dim == 2 on a free axis "me+1 mod 2" would also match the neighbour BELOW)
      const bool wrap = g.pbc[d] == 1 && np > 2;
      if (owner == me + 1 || (wrap && me == np - 1 && owner == 0)) dir += stride;
      else if (owner == me - 1 || (wrap && me == 0 && owner == np - 1)) dir -= stride;
      else lost = true;                                     // "Atom jumped multiple CPUs" (:170)
      v = 1;
    } else v = v - g.coff[d] + 1;
    c[d] = v;
  }
  int ci = (c[0] * g.cdim[1] + c[1]) * g.cdim[2] + c[2];
  if (away) {
    if (b.filter) ci = g.nall + 27;
    else { ci = g.nall + dir; if (lost) atomicAdd(&flags[FL_LOST], 1); }
  }
  cellid[i] = ci;
  atomicAdd(&cell_count[ci], 1);
}

__global__ void k_scatter(const int *cellid, long n, const int *cell_start, int *cell_fill, int *perm)
{
  long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = cellid[i];
  perm[cell_start[c] + atomicAdd(&cell_fill[c], 1)] = (int) i;
}

// Canonical order inside a cell: ascending atom number.  (The reference's order is history
// dependent, src/imd_alloc.c:106-123; results do not depend on it -- SURVEY.md section 9 item 1 -- but a
// canonical order makes every run bit-reproducible.)
__global__ void k_sort_cells(const int *cell_start, const int *cell_count, int *perm, const int *nummer, int nall)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nall) return;
  const int s0 = cell_start[c], n = cell_count[c];
  for (int a = 1; a < n; a++) {
    int pa = perm[s0 + a], ka = nummer[pa], b = a - 1;
    while (b >= 0 && nummer[perm[s0 + b]] > ka) { perm[s0 + b + 1] = perm[s0 + b]; b--; }
    perm[s0 + b + 1] = pa;
  }
}

__global__ void k_gather(const int *perm, long n, const double4 *pos, const double4 *mom, const int *nummer,
                         const int *cellid, double4 *pos2, double4 *mom2, int *nummer2, int *cellid2)
{
  long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (i >= n) return;
  int p = perm[i];
  pos2[i] = pos[p];
  mom2[i] = mom[p];
  nummer2[i] = nummer[p];
  cellid2[i] = cellid[p];
}

// =====================================================================================================
// make_nblist
// =====================================================================================================
// One thread per owned atom scans the 27 cells around its own.  The list is FULL (both directions
// of every pair) so that the force kernels never scatter; its symmetric closure is exactly the
// reference's half list (src/imd_forces_nbl.c:218-269), because every pair is tested with the
// reference's own operands and rounding:
//   * the reference tests a pair once, from the atom whose cell comes first in the half stencil
//     l=0..1, m=-l..1, n=(l==0?-m:-l)..1 (src/imd_geom_3d.c:895-899), as d = ort_j - ort_i where
//     ort_j is the buffer-cell image (x_j + shift, added stage-wise) if the neighbour cell is a
//     buffer cell;
//   * for an "upper" neighbour cell we are that first atom: d = pos[j] - x_i with pos[j] the image;
//   * for a "lower" neighbour cell the reference started from j and saw OUR image in the buffer
//     cell on the opposite side: d = image(x_i, -shift) - x_j(unshifted).  When the lower cell is a
//     real cell both are the same up to sign; when it is a buffer cell we rebuild exactly that.
__global__ void __launch_bounds__(128)
k_build_nbl(const double4 *__restrict__ pos, long n_own, Geom g, const int *__restrict__ cellid,
            const int *__restrict__ cell_start, const int *__restrict__ cell_count,
            const int *__restrict__ cell_code, const int *__restrict__ gsrc, const double4 *__restrict__ ghost_raw,
            int *__restrict__ nbl, int *__restrict__ nnb, int max_nb, int L,
            int *flags, int count_only)
{
  long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (i >= n_own) return;
  const double4 xi = pos[i];
  const int c1 = cellid[i];
  const int cz = c1 % g.cdim[2], cy = (c1 / g.cdim[2]) % g.cdim[1], cx = c1 / (g.cdim[2] * g.cdim[1]);
  int cnt = 0;
  for (int l = -1; l <= 1; l++)
    for (int m = -1; m <= 1; m++)
      for (int n = -1; n <= 1; n++) {
        const int c2 = ((cx + l) * g.cdim[1] + (cy + m)) * g.cdim[2] + (cz + n);
        const int nj = cell_count[c2];
        if (nj == 0) continue;
        const int j0 = cell_start[c2];
        const bool upper = (l > 0) || (l == 0 && (m > 0 || (m == 0 && n >= 0)));
        const bool ghost = j0 >= n_own;
        if (ghost && !upper) {
          // our image as the reference's buffer cell on the far side holds it
          const int code = cell_code[c2];               // shift of the images in c2
          const int inv = 26 - code;                    // (-sx,-sy,-sz)
          const double4 me = image_pos(xi, inv, g);
          for (int t = 0; t < nj; t++) {
            const int j = j0 + t;
            const int src = gsrc[j - n_own];              // the owner's unshifted copy: ours, or as received
            const double4 xj = src >= 0 ? pos[src] : ghost_raw[j - n_own];
            const double r2 = r2_exact(__dsub_rn(me.x, xj.x), __dsub_rn(me.y, xj.y), __dsub_rn(me.z, xj.z));
            if (r2 < g.cellsz) {
              if (!count_only && cnt < max_nb) nbl[nbl_index(i, cnt, L, max_nb / L)] = j;
              cnt++;
            }
          }
        } else {
          for (int t = 0; t < nj; t++) {
            const int j = j0 + t;
            if (j == i) continue;
            const double4 xj = pos[j];
            const double r2 = r2_exact(__dsub_rn(xj.x, xi.x), __dsub_rn(xj.y, xi.y), __dsub_rn(xj.z, xi.z));
            if (r2 < g.cellsz) {
              if (!count_only && cnt < max_nb) nbl[nbl_index(i, cnt, L, max_nb / L)] = j;
              cnt++;
            }
          }
        }
      }
  if (!count_only) nnb[i] = cnt < max_nb ? cnt : max_nb;
  atomicMax(&flags[FL_MAXNB], cnt);
  if (!count_only && cnt > max_nb) atomicExch(&flags[FL_NBL_OVERFLOW], 1);
}


// ---- the production build: single-precision pre-filter, exact test on the survivors, entries grouped by skin class ----
// Phase 1 walks the 27 cells with float positions, two candidates per packed FP32 instruction, against a
// cut-off widened by the float rounding bound, and records the ~16 % survivors as one bit per candidate in
// shared memory (W 32-bit words per cell and thread: 108 bytes per thread for cells of up to 32 atoms, so a
// full complement of warps fits beside a large L1).  Phase 2 walks the set bits and repeats the reference's
// exact FP64 test (operands and rounding as in k_build_nbl above): pairs inside the largest cut-off go
// straight to the list, pairs outside r_list lose their bit, skin pairs are counted per class.  Phase 3 walks
// the remaining (skin) bits again and writes them class by class behind the core entries.  The neighbour SET
// is that of the reference: the float test can only let extra candidates through, never reject a pair the
// exact test accepts.
struct ClassT { double t2[NBL_CLASSES]; };       // (rc + q*w)^2, q = 0..NBL_CLASSES-1

// Single-precision positions in packed pairs: record k holds atoms 2k and 2k+1 as (x0,x1) (y0,y1) (z0,z1) (pad),
// 32 bytes = one 256-bit load, already in the operand layout of the packed FP32 instructions (FADD2/FFMA2).
__global__ void k_make_posf(const double4 *pos, long n, float4 *posf)
{
  long k = blockIdx.x * (long) blockDim.x + threadIdx.x;      // pair index
  if (2 * k >= n) return;
  const double4 a = pos[2 * k];
  const double4 b = 2 * k + 1 < n ? pos[2 * k + 1] : a;
  posf[2 * k] = make_float4((float) a.x, (float) b.x, (float) a.y, (float) b.y);
  posf[2 * k + 1] = make_float4((float) a.z, (float) b.z, 0.f, 0.f);
}

struct PairRec { float2 x, y, z, w; };
__device__ __forceinline__ PairRec ld_pair(const float4 *posf, int k)
{
  PairRec r;
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(r.x.x), "=f"(r.x.y), "=f"(r.y.x), "=f"(r.y.y), "=f"(r.z.x), "=f"(r.z.y), "=f"(r.w.x), "=f"(r.w.y)
      : "l"(posf + 2 * (size_t) k));
  return r;
}

// pass bits of the n <= 32 candidates js, js+1, ...; xf, yf, zf hold MINUS the atom's own coordinates twice
__device__ __forceinline__ unsigned nbl_scan_word(const float4 *__restrict__ posf, int js, int n, float2 xf, float2 yf,
                                                  float2 zf, float cutf)
{
  unsigned word = 0u;
  int j = js;
  const int jend = js + n;
  if (j & 1) {                                                  // odd first atom: upper half of its pair
    const PairRec p = ld_pair(posf, j >> 1);
    const float dx = p.x.y + xf.x, dy = p.y.y + yf.x, dz = p.z.y + zf.x;
    word = fmaf(dz, dz, fmaf(dy, dy, dx * dx)) < cutf ? 1u : 0u;
    j++;
  }
  int np = (jend - j) >> 1;                                     // whole pairs
  int k = j >> 1, sh = j - js;
#define NBL_PAIR(P, T) do { \
    const float2 dx_ = __fadd2_rn((P).x, xf), dy_ = __fadd2_rn((P).y, yf), dz_ = __fadd2_rn((P).z, zf); \
    const float2 r2_ = __ffma2_rn(dz_, dz_, __ffma2_rn(dy_, dy_, __fmul2_rn(dx_, dx_))); \
    if (r2_.x < cutf) m8 |= 1u << (T); if (r2_.y < cutf) m8 |= 2u << (T); } while (0)
  for (; np >= 4; np -= 4, k += 4, sh += 8) {
    const PairRec p0 = ld_pair(posf, k), p1 = ld_pair(posf, k + 1), p2 = ld_pair(posf, k + 2), p3 = ld_pair(posf, k + 3);
    unsigned m8 = 0u;
    NBL_PAIR(p0, 0); NBL_PAIR(p1, 2); NBL_PAIR(p2, 4); NBL_PAIR(p3, 6);
    word |= m8 << sh;
  }
  if (np >= 2) {
    const PairRec p0 = ld_pair(posf, k), p1 = ld_pair(posf, k + 1);
    unsigned m8 = 0u;
    NBL_PAIR(p0, 0); NBL_PAIR(p1, 2);
    word |= m8 << sh;
    np -= 2; k += 2; sh += 4;
  }
  if (np) { const PairRec p0 = ld_pair(posf, k); unsigned m8 = 0u; NBL_PAIR(p0, 0); word |= m8 << sh; k++; sh += 2; }
#undef NBL_PAIR
  if (2 * k < jend) {                                           // even last atom: lower half of its pair
    const PairRec p = ld_pair(posf, k);
    const float dx = p.x.x + xf.x, dy = p.y.x + yf.x, dz = p.z.x + zf.x;
    if (fmaf(dz, dz, fmaf(dy, dy, dx * dx)) < cutf) word |= 1u << sh;
  }
  return word;
}

struct BuildCtx {                 // what the exact test of one candidate needs
  const double4 *pos, *ghost_raw; const int *cellid, *cell_code, *gsrc;
  double4 xi; int i, n_own;
};
// the reference's own test operands (see k_build_nbl): mirror = the pair is evaluated from the other atom's side
__device__ __forceinline__ double nbl_exact_r2(const BuildCtx &b, const Geom &g, int j, bool mirror, int &tj)
{
  if (mirror) {
    const int inv = 26 - b.cell_code[b.cellid[j]];             // our image as the buffer cell on the far side holds it
    const double4 me = image_pos(b.xi, inv, g);
    const int src = b.gsrc[j - b.n_own];
    const double4 xj = src >= 0 ? b.pos[src] : b.ghost_raw[j - b.n_own];
    tj = sorte_of(xj.w);
    return r2_exact(__dsub_rn(me.x, xj.x), __dsub_rn(me.y, xj.y), __dsub_rn(me.z, xj.z));
  }
  const double4 xj = ld_atom(b.pos + j);
  tj = sorte_of(xj.w);
  return r2_exact(__dsub_rn(xj.x, b.xi.x), __dsub_rn(xj.y, b.xi.y), __dsub_rn(xj.z, b.xi.z));
}

template <int BS, int WT>                                     // WT: words per cell as a compile-time constant (0: run time)
__global__ void __launch_bounds__(BS)
k_build_nbl2(const double4 *__restrict__ pos, const float4 *__restrict__ posf, int n_own, Geom g,
             const int *__restrict__ cellid, const int *__restrict__ cell_start, const int *__restrict__ cell_count,
             const int *__restrict__ cell_code, const int *__restrict__ gsrc, const double4 *__restrict__ ghost_raw,
             int *__restrict__ nbl, int *__restrict__ nnb, unsigned long long *__restrict__ nnbc, int max_nb, int L,
             int Wrt, float cutf, ClassT T, int *flags)
{
  extern __shared__ unsigned bits[];                          // [27*W][BS]
  const int W = WT > 0 ? WT : Wrt;                            // cells of <= 32 / <= 64 atoms: no run-time divisions
  const int i = blockIdx.x * BS + threadIdx.x;
  int total = 0;
  if (i < n_own) {
    BuildCtx b;
    b.pos = pos; b.ghost_raw = ghost_raw; b.cellid = cellid; b.cell_code = cell_code; b.gsrc = gsrc;
    b.xi = pos[i]; b.i = i; b.n_own = n_own;
    const float2 xf = make_float2(-(float) b.xi.x, -(float) b.xi.x), yf = make_float2(-(float) b.xi.y, -(float) b.xi.y),
                 zf = make_float2(-(float) b.xi.z, -(float) b.xi.z);
    const int c1 = cellid[i];
    const int cz = c1 % g.cdim[2], cy = (c1 / g.cdim[2]) % g.cdim[1], cx = c1 / (g.cdim[2] * g.cdim[1]);
    unsigned *const my = bits + threadIdx.x;
    // ---- phase 1 ----
    unsigned long long nz = 0ull;                             // words with survivors (WT = 1, 2: at most 54 words), so that the walks
    int cs = 0;                                               // below jump to the next one instead of scanning for it
    for (int l = -1; l <= 1; l++)
      for (int m = -1; m <= 1; m++) {
        const int crow = ((cx + l) * g.cdim[1] + (cy + m)) * g.cdim[2] + cz;
#pragma unroll
        for (int n = -1; n <= 1; n++, cs++) {
          const int c2 = crow + n;
          const int nj = cell_count[c2];
          const int j0 = cell_start[c2];
          if (nj > 32 * W) atomicMax(&flags[FL_CELLFULL], nj);
          for (int w = 0; w < W; w++) {
            const int lo = 32 * w, cntw = min(nj - lo, 32);
            const unsigned wd = cntw > 0 ? nbl_scan_word(posf, j0 + lo, cntw, xf, yf, zf, cutf) : 0u;
            my[(cs * W + w) * BS] = wd;
            if (WT > 0 && wd) nz |= 1ull << (cs * W + w);
          }
        }
      }
    // ---- phases 2 and 3 walk the set bits; j0/mirror of the current cell are refreshed when the word index moves on ----
    const int R = max_nb / L, nwords = 27 * W;
    int *const row0 = nbl + nbl_index(i, 0, L, R);             // L == 1: entry p lives at row0[32 p]
    int n0 = 0;
    unsigned long long counts = 0ull;
#define NBL_WALK(BODY) do { \
      int wi = -1, jb = 0; bool mirror = false; unsigned word = 0u, keep = 0u; \
      unsigned long long todo = nz; nz = 0ull; \
      for (;;) { \
        if (word == 0u) {                       /* this lane moves on to its next non-empty word */ \
          if (wi >= 0) { my[wi * BS] = keep; if (WT > 0 && keep) nz |= 1ull << wi; } \
          if (WT > 0) { if (todo == 0ull) break; wi = __ffsll((long long) todo) - 1; todo &= todo - 1; word = my[wi * BS]; } \
          else { for (++wi; wi < nwords; ++wi) { word = my[wi * BS]; if (word) break; } \
                 if (wi >= nwords) break; } \
          const int cs_ = wi / W, l_ = cs_ / 9 - 1, m_ = (cs_ / 3) % 3 - 1, n_ = cs_ % 3 - 1; \
          const int j0_ = cell_start[((cx + l_) * g.cdim[1] + (cy + m_)) * g.cdim[2] + cz + n_]; \
          const bool upper_ = (l_ > 0) || (l_ == 0 && (m_ > 0 || (m_ == 0 && n_ >= 0))); \
          mirror = j0_ >= n_own && !upper_; jb = j0_ + 32 * (wi % W); keep = word; \
        } \
        const int t = __ffs(word) - 1; word &= word - 1; \
        const int j = jb + t; \
        BODY \
      } } while (0)
    NBL_WALK({
      int tj;
      const double r2 = nbl_exact_r2(b, g, j, mirror, tj);
      const int ent = j | (tj << NBL_TSHIFT);                    // the neighbour's type rides in the entry (NBL_TSHIFT)
      if (!(r2 < g.cellsz) || j == i) keep &= ~(1u << t);
      else if (r2 <= T.t2[0]) {
        keep &= ~(1u << t);
        if (n0 < max_nb) { if (L == 1) row0[n0 * 32] = ent; else nbl[nbl_index(i, n0, L, R)] = ent; }
        n0++;
      } else {
        int c = 1;
#pragma unroll
        for (int k = 1; k < NBL_CLASSES; k++) c += r2 > T.t2[k] ? 1 : 0;
        counts += 1ull << (NBL_CBITS * c);
      }
    });
    unsigned long long cum = (unsigned long long) (n0 < max_nb ? n0 : max_nb), offs = 0ull;
    int run = n0;
#pragma unroll
    for (int c = 1; c <= NBL_CLASSES; c++) {
      offs |= (unsigned long long) (run & ((1 << NBL_CBITS) - 1)) << (NBL_CBITS * c);
      run += (int) ((counts >> (NBL_CBITS * c)) & ((1u << NBL_CBITS) - 1));
      cum |= (unsigned long long) ((run < max_nb ? run : max_nb) & ((1 << NBL_CBITS) - 1)) << (NBL_CBITS * c);
    }
    total = run;
    if (total <= max_nb && total > n0) {
      NBL_WALK({
        int tj;
        const double r2 = nbl_exact_r2(b, g, j, mirror, tj);
        const int ent = j | (tj << NBL_TSHIFT);
        int c = 1;
#pragma unroll
        for (int k = 1; k < NBL_CLASSES; k++) c += r2 > T.t2[k] ? 1 : 0;
        const int sh = NBL_CBITS * c;
        const int p = (int) ((offs >> sh) & ((1u << NBL_CBITS) - 1));
        offs += 1ull << sh;
        if (L == 1) row0[p * 32] = ent; else nbl[nbl_index(i, p, L, R)] = ent;
      });
    }
#undef NBL_WALK
    nnb[i] = total < max_nb ? total : max_nb;
    nnbc[i] = cum;
  }
  const int wmax = __reduce_max_sync(0xffffffffu, total);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&flags[FL_MAXNB], wmax);
    if (wmax > max_nb) atomicExch(&flags[FL_NBL_OVERFLOW], 1);
  }
}

__global__ void k_save_nblpos(const double4 *pos, long n, long stride, double *nblpos)
{
  long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 p = pos[i];
  nblpos[i] = p.x; nblpos[stride + i] = p.y; nblpos[2 * stride + i] = p.z;
}

__global__ void k_sum_int(const int *v, long n, unsigned long long *out)
{
  long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  unsigned long long x = (i < n) ? (unsigned long long) v[i] : 0ull;
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  if ((threadIdx.x & 31) == 0 && x) atomicAdd(out, x);
}

static int alloc_nbl(imdb200_sim *s, int max_nb)
{
  const int L = s->lanes;
  max_nb = ((max_nb + L - 1) / L) * L;
  long n_pad = ((s->cap_atoms + 31) / 32) * 32;
  if (s->nbl && max_nb <= s->max_nb && n_pad == s->n_pad) return 0;
  if (max_nb >= (1 << NBL_CBITS)) return imdb_fail(IMDB200_ERR_NBL, "more than %d neighbours per atom", (1 << NBL_CBITS) - 1);
  if (s->nbl) cudaFree(s->nbl);
  if (s->nnb) cudaFree(s->nnb);
  if (s->nnbc) cudaFree(s->nnbc);
  s->nbl = nullptr; s->nnb = nullptr; s->nnbc = nullptr;
  CUDA_TRY(cudaMalloc(&s->nbl, (size_t) n_pad * max_nb * sizeof(int)));
  // slots behind an atom's last entry are never walked, but imdb200_get_nblist copies whole rows out (initcheck-clean)
  CUDA_TRY(cudaMemsetAsync(s->nbl, 0, (size_t) n_pad * max_nb * sizeof(int), s->stream));
  CUDA_TRY(cudaMalloc(&s->nnb, (size_t) n_pad * sizeof(int)));
  CUDA_TRY(cudaMalloc(&s->nnbc, (size_t) n_pad * sizeof(unsigned long long)));
  s->max_nb = max_nb;
  s->n_pad = n_pad;
  return 0;
}

static int read_flags(imdb200_sim *s)
{
  CUDA_TRY(cudaMemcpyAsync(s->h_flags, s->d_flags, FL_COUNT * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return 0;
}

// Launch of the production list build.  The float cut-off is the exact one plus a bound on what rounding the
// coordinates to single precision can do to r^2: |delta r2| <= 2 r |delta d| with |delta d| <= 4 eps maxc per
// component (two rounded coordinates and the rounded difference), maxc = largest coordinate magnitude.
static int launch_build(imdb200_sim *s, long n)
{
  const Geom &g = s->geom;
  cudaStream_t st = s->stream;
  const long ntot = s->n_own + s->n_ghost;
  if (ntot >= (1L << 29)) return imdb_fail(IMDB200_ERR_ARG, "more than 2^29 atoms and images on one GPU");
  k_make_posf<<<cdiv(ntot / 2 + 1, 256), 256, 0, st>>>(s->pos, ntot, s->posf); LAUNCH_CHECK();
  double maxc = 0.0;
  for (int d = 0; d < 3; d++) {
    double c = 0.0;
    for (int b = 0; b < 3; b++) c += fabs(g.box[b][d]);
    if (c > maxc) maxc = c;
  }
  maxc *= 2.0;                                                 // images reach one cell beyond the box
  const double rl = sqrt(g.cellsz);
  const double slack = 2.0 * rl * 1.7320508 * (4.0 * 1.2e-7 * maxc) + 1e-5 * g.cellsz;
  const float cutf = (float) ((g.cellsz + slack) * (1.0 + 1e-6));
  ClassT T;
  const double rc = sqrt(s->cellsz0), w = s->cfg.nbl_margin / NBL_CLASSES;
  for (int k = 0; k < NBL_CLASSES; k++) T.t2[k] = (rc + k * w) * (rc + k * w);
  T.t2[0] = s->cellsz0 * (1.0 + 1e-12);      // the force kernels form r2 with FMAs: an ulp of slack
  const int max_nb = s->max_nb, L = s->lanes;
  if (s->cell_words < 1) s->cell_words = 1;
  const int W = s->cell_words;
  const size_t per_thread = (size_t) 27 * W * sizeof(unsigned);
#define BUILD(BS, WT) do { \
    const size_t sm = per_thread * BS; \
    if (sm > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(k_build_nbl2<BS, WT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm)); \
    k_build_nbl2<BS, WT><<<cdiv(n > 0 ? n : 1, BS), BS, sm, st>>>(s->pos, s->posf, (int) n, g, s->cellid, s->cell_start, s->cell_count, \
        s->cell_code, s->gsrc, s->ghost_raw, s->nbl, s->nnb, s->nnbc, max_nb, L, W, cutf, T, s->d_flags); \
    LAUNCH_CHECK(); } while (0)
  if (W == 1) BUILD(128, 1);
  else if (W == 2) BUILD(128, 2);          // e.g. B2 Ni-Al: 17 atoms per cell on average, up to 54 where three lattice planes fall into a cell
  else if (per_thread * 128 <= 64 * 1024) BUILD(128, 0);
  else if (per_thread * 32 <= 200 * 1024) BUILD(32, 0);
  else return imdb_fail(IMDB200_ERR_CELLS, "cells of more than %d atoms do not fit the list build", 32 * W);
#undef BUILD
  return 0;
}

// wrap, bin and sort the first n atoms into cell order; atoms of other domains end up behind the
// owned ones, grouped by leave direction.  h_extra receives cell_start[nall] (= atoms that stay) and
// the 28 extra bin counts.
static int bin_and_sort(imdb200_sim *s, long n, int filter, int *h_extra)
{
  const Geom &g = s->geom;
  cudaStream_t st = s->stream;
  const int nbins = g.nall + 28;
  CUDA_TRY(cudaMemsetAsync(s->cell_count, 0, (g.nall + NBIN_EXTRA) * sizeof(int), st));
  CUDA_TRY(cudaMemsetAsync(s->cell_fill, 0, (g.nall + NBIN_EXTRA) * sizeof(int), st));
  BinArgs b;
  for (int d = 0; d < 3; d++) { b.cpu_dim[d] = s->cfg.cpu_dim[d]; b.my_coord[d] = s->cfg.my_coord[d]; }
  b.filter = filter;
  const int nb = cdiv(n, 256);
  k_wrap_bin<<<nb, 256, 0, st>>>(s->pos, n, g, b, s->cellid, s->cell_count, s->d_flags); LAUNCH_CHECK();
  TRY(scan_exclusive(s, s->cell_count, s->cell_start, nbins, nullptr));
  k_scatter<<<nb, 256, 0, st>>>(s->cellid, n, s->cell_start, s->cell_fill, s->perm); LAUNCH_CHECK();
  k_sort_cells<<<cdiv(g.nall, 128), 128, 0, st>>>(s->cell_start, s->cell_count, s->perm, s->nummer, g.nall); LAUNCH_CHECK();
  k_gather<<<nb, 256, 0, st>>>(s->perm, n, s->pos, s->mom, s->nummer, s->cellid, s->pos_alt, s->mom_alt,
                               s->nummer_alt, s->cellid_alt);
  LAUNCH_CHECK();
  { double4 *t = s->pos; s->pos = s->pos_alt; s->pos_alt = t; }
  { double4 *t = s->mom; s->mom = s->mom_alt; s->mom_alt = t; }
  { int *t = s->nummer; s->nummer = s->nummer_alt; s->nummer_alt = t; }
  { int *t = s->cellid; s->cellid = s->cellid_alt; s->cellid_alt = t; }
  CUDA_TRY(cudaMemcpyAsync(h_extra, s->cell_start + g.nall, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(h_extra + 1, s->cell_count + g.nall, 28 * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(s->h_flags, s->d_flags, FL_COUNT * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (s->h_flags[FL_LOST]) return imdb_fail(IMDB200_ERR_CELLS, "Atom jumped multiple CPUs (%d atoms)", s->h_flags[FL_LOST]);
  return 0;
}

// nactive: the degrees of freedom that move -- 3 per atom of a real type, the restriction components of its virtual
// type otherwise (read_atoms, src/imd_io_3d.c:469-481; generate_atoms, src/imd_generate.c:450-451).  It enters the
// Nose-Hoover update (src/imd_integrate.c:1140) and the temperature the writers print.
__global__ void k_count_nactive(const double4 *pos, long n, int ntypes, const double *restr, int n_restr,
                                unsigned long long *out)
{
  long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  unsigned long long c = 0;
  if (i < n) {
    const int v = vsorte_of(pos[i].w);
    c = 3;
    if (v >= ntypes && restr && v < n_restr)
      c = (unsigned long long) ((long long) restr[3 * v] + (long long) restr[3 * v + 1] + (long long) restr[3 * v + 2]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

int cells_count_nactive(imdb200_sim *s)
{
  cudaStream_t st = s->stream;
  unsigned long long *d = (unsigned long long *) (s->d_scal + SC_COUNT - 1);   // scratch slot, as for the list length
  unsigned long long mine = 0;
  CUDA_TRY(cudaMemsetAsync(d, 0, sizeof(unsigned long long), st));
  if (s->n_own > 0) { k_count_nactive<<<cdiv(s->n_own, 256), 256, 0, st>>>(s->pos, s->n_own, s->cfg.ntypes, s->restr, s->n_restr, d); LAUNCH_CHECK(); }
  CUDA_TRY(cudaMemcpyAsync(&mine, d, sizeof(mine), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  long long tot = 0;
  TRY(comm_allgather_ll(s, (long long) mine, &tot));
  s->nactive = tot;
  s->nactive_dirty = 0;
  return 0;
}

// ---- boundary-first warp order (see imdb200_sim::worder) -----------------------------------------------
// a warp of 32 thread slots (32/L atoms) is a boundary warp when one of its atoms sits in the outermost layer of owned
// cells: only those atoms have images in their lists, and only they are sent to neighbours
__global__ void k_warp_flags(const int *cellid, long n_own, int L, Geom g, long n_warps, int *wflag)
{
  const long w = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (w >= n_warps) return;
  const long a0 = w * 32 / L, a1 = min(n_own, (w * 32 + 31) / L + 1);
  int f = 0;
  for (long i = a0; i < a1 && !f; i++) {
    const int c = cellid[i];
    const int cz = c % g.cdim[2], cy = (c / g.cdim[2]) % g.cdim[1], cx = c / (g.cdim[2] * g.cdim[1]);
    f = cx == 1 || cx == g.cdim[0] - 2 || cy == 1 || cy == g.cdim[1] - 2 || cz == 1 || cz == g.cdim[2] - 2;
  }
  wflag[w] = f;
}
__global__ void k_warp_order(const int *wflag, const int *wscan, long n_warps, const int *n_boundary, int *worder)
{
  const long w = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (w >= n_warps) return;
  const int before = wscan[w];                         // boundary warps in front of w
  worder[wflag[w] ? before : *n_boundary + (int) (w - before)] = (int) w;
}

int cells_build_worder(imdb200_sim *s)
{
  s->n_bwarps = 0;
  s->n_warps = (s->n_own * s->lanes + 31) / 32;
  if (s->nranks == 1 || s->n_warps == 0) return 0;
  cudaStream_t st = s->stream;
  if (s->n_warps + 1 > s->worder_cap) {
    CUDA_TRY(cudaStreamSynchronize(st));
    void *old[] = {s->worder, s->wflag, s->wscan};
    for (void *q : old) if (q) cudaFree(q);
    s->worder_cap = s->n_warps + s->n_warps / 8 + 1024;
    CUDA_TRY(cudaMalloc(&s->worder, s->worder_cap * sizeof(int)));
    CUDA_TRY(cudaMalloc(&s->wflag, s->worder_cap * sizeof(int)));
    CUDA_TRY(cudaMalloc(&s->wscan, s->worder_cap * sizeof(int)));
  }
  const int nb = cdiv(s->n_warps, 256);
  k_warp_flags<<<nb, 256, 0, st>>>(s->cellid, s->n_own, s->lanes, s->geom, s->n_warps, s->wflag); LAUNCH_CHECK();
  CUDA_TRY(cudaMemsetAsync(s->wflag + s->n_warps, 0, sizeof(int), st));
  TRY(scan_exclusive(s, s->wflag, s->wscan, (int) s->n_warps + 1, nullptr));       // entry n_warps = number of boundary warps
  k_warp_order<<<nb, 256, 0, st>>>(s->wflag, s->wscan, s->n_warps, s->wscan + s->n_warps, s->worder); LAUNCH_CHECK();
  int nbw = 0;
  CUDA_TRY(cudaMemcpyAsync(&nbw, s->wscan + s->n_warps, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  s->n_bwarps = nbw;
  return 0;
}

// fix_cells + send_cells + make_nblist, i.e. the `0 == have_valid_nbl` branch of calc_forces
// (src/imd_forces_nbl.c:304-317)
int cells_rebuild(imdb200_sim *s)
{
  const Geom &g = s->geom;
  cudaStream_t st = s->stream;
  if (s->n_own <= 0 && s->nranks == 1) return imdb_fail(IMDB200_ERR_ARG, "no atoms");
  s->p2p_step = 0;                            // every exchange of a step with a list build goes through NCCL
  // ---- fix_cells: wrap into the box, bin, sort into cell order, hand atoms over to their new owners ----
  CUDA_TRY(cudaMemsetAsync(s->d_flags, 0, FL_COUNT * sizeof(int), st));
  int *h_extra = s->h_starts;                 // pinned scratch: [0] atoms that stay, [1..28] extra bins
  RebuildClock clk;
  TRY(bin_and_sort(s, s->n_own, s->need_filter, h_extra));
  clk.lap(s, "bin_and_sort");
  long n = h_extra[0];
  s->need_filter = 0;
  if (s->nranks > 1) {
    long n_new = n;
    TRY(comm_migrate(s, h_extra + 1, n, &n_new));             // send_atoms (src/imd_fix_cells_3d.c:331-437)
    if (n_new != n) {                                           // arrivals were appended behind the sorted atoms
      TRY(bin_and_sort(s, n_new, 0, h_extra));
      if (h_extra[0] != n_new) return imdb_fail(IMDB200_ERR_CELLS, "received atoms that are not in this domain");
      n = n_new;
    }
    long long tot = 0;
    TRY(comm_allgather_ll(s, n, &tot));
    if (s->natoms_global == 0) s->natoms_global = tot;
    else if (tot != s->natoms_global) return imdb_fail(IMDB200_ERR_CELLS, "atom count changed from %lld to %lld in fix_cells", s->natoms_global, tot);
  } else s->natoms_global = n;
  s->n_own = n;
  if (s->nactive_dirty) TRY(cells_count_nactive(s));
  // ---- buffer cells: images of the boundary cells, ours or a neighbour's -----------------------------
  clk.lap(s, "migrate");
  TRY(comm_setup_ghosts(s));
  TRY(comm_ghost_pos(s));
  clk.lap(s, "ghost setup + positions");
  const int nb = cdiv(n > 0 ? n : 1, 256);
  // ---- make_nblist -------------------------------------------------------------------------------------
  const int L = s->lanes;
  int overflow_max = 0;                         // largest per-atom count the build that overflowed has seen
  for (int attempt = 0; attempt < 3; attempt++) {
    if (s->max_nb == 0) {
      // size the table from an exact count (estimate_nblist_size, src/imd_forces_nbl.c:74-128)
      CUDA_TRY(cudaMemsetAsync(&s->d_flags[FL_MAXNB], 0, sizeof(int), st));
      k_build_nbl<<<cdiv(n, 128), 128, 0, st>>>(s->pos, n, g, s->cellid, s->cell_start, s->cell_count, s->cell_code,
                                                s->gsrc, s->ghost_raw, nullptr, nullptr, 0, L, s->d_flags, 1); LAUNCH_CHECK();
      TRY(read_flags(s));
      int want = (int) (s->cfg.nbl_size * s->h_flags[FL_MAXNB]) + 2;
      TRY(alloc_nbl(s, want > 8 ? want : 8));
    } else if (overflow_max > 0) {
      // "neighbor table full": the build that overflowed has counted every atom's neighbours all the same
      int want = (int) (s->cfg.nbl_size * overflow_max) + 2;
      TRY(alloc_nbl(s, want > s->max_nb + 2 ? want : s->max_nb + 2));
    } else TRY(alloc_nbl(s, s->max_nb));
    CUDA_TRY(cudaMemsetAsync(&s->d_flags[FL_MAXNB], 0, sizeof(int), st));
    CUDA_TRY(cudaMemsetAsync(&s->d_flags[FL_NBL_OVERFLOW], 0, sizeof(int), st));
    CUDA_TRY(cudaMemsetAsync(&s->d_flags[FL_CELLFULL], 0, sizeof(int), st));
    TRY(launch_build(s, n));
    TRY(read_flags(s));
    clk.lap(s, "list build attempt");
    if (s->h_flags[FL_CELLFULL]) {               // a cell holds more atoms than the candidate bit words cover
      s->cell_words = (s->h_flags[FL_CELLFULL] + 31) / 32;
      attempt--;
      continue;
    }
    if (!s->h_flags[FL_NBL_OVERFLOW]) break;
    overflow_max = s->h_flags[FL_MAXNB];
    if (attempt == 2) return imdb_fail(IMDB200_ERR_NBL, "neighbor table full - increase nbl_size");
  }
  // total list length (last_nbl_len, src/imd_forces_nbl.c:270)
  unsigned long long *d_len = (unsigned long long *) (s->d_scal + SC_COUNT - 1);
  CUDA_TRY(cudaMemsetAsync(d_len, 0, sizeof(unsigned long long), st));
  k_sum_int<<<nb, 256, 0, st>>>(s->nnb, n, d_len); LAUNCH_CHECK();
  // NBL_POS <- ORT (src/imd_forces_nbl.c:142-154)
  k_save_nblpos<<<nb, 256, 0, st>>>(s->pos, n, s->cap_atoms, s->nblpos); LAUNCH_CHECK();
  unsigned long long len = 0;
  CUDA_TRY(cudaMemcpyAsync(&len, d_len, sizeof(len), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  s->nbl_len = (long long) len;
  clk.lap(s, "length + nblpos");
  TRY(cells_build_worder(s));
  clk.lap(s, "warp order");
  s->disp2 = 0.0;                 // NBL_POS == ORT
  TRY(step_snapshot_disp2(s, 1));
  s->skin_all = 0;
  s->have_valid_nbl = 1;
  s->nbl_count++;
  return 0;
}
