// forces_adp.cu -- the angular-dependent terms of ADP builds (`#ifdef ADP` branches of calc_forces,
// src/imd_forces_nbl.c:613-631, 919-929, 1096-1110, 1217-1255; tables adp_upot / adp_wpot, src/imd_potential.c:87-92).
//
// STATUS: green on B200 against fixtures of the reference's `adp` build, Cu and Ni-Al
// (tests/test_gpu_parity.py::test_cuda_adp_matches_reference_fixture).
//
// The terms ride on top of the EAM passes as two kernels of their own, so the benchmark kernels are untouched:
//   k_adp_pass1  after pass 1:  mu_i = sum_j u(r_ij) d_ij,  lambda_i = sum_j w(r_ij) d_ij (x) d_ij  (full list: every
//                atom gathers its own sums; the reference adds -+ the same terms to atom j, :617-629), then the ADP
//                energy 1/2 (sum dev(lambda)^2 + 2 sum offdiag^2 + |mu|^2) (:1096-1110)
//   halo         mu and lambda of the owners into the images (copy_dF in ADP builds, src/imd_comm_force_3d.c:1047-1057)
//   k_adp_pass2  after pass 2:  dipole force  (mu_i-mu_j) u + ((mu_i-mu_j).d) u' d  and quadrupole force
//                2 w v + ((v.d - nu r2) w' - 2 nu w) d  with v = (lambda_i+lambda_j) d, nu = tr/3 (:1217-1254)
// One thread per atom walks all stored list entries (no skin classes, no software pipeline: correctness first).
#include "internal.cuh"
#include <vector>

struct AdpArgs {
  const double4 *pos;
  double4 *frc;
  const int *nbl, *nnb;
  long n_own; int L, R;
  double *mu, *la; long stride;          // SoA: mu[c*stride + i], c < 3; la[c*stride + i], c < 6 = xx yy zz yz zx xy
  double *presstens; long pstride; int press;
  double *partial;
  int *flags;
};

// value and twice the derivative with respect to r^2 of one ADP table column (PAIR_INT in the mode of the build)
__device__ __forceinline__ void adp_lookup(const TabMeta &m, const double4 *coef, int col, double r2, double &val,
                                           double &grad, int &is_short)
{
  int k; double chi;
  tab_index(r2, m.begin[col], m.end[col], m.invstep[col], k, chi, is_short);
  const double2 c01 = ld2(reinterpret_cast<const double *>(coef + (size_t) k * m.ncols + col));
  const double2 c23 = ld2(reinterpret_cast<const double *>(coef + (size_t) k * m.ncols + col) + 2);
  const double4 c = make_double4(c01.x, c01.y, c23.x, c23.y);
  val = fma(chi, fma(chi, fma(chi, c.w, c.z), c.y), c.x);
  grad = 2.0 * m.invstep[col] * fma(chi, fma(3.0 * chi, c.w, c.z + c.z), c.y);
}

__global__ void __launch_bounds__(256) k_adp_pass1(AdpArgs a, DevTables T)
{
  const long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  double red[1] = {0.0};
  int is_short = 0;
  if (i < a.n_own) {
    const double4 xi = a.pos[i];
    const int it = sorte_of(xi.w), nt = T.ntypes, nn = a.nnb[i];
    double mx = 0.0, my = 0.0, mz = 0.0, lxx = 0.0, lyy = 0.0, lzz = 0.0, lyz = 0.0, lzx = 0.0, lxy = 0.0;
    for (int m = 0; m < nn; m++) {
      const int j = a.nbl[nbl_index(i, m, a.L, a.R)] & NBL_JMASK;
      const double4 xj = ld_atom(a.pos + j);
      const double dx = xj.x - xi.x, dy = xj.y - xi.y, dz = xj.z - xi.z;
      const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
      const int col = it * nt + sorte_of(xj.w);
      double v, g;
      if (r2 < T.adpu.end[col]) {                                        // :615-620
        adp_lookup(T.adpu, T.adpuK, col, r2, v, g, is_short);
        mx = fma(v, dx, mx); my = fma(v, dy, my); mz = fma(v, dz, mz);
      }
      if (r2 < T.adpw.end[col]) {                                        // :622-630
        adp_lookup(T.adpw, T.adpwK, col, r2, v, g, is_short);
        lxx = fma(v * dx, dx, lxx); lyy = fma(v * dy, dy, lyy); lzz = fma(v * dz, dz, lzz);
        lyz = fma(v * dy, dz, lyz); lzx = fma(v * dz, dx, lzx); lxy = fma(v * dx, dy, lxy);
      }
    }
    a.mu[i] = mx; a.mu[a.stride + i] = my; a.mu[2 * a.stride + i] = mz;
    a.la[i] = lxx; a.la[a.stride + i] = lyy; a.la[2 * a.stride + i] = lzz;
    a.la[3 * a.stride + i] = lyz; a.la[4 * a.stride + i] = lzx; a.la[5 * a.stride + i] = lxy;
    // ADP energy of the atom (:1096-1110)
    const double tr = (lxx + lyy + lzz) / 3.0;
    double pot = (lxx - tr) * (lxx - tr) + (lyy - tr) * (lyy - tr) + (lzz - tr) * (lzz - tr);
    pot += 2.0 * (lyz * lyz + lzx * lzx + lxy * lxy);
    pot += mx * mx + my * my + mz * mz;
    pot *= 0.5;
    double4 f = a.frc[i];
    f.w += pot;
    a.frc[i] = f;
    red[0] = pot;
  }
  if (is_short) atomicExch(&a.flags[FL_SHORT], 1);
  block_sum_store<1>(red, a.partial);
}

__global__ void __launch_bounds__(256) k_adp_pass2(AdpArgs a, DevTables T)
{
  const long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
  double red[1] = {0.0};
  int is_short = 0;
  if (i < a.n_own) {
    const double4 xi = a.pos[i];
    const int it = sorte_of(xi.w), nt = T.ntypes, nn = a.nnb[i];
    const long st = a.stride;
    const double mix = a.mu[i], miy = a.mu[st + i], miz = a.mu[2 * st + i];
    const double ixx = a.la[i], iyy = a.la[st + i], izz = a.la[2 * st + i];
    const double iyz = a.la[3 * st + i], izx = a.la[4 * st + i], ixy = a.la[5 * st + i];
    double fx = 0.0, fy = 0.0, fz = 0.0, vir = 0.0;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0, s5 = 0.0;
    for (int m = 0; m < nn; m++) {
      const int j = a.nbl[nbl_index(i, m, a.L, a.R)] & NBL_JMASK;
      const double4 xj = ld_atom(a.pos + j);
      const double dx = xj.x - xi.x, dy = xj.y - xi.y, dz = xj.z - xi.z;
      const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
      const int col1 = sorte_of(xj.w) * nt + it;                         // :1167
      double gx = 0.0, gy = 0.0, gz = 0.0;
      bool have = false;
      if (r2 < T.adpu.end[col1]) {                                       // dipole distortion :1217-1229
        double u, du;
        adp_lookup(T.adpu, T.adpuK, col1, r2, u, du, is_short);
        const double ux = mix - __ldg(a.mu + j), uy = miy - __ldg(a.mu + st + j), uz = miz - __ldg(a.mu + 2 * st + j);
        const double tmp = (ux * dx + uy * dy + uz * dz) * du;
        gx += ux * u + tmp * dx; gy += uy * u + tmp * dy; gz += uz * u + tmp * dz;
        have = true;
      }
      if (r2 < T.adpw.end[col1]) {                                       // quadrupole distortion :1231-1254
        double w, dw;
        adp_lookup(T.adpw, T.adpwK, col1, r2, w, dw, is_short);
        const double xx = ixx + __ldg(a.la + j), yy = iyy + __ldg(a.la + st + j), zz = izz + __ldg(a.la + 2 * st + j);
        const double yz = iyz + __ldg(a.la + 3 * st + j), zx = izx + __ldg(a.la + 4 * st + j), xy = ixy + __ldg(a.la + 5 * st + j);
        const double vx = xx * dx + xy * dy + zx * dz, vy = xy * dx + yy * dy + yz * dz, vz = zx * dx + yz * dy + zz * dz;
        const double nu = (xx + yy + zz) / 3.0;
        const double f1 = 2.0 * w;
        const double f2 = ((vx * dx + vy * dy + vz * dz) - nu * r2) * dw - nu * f1;
        gx += f1 * vx + f2 * dx; gy += f1 * vy + f2 * dy; gz += f1 * vz + f2 * dz;
        have = true;
      }
      if (have) {                                                        // :1267-1305
        fx += gx; fy += gy; fz += gz;
        vir += dx * gx + dy * gy + dz * gz;
        if (a.press) { s0 = fma(dx, gx, s0); s1 = fma(dy, gy, s1); s2 = fma(dz, gz, s2);
                       s3 = fma(dy, gz, s3); s4 = fma(dz, gx, s4); s5 = fma(dx, gy, s5); }
      }
    }
    double4 f = a.frc[i];
    f.x += fx; f.y += fy; f.z += fz;
    a.frc[i] = f;
    if (a.press) {
      double *p = a.presstens + i;
      p[0] -= 0.5 * s0; p[a.pstride] -= 0.5 * s1; p[2 * a.pstride] -= 0.5 * s2;
      p[3 * a.pstride] -= 0.5 * s3; p[4 * a.pstride] -= 0.5 * s4; p[5 * a.pstride] -= 0.5 * s5;
    }
    red[0] = -0.5 * vir;                                               // virial -= SPROD(d,force) once per pair (:1280)
  }
  if (is_short) atomicExch(&a.flags[FL_SHORT], 1);
  block_sum_store<1>(red, a.partial);
}

// ---- host side ------------------------------------------------------------------------------------------------
int adp_ensure_arrays(imdb200_sim *s)
{
  if (!s->tabs.have_adp || s->adp_cap == s->cap_atoms) return 0;
  if (s->adp_mu) cudaFree(s->adp_mu);
  if (s->adp_la) cudaFree(s->adp_la);
  s->adp_mu = s->adp_la = nullptr; s->adp_cap = 0;
  if (s->cap_atoms <= 0) return 0;
  CUDA_TRY(cudaMalloc(&s->adp_mu, 3 * s->cap_atoms * sizeof(double)));
  CUDA_TRY(cudaMalloc(&s->adp_la, 6 * s->cap_atoms * sizeof(double)));
  s->adp_cap = s->cap_atoms;
  return 0;
}

static AdpArgs adp_args(imdb200_sim *s)
{
  AdpArgs a;
  a.pos = s->pos; a.frc = s->frc; a.nbl = s->nbl; a.nnb = s->nnb;
  a.n_own = s->n_own; a.L = s->lanes; a.R = s->max_nb / s->lanes;
  a.mu = s->adp_mu; a.la = s->adp_la; a.stride = s->adp_cap;
  a.presstens = s->presstens; a.pstride = s->cap_atoms; a.press = s->press_calc;
  a.partial = s->d_partial; a.flags = s->d_flags;
  return a;
}

int forces_adp_pass1(imdb200_sim *s)
{
  TRY(adp_ensure_arrays(s));
  const int nb = cdiv(s->n_own, 256);
  if (nb == 0) return 0;
  k_adp_pass1<<<nb, 256, 0, s->stream>>>(adp_args(s), s->tabs); LAUNCH_CHECK();
  const int slots[1] = {SC_EPOT};
  return reduce_finish(s, nb, 1, slots, 1);            // adds to the EAM energy of pass 1
}

int forces_adp_pass2(imdb200_sim *s)
{
  const int nb = cdiv(s->n_own, 256);
  if (nb == 0) return 0;
  k_adp_pass2<<<nb, 256, 0, s->stream>>>(adp_args(s), s->tabs); LAUNCH_CHECK();
  const int slots[1] = {SC_VIRIAL};
  return reduce_finish(s, nb, 1, slots, 1);
}

int forces_adp_halo(imdb200_sim *s)
{
  TRY(comm_ghost_field(s, s->adp_mu, 3, s->adp_cap));
  return comm_ghost_field(s, s->adp_la, 6, s->adp_cap);
}
