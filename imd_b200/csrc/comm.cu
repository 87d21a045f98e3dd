// comm.cu -- buffer-cell (ghost) halo of one domain and its exchange between GPUs.
//
// Reference behaviour restated here (not its data structures):
//   setup_mpi_topology          src/imd_geom_mpi_3d.c:32-90    rank grid, x-major like MPI_Cart_create
//   send_cells + copy/pack/unpack_cell   src/imd_comm_force_3d.c:222-396, 726-889   ghost positions
//   copy_dF / pack_dF / unpack_dF        src/imd_comm_force_3d.c:1031-1250          ghost 2F'(rho)
//   send_forces + add/pack/unpack_forces src/imd_comm_force_3d.c:569-714, 897-1020  ghost sums back to owners
//   send_atoms / fix_cells               src/imd_fix_cells_3d.c:36-201, 331-437     atom migration
//   MPI_Allreduce sites                  src/imd_forces_nbl.c:1975-1994, 2032; src/imd_integrate.c:453, 1115
//
// The reference sweeps z, y, x with 6 (5) face messages per exchange so that edges and corners travel
// two and three times (Plimpton).  Three dependent message rounds per exchange are latency-bound on
// NVLink, so here every rank talks to its up-to-26 neighbours directly in ONE
// ncclGroup per exchange: the receive side of direction d is a contiguous slice of the atom arrays
// (ghost atoms are stored direction-major), so ncclRecv writes in place and only the send side needs a
// pack kernel.  A neighbour that is this rank itself (periodic wrap with cpu_dim == 1 on that axis) is
// served by a gather kernel instead of a message.  Senders transmit unshifted positions; the receiver
// adds the periodic image shift stage by stage exactly as the reference's three sweeps do, and keeps the
// unshifted copy for the neighbour-list build (cells.cu) -- that is what makes the neighbour set
// bit-identical to the single-process reference on any process grid.
//
// NCCL is bound at run time (dlopen) so that the library loads on hosts without it and uses the copy a
// host process (e.g. torch) has already loaded.
#include "internal.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

// ---- NCCL binding ----------------------------------------------------------------------------------------
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)(void);
  ncclResult_t (*GroupEnd)(void);
  const char *(*GetErrorString)(ncclResult_t);
  void *handle;
};
static NcclApi g_nccl = {};

static int nccl_load(void)
{
  if (g_nccl.handle) return 0;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);     // the copy the host process already uses
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return imdb_fail(IMDB200_ERR_COMM, "cannot load libnccl.so.2: %s", dlerror());
#define BIND(name) do { *(void **) &g_nccl.name = dlsym(h, "nccl" #name); \
  if (!g_nccl.name) return imdb_fail(IMDB200_ERR_COMM, "libnccl lacks nccl" #name); } while (0)
  BIND(GetUniqueId); BIND(CommInitRank); BIND(CommDestroy); BIND(Send); BIND(Recv); BIND(AllGather);
  BIND(GroupStart); BIND(GroupEnd); BIND(GetErrorString);
#undef BIND
  g_nccl.handle = h;
  return 0;
}
#define NCCL_TRY(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) \
  return imdb_fail(IMDB200_ERR_COMM, "%s:%d: %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); } while (0)

extern "C" int imdb200_comm_unique_id(void *id128)
{
  if (!id128) return imdb_fail(IMDB200_ERR_ARG, "null argument");
  TRY(nccl_load());
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  NCCL_TRY(g_nccl.GetUniqueId((ncclUniqueId *) id128));
  return 0;
}

extern "C" int imdb200_comm_init(imdb200_sim *s, const void *id128, int rank, int nranks)
{
  if (!s || !id128) return imdb_fail(IMDB200_ERR_ARG, "null argument");
  if (nranks != s->nranks) return imdb_fail(IMDB200_ERR_ARG, "communicator has %d ranks, cpu_dim needs %d", nranks, s->nranks);
  if (rank != s->rank) return imdb_fail(IMDB200_ERR_ARG, "rank %d does not match my_coord (x-major rank %d)", rank, s->rank);
  TRY(nccl_load());
  CUDA_TRY(cudaSetDevice(s->cfg.device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c;
  NCCL_TRY(g_nccl.CommInitRank(&c, nranks, id, rank));
  s->nccl_comm = c;
  if (!s->d_all) {
    CUDA_TRY(cudaMalloc(&s->d_all, (size_t) nranks * SC_COUNT * sizeof(double)));
    CUDA_TRY(cudaMalloc(&s->d_glob, SC_COUNT * sizeof(double)));
    CUDA_TRY(cudaMemset(s->d_glob, 0, SC_COUNT * sizeof(double)));
  }
  return comm_p2p_enable(s);               // peer-memory halo, only with IMDB200_HALO_P2P=1
}

void comm_free(imdb200_sim *s)
{
  comm_p2p_free(s);
  if (s->nccl_comm && g_nccl.handle) g_nccl.CommDestroy((ncclComm_t) s->nccl_comm);
  s->nccl_comm = nullptr;
  void *ptrs[] = {s->gcells, s->gcount, s->gstart, s->scells, s->scount, s->sstart, s->send_idx, s->sendbuf4,
                  s->sendbuf1, s->sendbufi, s->d_all};
  for (void *p : ptrs) if (p) cudaFree(p);
  if (s->d_glob && s->d_glob != s->d_scal) cudaFree(s->d_glob);
  if (s->h_starts) cudaFreeHost(s->h_starts);
  s->gcells = nullptr; s->gcount = s->gstart = s->scells = s->scount = s->sstart = s->send_idx = nullptr;
  s->sendbuf4 = nullptr; s->sendbuf1 = nullptr; s->sendbufi = nullptr; s->d_all = nullptr; s->d_glob = nullptr;
  s->h_starts = nullptr;
}

static int need_comm(imdb200_sim *s)
{
  if (s->nranks > 1 && !s->nccl_comm) return imdb_fail(IMDB200_ERR_COMM, "cpu_dim has %d ranks: call imdb200_comm_init first", s->nranks);
  return 0;
}

// ---- halo plan --------------------------------------------------------------------------------------------
// Region of direction d along one axis with s in {-1,0,1}: receive cells, their source cells on the
// neighbour, and the owned cells sent towards d.
static void axis_range(int sgn, int cdim, int *lo, int *hi) { if (sgn == 0) { *lo = 1; *hi = cdim - 2; } else *lo = *hi = (sgn < 0 ? 0 : cdim - 1); }

int comm_plan(imdb200_sim *s)
{
  const Geom &g = s->geom;
  std::vector<GhostCell> gc;
  std::vector<int> sc;
  int peers[27], codes[27];
  imdb200_halo_peers(s->cfg.cpu_dim, s->cfg.my_coord, g.pbc, peers, codes);
  for (int d = 0; d < 27; d++) {
    DirPlan &D = s->dir[d];
    memset(&D, 0, sizeof(D));
    D.peer = (d == 13) ? -1 : peers[d];
    D.scell_off = -1;
  }
  // Receive and send regions are laid out PEER-major so that everything exchanged with one neighbour rank is one
  // contiguous slice = one message per peer and exchange (up to 26 directions share 1..7 peers on small process
  // grids; a message costs ~6 us whatever its size).  Within a peer the receive regions follow ascending
  // direction and the send regions descending direction: what is sent towards d arrives as direction 26-d.
  int ro[26], so[26], nro = 0, nso = 0;
  imdb200_halo_message_order(peers, s->rank, ro, &nro, so, &nso);      // host/topology.c
  const std::vector<int> rorder(ro, ro + nro);
  const std::vector<int> sorder(so, so + nso);
  for (int pass = 0; pass < 2; pass++) {
    for (int d : (pass == 0 ? rorder : sorder)) {
      DirPlan &D = s->dir[d];
      const int sg[3] = {d % 3 - 1, (d / 3) % 3 - 1, d / 9 - 1};
      const int code = codes[d];
      const bool self = D.peer == s->rank;
      int lo[3], hi[3];
      for (int a = 0; a < 3; a++) axis_range(sg[a], g.cdim[a], &lo[a], &hi[a]);
      if (pass == 0) D.gcell_off = (int) gc.size(); else D.scell_off = (int) sc.size();
      for (int i = lo[0]; i <= hi[0]; i++)
        for (int j = lo[1]; j <= hi[1]; j++)
          for (int k = lo[2]; k <= hi[2]; k++) {
            const int c[3] = {i, j, k};
            int src[3], snd[3];
            for (int a = 0; a < 3; a++) {
              // receive cell -> its source on the neighbour; owned cell sent towards d
              src[a] = sg[a] == 0 ? c[a] : (sg[a] < 0 ? g.cdim[a] - 2 : 1);
              snd[a] = sg[a] == 0 ? c[a] : (sg[a] < 0 ? 1 : g.cdim[a] - 2);
            }
            if (pass == 0) {
              GhostCell x;
              x.dst = (i * g.cdim[1] + j) * g.cdim[2] + k;
              x.src = self ? (src[0] * g.cdim[1] + src[1]) * g.cdim[2] + src[2] : -1;
              x.code = code;
              x.peer = D.peer;
              gc.push_back(x);
            } else {
              sc.push_back((snd[0] * g.cdim[1] + snd[1]) * g.cdim[2] + snd[2]);
            }
          }
      if (pass == 0) D.ncells = (int) gc.size() - D.gcell_off;
    }
  }
  // one PeerPlan per neighbour rank: the cell slices are fixed by the grid, the atom slices are set at every rebuild
  s->n_peers = 0;
  for (int d : rorder) {
    const DirPlan &D = s->dir[d];
    if (D.peer == s->rank) continue;
    if (s->n_peers == 0 || s->peers[s->n_peers - 1].peer != D.peer) {
      PeerPlan &P = s->peers[s->n_peers++];
      memset(&P, 0, sizeof(P));
      P.peer = D.peer; P.gcell_off = D.gcell_off; P.scell_off = -1;
    }
    s->peers[s->n_peers - 1].ncells += D.ncells;
  }
  for (int d : sorder) {
    const DirPlan &D = s->dir[d];
    if (D.peer == s->rank) continue;
    for (int q = 0; q < s->n_peers; q++)
      if (s->peers[q].peer == D.peer && s->peers[q].scell_off < 0) s->peers[q].scell_off = D.scell_off;
  }
  void *old[] = {s->gcells, s->gcount, s->gstart, s->scells, s->scount, s->sstart};
  for (void *p : old) if (p) cudaFree(p);
  if (s->h_starts) cudaFreeHost(s->h_starts);
  s->gcells = nullptr; s->gcount = s->gstart = s->scells = s->scount = s->sstart = nullptr; s->h_starts = nullptr;
  s->n_gcells = (int) gc.size();
  s->n_scells = (int) sc.size();
  CUDA_TRY(cudaMalloc(&s->gcells, (gc.size() + 1) * sizeof(GhostCell)));
  CUDA_TRY(cudaMalloc(&s->gcount, (gc.size() + 1) * sizeof(int)));
  CUDA_TRY(cudaMalloc(&s->gstart, (gc.size() + 1) * sizeof(int)));
  CUDA_TRY(cudaMalloc(&s->scells, (sc.size() + 1) * sizeof(int)));
  CUDA_TRY(cudaMalloc(&s->scount, (sc.size() + 1) * sizeof(int)));
  CUDA_TRY(cudaMalloc(&s->sstart, (sc.size() + 1) * sizeof(int)));
  if (!gc.empty()) CUDA_TRY(cudaMemcpy(s->gcells, gc.data(), gc.size() * sizeof(GhostCell), cudaMemcpyHostToDevice));
  if (!sc.empty()) CUDA_TRY(cudaMemcpy(s->scells, sc.data(), sc.size() * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMallocHost(&s->h_starts, (gc.size() + sc.size() + 128) * sizeof(int)));
  return 0;
}

// ---- kernels ----------------------------------------------------------------------------------------------
__global__ void k_cell_counts(const int *cells, int n, const int *cell_count, int *out)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = cell_count[cells[i]];
}

__global__ void k_ghost_count_self(const GhostCell *gc, int ng, const int *cell_count, int *gcount)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ng && gc[i].src >= 0) gcount[i] = cell_count[gc[i].src];
}

// one warp per ghost cell: publish the cell's range and record where every image comes from
__global__ void k_ghost_fill(const GhostCell *gc, int ng, const int *gstart, const int *gcount, long n_own,
                             int *cell_start, int *cell_count, int *cell_code, int *gsrc, int *cellid_g)
{
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= ng) return;
  GhostCell x = gc[w];
  const int cnt = gcount[w], st = gstart[w], s0 = x.src >= 0 ? cell_start[x.src] : -1;
  for (int t = lane; t < cnt; t += 32) { gsrc[st + t] = s0 >= 0 ? s0 + t : -1; cellid_g[st + t] = x.dst; }
  if (lane == 0) { cell_start[x.dst] = (int) n_own + st; cell_count[x.dst] = cnt; cell_code[x.dst] = x.code; }
}

__global__ void k_send_fill(const int *scells, int ns, const int *sstart, const int *scount, const int *cell_start,
                            int *send_idx)
{
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= ns) return;
  const int cnt = scount[w], st = sstart[w], s0 = cell_start[scells[w]];
  for (int t = lane; t < cnt; t += 32) send_idx[st + t] = s0 + t;
}

__global__ void k_pack4(const double4 *src, const int *idx, long n, double4 *out)
{
  long t = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (t < n) out[t] = src[idx[t]];
}
__global__ void k_pack1(const double *src, const int *idx, long n, double *out)
{
  long t = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (t < n) out[t] = src[idx[t]];
}
__global__ void k_packi(const int *src, const int *idx, long n, int *out)
{
  long t = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (t < n) out[t] = src[idx[t]];
}

// ghost position = image of the owner's position: gathered here (own atoms) or as received (ghost_raw)
__global__ void k_ghost_pos(double4 *pos, long n_own, long n_ghost, const int *gsrc, const double4 *ghost_raw,
                            const int *cellid_g, const int *cell_code, Geom g)
{
  long t = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (t >= n_ghost) return;
  const int src = gsrc[t];
  pos[n_own + t] = image_pos(src >= 0 ? pos[src] : ghost_raw[t], cell_code[cellid_g[t]], g);
}

__global__ void k_ghost_dF(double *dF, double4 *posdf, const double4 *pos, long n_own, long n_ghost, const int *gsrc)
{
  long t = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (t >= n_ghost) return;
  const int src = gsrc[t];
  const double d = src >= 0 ? dF[src] : dF[n_own + t];          // remote values were received in place
  dF[n_own + t] = d;
  if (posdf) { const double4 p = pos[n_own + t]; posdf[n_own + t] = make_double4(p.x, p.y, p.z, d); }
}

__global__ void k_ghost_num_self(int *ghost_num, const int *nummer, long n_ghost, const int *gsrc)
{
  long t = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (t < n_ghost && gsrc[t] >= 0) ghost_num[t] = nummer[gsrc[t]];
}

// reverse direction: add what the images accumulated to their owners (add_forces / unpack_forces)
__global__ void k_reverse_self(double *field, int ncomp, long stride, long n_own, long n_ghost, const int *gsrc)
{
  long t = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (t >= n_ghost) return;
  const int src = gsrc[t];
  if (src < 0) return;
  for (int c = 0; c < ncomp; c++) atomicAdd(&field[c * stride + src], field[c * stride + n_own + t]);
}
__global__ void k_reverse_unpack(double *field, int ncomp, long stride, const int *send_idx, long n_send, const double *buf)
{
  long t = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (t >= n_send) return;
  for (int c = 0; c < ncomp; c++) atomicAdd(&field[c * stride + send_idx[t]], buf[c * n_send + t]);
}

// ---- exchanges ----------------------------------------------------------------------------------------------
// forward: owner -> images.  One group, one send and one receive per neighbour rank (see comm_plan).
template <typename T>
static int exchange_forward(imdb200_sim *s, const T *sendbuf, T *recv_base, ncclDataType_t ty, int per_elem)
{
  ncclComm_t c = (ncclComm_t) s->nccl_comm;
  NCCL_TRY(g_nccl.GroupStart());
  for (int q = 0; q < s->n_peers; q++) {
    const PeerPlan &P = s->peers[q];
    if (P.send_cnt) NCCL_TRY(g_nccl.Send(sendbuf + (size_t) P.send_off * per_elem, (size_t) P.send_cnt * per_elem, ty, P.peer, c, s->stream));
    if (P.recv_cnt) NCCL_TRY(g_nccl.Recv(recv_base + (size_t) P.recv_off * per_elem, (size_t) P.recv_cnt * per_elem, ty, P.peer, c, s->stream));
  }
  NCCL_TRY(g_nccl.GroupEnd());
  return 0;
}

static int ensure_send_capacity(imdb200_sim *s, long n)
{
  if (n <= s->cap_send) return 0;
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  void *old[] = {s->send_idx, s->sendbuf4, s->sendbuf1, s->sendbufi};
  for (void *p : old) if (p) cudaFree(p);
  const long cap = n + n / 4 + 1024;
  CUDA_TRY(cudaMalloc(&s->send_idx, cap * sizeof(int)));
  CUDA_TRY(cudaMalloc(&s->sendbuf4, cap * sizeof(double4)));
  CUDA_TRY(cudaMalloc(&s->sendbuf1, cap * 8 * sizeof(double)));    // up to 8 components in comm_reverse_add
  CUDA_TRY(cudaMalloc(&s->sendbufi, cap * sizeof(int)));
  s->cap_send = cap;
  return 0;
}

static int allgather_bytes(imdb200_sim *s, const void *mine, void *all, size_t bytes)
{
  NCCL_TRY(g_nccl.AllGather(mine, all, bytes, ncclInt8, (ncclComm_t) s->nccl_comm, s->stream));
  return 0;
}

// At a rebuild: how many atoms sit in every buffer cell, where the images go, what we have to send.
int comm_setup_ghosts(imdb200_sim *s)
{
  TRY(need_comm(s));
  const Geom &g = s->geom;
  cudaStream_t st = s->stream;
  const long n = s->n_own;
  CUDA_TRY(cudaMemsetAsync(s->cell_code, 0, (g.nall + NBIN_EXTRA) * sizeof(int), st));
  s->n_ghost = 0; s->n_send = 0;
  for (int d = 0; d < 27; d++) { s->dir[d].recv_off = s->dir[d].recv_cnt = s->dir[d].send_off = s->dir[d].send_cnt = 0; }
  for (int q = 0; q < s->n_peers; q++) { s->peers[q].recv_off = s->peers[q].recv_cnt = s->peers[q].send_off = s->peers[q].send_cnt = 0; }
  if (!s->n_gcells) return 0;
  const int ng = s->n_gcells, ns = s->n_scells;
  CUDA_TRY(cudaMemsetAsync(s->gcount, 0, (ng + 1) * sizeof(int), st));
  k_ghost_count_self<<<cdiv(ng, 256), 256, 0, st>>>(s->gcells, ng, s->cell_count, s->gcount); LAUNCH_CHECK();
  if (ns) {
    // per-cell populations of the cells we send; the neighbour needs them to lay out its buffer cells
    CUDA_TRY(cudaMemsetAsync(s->scount, 0, (ns + 1) * sizeof(int), st));
    k_cell_counts<<<cdiv(ns, 256), 256, 0, st>>>(s->scells, ns, s->cell_count, s->scount); LAUNCH_CHECK();
    ncclComm_t c = (ncclComm_t) s->nccl_comm;
    NCCL_TRY(g_nccl.GroupStart());
    for (int q = 0; q < s->n_peers; q++) {
      const PeerPlan &P = s->peers[q];
      NCCL_TRY(g_nccl.Send(s->scount + P.scell_off, P.ncells, ncclInt32, P.peer, c, st));
      NCCL_TRY(g_nccl.Recv(s->gcount + P.gcell_off, P.ncells, ncclInt32, P.peer, c, st));
    }
    NCCL_TRY(g_nccl.GroupEnd());
    TRY(scan_exclusive(s, s->scount, s->sstart, ns + 1, nullptr));
  }
  TRY(scan_exclusive(s, s->gcount, s->gstart, ng + 1, nullptr));      // entry ng = total
  int *h_g = s->h_starts, *h_s = s->h_starts + ng + 1;
  CUDA_TRY(cudaMemcpyAsync(h_g, s->gstart, (ng + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (ns) CUDA_TRY(cudaMemcpyAsync(h_s, s->sstart, (ns + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  s->n_ghost = h_g[ng];
  s->n_send = ns ? h_s[ns] : 0;
  for (int d = 0; d < 27; d++) {
    DirPlan &D = s->dir[d];
    if (D.peer < 0) continue;
    D.recv_off = h_g[D.gcell_off];
    D.recv_cnt = h_g[D.gcell_off + D.ncells] - D.recv_off;
    if (D.scell_off >= 0) { D.send_off = h_s[D.scell_off]; D.send_cnt = h_s[D.scell_off + D.ncells] - D.send_off; }
  }
  for (int q = 0; q < s->n_peers; q++) {
    PeerPlan &P = s->peers[q];
    P.recv_off = h_g[P.gcell_off]; P.recv_cnt = h_g[P.gcell_off + P.ncells] - P.recv_off;
    P.send_off = h_s[P.scell_off]; P.send_cnt = h_s[P.scell_off + P.ncells] - P.send_off;
  }
  TRY(cells_ensure_capacity(s, n + s->n_ghost + 1));
  TRY(ensure_send_capacity(s, s->n_send));
  const int nbw = cdiv((long) ng * 32, 256);
  k_ghost_fill<<<nbw, 256, 0, st>>>(s->gcells, ng, s->gstart, s->gcount, n, s->cell_start, s->cell_count, s->cell_code,
                                    s->gsrc, s->cellid + n); LAUNCH_CHECK();
  if (ns) { k_send_fill<<<cdiv((long) ns * 32, 256), 256, 0, st>>>(s->scells, ns, s->sstart, s->scount, s->cell_start, s->send_idx); LAUNCH_CHECK(); }
  // atom numbers of the images (imdb200_get_nblist reports pairs by NUMMER)
  if (s->n_ghost) { k_ghost_num_self<<<cdiv(s->n_ghost, 256), 256, 0, st>>>(s->ghost_num, s->nummer, s->n_ghost, s->gsrc); LAUNCH_CHECK(); }
  if (s->n_send) {
    k_packi<<<cdiv(s->n_send, 256), 256, 0, st>>>(s->nummer, s->send_idx, s->n_send, s->sendbufi); LAUNCH_CHECK();
    TRY(exchange_forward<int>(s, s->sendbufi, s->ghost_num, ncclInt32, 1));
  }
  if (s->p2p_on) TRY(comm_p2p_setup(s, allgather_bytes));
  return 0;
}

// send_cells(copy_cell, pack_cell, unpack_cell): positions (and types) of the owners into the buffer cells
int comm_ghost_pos(imdb200_sim *s)
{
  const bool p2p = s->p2p_step && comm_p2p_ready(s);
  if (p2p) TRY(comm_p2p_positions(s));     // direct stores into the neighbours' ghost_raw + stream flags
  if (s->n_ghost == 0) return 0;
  if (s->n_send && !p2p) {
    TRY(need_comm(s));
    k_pack4<<<cdiv(s->n_send, 256), 256, 0, s->stream>>>(s->pos, s->send_idx, s->n_send, s->sendbuf4); LAUNCH_CHECK();
    TRY(exchange_forward<double>(s, (const double *) s->sendbuf4, (double *) s->ghost_raw, ncclFloat64, 4));
  }
  k_ghost_pos<<<cdiv(s->n_ghost, 256), 256, 0, s->stream>>>(s->pos, s->n_own, s->n_ghost, s->gsrc, s->ghost_raw,
                                                              s->cellid + s->n_own, s->cell_code, s->geom);
  LAUNCH_CHECK();
  return 0;
}

// the images from what has arrived (the tail of comm_ghost_pos / comm_ghost_dF, for the overlapped peer-memory exchange)
int comm_ghost_pos_finish(imdb200_sim *s)
{
  if (s->n_ghost == 0) return 0;
  k_ghost_pos<<<cdiv(s->n_ghost, 256), 256, 0, s->stream>>>(s->pos, s->n_own, s->n_ghost, s->gsrc, s->ghost_raw,
                                                              s->cellid + s->n_own, s->cell_code, s->geom);
  LAUNCH_CHECK();
  return 0;
}
int comm_ghost_dF_finish(imdb200_sim *s)
{
  if (s->n_ghost == 0) return 0;
  k_ghost_dF<<<cdiv(s->n_ghost, 256), 256, 0, s->stream>>>(s->dF, s->posdf, s->pos, s->n_own,
                                                             s->n_ghost, s->gsrc);
  LAUNCH_CHECK();
  return 0;
}

// send_cells(copy_dF, pack_dF, unpack_dF): 2F'(rho) of the owners into the buffer cells
int comm_ghost_dF(imdb200_sim *s)
{
  const bool p2p = s->p2p_step && comm_p2p_ready(s);
  if (p2p) TRY(comm_p2p_dF(s));
  if (s->n_ghost == 0) return 0;
  if (s->n_send && !p2p) {
    TRY(need_comm(s));
    k_pack1<<<cdiv(s->n_send, 256), 256, 0, s->stream>>>(s->dF, s->send_idx, s->n_send, s->sendbuf1); LAUNCH_CHECK();
    TRY(exchange_forward<double>(s, s->sendbuf1, s->dF + s->n_own, ncclFloat64, 1));
  }
  k_ghost_dF<<<cdiv(s->n_ghost, 256), 256, 0, s->stream>>>(s->dF, s->posdf, s->pos, s->n_own,
                                                             s->n_ghost, s->gsrc);
  LAUNCH_CHECK();
  return 0;
}

// owners -> images for a caller-chosen SoA field (ADP: mu and lambda travel with EAM_DF, src/imd_comm_force_3d.c:1047-1057)
__global__ void k_pack_soa(const double *src, long stride, int ncomp, const int *idx, long n, double *out)
{
  long t = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int i = idx[t];
  for (int c = 0; c < ncomp; c++) out[(size_t) c * n + t] = src[(size_t) c * stride + i];
}
__global__ void k_ghost_soa_self(double *field, long stride, int ncomp, long n_own, long n_ghost, const int *gsrc)
{
  long t = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (t >= n_ghost) return;
  const int src = gsrc[t];
  if (src < 0) return;                                  // remote values were received in place
  for (int c = 0; c < ncomp; c++) field[(size_t) c * stride + n_own + t] = field[(size_t) c * stride + src];
}

int comm_ghost_field(imdb200_sim *s, double *field, int ncomp, long stride)
{
  if (s->n_ghost == 0) return 0;
  if (ncomp < 1 || ncomp > 8) return imdb_fail(IMDB200_ERR_ARG, "comm_ghost_field: 1..8 components");
  if (s->n_send) {
    TRY(need_comm(s));
    ncclComm_t c = (ncclComm_t) s->nccl_comm;
    k_pack_soa<<<cdiv(s->n_send, 256), 256, 0, s->stream>>>(field, stride, ncomp, s->send_idx, s->n_send, s->sendbuf1); LAUNCH_CHECK();
    NCCL_TRY(g_nccl.GroupStart());
    for (int comp = 0; comp < ncomp; comp++)
      for (int q = 0; q < s->n_peers; q++) {
        const PeerPlan &P = s->peers[q];
        if (P.send_cnt) NCCL_TRY(g_nccl.Send(s->sendbuf1 + (size_t) comp * s->n_send + P.send_off, P.send_cnt, ncclFloat64, P.peer, c, s->stream));
        if (P.recv_cnt) NCCL_TRY(g_nccl.Recv(field + (size_t) comp * stride + s->n_own + P.recv_off, P.recv_cnt, ncclFloat64, P.peer, c, s->stream));
      }
    NCCL_TRY(g_nccl.GroupEnd());
  }
  k_ghost_soa_self<<<cdiv(s->n_ghost, 256), 256, 0, s->stream>>>(field, stride, ncomp, s->n_own, s->n_ghost, s->gsrc);
  LAUNCH_CHECK();
  return 0;
}

// EEAM builds copy EAM_DM together with EAM_DF (src/imd_comm_force_3d.c:1044-1046, 1116-1118, 1155-1157)
int comm_ghost_dM(imdb200_sim *s)
{
  if (s->n_ghost == 0) return 0;
  if (s->n_send) {
    TRY(need_comm(s));
    k_pack1<<<cdiv(s->n_send, 256), 256, 0, s->stream>>>(s->dM, s->send_idx, s->n_send, s->sendbuf1); LAUNCH_CHECK();
    TRY(exchange_forward<double>(s, s->sendbuf1, s->dM + s->n_own, ncclFloat64, 1));
  }
  k_ghost_dF<<<cdiv(s->n_ghost, 256), 256, 0, s->stream>>>(s->dM, nullptr, s->pos, s->n_own, s->n_ghost, s->gsrc);
  LAUNCH_CHECK();
  return 0;
}

// send_forces(add_*, pack_*, unpack_*): field[c*stride + i], c < ncomp.  The image entries
// [n_own, n_own+n_ghost) are added to their owners; the full-list force kernels never write to images,
// so the step loop does not need this -- it is the reference's reverse path for callers that accumulate
// on images (half lists, per-atom tallies).
int comm_reverse_add(imdb200_sim *s, double *field, int ncomp, long stride)
{
  if (s->n_ghost == 0) return 0;
  if (ncomp < 1 || ncomp > 8) return imdb_fail(IMDB200_ERR_ARG, "comm_reverse_add: 1..8 components");
  cudaStream_t st = s->stream;
  if (s->n_send) {
    TRY(need_comm(s));
    ncclComm_t c = (ncclComm_t) s->nccl_comm;
    // the roles swap: every receive slice of the forward exchange is now sent, component by component
    NCCL_TRY(g_nccl.GroupStart());
    for (int comp = 0; comp < ncomp; comp++) {
      for (int q = 0; q < s->n_peers; q++) {
        const PeerPlan &P = s->peers[q];
        if (P.recv_cnt) NCCL_TRY(g_nccl.Send(field + comp * stride + s->n_own + P.recv_off, P.recv_cnt, ncclFloat64, P.peer, c, st));
        if (P.send_cnt) NCCL_TRY(g_nccl.Recv(s->sendbuf1 + comp * s->n_send + P.send_off, P.send_cnt, ncclFloat64, P.peer, c, st));
      }
    }
    NCCL_TRY(g_nccl.GroupEnd());
    k_reverse_unpack<<<cdiv(s->n_send, 256), 256, 0, st>>>(field, ncomp, stride, s->send_idx, s->n_send, s->sendbuf1); LAUNCH_CHECK();
  }
  k_reverse_self<<<cdiv(s->n_ghost, 256), 256, 0, st>>>(field, ncomp, stride, s->n_own, s->n_ghost, s->gsrc); LAUNCH_CHECK();
  return 0;
}

// ---- atom migration (send_atoms) ------------------------------------------------------------------------------
// After the sort, the atoms that left the domain sit behind the n_stay owned ones, grouped by direction.
// Exchange the 26 counts, then the records (position+types, momentum+mass, number).
int comm_migrate(imdb200_sim *s, const int *h_counts, long n_stay, long *n_new)
{
  TRY(need_comm(s));
  ncclComm_t c = (ncclComm_t) s->nccl_comm;
  cudaStream_t st = s->stream;
  int *d_cnt = s->sendbufi ? s->sendbufi : nullptr;
  if (!d_cnt) { TRY(ensure_send_capacity(s, 1024)); d_cnt = s->sendbufi; }
  int *h = s->h_starts + 64;                 // [0..26] send counts, [27..53] receive counts
  long n_leave = 0;
  for (int d = 0; d < 27; d++) { h[d] = d == 13 ? 0 : h_counts[d]; n_leave += h[d]; h[27 + d] = 0; }
  CUDA_TRY(cudaMemcpyAsync(d_cnt, h, 54 * sizeof(int), cudaMemcpyHostToDevice, st));
  NCCL_TRY(g_nccl.GroupStart());
  for (int d = 0; d < 27; d++) {
    const DirPlan &D = s->dir[d];
    if (D.peer < 0 || D.peer == s->rank) continue;
    NCCL_TRY(g_nccl.Send(d_cnt + d, 1, ncclInt32, D.peer, c, st));
  }
  for (int d = 26; d >= 0; d--) {
    const DirPlan &D = s->dir[d];
    if (D.peer < 0 || D.peer == s->rank) continue;
    NCCL_TRY(g_nccl.Recv(d_cnt + 27 + d, 1, ncclInt32, D.peer, c, st));
  }
  NCCL_TRY(g_nccl.GroupEnd());
  CUDA_TRY(cudaMemcpyAsync(h + 27, d_cnt + 27, 27 * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  long n_arrive = 0;
  for (int d = 0; d < 27; d++) {
    if (h[d] && (s->dir[d].peer < 0 || s->dir[d].peer == s->rank))
      return imdb_fail(IMDB200_ERR_CELLS, "%d atoms leave towards direction %d where there is no neighbour", h[d], d);
    n_arrive += h[27 + d];
  }
  *n_new = n_stay + n_arrive;
  if (n_leave == 0 && n_arrive == 0) return 0;
  // arrivals land behind the leavers, then slide down over them
  s->n_own = n_stay + n_leave;
  TRY(cells_ensure_capacity(s, n_stay + n_leave + n_arrive + 1));
  NCCL_TRY(g_nccl.GroupStart());
  long off = n_stay;
  for (int d = 0; d < 27; d++) {
    const DirPlan &D = s->dir[d];
    if (h[d] == 0) continue;
    NCCL_TRY(g_nccl.Send(s->pos + off, (size_t) h[d] * 4, ncclFloat64, D.peer, c, st));
    NCCL_TRY(g_nccl.Send(s->mom + off, (size_t) h[d] * 4, ncclFloat64, D.peer, c, st));
    NCCL_TRY(g_nccl.Send(s->nummer + off, (size_t) h[d], ncclInt32, D.peer, c, st));
    off += h[d];
  }
  long roff = n_stay + n_leave;
  for (int d = 26; d >= 0; d--) {
    const DirPlan &D = s->dir[d];
    if (h[27 + d] == 0) continue;
    NCCL_TRY(g_nccl.Recv(s->pos + roff, (size_t) h[27 + d] * 4, ncclFloat64, D.peer, c, st));
    NCCL_TRY(g_nccl.Recv(s->mom + roff, (size_t) h[27 + d] * 4, ncclFloat64, D.peer, c, st));
    NCCL_TRY(g_nccl.Recv(s->nummer + roff, (size_t) h[27 + d], ncclInt32, D.peer, c, st));
    roff += h[27 + d];
  }
  NCCL_TRY(g_nccl.GroupEnd());
  if (n_arrive && n_leave) {
    const long src = n_stay + n_leave;
    CUDA_TRY(cudaMemcpyAsync(s->pos_alt, s->pos + src, n_arrive * sizeof(double4), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s->pos + n_stay, s->pos_alt, n_arrive * sizeof(double4), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s->mom_alt, s->mom + src, n_arrive * sizeof(double4), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s->mom + n_stay, s->mom_alt, n_arrive * sizeof(double4), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s->nummer_alt, s->nummer + src, n_arrive * sizeof(int), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s->nummer + n_stay, s->nummer_alt, n_arrive * sizeof(int), cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

// ---- the MPI_Allreduce sites: every rank gets every rank's scalar block and combines them in rank order --------
__global__ void k_combine_scalars(const double *all, int nranks, double *glob, int *flags)
{
  const int v = threadIdx.x;
  if (v >= SC_COUNT) return;
  double x = all[v];
  for (int r = 1; r < nranks; r++) {
    const double y = all[r * SC_COUNT + v];
    if (v == SC_MAXD2 || v == SC_SHORT) x = fmax(x, y);   // MPI_MAX of check_nblist (src/imd_forces_nbl.c:2032)
    else if (v != SC_ETA) x += y;              // MPI_SUM; eta is identical on every rank
  }
  glob[v] = x;
  if (v == SC_SHORT && x > 0.0) flags[FL_SHORT] = 1;      // a short distance seen by any rank is seen by all
}
__global__ void k_short_to_scal(const int *flags, double *scal) { scal[SC_SHORT] = flags[FL_SHORT] ? 1.0 : 0.0; }

int comm_sync_scalars(imdb200_sim *s)
{
  if (s->nranks == 1) return 0;
  TRY(need_comm(s));
  k_short_to_scal<<<1, 1, 0, s->stream>>>(s->d_flags, s->d_scal); LAUNCH_CHECK();
  NCCL_TRY(g_nccl.AllGather(s->d_scal, s->d_all, SC_COUNT, ncclFloat64, (ncclComm_t) s->nccl_comm, s->stream));
  k_combine_scalars<<<1, 32, 0, s->stream>>>(s->d_all, s->nranks, s->d_glob, s->d_flags); LAUNCH_CHECK();
  return 0;
}

int comm_allgather_ll(imdb200_sim *s, long long mine, long long *total)
{
  if (s->nranks == 1) { *total = mine; return 0; }
  TRY(need_comm(s));
  long long *d = (long long *) s->d_all;       // scratch, nranks*SC_COUNT doubles >= nranks+1 values
  CUDA_TRY(cudaMemcpyAsync(d + s->nranks, &mine, sizeof(mine), cudaMemcpyHostToDevice, s->stream));
  NCCL_TRY(g_nccl.AllGather(d + s->nranks, d, 1, ncclInt64, (ncclComm_t) s->nccl_comm, s->stream));
  std::vector<long long> h(s->nranks);
  CUDA_TRY(cudaMemcpyAsync(h.data(), d, s->nranks * sizeof(long long), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  *total = 0;
  for (long long x : h) *total += x;
  return 0;
}
