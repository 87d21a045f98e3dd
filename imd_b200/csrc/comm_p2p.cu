// comm_p2p.cu -- halo exchange over peer memory instead of NCCL messages: the pack kernel of the
// boundary cells stores the records straight into the neighbour's ghost region over NVLink, completion is signalled
// with stream memory operations (cuStreamWriteValue64 / cuStreamWaitValue64), so no communication kernel competes
// with the persistent force CTAs and no SM spins while it waits.
//
// STATUS: on by default when every rank can map every other rank's memory (IMDB200_HALO_P2P=0 keeps the NCCL messages;
// on every rank or on none: the ranks agree on it at imdb200_comm_init).  Green on 2 B200s against the same reference fixtures as the NCCL path
// (tests/test_multi_gpu.py::test_two_domains_peer_memory_halo, profiles/r2_pytest_mgpu_2.log).
//
// Replaces, for the steps between two list builds: send_cells(copy_cell, pack_cell, unpack_cell) and
// send_cells(copy_dF, pack_dF, unpack_dF) (src/imd_comm_force_3d.c:222-396, 726-778, 1031-1060).  The step with a list
// build keeps the NCCL path (its exchanges of counts and atom numbers are messages anyway).
//
// Buffers (all plain cudaMalloc allocations of this rank, exported with cudaIpcGetMemHandle at every rebuild and
// re-opened by a peer only when the handle changed):
//   ghost_raw   double4[cap]   unshifted positions of the images, filled by the owners' ranks
//   dF          double[cap]    F'(rho) of owners and images
//   flags       u64[2][nranks] flags[kind][r] = number of the last exchange of that kind rank r has completed into us
// Buffer re-use is ordered by the data flow of EAM; pair-only runs keep the NCCL path.
#include "internal.cuh"
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define P2P_MAXRANKS 64

typedef int (*cuStreamOp64_t)(cudaStream_t, unsigned long long, unsigned long long, unsigned int);
static cuStreamOp64_t g_cuWrite64 = nullptr, g_cuWait64 = nullptr;

static int driver_load(void)
{
  if (g_cuWrite64 && g_cuWait64) return 0;
  void *h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return imdb_fail(IMDB200_ERR_COMM, "cannot load libcuda.so.1: %s", dlerror());
  g_cuWrite64 = (cuStreamOp64_t) dlsym(h, "cuStreamWriteValue64_v2");
  g_cuWait64 = (cuStreamOp64_t) dlsym(h, "cuStreamWaitValue64_v2");
  if (!g_cuWrite64) g_cuWrite64 = (cuStreamOp64_t) dlsym(h, "cuStreamWriteValue64");
  if (!g_cuWait64) g_cuWait64 = (cuStreamOp64_t) dlsym(h, "cuStreamWaitValue64");
  if (!g_cuWrite64 || !g_cuWait64) return imdb_fail(IMDB200_ERR_COMM, "libcuda lacks cuStreamWriteValue64 / cuStreamWaitValue64");
  return 0;
}

// what every rank publishes at a rebuild
struct P2PRecord {
  cudaIpcMemHandle_t h_raw, h_dF, h_flags;
  long long n_own;
  int regrown, pad0;                   // this rank re-allocated an exported buffer since the last setup
  int recv_off[P2P_MAXRANKS];          // where rank r's slice starts in MY ghost region, -1: r sends nothing to me
  int pad[2];
};

struct P2PPeer {
  int rank;
  cudaIpcMemHandle_t h_raw, h_dF, h_flags;   // handles currently mapped
  double4 *raw; double *dF; unsigned long long *flags;
  int have;                            // mapped
  long long n_own; int recv_off;       // the peer's numbers for the slice coming from me
};

struct P2PState {
  void *graveyard[16]; int n_grave;    // exported buffers that were re-allocated: freed once every peer has unmapped them
  unsigned long long *flags;           // mine
  P2PRecord *d_rec;                    // [nranks + 1]: all-gather target + my record behind it
  P2PPeer peer[26];
  int npeer;
  unsigned long long step[2];
  int ready;
};

struct P2PTable {                      // per launch: where every slice of the send list goes
  int n;
  int send_off[26], send_cnt[26];
  void *dst[26];
};

template <typename T> __global__ void k_pack_p2p(const T *src, const int *idx, long n_send, P2PTable tab)
{
  const long t = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (t >= n_send) return;
  for (int q = 0; q < tab.n; q++) {
    const long r = t - tab.send_off[q];
    if (r >= 0 && r < tab.send_cnt[q]) { reinterpret_cast<T *>(tab.dst[q])[r] = src[idx[t]]; return; }
  }
}

extern "C" int comm_p2p_enable(imdb200_sim *s)
{
  const char *e = getenv("IMDB200_HALO_P2P");
  s->p2p_on = 0;
  if (s->nranks <= 1) return 0;
  // On by default where every rank can map every other rank's memory (one NVLink / NVSwitch node); IMDB200_HALO_P2P=0
  // keeps the NCCL messages.  Every rank must take the same path (a rank without it would skip the all-gathers of
  // comm_p2p_setup), so the ranks agree first: device ordinals are gathered, peer access is probed, the votes are summed.
  long long want = ((!e || atoi(e) != 0) && s->nranks <= P2P_MAXRANKS && driver_load() == 0) ? 1 : 0, all = 0;
  {
    std::vector<long long> devs(s->nranks, -1);
    for (int r = 0; r < s->nranks; r++) {                // rank r's device ordinal, one value per collective (all ranks take part)
      long long v = 0;
      TRY(comm_allgather_ll(s, s->rank == r ? (long long) s->cfg.device + 1 : 0, &v));
      devs[r] = v - 1;
    }
    for (int r = 0; r < s->nranks && want; r++) {
      if (r == s->rank) continue;
      int can = 0;
      if (devs[r] == s->cfg.device || cudaDeviceCanAccessPeer(&can, s->cfg.device, (int) devs[r]) != cudaSuccess || !can) want = 0;
    }
    cudaGetLastError();
  }
  TRY(comm_allgather_ll(s, want, &all));
  if (all != s->nranks) return 0;
  P2PState *st = (P2PState *) calloc(1, sizeof(P2PState));
  if (!st) return imdb_fail(IMDB200_ERR_COMM, "out of memory");
  CUDA_TRY(cudaMalloc(&st->flags, 2 * P2P_MAXRANKS * sizeof(unsigned long long)));
  CUDA_TRY(cudaMemset(st->flags, 0, 2 * P2P_MAXRANKS * sizeof(unsigned long long)));
  CUDA_TRY(cudaMalloc(&st->d_rec, (size_t) (s->nranks + 1) * sizeof(P2PRecord)));
  s->p2p = st;
  s->p2p_on = 1;
  return 0;
}

// cells.cu re-allocates a per-atom array that peers may have mapped (ghost_raw, dF): the old allocation must outlive
// their mappings (freeing exported memory before the importer closes it is undefined), so it is parked here and
// freed at the end of the next comm_p2p_setup, after every peer has re-mapped.  Returns 1 when it took the pointer.
int comm_p2p_defer_free(imdb200_sim *s, void *ptr)
{
  P2PState *st = (P2PState *) s->p2p;
  if (!s->p2p_on || !st || !ptr || st->n_grave >= 16) return 0;
  st->graveyard[st->n_grave++] = ptr;
  return 1;
}

void comm_p2p_free(imdb200_sim *s)
{
  P2PState *st = (P2PState *) s->p2p;
  if (!st) return;
  for (int i = 0; i < st->n_grave; i++) cudaFree(st->graveyard[i]);
  for (int q = 0; q < 26; q++) {
    P2PPeer &P = st->peer[q];
    if (P.raw) cudaIpcCloseMemHandle(P.raw);
    if (P.dF) cudaIpcCloseMemHandle(P.dF);
    if (P.flags) cudaIpcCloseMemHandle(P.flags);
  }
  if (st->flags) cudaFree(st->flags);
  if (st->d_rec) cudaFree(st->d_rec);
  free(st);
  s->p2p = nullptr; s->p2p_on = 0;
}

static int remap(void **mapped, cudaIpcMemHandle_t *cur, const cudaIpcMemHandle_t &now, int have)
{
  if (have && *mapped && memcmp(cur, &now, sizeof(now)) == 0) return 0;
  if (*mapped) { cudaIpcCloseMemHandle(*mapped); *mapped = nullptr; }
  cudaError_t e = cudaIpcOpenMemHandle(mapped, now, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return imdb_fail(IMDB200_ERR_COMM, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  *cur = now;
  return 0;
}

// At a rebuild, after comm_setup_ghosts: publish handles, atom counts and receive offsets; map what changed.
// `allgather` is comm.cu's byte all-gather over NCCL (collective, stream-ordered).
int comm_p2p_setup(imdb200_sim *s, int (*allgather)(imdb200_sim *, const void *, void *, size_t))
{
  P2PState *st = (P2PState *) s->p2p;
  if (!s->p2p_on || !st) return 0;
  st->ready = 0;
  if (!s->tabs.have_eam) return 0;                        // buffer re-use relies on the F' exchange, see the header
  P2PRecord mine;
  memset(&mine, 0, sizeof(mine));
  CUDA_TRY(cudaIpcGetMemHandle(&mine.h_raw, s->ghost_raw));
  CUDA_TRY(cudaIpcGetMemHandle(&mine.h_dF, s->dF));
  CUDA_TRY(cudaIpcGetMemHandle(&mine.h_flags, st->flags));
  mine.n_own = s->n_own;
  mine.regrown = st->n_grave > 0;
  for (int r = 0; r < P2P_MAXRANKS; r++) mine.recv_off[r] = -1;
  for (int q = 0; q < s->n_peers; q++) mine.recv_off[s->peers[q].peer] = s->peers[q].recv_cnt ? s->peers[q].recv_off : -1;
  P2PRecord *d_mine = st->d_rec + s->nranks;
  CUDA_TRY(cudaMemcpyAsync(d_mine, &mine, sizeof(mine), cudaMemcpyHostToDevice, s->stream));
  TRY(allgather(s, d_mine, st->d_rec, sizeof(P2PRecord)));
  std::vector<P2PRecord> all(s->nranks);
  CUDA_TRY(cudaMemcpyAsync(all.data(), st->d_rec, (size_t) s->nranks * sizeof(P2PRecord), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  st->npeer = s->n_peers;
  for (int q = 0; q < s->n_peers; q++) {
    P2PPeer &P = st->peer[q];
    const int r = s->peers[q].peer;
    if (P.have && P.rank != r) {                          // the process grid changed under us: drop the old mappings
      if (P.raw) cudaIpcCloseMemHandle(P.raw);
      if (P.dF) cudaIpcCloseMemHandle(P.dF);
      if (P.flags) cudaIpcCloseMemHandle(P.flags);
      P.raw = nullptr; P.dF = nullptr; P.flags = nullptr; P.have = 0;
    }
    P.rank = r;
    TRY(remap((void **) &P.raw, &P.h_raw, all[r].h_raw, P.have));
    TRY(remap((void **) &P.dF, &P.h_dF, all[r].h_dF, P.have));
    TRY(remap((void **) &P.flags, &P.h_flags, all[r].h_flags, P.have));
    P.have = 1;
    P.n_own = all[r].n_own;
    P.recv_off = all[r].recv_off[s->rank];
    if (s->peers[q].send_cnt && P.recv_off < 0) return imdb_fail(IMDB200_ERR_COMM, "peer %d does not expect the slice rank %d sends", r, s->rank);
  }
  // a rank re-allocated exported memory: once EVERY rank is past its re-mapping (a second, empty collective is the
  // barrier) the old allocations have no importer left and can go
  int any_regrown = 0;
  for (int r = 0; r < s->nranks; r++) any_regrown |= all[r].regrown;
  if (any_regrown) {
    long long one = 1, sum = 0;
    TRY(comm_allgather_ll(s, one, &sum));
    for (int i = 0; i < st->n_grave; i++) cudaFree(st->graveyard[i]);
    st->n_grave = 0;
  }
  st->ready = 1;
  return 0;
}

// First half of an exchange: store my boundary values (kind 0: positions, 1: F') into every peer's ghost region and
// tell the peer, in stream order, that my slice of exchange number `step` is complete in its memory.
int comm_p2p_begin(imdb200_sim *s, int kind)
{
  P2PState *st = (P2PState *) s->p2p;
  P2PTable tab;
  memset(&tab, 0, sizeof(tab));
  tab.n = st->npeer;
  for (int q = 0; q < st->npeer; q++) {
    const PeerPlan &pl = s->peers[q];
    const P2PPeer &P = st->peer[q];
    tab.send_off[q] = pl.send_off; tab.send_cnt[q] = pl.send_cnt;
    tab.dst[q] = kind == 0 ? (void *) (P.raw + P.recv_off) : (void *) (P.dF + P.n_own + P.recv_off);
  }
  if (s->n_send) {
    const int nb = cdiv(s->n_send, 256);
    if (kind == 0) k_pack_p2p<double4><<<nb, 256, 0, s->stream>>>(s->pos, s->send_idx, s->n_send, tab);
    else k_pack_p2p<double><<<nb, 256, 0, s->stream>>>(s->dF, s->send_idx, s->n_send, tab);
    LAUNCH_CHECK();
  }
  const unsigned long long step = ++st->step[kind];
  for (int q = 0; q < st->npeer; q++) {
    unsigned long long *f = st->peer[q].flags + (size_t) kind * P2P_MAXRANKS + s->rank;
    if (g_cuWrite64(s->stream, (unsigned long long) (uintptr_t) f, step, 0u) != 0)
      return imdb_fail(IMDB200_ERR_COMM, "cuStreamWriteValue64 failed");
  }
  return 0;
}

// Second half: wait, in stream order and without occupying an SM, until every peer has told me the same
// (CU_STREAM_WAIT_VALUE_GEQ = 0).  Whatever is queued between the two halves overlaps the transfer.
int comm_p2p_end(imdb200_sim *s, int kind)
{
  P2PState *st = (P2PState *) s->p2p;
  const unsigned long long step = st->step[kind];
  for (int q = 0; q < st->npeer; q++) {
    unsigned long long *f = st->flags + (size_t) kind * P2P_MAXRANKS + st->peer[q].rank;
    if (g_cuWait64(s->stream, (unsigned long long) (uintptr_t) f, step, 0u) != 0)
      return imdb_fail(IMDB200_ERR_COMM, "cuStreamWaitValue64 failed");
  }
  return 0;
}

static int exchange_p2p(imdb200_sim *s, int kind) { TRY(comm_p2p_begin(s, kind)); return comm_p2p_end(s, kind); }

int comm_p2p_ready(const imdb200_sim *s) { return s->p2p_on && s->p2p && ((const P2PState *) s->p2p)->ready; }
int comm_p2p_positions(imdb200_sim *s) { return exchange_p2p(s, 0); }     // then k_ghost_pos, as after the NCCL exchange
int comm_p2p_dF(imdb200_sim *s) { return exchange_p2p(s, 1); }            // then k_ghost_dF
