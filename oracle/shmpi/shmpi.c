/* oracle/shmpi/shmpi.c -- TEST INFRASTRUCTURE ONLY: the shared-memory MPI subset declared in mpi.h.
 *
 * Written for one purpose: run the unmodified reference's MPI build (`-DMPI`, src/imd_mpi_util.c,
 * src/imd_comm_force_3d.c, src/imd_fix_cells_3d.c) on the host cores of one box without an MPI installation.
 *
 *   SHMPI_NP=<ranks>      number of ranks MPI_Init() creates by fork() (default 1)
 *   SHMPI_RING_KB=<kb>    capacity of each rank-pair ring (default 256); larger messages stream through it
 *   SHMPI_PIN=0           do not pin rank r to the r-th CPU of the inherited affinity mask
 *
 * Semantics kept: non-overtaking order per (source, destination), tag and MPI_ANY_SOURCE / MPI_ANY_TAG
 * matching, unexpected-message buffering, MPI_Get_count, truncation is fatal.  Collectives combine the
 * contributions in rank order on every rank, so all ranks obtain bit-identical results.
 */
#define _GNU_SOURCE
#include "mpi.h"

#include <errno.h>
#include <sched.h>
#include <signal.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/prctl.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#define SLOT_BYTES (1L << 16)
#define MAXP 1024
#define HDR_BYTES 16L

typedef struct {
  _Atomic unsigned long head; /* bytes consumed, written by the receiver */
  char pad1[56];
  _Atomic unsigned long tail; /* bytes produced, written by the sender */
  char pad2[56];
} ring_ctl;

typedef struct {
  int np;
  long ring_bytes;
  _Atomic int bar_count;
  _Atomic int bar_sense;
  _Atomic int aborted;
  _Atomic int finalized;
  pid_t pids[MAXP];
} shm_hdr;

struct shmpi_request {
  int is_send, peer, tag;
  char *buf;
  long nbytes; /* send: message size; recv: capacity */
  long moved;
  int hdr_done, complete;
  MPI_Status st;
  struct shmpi_request *next;
};

typedef struct unexp {
  int src, tag, complete;
  long nbytes;
  char *data;
  struct shmpi_request *claimed_by;
  struct unexp *next;
} unexp;

typedef struct {
  int active, tag;
  long nbytes, got, cap;
  char *dst;
  struct shmpi_request *req;
  unexp *ux;
} incoming;

static shm_hdr *H;
static char *slots;
static ring_ctl *rctl;
static char *rdata;
static int np = 1, me = 0, local_sense = 0, initialised = 0;
static long RING;
static struct shmpi_request **sq_head, **sq_tail; /* per destination */
static struct shmpi_request *rq_head, *rq_tail;   /* posted receives, post order */
static unexp *ux_head, *ux_tail;
static incoming *in;
static int cart_nd, cart_dims[8], cart_per[8];

static void die(const char *msg)
{
  fprintf(stderr, "shmpi[%d]: %s\n", me, msg);
  fflush(stderr);
  MPI_Abort(MPI_COMM_WORLD, 4);
}

static inline ring_ctl *RC(int src, int dst) { return &rctl[(long)src * np + dst]; }
static inline char *RD(int src, int dst) { return rdata + ((long)src * np + dst) * RING; }

static void ring_put(int dst, const char *p, long n)
{
  ring_ctl *c = RC(me, dst);
  char *d = RD(me, dst);
  unsigned long t = atomic_load_explicit(&c->tail, memory_order_relaxed);
  long o = (long)(t % (unsigned long)RING), first = n < RING - o ? n : RING - o;
  memcpy(d + o, p, (size_t)first);
  if (n > first) memcpy(d, p + first, (size_t)(n - first));
  atomic_store_explicit(&c->tail, t + (unsigned long)n, memory_order_release);
}

static void ring_get(int src, char *p, long n)
{
  ring_ctl *c = RC(src, me);
  char *d = RD(src, me);
  unsigned long h = atomic_load_explicit(&c->head, memory_order_relaxed);
  long o = (long)(h % (unsigned long)RING), first = n < RING - o ? n : RING - o;
  memcpy(p, d + o, (size_t)first);
  if (n > first) memcpy(p + first, d, (size_t)(n - first));
  atomic_store_explicit(&c->head, h + (unsigned long)n, memory_order_release);
}

static inline long ring_space(int dst)
{
  ring_ctl *c = RC(me, dst);
  return RING - (long)(atomic_load_explicit(&c->tail, memory_order_relaxed) -
                       atomic_load_explicit(&c->head, memory_order_acquire));
}

static inline long ring_avail(int src)
{
  ring_ctl *c = RC(src, me);
  return (long)(atomic_load_explicit(&c->tail, memory_order_acquire) -
                atomic_load_explicit(&c->head, memory_order_relaxed));
}

/* ---- progress engine ------------------------------------------------------------------------------ */
static int progress_send(int dst)
{
  int moved = 0;
  struct shmpi_request *r;
  while ((r = sq_head[dst]) != NULL) {
    long sp = ring_space(dst);
    if (!r->hdr_done) {
      long hdr[2];
      if (sp < HDR_BYTES) break;
      hdr[0] = r->tag;
      hdr[1] = r->nbytes;
      ring_put(dst, (const char *)hdr, HDR_BYTES);
      r->hdr_done = 1;
      sp -= HDR_BYTES;
      moved = 1;
    }
    if (r->moved < r->nbytes) {
      long n = r->nbytes - r->moved;
      if (n > sp) n = sp;
      if (n > 0) {
        ring_put(dst, r->buf + r->moved, n);
        r->moved += n;
        moved = 1;
      }
    }
    if (r->moved < r->nbytes) break;
    r->complete = 1;
    sq_head[dst] = r->next;
    if (!sq_head[dst]) sq_tail[dst] = NULL;
  }
  return moved;
}

static void finish_recv(struct shmpi_request *r, int src, int tag, long nbytes)
{
  r->st.MPI_SOURCE = src;
  r->st.MPI_TAG = tag;
  r->st.MPI_ERROR = MPI_SUCCESS;
  r->st.shmpi_bytes = nbytes;
  r->complete = 1;
}

static void ux_remove(unexp *u)
{
  unexp **pp = &ux_head, *prev = NULL;
  while (*pp && *pp != u) { prev = *pp; pp = &(*pp)->next; }
  if (*pp) {
    *pp = u->next;
    if (ux_tail == u) ux_tail = prev;
  }
  free(u->data);
  free(u);
}

static int progress_recv(int src)
{
  int moved = 0;
  incoming *m = &in[src];
  for (;;) {
    long av = ring_avail(src);
    if (!m->active) {
      long hdr[2];
      struct shmpi_request *r, *prev = NULL;
      if (av < HDR_BYTES) break;
      ring_get(src, (char *)hdr, HDR_BYTES);
      av -= HDR_BYTES;
      moved = 1;
      m->active = 1;
      m->tag = (int)hdr[0];
      m->nbytes = hdr[1];
      m->got = 0;
      m->req = NULL;
      m->ux = NULL;
      for (r = rq_head; r; prev = r, r = r->next)
        if ((r->peer == src || r->peer == MPI_ANY_SOURCE) && (r->tag == m->tag || r->tag == MPI_ANY_TAG)) break;
      if (r) { /* matched a posted receive: stream straight into the user buffer */
        if (prev) prev->next = r->next; else rq_head = r->next;
        if (rq_tail == r) rq_tail = prev;
        r->next = NULL;
        if (m->nbytes > r->nbytes) die("message truncated (receive buffer too small)");
        m->req = r;
        m->dst = r->buf;
      } else {
        unexp *u = (unexp *)calloc(1, sizeof(unexp));
        u->src = src;
        u->tag = m->tag;
        u->nbytes = m->nbytes;
        u->data = (char *)malloc((size_t)(m->nbytes > 0 ? m->nbytes : 1));
        if (ux_tail) ux_tail->next = u; else ux_head = u;
        ux_tail = u;
        m->ux = u;
        m->dst = u->data;
      }
    }
    if (m->got < m->nbytes) {
      long n = m->nbytes - m->got;
      if (n > av) n = av;
      if (n <= 0) break;
      ring_get(src, m->dst + m->got, n);
      m->got += n;
      moved = 1;
    }
    if (m->got < m->nbytes) break;
    if (m->req) {
      finish_recv(m->req, src, m->tag, m->nbytes);
    } else {
      unexp *u = m->ux;
      u->complete = 1;
      if (u->claimed_by) {
        memcpy(u->claimed_by->buf, u->data, (size_t)u->nbytes);
        finish_recv(u->claimed_by, src, u->tag, u->nbytes);
        ux_remove(u);
      }
    }
    m->active = 0;
  }
  return moved;
}

static int progress_all(void)
{
  int moved = 0, p;
  for (p = 0; p < np; p++) {
    if (sq_head[p]) moved |= progress_send(p);
    moved |= progress_recv(p);
  }
  return moved;
}

static void check_abort(void)
{
  if (atomic_load_explicit(&H->aborted, memory_order_relaxed)) _exit(3);
}

static inline void idle(unsigned *spins)
{
  if (++*spins > 2000) {
    check_abort();
    sched_yield();
    *spins = 1000;
  } else {
    __builtin_ia32_pause();
  }
}

static void barrier(void)
{
  int s = !local_sense;
  unsigned spins = 0;
  local_sense = s;
  if (np == 1) return;
  if (atomic_fetch_add(&H->bar_count, 1) == np - 1) {
    atomic_store(&H->bar_count, 0);
    atomic_store(&H->bar_sense, s);
  } else {
    while (atomic_load_explicit(&H->bar_sense, memory_order_acquire) != s) {
      if (!progress_all()) idle(&spins);
    }
  }
}

static void wait_req(struct shmpi_request *r)
{
  unsigned spins = 0;
  while (!r->complete)
    if (!progress_all()) idle(&spins);
}

/* ---- start-up / shut-down ------------------------------------------------------------------------- */
static void on_sigchld(int sig)
{
  int st, saved = errno;
  pid_t p;
  (void)sig;
  while ((p = waitpid(-1, &st, WNOHANG)) > 0)
    if (H && (WIFSIGNALED(st) || (WIFEXITED(st) && WEXITSTATUS(st) != 0 && !atomic_load(&H->finalized))))
      atomic_store(&H->aborted, 1);
  errno = saved;
}

int MPI_Init(int *argc, char ***argv)
{
  const char *e = getenv("SHMPI_NP");
  long kb = 256;
  size_t total, off_slots, off_ctl, off_data;
  int r;
  (void)argc; (void)argv;
  if (initialised) return MPI_SUCCESS;
  np = e ? atoi(e) : 1;
  if (np < 1 || np > MAXP) { fprintf(stderr, "shmpi: SHMPI_NP out of range\n"); exit(2); }
  if ((e = getenv("SHMPI_RING_KB")) != NULL && atol(e) >= 4) kb = atol(e);
  RING = kb * 1024;
  off_slots = (sizeof(shm_hdr) + 4095) & ~(size_t)4095;
  off_ctl = off_slots + (size_t)np * SLOT_BYTES;
  off_data = (off_ctl + (size_t)np * np * sizeof(ring_ctl) + 4095) & ~(size_t)4095;
  total = off_data + (size_t)np * np * (size_t)RING;
  {
    char *base = (char *)mmap(NULL, total, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (base == MAP_FAILED) { perror("shmpi: mmap"); exit(2); }
    H = (shm_hdr *)base;
    slots = base + off_slots;
    rctl = (ring_ctl *)(base + off_ctl);
    rdata = base + off_data;
  }
  H->np = np;
  H->ring_bytes = RING;
  H->pids[0] = getpid();
  fflush(NULL);
  if (np > 1) {
    struct sigaction sa;
    memset(&sa, 0, sizeof sa);
    sa.sa_handler = on_sigchld;
    sa.sa_flags = SA_RESTART | SA_NOCLDSTOP;
    sigaction(SIGCHLD, &sa, NULL);
  }
  for (r = 1; r < np; r++) {
    pid_t p = fork();
    if (p < 0) { perror("shmpi: fork"); exit(2); }
    if (p == 0) {
      signal(SIGCHLD, SIG_DFL);
      prctl(PR_SET_PDEATHSIG, SIGKILL);
      me = r;
      break;
    }
    H->pids[r] = p;
  }
  e = getenv("SHMPI_PIN");
  if (!(e && e[0] == '0')) {
    cpu_set_t set, one;
    if (sched_getaffinity(0, sizeof set, &set) == 0) {
      int n = CPU_COUNT(&set), want = n > 0 ? me % n : 0, c, k = 0;
      for (c = 0; c < CPU_SETSIZE; c++)
        if (CPU_ISSET(c, &set) && k++ == want) {
          CPU_ZERO(&one);
          CPU_SET(c, &one);
          sched_setaffinity(0, sizeof one, &one);
          break;
        }
    }
  }
  sq_head = (struct shmpi_request **)calloc((size_t)np, sizeof *sq_head);
  sq_tail = (struct shmpi_request **)calloc((size_t)np, sizeof *sq_tail);
  in = (incoming *)calloc((size_t)np, sizeof *in);
  initialised = 1;
  barrier();
  return MPI_SUCCESS;
}

int MPI_Finalize(void)
{
  if (!initialised) return MPI_SUCCESS;
  barrier();
  atomic_store(&H->finalized, 1);
  if (me == 0 && np > 1) {
    int st;
    signal(SIGCHLD, SIG_DFL);
    while (waitpid(-1, &st, 0) > 0 || errno == EINTR) {}
  }
  initialised = 0;
  if (me != 0) {
    fflush(NULL);
    _exit(0); /* children leave here: rank 0 alone returns into the caller's exit path */
  }
  return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm comm, int code)
{
  int r;
  (void)comm;
  fflush(NULL);
  if (H) {
    atomic_store(&H->aborted, 1);
    for (r = 0; r < np; r++)
      if (r != me && H->pids[r] > 0) kill(H->pids[r], SIGTERM);
  }
  _exit(code ? code : 1);
}

int MPI_Comm_size(MPI_Comm comm, int *size) { (void)comm; *size = np; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { (void)comm; *rank = me; return MPI_SUCCESS; }
int MPI_Barrier(MPI_Comm comm) { (void)comm; barrier(); return MPI_SUCCESS; }

double MPI_Wtime(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ---- collectives ---------------------------------------------------------------------------------- */
#define ELSIZE(dt) ((long)((dt) & 0xff))
#define KIND(dt) ((dt) >> 8)

#define COMBINE(T)                                                             \
  do {                                                                         \
    T *a = (T *)acc;                                                           \
    const T *b = (const T *)src;                                               \
    long i;                                                                    \
    switch (op) {                                                              \
      case MPI_SUM: for (i = 0; i < n; i++) a[i] = a[i] + b[i]; break;         \
      case MPI_PROD: for (i = 0; i < n; i++) a[i] = a[i] * b[i]; break;        \
      case MPI_MAX: for (i = 0; i < n; i++) if (b[i] > a[i]) a[i] = b[i]; break; \
      case MPI_MIN: for (i = 0; i < n; i++) if (b[i] < a[i]) a[i] = b[i]; break; \
      default: die("unsupported reduction operation");                         \
    }                                                                          \
  } while (0)

static void combine(void *acc, const void *src, long n, MPI_Datatype dt, MPI_Op op)
{
  switch (KIND(dt)) {
    case 1: COMBINE(char); break;
    case 2: case 9: COMBINE(unsigned char); break;
    case 3: COMBINE(short); break;
    case 4: COMBINE(int); break;
    case 5: case 11: COMBINE(long); break;
    case 6: COMBINE(float); break;
    case 7: COMBINE(double); break;
    case 8: COMBINE(unsigned); break;
    case 10: COMBINE(unsigned long); break;
    default: die("unsupported datatype in reduction");
  }
}

static void reduce_impl(const void *sbuf, void *rbuf, int count, MPI_Datatype dt, MPI_Op op, int root)
{
  long el = ELSIZE(dt), per = SLOT_BYTES / el, off;
  for (off = 0; off < count; off += per) {
    long n = count - off < per ? count - off : per;
    int r;
    memcpy(slots + (long)me * SLOT_BYTES, (const char *)sbuf + off * el, (size_t)(n * el));
    barrier();
    if (root < 0 || root == me) {
      char *acc = (char *)rbuf + off * el;
      memcpy(acc, slots, (size_t)(n * el));
      for (r = 1; r < np; r++) combine(acc, slots + (long)r * SLOT_BYTES, n, dt, op);
    }
    barrier();
  }
}

int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm)
{
  (void)comm;
  reduce_impl(sbuf, rbuf, count, dt, op, -1);
  return MPI_SUCCESS;
}

int MPI_Reduce(const void *sbuf, void *rbuf, int count, MPI_Datatype dt, MPI_Op op, int root, MPI_Comm comm)
{
  (void)comm;
  reduce_impl(sbuf, rbuf, count, dt, op, root);
  return MPI_SUCCESS;
}

int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm comm)
{
  long total = (long)count * ELSIZE(dt), off;
  (void)comm;
  if (np == 1) return MPI_SUCCESS;
  for (off = 0; off < total; off += SLOT_BYTES) {
    long n = total - off < SLOT_BYTES ? total - off : SLOT_BYTES;
    if (me == root) memcpy(slots, (char *)buf + off, (size_t)n);
    barrier();
    if (me != root) memcpy((char *)buf + off, slots, (size_t)n);
    barrier();
  }
  return MPI_SUCCESS;
}

/* ---- point to point ------------------------------------------------------------------------------- */
static struct shmpi_request *new_req(int is_send, void *buf, long nbytes, int peer, int tag)
{
  struct shmpi_request *r = (struct shmpi_request *)calloc(1, sizeof *r);
  r->is_send = is_send;
  r->buf = (char *)buf;
  r->nbytes = nbytes;
  r->peer = peer;
  r->tag = tag;
  return r;
}

int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req)
{
  struct shmpi_request *r;
  (void)comm;
  if (dest < 0 || dest >= np) die("MPI_Isend: bad destination");
  r = new_req(1, (void *)buf, (long)count * ELSIZE(dt), dest, tag);
  if (sq_tail[dest]) sq_tail[dest]->next = r; else sq_head[dest] = r;
  sq_tail[dest] = r;
  progress_send(dest);
  *req = r;
  return MPI_SUCCESS;
}

int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Request *req)
{
  struct shmpi_request *r;
  unexp *u;
  (void)comm;
  if (src != MPI_ANY_SOURCE && (src < 0 || src >= np)) die("MPI_Irecv: bad source");
  r = new_req(0, buf, (long)count * ELSIZE(dt), src, tag);
  *req = r;
  for (u = ux_head; u; u = u->next)
    if (!u->claimed_by && (src == u->src || src == MPI_ANY_SOURCE) && (tag == u->tag || tag == MPI_ANY_TAG)) break;
  if (u) {
    if (u->nbytes > r->nbytes) die("message truncated (receive buffer too small)");
    if (u->complete) {
      memcpy(r->buf, u->data, (size_t)u->nbytes);
      finish_recv(r, u->src, u->tag, u->nbytes);
      ux_remove(u);
    } else {
      u->claimed_by = r;
    }
    return MPI_SUCCESS;
  }
  if (rq_tail) rq_tail->next = r; else rq_head = r;
  rq_tail = r;
  return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request *req, MPI_Status *st)
{
  struct shmpi_request *r = *req;
  if (!r) return MPI_SUCCESS;
  wait_req(r);
  if (st && !r->is_send) *st = r->st;
  free(r);
  *req = MPI_REQUEST_NULL;
  return MPI_SUCCESS;
}

int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *sts)
{
  int i;
  for (i = 0; i < n; i++) MPI_Wait(&reqs[i], sts ? &sts[i] : NULL);
  return MPI_SUCCESS;
}

int MPI_Waitany(int n, MPI_Request *reqs, int *index, MPI_Status *st)
{
  unsigned spins = 0;
  int i, live = 0;
  for (i = 0; i < n; i++) live += reqs[i] != NULL;
  if (!live) { *index = MPI_UNDEFINED; return MPI_SUCCESS; }
  for (;;) {
    for (i = 0; i < n; i++)
      if (reqs[i] && reqs[i]->complete) {
        *index = i;
        return MPI_Wait(&reqs[i], st);
      }
    if (!progress_all()) idle(&spins);
  }
}

int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm)
{
  MPI_Request r;
  MPI_Isend(buf, count, dt, dest, tag, comm, &r);
  return MPI_Wait(&r, NULL);
}

int MPI_Recv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Status *st)
{
  MPI_Request r;
  MPI_Irecv(buf, count, dt, src, tag, comm, &r);
  return MPI_Wait(&r, st);
}

int MPI_Sendrecv(const void *sbuf, int scount, MPI_Datatype sdt, int dest, int stag,
                 void *rbuf, int rcount, MPI_Datatype rdt, int src, int rtag, MPI_Comm comm, MPI_Status *st)
{
  MPI_Request r[2];
  MPI_Irecv(rbuf, rcount, rdt, src, rtag, comm, &r[0]);
  MPI_Isend(sbuf, scount, sdt, dest, stag, comm, &r[1]);
  MPI_Wait(&r[0], st);
  return MPI_Wait(&r[1], NULL);
}

int MPI_Get_count(const MPI_Status *st, MPI_Datatype dt, int *count)
{
  *count = (int)(st->shmpi_bytes / ELSIZE(dt));
  return MPI_SUCCESS;
}

/* ---- Cartesian topology (row-major ranks, like every MPI implementation without reordering) -------------- */
int MPI_Cart_create(MPI_Comm comm, int ndims, const int *dims, const int *periods, int reorder, MPI_Comm *cart)
{
  int i;
  long prod = 1;
  (void)comm; (void)reorder;
  if (ndims > 8) die("MPI_Cart_create: too many dimensions");
  cart_nd = ndims;
  for (i = 0; i < ndims; i++) { cart_dims[i] = dims[i]; cart_per[i] = periods[i]; prod *= dims[i]; }
  if (prod != np) die("MPI_Cart_create: grid size differs from the number of ranks (set SHMPI_NP = product of cpu_dim)");
  *cart = (MPI_Comm)2;
  return MPI_SUCCESS;
}

int MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int *coords)
{
  int i;
  (void)comm;
  for (i = cart_nd - 1; i >= 0; i--) {
    if (i < maxdims) coords[i] = rank % cart_dims[i];
    rank /= cart_dims[i];
  }
  return MPI_SUCCESS;
}

int MPI_Cart_rank(MPI_Comm comm, const int *coords, int *rank)
{
  int i, r = 0;
  (void)comm;
  for (i = 0; i < cart_nd; i++) {
    int c = coords[i];
    if (cart_per[i]) c = ((c % cart_dims[i]) + cart_dims[i]) % cart_dims[i];
    else if (c < 0 || c >= cart_dims[i]) die("MPI_Cart_rank: coordinate outside a non-periodic grid");
    r = r * cart_dims[i] + c;
  }
  *rank = r;
  return MPI_SUCCESS;
}

int MPI_Alloc_mem(MPI_Aint size, MPI_Info info, void *baseptr)
{
  (void)info;
  *(void **)baseptr = malloc((size_t)(size > 0 ? size : 1));
  return *(void **)baseptr ? MPI_SUCCESS : 1;
}

int MPI_Free_mem(void *base) { free(base); return MPI_SUCCESS; }
