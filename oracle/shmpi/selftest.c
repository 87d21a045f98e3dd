/* oracle/shmpi/selftest.c -- TEST INFRASTRUCTURE ONLY: exercises the MPI subset of shmpi.c the way IMD uses it
 * (tests/test_ref_mpi.py compiles and runs it with SHMPI_NP = 1, 2, 5):
 *   collectives longer than one slot, messages longer than one ring, tag matching out of order (unexpected-message
 *   queue), MPI_ANY_SOURCE / MPI_ANY_TAG with MPI_Get_count, MPI_Waitany, MPI_Sendrecv around a ring, self-sends,
 *   the Cartesian topology calls.  Prints "SHMPI_SELFTEST_OK <np>" from rank 0 and exits 0 when everything holds. */
#include "mpi.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "rank %d: check failed at line %d: %s\n", me, __LINE__, #c); MPI_Abort(MPI_COMM_WORLD, 9); } } while (0)

int main(int argc, char **argv)
{
  int me, np, i, r;
  MPI_Init(&argc, &argv);
  MPI_Comm_rank(MPI_COMM_WORLD, &me);
  MPI_Comm_size(MPI_COMM_WORLD, &np);

  /* ---- collectives, longer than one 64 KB slot ---- */
  {
    const int n = 20000;
    double *a = malloc(n * sizeof(double)), *b = malloc(n * sizeof(double));
    int imax = me * 7 % 5, gmax = -1, want = 0;
    long lsum = me + 1, gl = 0;
    for (i = 0; i < n; i++) a[i] = (me + 1) * 0.5 + i;
    MPI_Allreduce(a, b, n, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
    for (i = 0; i < n; i += 997) CHECK(b[i] == 0.25 * np * (np + 1) + (double) np * i);
    MPI_Allreduce(a, a, n, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD);          /* aliased buffers */
    CHECK(a[n - 1] == np * 0.5 + (n - 1));
    MPI_Allreduce(&imax, &gmax, 1, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
    for (r = 0; r < np; r++) if (r * 7 % 5 > want) want = r * 7 % 5;
    CHECK(gmax == want);
    MPI_Reduce(&lsum, &gl, 1, MPI_LONG, MPI_SUM, np - 1, MPI_COMM_WORLD);
    if (me == np - 1) CHECK(gl == (long) np * (np + 1) / 2);
    for (i = 0; i < n; i++) a[i] = me == np - 1 ? 3.0 * i : -1.0;
    MPI_Bcast(a, n, MPI_DOUBLE, np - 1, MPI_COMM_WORLD);
    CHECK(a[n - 1] == 3.0 * (n - 1) && a[0] == 0.0);
    free(a); free(b);
  }

  /* ---- ring exchange with messages four times the ring size, both directions at once ---- */
  {
    const int n = 1 << 17;                                 /* 1 MB of doubles */
    double *s1 = malloc(n * sizeof(double)), *s2 = malloc(n * sizeof(double));
    double *r1 = malloc(n * sizeof(double)), *r2 = malloc(n * sizeof(double));
    const int up = (me + 1) % np, dn = (me + np - 1) % np;
    MPI_Request q[4];
    MPI_Status st[4];
    int cnt;
    for (i = 0; i < n; i++) { s1[i] = me * 1e6 + i; s2[i] = -(me * 1e6 + i); }
    MPI_Irecv(r1, n, MPI_DOUBLE, dn, 11, MPI_COMM_WORLD, &q[0]);
    MPI_Irecv(r2, n, MPI_DOUBLE, up, 12, MPI_COMM_WORLD, &q[1]);
    MPI_Isend(s1, n, MPI_DOUBLE, up, 11, MPI_COMM_WORLD, &q[2]);
    MPI_Isend(s2, n - 5, MPI_DOUBLE, dn, 12, MPI_COMM_WORLD, &q[3]);
    MPI_Waitall(4, q, st);
    CHECK(r1[n - 1] == dn * 1e6 + (n - 1) && r2[7] == -(up * 1e6 + 7));
    MPI_Get_count(&st[1], MPI_DOUBLE, &cnt);
    CHECK(cnt == n - 5 && st[1].MPI_SOURCE == up && st[1].MPI_TAG == 12);
    /* Sendrecv around the ring */
    MPI_Sendrecv(s1, 1000, MPI_DOUBLE, up, 5, r1, n, MPI_DOUBLE, dn, 5, MPI_COMM_WORLD, &st[0]);
    MPI_Get_count(&st[0], MPI_DOUBLE, &cnt);
    CHECK(cnt == 1000 && r1[999] == dn * 1e6 + 999);
    free(s1); free(s2); free(r1); free(r2);
  }

  /* ---- tags received in the opposite order of sending: the first message waits in the unexpected queue ---- */
  {
    const int up = (me + 1) % np, dn = (me + np - 1) % np;
    int a = 100 + me, b = 200 + me, x = 0, y = 0;
    MPI_Status st;
    MPI_Request q[2];
    MPI_Isend(&a, 1, MPI_INT, up, 1, MPI_COMM_WORLD, &q[0]);
    MPI_Isend(&b, 1, MPI_INT, up, 2, MPI_COMM_WORLD, &q[1]);
    MPI_Recv(&y, 1, MPI_INT, dn, 2, MPI_COMM_WORLD, &st);
    MPI_Recv(&x, 1, MPI_INT, dn, 1, MPI_COMM_WORLD, &st);
    MPI_Waitall(2, q, MPI_STATUSES_IGNORE);
    CHECK(x == 100 + dn && y == 200 + dn);
  }

  /* ---- gather to rank 0 with MPI_ANY_SOURCE / MPI_ANY_TAG, as the writers do (src/imd_io.c:136-147) ---- */
  {
    char buf[64];
    if (me != 0) {
      int len = snprintf(buf, sizeof buf, "hello from %d", me) + 1;
      MPI_Send(buf, len, MPI_CHAR, 0, 40 + me, MPI_COMM_WORLD);
    } else {
      int seen = 0;
      for (r = 1; r < np; r++) {
        MPI_Status st;
        int len, who;
        MPI_Recv(buf, sizeof buf, MPI_CHAR, MPI_ANY_SOURCE, MPI_ANY_TAG, MPI_COMM_WORLD, &st);
        MPI_Get_count(&st, MPI_CHAR, &len);
        who = st.MPI_SOURCE;
        CHECK(st.MPI_TAG == 40 + who && len == (int) strlen(buf) + 1 && atoi(buf + 11) == who);
        seen |= 1 << who;
      }
      CHECK(seen == (1 << np) - 2);
    }
    MPI_Barrier(MPI_COMM_WORLD);     /* MPI_ANY_TAG would also match the traffic of the next section */
  }

  /* ---- Waitany over receives from every other rank, plus a send to oneself ---- */
  {
    int *in = calloc(np, sizeof(int)), self = -1, out = 1000 + me, done = 0, idx;
    MPI_Request *q = calloc(np + 1, sizeof(MPI_Request));
    MPI_Status st;
    for (r = 0; r < np; r++) if (r != me) MPI_Irecv(&in[r], 1, MPI_INT, r, 77, MPI_COMM_WORLD, &q[r]);
    MPI_Irecv(&self, 1, MPI_INT, me, 78, MPI_COMM_WORLD, &q[np]);
    for (r = 0; r < np; r++) if (r != me) MPI_Send(&out, 1, MPI_INT, r, 77, MPI_COMM_WORLD);
    MPI_Send(&out, 1, MPI_INT, me, 78, MPI_COMM_WORLD);
    for (;;) {
      MPI_Waitany(np + 1, q, &idx, &st);
      if (idx == MPI_UNDEFINED) break;
      done++;
      if (idx < np) CHECK(in[idx] == 1000 + idx && st.MPI_SOURCE == idx);
    }
    CHECK(done == np && self == 1000 + me);
    free(in); free(q);
  }

  /* ---- Cartesian topology: row-major ranks, periodic wrap ---- */
  {
    int dims[3] = {1, 1, np}, per[3] = {1, 1, 1}, c[3], rk;
    MPI_Comm cart;
    if (np % 2 == 0) { dims[0] = 2; dims[2] = np / 2; }
    MPI_Cart_create(MPI_COMM_WORLD, 3, dims, per, 1, &cart);
    MPI_Cart_coords(cart, me, 3, c);
    CHECK((c[0] * dims[1] + c[1]) * dims[2] + c[2] == me);
    c[2] -= 1; c[0] += dims[0];
    MPI_Cart_rank(cart, c, &rk);
    MPI_Cart_coords(cart, rk, 3, c);
    MPI_Cart_coords(cart, me, 3, dims);                    /* dims now holds my coordinates */
    CHECK(c[0] == dims[0] && c[1] == dims[1]);
  }

  MPI_Barrier(MPI_COMM_WORLD);
  if (me == 0) printf("SHMPI_SELFTEST_OK %d %.3f\n", np, MPI_Wtime() > 0.0 ? 1.0 : 0.0);
  MPI_Finalize();
  return 0;
}
