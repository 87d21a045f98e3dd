/* topology.c -- host-side (plain C, no CUDA) process-grid helpers kept from IMD's setup layer.
 *
 *   imdb200_calc_cpu_dim    calc_cpu_dim       src/imd_geom_mpi_3d.c:201-266
 *   imdb200_cart_coords/rank  MPI_Cart_coords / MPI_Cart_rank as used by setup_mpi_topology
 *                             (src/imd_geom_mpi_3d.c:43-53): row-major, last index fastest
 *   imdb200_halo_peers      the 26 neighbour ranks of setup_mpi_topology (:57-88, cpu_grid_coord :98-110)
 *                           plus the periodic image shift each neighbour's atoms need on arrival
 *                           (the shift vectors send_cells applies on boundary ranks,
 *                           src/imd_comm_force_3d.c:248-265).
 *   imdb200_halo_message_order  the order in which the 26 halo regions are laid out so that all regions exchanged
 *                           with one neighbour rank are adjacent: one message per peer instead of the reference's
 *                           one per sweep direction (src/imd_comm_force_3d.c:268-395)
 */
#include "../../include/imd_b200.h"
#include <math.h>

int imdb200_cart_rank(const int coord[3], const int cpu_dim[3])
{
  return (coord[0] * cpu_dim[1] + coord[1]) * cpu_dim[2] + coord[2];
}

void imdb200_cart_coords(int rank, const int cpu_dim[3], int coord[3])
{
  coord[2] = rank % cpu_dim[2];
  coord[1] = (rank / cpu_dim[2]) % cpu_dim[1];
  coord[0] = rank / (cpu_dim[2] * cpu_dim[1]);
}

/* Factorise num_cpus evenly; the largest factor goes to the axis with the largest requested cpu_dim. */
void imdb200_calc_cpu_dim(int num_cpus, int cpu_dim[3])
{
  int order[3] = {0, 1, 2}, f[3], n = num_cpus, trial, t;
  /* sort axes by requested size, largest first (stable like the reference's three swaps) */
  if (cpu_dim[order[2]] > cpu_dim[order[1]]) { t = order[1]; order[1] = order[2]; order[2] = t; }
  if (cpu_dim[order[1]] > cpu_dim[order[0]]) { t = order[0]; order[0] = order[1]; order[1] = t; }
  if (cpu_dim[order[2]] > cpu_dim[order[1]]) { t = order[1]; order[1] = order[2]; order[2] = t; }
  trial = (int) ceil(pow((double) n, 1.0 / 3.0));
  for (f[0] = trial; f[0] > 0; f[0]--) if (n % f[0] == 0) break;
  n /= f[0];
  trial = (int) ceil(sqrt((double) n));
  for (f[1] = trial; f[1] > 0; f[1]--) if (n % f[1] == 0) break;
  f[2] = n / f[1];
  if (f[2] > f[1]) { t = f[1]; f[1] = f[2]; f[2] = t; }
  if (f[1] > f[0]) { t = f[0]; f[0] = f[1]; f[1] = t; }
  if (f[2] > f[1]) { t = f[1]; f[1] = f[2]; f[2] = t; }
  cpu_dim[order[0]] = f[0]; cpu_dim[order[1]] = f[1]; cpu_dim[order[2]] = f[2];
}

/* Direction d = (sx+1) + 3*(sy+1) + 9*(sz+1), s in {-1,0,1}.  peer[d]: rank that owns the cells behind
 * face/edge/corner d, -1 where the box ends at a free surface, own rank for a periodic wrap onto
 * itself.  code[d]: image shift (same encoding) to add to that neighbour's positions. */
void imdb200_halo_peers(const int cpu_dim[3], const int my_coord[3], const int pbc_dirs[3], int peer[27],
                        int code[27])
{
  int d, a;
  for (d = 0; d < 27; d++) {
    int sg[3], pc[3], sh[3], ok = 1;
    sg[0] = d % 3 - 1; sg[1] = (d / 3) % 3 - 1; sg[2] = d / 9 - 1;
    peer[d] = -1; code[d] = 13;
    if (d == 13) continue;
    for (a = 0; a < 3; a++) {
      int c = my_coord[a] + sg[a];
      sh[a] = 0;
      if (c < 0 || c >= cpu_dim[a]) {
        if (!pbc_dirs[a]) ok = 0;
        sh[a] = sg[a];
        c = (c + cpu_dim[a]) % cpu_dim[a];
      }
      pc[a] = c;
    }
    if (!ok) continue;
    peer[d] = imdb200_cart_rank(pc, cpu_dim);
    code[d] = (sh[0] + 1) + 3 * (sh[1] + 1) + 9 * (sh[2] + 1);
  }
}

/* Peer-major layout of the halo regions.  recv_order lists the directions whose buffer cells this rank fills
 * (every d with peer[d] >= 0, own rank included), grouped by peer rank ascending and inside a peer by ascending
 * direction.  send_order lists the directions it sends towards (peer[d] >= 0 and != my_rank), grouped the same way
 * but inside a peer by DESCENDING direction: what is sent towards d arrives at the peer as direction 26-d, so the
 * peer's ascending receive order is the sender's descending send order and the slice exchanged between two ranks is
 * contiguous and identically ordered on both sides. */
void imdb200_halo_message_order(const int peer[27], int my_rank, int recv_order[26], int *n_recv,
                                int send_order[26], int *n_send)
{
  int d, i, j, nr = 0, ns = 0;
  for (d = 0; d < 27; d++) {
    if (d == 13 || peer[d] < 0) continue;
    recv_order[nr++] = d;
    if (peer[d] != my_rank) send_order[ns++] = d;
  }
  for (i = 1; i < nr; i++) {                       /* insertion sorts: (peer, d) ascending */
    int v = recv_order[i];
    for (j = i - 1; j >= 0 && (peer[recv_order[j]] > peer[v] || (peer[recv_order[j]] == peer[v] && recv_order[j] > v)); j--)
      recv_order[j + 1] = recv_order[j];
    recv_order[j + 1] = v;
  }
  for (i = 1; i < ns; i++) {                       /* (peer ascending, d descending) */
    int v = send_order[i];
    for (j = i - 1; j >= 0 && (peer[send_order[j]] > peer[v] || (peer[send_order[j]] == peer[v] && send_order[j] < v)); j--)
      send_order[j + 1] = send_order[j];
    send_order[j + 1] = v;
  }
  *n_recv = nr; *n_send = ns;
}
