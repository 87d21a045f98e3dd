/* pot_table.c -- host-side (plain C, no CUDA) reader for IMD potential tables.
 *
 * Keeps IMD's potential-table API: same two file formats, same post-processing, result in a
 * struct with the layout of pot_table_t (src/types.h:416-428), so that a table read here and a
 * table read by IMD's read_pot_table() (src/imd_potential.c:161-282) are interchangeable inputs
 * of imdb200_set_potentials().
 *
 *   format 1 (src/imd_potential.c:297-376): lines  r2 V_0 V_1 ... V_{ncols-1},  equidistant in r2;
 *            `end` of a column is the r2 of its last non-zero sample.
 *   format 2 (src/imd_potential.c:394-462): ncols lines  begin end step,  then the columns one
 *            after the other, len = (int)(1 + (end-begin)/step + 0.49) values each.
 *   header  "#F <format> <ncols>" ... "#E"; a file without header is read in default_format.
 *   radial tables are shifted so that the last sample is zero; every column gets two extra rows
 *   continuing the last parabola (init_threepoint, src/imd_potential.c:1256-1272).
 */
#include "../../include/imd_b200.h"
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int fail(const char *what, const char *file)
{
  fprintf(stderr, "imd_b200: %s %s\n", what, file);
  return IMDB200_ERR_IO;
}

/* slurp all numbers that follow the header */
static double *read_numbers(FILE *f, long *count)
{
  long cap = 1 << 16, n = 0;
  double *v = (double *) malloc(cap * sizeof(double)), x;
  while (fscanf(f, "%lf", &x) == 1) {
    if (n == cap) { cap *= 2; v = (double *) realloc(v, cap * sizeof(double)); }
    v[n++] = x;
  }
  *count = n;
  return v;
}

static void alloc_info(imdb200_pot_table *pt, int ncols)
{
  pt->ncols = ncols;
  pt->maxsteps = 0;
  pt->begin = (double *) calloc(ncols, sizeof(double));
  pt->end = (double *) calloc(ncols, sizeof(double));
  pt->step = (double *) calloc(ncols, sizeof(double));
  pt->invstep = (double *) calloc(ncols, sizeof(double));
  pt->len = (int *) calloc(ncols, sizeof(int));
  pt->table = NULL;
}

void imdb200_free_pot_table(imdb200_pot_table *pt)
{
  if (!pt) return;
  free(pt->begin); free(pt->end); free(pt->step); free(pt->invstep); free(pt->len); free(pt->table);
  memset(pt, 0, sizeof(*pt));
}

int imdb200_read_pot_table(imdb200_pot_table *pt, const char *filename, int ncols, int radial,
                           int ntypes, int default_format, double *cellsz)
{
  char line[1024];
  int format = default_format, have_header = 0, have_format = 0, done = 0, size = ncols;
  long nnum = 0, pos;
  double *num;
  FILE *f = fopen(filename, "r");
  (void) ntypes;
  if (!f) return fail("Could not open potential file", filename);

  /* header: '#' lines up to "#E" */
  while (!done) {
    pos = ftell(f);
    if (!fgets(line, sizeof(line), f)) { fclose(f); return fail("Unexpected end of file in", filename); }
    if (line[0] == '#') {
      have_header = 1;
      if (line[1] == 'E') done = 1;
      if (line[1] == 'F') {
        if (sscanf(line + 2, "%d %d", &format, &size) != 2) { fclose(f); return fail("Corrupted format header line in file", filename); }
        if (size != ncols) { fclose(f); fprintf(stderr, "Should be %d, is %d\n", ncols, size); return fail("Wrong number of data columns in file", filename); }
        if (format != 1 && format != 2) { fclose(f); return fail("Unrecognized format specified for file", filename); }
        have_format = 1;
      }
    } else if (have_header) { fclose(f); return fail("Corrupted header in file", filename); }
    else { fseek(f, pos, SEEK_SET); done = 1; }
  }
  if (have_header && !have_format) { fclose(f); return fail("Format not specified in header of file", filename); }
  num = read_numbers(f, &nnum);
  fclose(f);

  alloc_info(pt, ncols);
  if (format == 1) {
    const long rows = nnum / (ncols + 1);
    long k; int c;
    if (rows < 3) { free(num); return fail("too few samples in", filename); }
    pt->maxsteps = (int) (((rows + 49) / 50) * 50);          /* grows in PSTEP = 50 blocks (src/config.h:251) */
    pt->table = (double *) calloc((size_t) (pt->maxsteps + 2) * ncols, sizeof(double));
    for (k = 0; k < rows; k++)
      for (c = 0; c < ncols; c++) {
        const double v = num[k * (ncols + 1) + 1 + c];
        pt->table[k * ncols + c] = v;
        if (v != 0.0) { pt->end[c] = num[k * (ncols + 1)]; pt->len[c] = (int) k + 1; }
      }
    {
      const double r2_start = num[0], r2_last = num[(rows - 1) * (ncols + 1)];
      const double r2_step = (r2_last - r2_start) / (rows - 1);
      for (c = 0; c < ncols; c++) {
        const double delta = pt->table[(rows - 1) * ncols + c];
        pt->begin[c] = r2_start; pt->step[c] = r2_step; pt->invstep[c] = 1.0 / r2_step;
        if (radial) {
          if (delta != 0.0) {
            printf("Potential %1d%1d shifted by %e\n", c / ntypes, c % ntypes, delta);
            for (k = 0; k < rows; k++) pt->table[k * ncols + c] -= delta;
          }
          if (cellsz && pt->end[c] > *cellsz) *cellsz = pt->end[c];
        }
      }
    }
  } else {
    long at = 3L * ncols, k; int c;
    if (nnum < at) { free(num); return fail("Info line corrupt in", filename); }
    for (c = 0; c < ncols; c++) {
      double numstep;
      pt->begin[c] = num[3 * c]; pt->end[c] = num[3 * c + 1]; pt->step[c] = num[3 * c + 2];
      if (radial && cellsz && pt->end[c] > *cellsz) *cellsz = pt->end[c];
      pt->invstep[c] = 1.0 / pt->step[c];
      numstep = 1 + (pt->end[c] - pt->begin[c]) / pt->step[c];
      pt->len[c] = (int) (numstep + 0.49);
      if (pt->len[c] > pt->maxsteps) pt->maxsteps = pt->len[c];
      if (numstep - pt->len[c] >= 0.1 || pt->len[c] - numstep >= 0.1)
        fprintf(stderr, "WARNING: numstep = %f rounded to %d in file %s.\n", numstep, pt->len[c], filename);
    }
    pt->table = (double *) calloc((size_t) (pt->maxsteps + 2) * ncols, sizeof(double));
    for (c = 0; c < ncols; c++) {
      if (at + pt->len[c] > nnum) { free(num); return fail("wrong format in file", filename); }
      for (k = 0; k < pt->len[c]; k++) pt->table[k * ncols + c] = num[at + k];
      at += pt->len[c];
    }
    if (radial)
      for (c = 0; c < ncols; c++) {
        const double delta = pt->table[(long) (pt->len[c] - 1) * ncols + c];
        if (delta != 0.0) {
          printf("Potential %1d%1d shifted by %e\n", c / ntypes, c % ntypes, delta);
          for (k = 0; k < pt->len[c]; k++) pt->table[k * ncols + c] -= delta;
        }
      }
  }
  free(num);
  /* two pad rows per column: continue the last interpolation parabola */
  {
    int c;
    for (c = 0; c < ncols; c++) {
      double *y = pt->table + c;
      const long n = pt->len[c], nc = ncols;
      if (n < 3) continue;
      y[n * nc] = 3 * y[(n - 1) * nc] - 3 * y[(n - 2) * nc] + y[(n - 3) * nc];
      y[(n + 1) * nc] = 6 * y[(n - 1) * nc] - 8 * y[(n - 2) * nc] + 3 * y[(n - 3) * nc];
    }
  }
  return 0;
}
