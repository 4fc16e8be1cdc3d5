/* lbm_io.c -- parsers and writers of the d2q9-bgk file contract (see lbm_io.h). */
#define _POSIX_C_SOURCE 200809L
#include "lbm_io.h"

#include <ctype.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void die(const char* message, const int line, const char* file)
{
  fprintf(stderr, "Error at line %d of file %s:\n", line, file);
  fprintf(stderr, "%s\n", message);
  fflush(stderr);
  exit(EXIT_FAILURE);
}

static int read_int(FILE* fp, int* out)
{
  return fscanf(fp, "%d\n", out) == 1;
}

/* one real value, kept as text so it can be converted the way %f and %lf would */
static int read_real(FILE* fp, float* f, double* d)
{
  char tok[128];
  if (fscanf(fp, "%127s", tok) != 1) return 0;
  char* end = NULL;
  errno = 0;
  *d = strtod(tok, &end);
  if (end == tok) return 0;
  *f = strtof(tok, NULL);
  return 1;
}

void lbm_read_params(const char* paramfile, lbm_param* pf, lbm_param_f64* pd)
{
  char message[1024];
  FILE* fp = fopen(paramfile, "r");
  if (fp == NULL) {
    snprintf(message, sizeof message, "could not open input parameter file: %s", paramfile);
    die(message, __LINE__, __FILE__);
  }
  if (!read_int(fp, &pf->nx)) die("could not read param file: nx", __LINE__, __FILE__);
  if (!read_int(fp, &pf->ny)) die("could not read param file: ny", __LINE__, __FILE__);
  if (!read_int(fp, &pf->maxIters)) die("could not read param file: maxIters", __LINE__, __FILE__);
  if (!read_int(fp, &pf->reynolds_dim)) die("could not read param file: reynolds_dim", __LINE__, __FILE__);
  if (!read_real(fp, &pf->density, &pd->density)) die("could not read param file: density", __LINE__, __FILE__);
  if (!read_real(fp, &pf->accel, &pd->accel)) die("could not read param file: accel", __LINE__, __FILE__);
  if (!read_real(fp, &pf->omega, &pd->omega)) die("could not read param file: omega", __LINE__, __FILE__);
  fclose(fp);
  pd->nx = pf->nx; pd->ny = pf->ny; pd->maxIters = pf->maxIters; pd->reynolds_dim = pf->reynolds_dim;
}

/* Obstacle list -> bit mask.  The reference reads "%d %d %d\n" triples until EOF and
 * dies on a short triple, an out-of-range coordinate or a third value other than 1
 * (d2q9-bgk.c:2844-2853).  Synthetic grids have millions of lines, so the file is read
 * in blocks and tokenised by hand instead of one fscanf per line. */
uint32_t* lbm_read_obstacle_bits(const char* obstaclefile, int nx, int ny)
{
  char message[1024];
  FILE* fp = fopen(obstaclefile, "r");
  if (fp == NULL) {
    snprintf(message, sizeof message, "could not open input obstacles file: %s", obstaclefile);
    die(message, __LINE__, __FILE__);
  }
  const size_t wpr = (size_t)((nx + 31) / 32);
  uint32_t* bits = (uint32_t*)calloc(wpr * (size_t)ny, sizeof(uint32_t));
  if (bits == NULL) die("cannot allocate column memory for obstacles", __LINE__, __FILE__);

  enum { BUF = 1 << 20 };
  char* buf = (char*)malloc(BUF);
  if (buf == NULL) die("cannot allocate column memory for obstacles", __LINE__, __FILE__);
  long long vals[3];
  int nvals = 0;          /* values of the current triple already parsed */
  int in_tok = 0, neg = 0, digits = 0;
  long long cur = 0;
  size_t got;
  int eof = 0;
  while (!eof) {
    got = fread(buf, 1, BUF, fp);
    if (got < BUF) { eof = 1; buf[got++] = '\n'; }   /* sentinel ends a last token */
    for (size_t i = 0; i < got; i++) {
      const char c = buf[i];
      if (c >= '0' && c <= '9') {
        if (!in_tok) { in_tok = 1; neg = 0; cur = 0; digits = 0; }
        if (cur < (1LL << 40)) cur = cur * 10 + (c - '0');
        digits++;
      } else if ((c == '-' || c == '+') && !in_tok) {
        in_tok = 1; neg = (c == '-'); cur = 0; digits = 0;
      } else if (isspace((unsigned char)c)) {
        if (in_tok) {
          if (digits == 0) die("expected 3 values per line in obstacle file", __LINE__, __FILE__);
          vals[nvals++] = neg ? -cur : cur;
          in_tok = 0;
          if (nvals == 3) {
            if (vals[0] < 0 || vals[0] > nx - 1) die("obstacle x-coord out of range", __LINE__, __FILE__);
            if (vals[1] < 0 || vals[1] > ny - 1) die("obstacle y-coord out of range", __LINE__, __FILE__);
            if (vals[2] != 1) die("obstacle blocked value should be 1", __LINE__, __FILE__);
            bits[(size_t)vals[1] * wpr + (size_t)(vals[0] >> 5)] |= 1u << (vals[0] & 31);
            nvals = 0;
          }
        }
      } else {
        die("expected 3 values per line in obstacle file", __LINE__, __FILE__);
      }
    }
  }
  if (nvals != 0) die("expected 3 values per line in obstacle file", __LINE__, __FILE__);
  free(buf);
  fclose(fp);
  return bits;
}

void lbm_write_final_state_rows(void* fpv, int nx, long long row0, long long nrows,
                                const double* u_x, const double* u_y, const double* u,
                                const double* pressure, const uint32_t* obstacle_bits)
{
  FILE* fp = (FILE*)fpv;
  for (long long r = 0; r < nrows; r++) {
    const long long jj = row0 + r;
    for (int ii = 0; ii < nx; ii++) {
      const size_t n = (size_t)r * (size_t)nx + (size_t)ii;
      /* last column: the cell's own obstacle flag.  The reference prints
       * obstacles[ii*nx + jj] (d2q9-bgk.c:2978), a transposed index that differs from
       * the golden files in check/ for non-symmetric masks; the golden files hold the
       * cell's own flag, which is what is written here (check.py ignores the column). */
      fprintf(fp, "%d %lld %.12E %.12E %.12E %.12E %d\n", ii, jj, u_x[n], u_y[n], u[n], pressure[n],
              lbm_obstacle_bit(obstacle_bits, nx, ii, (int)jj));
    }
  }
}

void lbm_write_av_vels(const char* path, int n, const double* av_vels)
{
  FILE* fp = fopen(path, "w");
  if (fp == NULL) die("could not open file output file", __LINE__, __FILE__);
  static char big[1 << 20];
  setvbuf(fp, big, _IOFBF, sizeof big);
  for (int ii = 0; ii < n; ii++) fprintf(fp, "%d:\t%.12E\n", ii, av_vels[ii]);
  fclose(fp);
}
