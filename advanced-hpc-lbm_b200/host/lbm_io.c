/* lbm_io.c -- parsers and writers of the d2q9-bgk file contract (see lbm_io.h). */
#define _POSIX_C_SOURCE 200809L
#include "lbm_io.h"

#include <ctype.h>
#include <errno.h>
#include <math.h>
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

/* ---------------------------------------------------------------------------------
 * Fast, exact "%.12E".  The output files are millions of lines of four such numbers;
 * once the step loop runs on the GPU, fprintf's general-purpose conversion is what the
 * program spends its time in.  v = m * 2^e exactly, so
 * v * 10^p = m * 5^p * 2^(e+p) is an integer shift away from a 128-bit product; rounding
 * half-to-even on that exact value gives the same 13 digits glibc prints.  Values outside
 * the range the 128-bit product covers fall back to snprintf.
 * tests/test_host_format.py compares it with printf on millions of values.
 * --------------------------------------------------------------------------------- */
typedef unsigned __int128 u128;

static u128 pow5_table[56];
static int pow5_ready = 0;

static void pow5_init(void)
{
  pow5_table[0] = 1;
  for (int i = 1; i < 56; i++) pow5_table[i] = pow5_table[i - 1] * 5;
  pow5_ready = 1;
}

int lbm_format_e12(char* out, double v)
{
  union { double d; uint64_t u; } x;
  x.d = v;
  const int neg = (int)(x.u >> 63);
  const uint64_t frac = x.u & ((1ULL << 52) - 1);
  const int be = (int)((x.u >> 52) & 0x7ff);
  if (be == 0x7ff) return sprintf(out, "%.12E", v);
  char* o = out;
  if (neg) *o++ = '-';
  if (be == 0 && frac == 0) {
    memcpy(o, "0.000000000000E+00", 18);
    o += 18;
    *o = '\0';
    return (int)(o - out);
  }
  if (!pow5_ready) pow5_init();
  uint64_t m = (be == 0) ? frac : (frac | (1ULL << 52));
  int e = (be == 0) ? -1074 : be - 1075;
  const int tz = __builtin_ctzll(m);
  m >>= tz;
  e += tz;
  const int bl = 64 - __builtin_clzll(m);
  int k = (int)floor((double)(e + bl - 1) * 0.30102999566398120);
  const uint64_t lo_lim = 1000000000000ULL, hi_lim = 10000000000000ULL;
  uint64_t n = 0;
  for (int tries = 0; ; tries++) {
    const int p = 12 - k;
    if (tries > 3 || p < 0 || p > 55) return sprintf(out, "%.12E", v);
    /* bits of m * 5^p: bl + ceil(p * log2(5)) */
    const int need = bl + (int)(p * 2.3219280948873623) + 1;
    if (need > 126) return sprintf(out, "%.12E", v);
    const u128 a = (u128)m * pow5_table[p];
    const int sh = e + p;
    u128 q;
    if (sh >= 0) {
      if (need + sh > 126) return sprintf(out, "%.12E", v);
      q = a << sh;
    } else {
      const int s = -sh;
      if (s >= 127) return sprintf(out, "%.12E", v);
      q = a >> s;
      const u128 rem = a & (((u128)1 << s) - 1);
      const u128 half = (u128)1 << (s - 1);
      if (rem > half || (rem == half && (q & 1))) q++;
    }
    if (q < lo_lim) { k--; continue; }
    if (q >= hi_lim) { k++; continue; }
    n = (uint64_t)q;
    break;
  }
  char digits[13];
  for (int i = 12; i >= 0; i--) { digits[i] = (char)('0' + n % 10); n /= 10; }
  *o++ = digits[0];
  *o++ = '.';
  memcpy(o, digits + 1, 12);
  o += 12;
  *o++ = 'E';
  int ex = k;
  if (ex < 0) { *o++ = '-'; ex = -ex; } else { *o++ = '+'; }
  if (ex >= 100) { *o++ = (char)('0' + ex / 100); ex %= 100; }
  *o++ = (char)('0' + ex / 10);
  *o++ = (char)('0' + ex % 10);
  *o = '\0';
  return (int)(o - out);
}

static char* put_int(char* o, long long v)
{
  char tmp[24];
  int n = 0;
  if (v < 0) { *o++ = '-'; v = -v; }
  do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (n) *o++ = tmp[--n];
  return o;
}

/* formats rows [r_begin, r_end) of a block into buf; returns the number of bytes */
static size_t format_rows(char* buf, int nx, long long row0, long long r_begin, long long r_end,
                          const double* u_x, const double* u_y, const double* u, const double* pressure,
                          const uint32_t* obstacle_bits)
{
  char* o = buf;
  for (long long r = r_begin; r < r_end; r++) {
    const long long jj = row0 + r;
    for (int ii = 0; ii < nx; ii++) {
      const size_t n = (size_t)r * (size_t)nx + (size_t)ii;
      /* "%d %d %.12E %.12E %.12E %.12E %d\n" (d2q9-bgk.c:2978).  Last column: the cell's
       * own obstacle flag.  The reference prints obstacles[ii*nx + jj], a transposed index
       * that differs from the golden files in check/ for non-symmetric masks; the golden
       * files hold the cell's own flag, which is what is written here (check.py ignores
       * the column). */
      o = put_int(o, ii); *o++ = ' ';
      o = put_int(o, jj); *o++ = ' ';
      o += lbm_format_e12(o, u_x[n]); *o++ = ' ';
      o += lbm_format_e12(o, u_y[n]); *o++ = ' ';
      o += lbm_format_e12(o, u[n]); *o++ = ' ';
      o += lbm_format_e12(o, pressure[n]); *o++ = ' ';
      *o++ = (char)('0' + lbm_obstacle_bit(obstacle_bits, nx, ii, (int)jj));
      *o++ = '\n';
    }
  }
  return (size_t)(o - buf);
}

/* The block is cut into pieces that are formatted in parallel (OpenMP, if the host was
 * built with it) and written in order: the text is identical for any thread count. */
void lbm_write_final_state_rows(void* fpv, int nx, long long row0, long long nrows,
                                const double* u_x, const double* u_y, const double* u,
                                const double* pressure, const uint32_t* obstacle_bits)
{
  FILE* fp = (FILE*)fpv;
  enum { LINE_MAX_BYTES = 128 };          /* 2 ints (<= 11 chars) + 4 x 19 chars + flag + separators */
  if (!pow5_ready) pow5_init();           /* before the threads start */
  long long pieces = nrows < 64 ? nrows : 64;
  if (pieces < 1) return;
  char** bufs = (char**)calloc((size_t)pieces, sizeof(char*));
  size_t* lens = (size_t*)calloc((size_t)pieces, sizeof(size_t));
  if (bufs == NULL || lens == NULL) die("cannot allocate memory for output rows", __LINE__, __FILE__);
  int failed = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(|:failed)
  for (long long p = 0; p < pieces; p++) {
    const long long rb = nrows * p / pieces, re = nrows * (p + 1) / pieces;
    bufs[p] = (char*)malloc((size_t)(re - rb) * (size_t)nx * LINE_MAX_BYTES + 1);
    if (bufs[p] == NULL) { failed |= 1; continue; }
    lens[p] = format_rows(bufs[p], nx, row0, rb, re, u_x, u_y, u, pressure, obstacle_bits);
  }
  if (failed) die("cannot allocate memory for output rows", __LINE__, __FILE__);
  for (long long p = 0; p < pieces; p++) {
    if (lens[p] && fwrite(bufs[p], 1, lens[p], fp) != lens[p]) die("could not write output file", __LINE__, __FILE__);
    free(bufs[p]);
  }
  free(bufs);
  free(lens);
}

void lbm_write_av_vels(const char* path, int n, const double* av_vels)
{
  FILE* fp = fopen(path, "w");
  if (fp == NULL) die("could not open file output file", __LINE__, __FILE__);
  static char big[1 << 20];
  setvbuf(fp, big, _IOFBF, sizeof big);
  char line[64];
  for (int ii = 0; ii < n; ii++) {
    /* "%d:\t%.12E\n" (d2q9-bgk.c:2993) */
    char* o = put_int(line, ii);
    *o++ = ':'; *o++ = '\t';
    o += lbm_format_e12(o, av_vels[ii]);
    *o++ = '\n';
    fwrite(line, 1, (size_t)(o - line), fp);
  }
  fclose(fp);
}
