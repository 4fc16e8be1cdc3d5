/* lbm_io.h -- file contract of the reference CLI, host side (plain C99).
 *
 * Input:  7-value param file (d2q9-bgk.c:2736-2762), "x y 1" obstacle file
 *         (d2q9-bgk.c:2844-2857).
 * Output: final_state.dat "%d %d %.12E %.12E %.12E %.12E %d\n" (d2q9-bgk.c:2978) and
 *         av_vels.dat "%d:\t%.12E\n" (d2q9-bgk.c:2993).
 * Errors: die() prints "Error at line %d of file %s:\n%s\n" to stderr and exits with
 *         EXIT_FAILURE (d2q9-bgk.c:3001-3007).
 */
#ifndef LBM_IO_H
#define LBM_IO_H

#include <stddef.h>
#include <stdint.h>
#include "lbm_gpu.h"

void die(const char* message, const int line, const char* file);

/* reads the 7 values; fills both the float (%f) and the double (%lf) view */
void lbm_read_params(const char* paramfile, lbm_param* pf, lbm_param_f64* pd);

/* parses the obstacle list straight into the LBM_GPU_OBST_BITS layout
 * (((nx+31)/32) uint32 words per row); returns a malloc'ed array */
uint32_t* lbm_read_obstacle_bits(const char* obstaclefile, int nx, int ny);

static inline int lbm_obstacle_bit(const uint32_t* bits, int nx, int ii, int jj)
{
  return (int)((bits[(size_t)jj * (size_t)((nx + 31) / 32) + (size_t)(ii >> 5)] >> (ii & 31)) & 1u);
}

/* appends rows [row0,row0+nrows) of final_state.dat; fields are dense nx-wide rows */
void lbm_write_final_state_rows(void* fp, int nx, long long row0, long long nrows,
                                const double* u_x, const double* u_y, const double* u,
                                const double* pressure, const uint32_t* obstacle_bits);

void lbm_write_av_vels(const char* path, int n, const double* av_vels);

/* exact equivalent of sprintf(out, "%.12E", v); returns the number of characters */
int lbm_format_e12(char* out, double v);

#endif
