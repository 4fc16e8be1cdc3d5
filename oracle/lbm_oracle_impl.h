/*
 * lbm_oracle_impl.h -- body of the CPU oracle, included once per real type.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See lbm_oracle.c for the header that
 * explains who may use this.  The including file defines
 *     REAL        float | double
 *     SQRT        sqrtf | sqrt
 *     NAME(x)     x##_f32 | x##_f64
 *     R(x)        literal of type REAL
 *
 * Everything here is a restatement of the *maths* of /root/reference/d2q9-bgk.c
 * (citations are file:line into that file).  The reference unrolls the periodic
 * wrap into nine copies of one cell update (d2q9-bgk.c:262-1810); this file has
 * one cell update and generic wrap arithmetic instead.  The floating-point
 * expression trees (operand order, where the divides are) mirror the reference
 * source so that a build without fast-math (-O2 -ffp-contract=off) gives the same
 * bits as the reference built the same way.
 *
 * Cell layout is the reference's array-of-structs t_speed (d2q9-bgk.c:76-79):
 * cells[(ii + jj*nx)*9 + k], obstacles[ii + jj*nx] as int (d2q9-bgk.c:2797).
 */

/* ---- accelerate_flow: d2q9-bgk.c:1888-1918, fused copy :229-260 -------------- */
void NAME(oracle_accelerate_flow)(int nx, int ny, REAL density, REAL accel,
                                  REAL* cells, const int* obstacles)
{
  const REAL w1 = density * accel / R(9.0);
  const REAL w2 = density * accel / R(36.0);
  const int jj = ny - 2;                                  /* :240 */
  for (int ii = 0; ii < nx; ii++) {
    REAL* c = cells + ((size_t)ii + (size_t)jj * nx) * 9;
    if (!obstacles[(size_t)ii + (size_t)jj * nx]
        && (c[3] - w1) > R(0.0) && (c[6] - w2) > R(0.0) && (c[7] - w2) > R(0.0)) {  /* :246-249 */
      c[1] += w1; c[5] += w2; c[8] += w2;                 /* :252-254 */
      c[3] -= w1; c[6] -= w2; c[7] -= w2;                 /* :256-258 */
    }
  }
}

/* ---- propagate: d2q9-bgk.c:2123-2152 (pull streaming, periodic wrap) --------- */
void NAME(oracle_propagate)(int nx, int ny, const REAL* cells, REAL* tmp_cells)
{
#pragma omp parallel for schedule(static)
  for (int jj = 0; jj < ny; jj++) {
    const int y_n = (jj + 1) % ny;                        /* :2132 */
    const int y_s = (jj == 0) ? (ny - 1) : (jj - 1);      /* :2134 */
    for (int ii = 0; ii < nx; ii++) {
      const int x_e = (ii + 1) % nx;                      /* :2133 */
      const int x_w = (ii == 0) ? (nx - 1) : (ii - 1);    /* :2135 */
      REAL* t = tmp_cells + ((size_t)ii + (size_t)jj * nx) * 9;
#define SRC(x, y, k) cells[((size_t)(x) + (size_t)(y) * nx) * 9 + (k)]
      t[0] = SRC(ii,  jj,  0);                            /* :2139-2147 */
      t[1] = SRC(x_w, jj,  1);
      t[2] = SRC(ii,  y_s, 2);
      t[3] = SRC(x_e, jj,  3);
      t[4] = SRC(ii,  y_n, 4);
      t[5] = SRC(x_w, y_s, 5);
      t[6] = SRC(x_e, y_s, 6);
      t[7] = SRC(x_e, y_n, 7);
      t[8] = SRC(x_w, y_n, 8);
#undef SRC
    }
  }
}

/* one obstacle cell: bounce-back of the pulled values p[] (d2q9-bgk.c:971-981) */
static inline void NAME(cell_rebound)(const REAL* p, REAL* out)
{
  out[0] = p[0];
  out[1] = p[3]; out[2] = p[4]; out[3] = p[1]; out[4] = p[2];
  out[5] = p[7]; out[6] = p[8]; out[7] = p[5]; out[8] = p[6];
}

/* one fluid cell: BGK relaxation of the pulled values p[] (d2q9-bgk.c:983-1100) */
static inline void NAME(cell_collide)(const REAL* p, REAL omega, REAL* out)
{
  const REAL c_sq = R(1.0) / R(3.0);                      /* :984-987 */
  const REAL w0 = R(4.0) / R(9.0);
  const REAL w1 = R(1.0) / R(9.0);
  const REAL w2 = R(1.0) / R(36.0);

  REAL local_density = R(0.0);                            /* :988-998 */
  for (int k = 0; k < 9; k++) local_density += p[k];

  const REAL u_x = (p[1] + p[5] + p[8] - (p[3] + p[6] + p[7])) / local_density;  /* :1002-1008 */
  const REAL u_y = (p[2] + p[5] + p[6] - (p[4] + p[7] + p[8])) / local_density;  /* :1010-1016 */
  const REAL u_sq = u_x * u_x + u_y * u_y;                /* :1019 */

  REAL u[9];                                              /* :1022-1030 */
  u[1] =   u_x;        u[2] =         u_y;
  u[3] = - u_x;        u[4] =       - u_y;
  u[5] =   u_x + u_y;  u[6] = - u_x + u_y;
  u[7] = - u_x - u_y;  u[8] =   u_x - u_y;

  REAL d_equ[9];                                          /* :1033-1062 */
  d_equ[0] = w0 * local_density * (R(1.0) - u_sq / (R(2.0) * c_sq));
  for (int k = 1; k < 9; k++) {
    const REAL w = (k < 5) ? w1 : w2;
    d_equ[k] = w * local_density * (R(1.0) + u[k] / c_sq
                                    + (u[k] * u[k]) / (R(2.0) * c_sq * c_sq)
                                    - u_sq / (R(2.0) * c_sq));
  }
  for (int k = 0; k < 9; k++)                             /* :1066-1100 */
    out[k] = p[k] + omega * (d_equ[k] - p[k]);
}

/* |u| of one cell from its nine speeds (d2q9-bgk.c:1104-1128, :2681-2705, :2948-2972) */
static inline REAL NAME(cell_speed)(const REAL* f, REAL* ux_out, REAL* uy_out, REAL* rho_out)
{
  REAL local_density = R(0.0);
  for (int k = 0; k < 9; k++) local_density += f[k];
  const REAL u_x = (f[1] + f[5] + f[8] - (f[3] + f[6] + f[7])) / local_density;
  const REAL u_y = (f[2] + f[5] + f[6] - (f[4] + f[7] + f[8])) / local_density;
  if (ux_out) *ux_out = u_x;
  if (uy_out) *uy_out = u_y;
  if (rho_out) *rho_out = local_density;
  return SQRT((u_x * u_x) + (u_y * u_y));
}

/* ---- rebound: d2q9-bgk.c:2199-2228.  Called after propagate: swaps the opposite
 * directions of obstacle cells IN PLACE in the propagated grid `tmp_cells` (the
 * reference's `cells` argument is unused there too). --------------------------- */
void NAME(oracle_rebound)(int nx, int ny, REAL* cells, REAL* tmp_cells, const int* obstacles)
{
  (void)cells;
  for (size_t n = 0; n < (size_t)nx * ny; n++)
    if (obstacles[n]) {
      REAL p[9], o[9];
      for (int k = 0; k < 9; k++) p[k] = tmp_cells[n * 9 + k];
      NAME(cell_rebound)(p, o);
      for (int k = 0; k < 9; k++) tmp_cells[n * 9 + k] = o[k];
    }
}

/* ---- collision: d2q9-bgk.c:2554-2663.  Called after propagate/rebound: relaxes the
 * fluid cells IN PLACE in `tmp_cells` (:2646-2651). --------------------------- */
void NAME(oracle_collision)(int nx, int ny, REAL omega, REAL* cells, REAL* tmp_cells,
                            const int* obstacles)
{
  (void)cells;
#pragma omp parallel for schedule(static)
  for (long n = 0; n < (long)nx * ny; n++)
    if (!obstacles[n]) {
      REAL p[9], o[9];
      for (int k = 0; k < 9; k++) p[k] = tmp_cells[(size_t)n * 9 + k];
      NAME(cell_collide)(p, omega, o);
      for (int k = 0; k < 9; k++) tmp_cells[(size_t)n * 9 + k] = o[k];
    }
}

/* ---- av_velocity: d2q9-bgk.c:2665-2714.  Serial REAL accumulator in row-major
 * order like the reference; *tot_u_f64 (optional) gets a double accumulation of
 * the same per-cell values, which is what large grids must be compared with
 * (the fp32 serial sum loses all accuracy by 16384^2). ------ */
REAL NAME(oracle_av_velocity)(int nx, int ny, const REAL* cells, const int* obstacles,
                              double* tot_u_f64, long* tot_cells_out)
{
  long tot_cells = 0;
  REAL tot_u = R(0.0);
  double tot_d = 0.0;
  for (size_t n = 0; n < (size_t)nx * ny; n++) {
    if (!obstacles[n]) {
      const REAL s = NAME(cell_speed)(cells + n * 9, 0, 0, 0);
      tot_u += s;
      tot_d += (double)s;
      ++tot_cells;
    }
  }
  if (tot_u_f64) *tot_u_f64 = tot_d;
  if (tot_cells_out) *tot_cells_out = tot_cells;
  return tot_u / (REAL)tot_cells;
}

/* ---- the fused live step: timestep_new2, d2q9-bgk.c:228-1813 -----------------
 * accelerate row ny-2 of `cells` in place, then for every cell pull + (rebound |
 * collide) into `tmp_cells`, then the step's average velocity from the values
 * just stored.  `speed_scratch` (nx*ny REALs, may be NULL => allocated here) holds
 * the per-cell |u| so the row loop can run in parallel while the accumulation
 * stays in the reference's serial row-major order.                            */
REAL NAME(oracle_timestep)(int nx, int ny, REAL density, REAL accel, REAL omega,
                           REAL* cells, REAL* tmp_cells, const int* obstacles,
                           REAL* speed_scratch, double* tot_u_f64)
{
  NAME(oracle_accelerate_flow)(nx, ny, density, accel, cells, obstacles);   /* :229-260 */

  REAL* scratch = speed_scratch ? speed_scratch : (REAL*)malloc(sizeof(REAL) * (size_t)nx * ny);

#pragma omp parallel for schedule(static)
  for (int jj = 0; jj < ny; jj++) {
    const int y_n = (jj + 1) % ny;
    const int y_s = (jj == 0) ? (ny - 1) : (jj - 1);
    for (int ii = 0; ii < nx; ii++) {
      const int x_e = (ii + 1) % nx;
      const int x_w = (ii == 0) ? (nx - 1) : (ii - 1);
      const size_t n = (size_t)ii + (size_t)jj * nx;
      REAL p[9];
#define SRC(x, y, k) cells[((size_t)(x) + (size_t)(y) * nx) * 9 + (k)]
      p[0] = SRC(ii,  jj,  0);                            /* :990-998 */
      p[1] = SRC(x_w, jj,  1);
      p[2] = SRC(ii,  y_s, 2);
      p[3] = SRC(x_e, jj,  3);
      p[4] = SRC(ii,  y_n, 4);
      p[5] = SRC(x_w, y_s, 5);
      p[6] = SRC(x_e, y_s, 6);
      p[7] = SRC(x_e, y_n, 7);
      p[8] = SRC(x_w, y_n, 8);
#undef SRC
      REAL* out = tmp_cells + n * 9;
      if (obstacles[n]) {
        NAME(cell_rebound)(p, out);                       /* :971-981 */
        scratch[n] = R(0.0);
      } else {
        NAME(cell_collide)(p, omega, out);                /* :983-1100 */
        scratch[n] = NAME(cell_speed)(out, 0, 0, 0);      /* :1104-1128 */
      }
    }
  }

  long tot_cells = 0;
  REAL tot_u = R(0.0);
  double tot_d = 0.0;
  for (size_t n = 0; n < (size_t)nx * ny; n++) {
    if (!obstacles[n]) {
      tot_u += scratch[n];                                /* :1128-1130 */
      tot_d += (double)scratch[n];
      ++tot_cells;
    }
  }
  if (!speed_scratch) free(scratch);
  if (tot_u_f64) *tot_u_f64 = tot_d;
  return tot_u / (REAL)tot_cells;                         /* :1811 */
}

/* ---- the un-fused semantic step: timestep_old order, d2q9-bgk.c:1824-1831, with
 * av_velocity taken on the grid that holds the result (`tmp_cells`, which main swaps
 * into `cells` at :190).  Same result grid as the fused step. ------------------ */
REAL NAME(oracle_timestep_unfused)(int nx, int ny, REAL density, REAL accel, REAL omega,
                                   REAL* cells, REAL* tmp_cells, const int* obstacles)
{
  NAME(oracle_accelerate_flow)(nx, ny, density, accel, cells, obstacles);
  NAME(oracle_propagate)(nx, ny, cells, tmp_cells);
  NAME(oracle_rebound)(nx, ny, cells, tmp_cells, obstacles);
  NAME(oracle_collision)(nx, ny, omega, cells, tmp_cells, obstacles);
  return NAME(oracle_av_velocity)(nx, ny, tmp_cells, obstacles, 0, 0);
}

/* ---- rest-state initialisation: d2q9-bgk.c:2802-2823 ------------------------- */
void NAME(oracle_init_cells)(int nx, int ny, REAL density, REAL* cells)
{
  const REAL w0 = density * R(4.0) / R(9.0);
  const REAL w1 = density / R(9.0);
  const REAL w2 = density / R(36.0);
  for (size_t n = 0; n < (size_t)nx * ny; n++) {
    REAL* c = cells + n * 9;
    c[0] = w0;
    c[1] = c[2] = c[3] = c[4] = w1;
    c[5] = c[6] = c[7] = c[8] = w2;
  }
}

/* ---- the step loop of main: d2q9-bgk.c:180-201.  On return the newest state is
 * in `cells` if iters is even, in `tmp_cells` if odd (pointer swap :190); the
 * function returns which (0 = cells, 1 = tmp_cells).  av_vels_f64 (optional)
 * receives tot_u_f64 / tot_cells per step. ------------------------------------ */
int NAME(oracle_run)(int nx, int ny, int iters, REAL density, REAL accel, REAL omega,
                     REAL* cells, REAL* tmp_cells, const int* obstacles,
                     REAL* av_vels, double* av_vels_f64)
{
  REAL* scratch = (REAL*)malloc(sizeof(REAL) * (size_t)nx * ny);
  long free_cells = 0;
  for (size_t n = 0; n < (size_t)nx * ny; n++) free_cells += !obstacles[n];
  REAL* a = cells;
  REAL* b = tmp_cells;
  for (int tt = 0; tt < iters; tt++) {
    double tot_d = 0.0;
    const REAL av = NAME(oracle_timestep)(nx, ny, density, accel, omega, a, b, obstacles, scratch, &tot_d);
    if (av_vels) av_vels[tt] = av;
    if (av_vels_f64) av_vels_f64[tt] = tot_d / (double)free_cells;
    REAL* t = a; a = b; b = t;
  }
  free(scratch);
  return (a == cells) ? 0 : 1;
}

/* ---- final_state fields: write_values, d2q9-bgk.c:2935-2976 ------------------
 * Obstacle cells: u_x = u_y = u = 0, pressure = density * c_sq (:2940-2944). */
void NAME(oracle_final_state)(int nx, int ny, REAL density, const REAL* cells, const int* obstacles,
                              REAL* u_x, REAL* u_y, REAL* u, REAL* pressure)
{
  const REAL c_sq = R(1.0) / R(3.0);
  for (size_t n = 0; n < (size_t)nx * ny; n++) {
    if (obstacles[n]) {
      u_x[n] = u_y[n] = u[n] = R(0.0);
      pressure[n] = density * c_sq;
    } else {
      REAL ux, uy, rho;
      u[n] = NAME(cell_speed)(cells + n * 9, &ux, &uy, &rho);
      u_x[n] = ux; u_y[n] = uy;
      pressure[n] = rho * c_sq;
    }
  }
}

/* ---- calc_reynolds: d2q9-bgk.c:2893-2898 -------------------------------------- */
REAL NAME(oracle_calc_reynolds)(int nx, int ny, REAL omega, int reynolds_dim,
                                const REAL* cells, const int* obstacles)
{
  const REAL viscosity = R(1.0) / R(6.0) * (R(2.0) / omega - R(1.0));
  return NAME(oracle_av_velocity)(nx, ny, cells, obstacles, 0, 0) * reynolds_dim / viscosity;
}

/* ---- total_density: d2q9-bgk.c:2900-2916 (mass conservation check) ----------- */
double NAME(oracle_total_density)(int nx, int ny, const REAL* cells)
{
  double total = 0.0;
  for (size_t n = 0; n < (size_t)nx * ny * 9; n++) total += (double)cells[n];
  return total;
}
