// lbm_kernels.cuh -- sm_100a device code of the d2q9-bgk timestep loop.
//
// One fused kernel per timestep does what the reference's timestep_new2 does
// (d2q9-bgk.c:228-1813): accelerate_flow (:229-260), pull-propagate (:990-998),
// rebound (:971-981), BGK collision (:983-1100) and the av_velocity contribution
// (:1103-1130), plus -- new here -- the halo push into the neighbouring slabs.
//
// Data layout in HBM (one "slab" = the rows one GPU holds):
//   lattice  2 buffers x 9 planes x rows x pitch   structure-of-arrays;
//   window   2 parities x 2 directions x 3 planes x pitch: the halo rows.  "from below"
//            holds speeds 2,5,6 of the row under the slab's first row, "from above"
//            speeds 4,7,8 of the row over its last row (periodic in y, so with one slab
//            they are the slab's own opposite edge rows).  Written by the NEIGHBOURS'
//            step kernels (peer stores over NVLink), read by this slab's edge rows.  It
//            is the only memory other GPUs / processes map, together with the flags;
//   mask     rows x pitch/32 uint32, bit = 1 for an obstacle cell;
//   side     2 x 6 x pitch: row ny-2 of planes 1,3,5,6,7,8 AFTER accelerate_flow.
//            The lattice itself always holds the un-accelerated state; readers that
//            pull from row ny-2 take those six planes from `side` instead.  This is
//            the same arithmetic as accelerating in place before streaming
//            (d2q9-bgk.c:229-260) without a separate kernel or a pre-pass.
#pragma once
#include <cuda.h>            // CUtensorMap (type only; the encoder is fetched at run time)
#include <cuda_runtime.h>
#include <stdint.h>

// Compile-time tuning knobs of the step kernels (defaults = the measured best, see
// profiles/r01_kernel_variants.md; tools/build_variants.py builds the alternatives).
#ifndef LBM_BLOCK_THREADS
#define LBM_BLOCK_THREADS 128       // threads per block of K1a/K1b/K5
#endif
#ifndef LBM_MIN_BLOCKS
#define LBM_MIN_BLOCKS 7            // K1a fp32: 7 blocks of 128 threads per SM = 72 registers, no spills,
#endif                              // 896 threads/SM: +8.7 % over 3 x 256 threads at 80 registers
#ifndef LBM_LOAD_MODE
#define LBM_LOAD_MODE 0             // 0 plain, 1 ld.global.cs (evict-first), 2 ld.global.nc, 3 nc + L1::no_allocate
#endif
#ifndef LBM_PERSIST_MIN_BLOCKS
#define LBM_PERSIST_MIN_BLOCKS 6    // K5: keeps it at <= 85 registers so 6 blocks/SM are resident
#endif
#ifndef LBM_AV_MODE
#define LBM_AV_MODE 0               // 0 block reduction + one atomic per block; 1 = NO av sums (experiment only)
#endif
#ifndef LBM_K5_EXPERIMENT
#define LBM_K5_EXPERIMENT 0         // timing experiments only: 1 = no grid barrier (wrong results), 2 = barrier only
#endif
#ifndef LBM_APPROX_MODE
#define LBM_APPROX_MODE 1           // default (non-strict) fp32 build: 0 = IEEE 1/x and sqrt; 1 = rcp.approx / sqrt.approx
#endif                              // for |u| of the av sum only (never feeds back into the lattice): -12 % instructions,
                                    // +4-7 % on L2-resident grids; 2 = also the collision's 1/rho (measured: no further gain)
#ifndef LBM_STORE_MODE
#define LBM_STORE_MODE 0            // 0 plain, 1 st.global.cs (streaming), 2 st.global.cg
#endif

namespace lbm {

// ------------------------------------------------------------------------------------
// arithmetic policy: STRICT mirrors the reference's C expression trees with
// round-to-nearest single operations (never contracted into FMA) so the result is
// bit-identical to a gcc -O2 -ffp-contract=off build; the default lets the compiler
// contract and uses an algebraically equal, cheaper form of the equilibrium.
// ------------------------------------------------------------------------------------
template <typename real, bool STRICT> struct Ops;

// 1/x and sqrt of the default build; the approximate forms are an experiment knob
__device__ __forceinline__ float fast_rcp(float x, int level) {
#if LBM_APPROX_MODE >= 1
  if (LBM_APPROX_MODE >= level) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#endif
  (void)level;
  return 1.0f / x;
}
__device__ __forceinline__ double fast_rcp(double x, int) { return 1.0 / x; }
__device__ __forceinline__ float fast_sqrt(float x) {
#if LBM_APPROX_MODE >= 1
  float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
  return sqrtf(x);
#endif
}
__device__ __forceinline__ double fast_sqrt(double x) { return ::sqrt(x); }

template <> struct Ops<float, true> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
};
template <> struct Ops<double, true> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
};
template <> struct Ops<float, false> {
  static __device__ __forceinline__ float add(float a, float b) { return a + b; }
  static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
  static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
  static __device__ __forceinline__ float div(float a, float b) { return a / b; }
  static __device__ __forceinline__ float sqrt(float a) { return sqrtf(a); }
};
template <> struct Ops<double, false> {
  static __device__ __forceinline__ double add(double a, double b) { return a + b; }
  static __device__ __forceinline__ double sub(double a, double b) { return a - b; }
  static __device__ __forceinline__ double mul(double a, double b) { return a * b; }
  static __device__ __forceinline__ double div(double a, double b) { return a / b; }
  static __device__ __forceinline__ double sqrt(double a) { return ::sqrt(a); }
};

// |u| is accumulated as an exact 128-bit fixed-point sum (unit 2^-52): integer adds
// are associative, so the per-step average does not depend on block scheduling, grid
// shape or on how the rows are split over GPUs.
#define LBM_FIX_SCALE 4503599627370496.0 /* 2^52 */
// A cell whose |u| is NaN or beyond any physical value (the lattice has blown up) cannot
// be represented in the fixed-point sum: the step's high word gets this bit and the host
// reports that step's average as NaN, which is what the reference's float sum would give.
#define LBM_SPEED_LIMIT 2048.0
#define LBM_NONFINITE_MARK (1ULL << 63)
__device__ __forceinline__ unsigned long long to_fixed(float s) {
  return __float2ull_rn(s * 4503599627370496.0f);
}
__device__ __forceinline__ unsigned long long to_fixed(double s) {
  return __double2ull_rn(s * 4503599627370496.0);
}

// ------------------------------------------------------------------------------------
// one cell: p[0..8] pulled values -> o[0..8] stored values, returns |u| of the stored
// values (0 for an obstacle cell).
// ------------------------------------------------------------------------------------
template <typename real, bool STRICT>
__device__ __forceinline__ real cell_update(const real (&p)[9], const bool obstacle, const real omega,
                                            real (&o)[9]) {
  typedef Ops<real, STRICT> M;
  real c[9];
  real speed;
  if (STRICT) {
    // d2q9-bgk.c:983-1100, expression trees as written there
    const real c_sq = (real)1 / (real)3;
    const real w0 = (real)4 / (real)9;
    const real w1 = (real)1 / (real)9;
    const real w2 = (real)1 / (real)36;
    const real two_csq = (real)2 * c_sq;
    const real two_csq_csq = (real)2 * c_sq * c_sq;
    real rho = M::add((real)0, p[0]);
#pragma unroll
    for (int k = 1; k < 9; k++) rho = M::add(rho, p[k]);
    const real ux = M::div(M::sub(M::add(M::add(p[1], p[5]), p[8]), M::add(M::add(p[3], p[6]), p[7])), rho);
    const real uy = M::div(M::sub(M::add(M::add(p[2], p[5]), p[6]), M::add(M::add(p[4], p[7]), p[8])), rho);
    const real usq = M::add(M::mul(ux, ux), M::mul(uy, uy));
    real u[9];
    u[1] = ux;                 u[2] = uy;
    u[3] = -ux;                u[4] = -uy;
    u[5] = M::add(ux, uy);     u[6] = M::add(-ux, uy);
    u[7] = M::sub(-ux, uy);    u[8] = M::sub(ux, uy);
    const real t_usq = M::div(usq, two_csq);
    real d[9];
    d[0] = M::mul(M::mul(w0, rho), M::sub((real)1, t_usq));
#pragma unroll
    for (int k = 1; k < 9; k++) {
      const real w = (k < 5) ? w1 : w2;
      const real poly = M::sub(M::add(M::add((real)1, M::div(u[k], c_sq)),
                                      M::div(M::mul(u[k], u[k]), two_csq_csq)), t_usq);
      d[k] = M::mul(M::mul(w, rho), poly);
    }
#pragma unroll
    for (int k = 0; k < 9; k++) c[k] = M::add(p[k], M::mul(omega, M::sub(d[k], p[k])));
    // d2q9-bgk.c:1103-1128: velocity recomputed from the values just stored
    real rho2 = M::add((real)0, c[0]);
#pragma unroll
    for (int k = 1; k < 9; k++) rho2 = M::add(rho2, c[k]);
    const real vx = M::div(M::sub(M::add(M::add(c[1], c[5]), c[8]), M::add(M::add(c[3], c[6]), c[7])), rho2);
    const real vy = M::div(M::sub(M::add(M::add(c[2], c[5]), c[6]), M::add(M::add(c[4], c[7]), c[8])), rho2);
    speed = M::sqrt(M::add(M::mul(vx, vx), M::mul(vy, vy)));
  } else {
    // same maths, cheaper form: one reciprocal, 1/c_sq = 3, 1/(2 c_sq^2) = 4.5,
    // 1/(2 c_sq) = 1.5; the compiler is free to contract into FMA.
    const real w0 = (real)(4.0 / 9.0), w1 = (real)(1.0 / 9.0), w2 = (real)(1.0 / 36.0);
    const real e = (p[1] + p[5]) + p[8];
    const real w = (p[3] + p[6]) + p[7];
    const real n = (p[2] + p[5]) + p[6];
    const real s = (p[4] + p[7]) + p[8];
    const real rho = ((p[0] + p[2]) + (p[4] + e)) + w;
    const real inv = fast_rcp(rho, 2);
    const real ux = (e - w) * inv;
    const real uy = (n - s) * inv;
    const real base = (real)1 - (real)1.5 * (ux * ux + uy * uy);
    const real r0 = omega * w0 * rho, r1 = omega * w1 * rho, r2 = omega * w2 * rho;
    const real keep = (real)1 - omega;
    const real upv = ux + uy, umv = ux - uy;
#define LBM_EQ(uk) (base + (uk) * ((real)3 + (real)4.5 * (uk)))
    c[0] = keep * p[0] + r0 * base;
    c[1] = keep * p[1] + r1 * LBM_EQ(ux);
    c[2] = keep * p[2] + r1 * LBM_EQ(uy);
    c[3] = keep * p[3] + r1 * LBM_EQ(-ux);
    c[4] = keep * p[4] + r1 * LBM_EQ(-uy);
    c[5] = keep * p[5] + r2 * LBM_EQ(upv);
    c[6] = keep * p[6] + r2 * LBM_EQ(-umv);
    c[7] = keep * p[7] + r2 * LBM_EQ(-upv);
    c[8] = keep * p[8] + r2 * LBM_EQ(umv);
#undef LBM_EQ
    const real e2 = (c[1] + c[5]) + c[8];
    const real w_2 = (c[3] + c[6]) + c[7];
    const real n2 = (c[2] + c[5]) + c[6];
    const real s2 = (c[4] + c[7]) + c[8];
    const real rho2 = ((c[0] + c[2]) + (c[4] + e2)) + w_2;
    const real inv2 = fast_rcp(rho2, 1);
    const real vx = (e2 - w_2) * inv2;
    const real vy = (n2 - s2) * inv2;
    speed = fast_sqrt(vx * vx + vy * vy);
  }
  // obstacle: bounce-back of the pulled values (d2q9-bgk.c:971-981), no average
  o[0] = obstacle ? p[0] : c[0];
  o[1] = obstacle ? p[3] : c[1];
  o[2] = obstacle ? p[4] : c[2];
  o[3] = obstacle ? p[1] : c[3];
  o[4] = obstacle ? p[2] : c[4];
  o[5] = obstacle ? p[7] : c[5];
  o[6] = obstacle ? p[8] : c[6];
  o[7] = obstacle ? p[5] : c[7];
  o[8] = obstacle ? p[6] : c[8];
  return obstacle ? (real)0 : speed;
}

// accelerate_flow on one cell's speeds (d2q9-bgk.c:246-258); a[] = {f1,f3,f5,f6,f7,f8}.
template <typename real, bool STRICT>
__device__ __forceinline__ void cell_accelerate(real& f1, real& f3, real& f5, real& f6, real& f7, real& f8,
                                                const bool obstacle, const real aw1, const real aw2) {
  typedef Ops<real, true> M;   // single adds/subs: nothing to contract, always exact ops
  const bool go = !obstacle && M::sub(f3, aw1) > (real)0 && M::sub(f6, aw2) > (real)0 &&
                  M::sub(f7, aw2) > (real)0;
  if (go) {
    f1 = M::add(f1, aw1); f5 = M::add(f5, aw2); f8 = M::add(f8, aw2);
    f3 = M::sub(f3, aw1); f6 = M::sub(f6, aw2); f7 = M::sub(f7, aw2);
  }
}

// |u| etc. of a stored cell (d2q9-bgk.c:2681-2705 / :2948-2972), reference tree order.
template <typename real>
__device__ __forceinline__ real cell_macroscopic(const real (&f)[9], real& ux, real& uy, real& rho) {
  typedef Ops<real, true> M;
  rho = M::add((real)0, f[0]);
#pragma unroll
  for (int k = 1; k < 9; k++) rho = M::add(rho, f[k]);
  ux = M::div(M::sub(M::add(M::add(f[1], f[5]), f[8]), M::add(M::add(f[3], f[6]), f[7])), rho);
  uy = M::div(M::sub(M::add(M::add(f[2], f[5]), f[6]), M::add(M::add(f[4], f[7]), f[8])), rho);
  return M::sqrt(M::add(M::mul(ux, ux), M::mul(uy, uy)));
}

// ------------------------------------------------------------------------------------
// kernel arguments
// ------------------------------------------------------------------------------------
template <typename real>
struct StepArgs {
  const real* src;             // lattice buffer read this step (plane 0, local row 0)
  real* dst;                   // lattice buffer written this step
  const real* side_src;        // accelerated row ny-2, planes {1,3,5,6,7,8} x pitch (read)
  real* side_dst;              // same, written for the next step
  const uint32_t* mask;        // rows x mask_pitch words
  unsigned long long* av;      // this step's |u| sums: LBM_AV_SLOTS lines of {low halves, high halves, pad}
  // halo rows: 3 x pitch each, see "window" above
  const real* halo_s;          // own window, src parity: speeds {2,5,6} of the row below row 0
  const real* halo_n;          // own window, src parity: speeds {4,7,8} of the row above row rows-1
  real* push_up;               // neighbour above's window, dst parity, "from below": {2,5,6}
  real* push_dn;               // neighbour below's window, dst parity, "from above": {4,7,8}
  // cross-slab ordering (only when MULTI)
  volatile unsigned long long* flag_from_below;  // local: steps completed by neighbour below
  volatile unsigned long long* flag_from_above;
  unsigned long long* up_flag;                   // neighbour above's flag_from_below
  unsigned long long* dn_flag;                   // neighbour below's flag_from_above
  unsigned long long* boundary_done;             // local counter of finished boundary blocks
  unsigned long long step;                       // global index of this step (0-based)
  long long plane_stride;      // rows * pitch
  int nx;
  int rows;                    // local rows
  int pitch;                   // elements per row, multiple of 32
  int mask_pitch;              // words per mask row
  int accel_row;               // local row holding global row ny-2, or LBM_NO_ROW
  int tiles_x, tiles_y;
  real omega;
  real aw1, aw2;               // density*accel/9, density*accel/36 (d2q9-bgk.c:230-231)
};

#define LBM_NO_ROW (-1000)

template <typename real> struct alignas(4 * sizeof(real)) Vec4 { real x, y, z, w; };

__device__ __forceinline__ Vec4<float> ld4(const float* p) {
#if LBM_LOAD_MODE == 1
  const float4 t = __ldcs(reinterpret_cast<const float4*>(p));
#elif LBM_LOAD_MODE == 2
  const float4 t = __ldg(reinterpret_cast<const float4*>(p));
#elif LBM_LOAD_MODE == 3
  float4 t;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "l"(p));
#else
  const float4 t = *reinterpret_cast<const float4*>(p);
#endif
  Vec4<float> v; v.x = t.x; v.y = t.y; v.z = t.z; v.w = t.w;
  return v;
}
__device__ __forceinline__ Vec4<double> ld4(const double* p) { return *reinterpret_cast<const Vec4<double>*>(p); }
__device__ __forceinline__ void st4(float* p, const Vec4<float>& v) {
#if LBM_STORE_MODE == 1
  __stcs(reinterpret_cast<float4*>(p), make_float4(v.x, v.y, v.z, v.w));
#elif LBM_STORE_MODE == 2
  __stcg(reinterpret_cast<float4*>(p), make_float4(v.x, v.y, v.z, v.w));
#else
  *reinterpret_cast<Vec4<float>*>(p) = v;
#endif
}
__device__ __forceinline__ void st4(double* p, const Vec4<double>& v) { *reinterpret_cast<Vec4<double>*>(p) = v; }

// L2-only loads (ld.global.cg): the persistent kernel re-reads, step after step, memory
// that other SMs wrote in the previous step, so nothing may be served from a stale L1 line.
__device__ __forceinline__ Vec4<float> ld4cg(const float* p) {
  const float4 t = __ldcg(reinterpret_cast<const float4*>(p));
  Vec4<float> v; v.x = t.x; v.y = t.y; v.z = t.z; v.w = t.w;
  return v;
}
__device__ __forceinline__ Vec4<double> ld4cg(const double* p) {
  const double2 a = __ldcg(reinterpret_cast<const double2*>(p));
  const double2 b = __ldcg(reinterpret_cast<const double2*>(p) + 1);
  Vec4<double> v; v.x = a.x; v.y = a.y; v.z = b.x; v.w = b.y;
  return v;
}
template <typename real, bool CG>
__device__ __forceinline__ Vec4<real> load4(const real* p) { return CG ? ld4cg(p) : ld4(p); }
template <typename real, bool CG>
__device__ __forceinline__ real load1(const real* p) { return CG ? __ldcg(p) : *p; }

__device__ __forceinline__ unsigned long long ld_acquire_sys(const volatile unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Exact sum of the per-thread fixed-point |u| (unit 2^-52).  The sum of a warp (shuffle
// tree) or of a block is added to the step's accumulators as two fire-and-forget atomics
// (RED): the low and the high 32-bit half go to two 64-bit words, so that
// total = (sum of high halves << 32) + sum of low halves with no carry to propagate and no
// return value to wait for.  LBM_AV_SLOTS such pairs per step, each in its own 128-byte
// line: one L2 line takes only about 0.47 G atomics/s (measured).
//   block_accumulate  one pair of atomics per block (shared memory + __syncthreads): the
//                     per-step kernels, where a 16384^2 step has 2.1 M warps -- one pair
//                     per WARP was measured 8 % slower there (L2 atomic traffic);
//   warp_accumulate   one pair per warp, no barrier: the persistent kernel, few warps and
//                     a grid barrier right behind (128^2: 3.0 -> 2.6 us per step).
#define LBM_AV_SLOTS 8
#define LBM_AV_STRIDE 16                 /* words between slots: one 128-byte line each */

__device__ __forceinline__ void av_add(unsigned long long q, unsigned long long* av_step, const unsigned slot) {
  unsigned long long* p = av_step + LBM_AV_STRIDE * (slot & (LBM_AV_SLOTS - 1));
  atomicAdd(p, q & 0xffffffffULL);
  atomicAdd(p + 1, q >> 32);
}

__device__ __forceinline__ void warp_accumulate(unsigned long long q, unsigned long long* av_step) {
#if LBM_AV_MODE == 1
  if (q == 0xffffffffffffffffULL) *av_step = q;   // keeps q alive, never true
  return;
#endif
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) q += __shfl_down_sync(0xffffffffu, q, off);
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if ((tid & 31) == 0 && q != 0ULL) av_add(q, av_step, blockIdx.x + (tid >> 5));
}

__device__ __forceinline__ void block_accumulate(unsigned long long q, unsigned long long* av_step) {
#if LBM_AV_MODE == 1
  if (q == 0xffffffffffffffffULL) *av_step = q;   // keeps q alive, never true
  return;
#endif
  __shared__ unsigned long long warp_sums[32];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) q += __shfl_down_sync(0xffffffffu, q, off);
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nwarps = (blockDim.x * blockDim.y + 31) >> 5;
  if ((tid & 31) == 0) warp_sums[tid >> 5] = q;
  __syncthreads();
  if (tid < 32) {
    unsigned long long v = (tid < nwarps) ? warp_sums[tid] : 0ULL;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (tid == 0 && v != 0ULL) av_add(v, av_step, blockIdx.x);
  }
}

// After a run: fold the LBM_AV_SLOTS padded slots of every step into one {low, high} pair,
// so that 16 bytes per step cross PCIe instead of 1 KiB.  The non-finite mark (bit 63 of a
// high word) is OR-ed, everything else summed.
__global__ void lbm_compact_av(const unsigned long long* __restrict__ av, unsigned long long* __restrict__ out,
                               int n_steps) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_steps) return;
  const unsigned long long* p = av + (size_t)t * (LBM_AV_STRIDE * LBM_AV_SLOTS);
  unsigned long long lo = 0ULL, hi = 0ULL, mark = 0ULL;
#pragma unroll
  for (int k = 0; k < LBM_AV_SLOTS; k++) {
    lo += p[LBM_AV_STRIDE * k];
    const unsigned long long h = p[LBM_AV_STRIDE * k + 1];
    hi += h & ~LBM_NONFINITE_MARK;
    mark |= h & LBM_NONFINITE_MARK;
  }
  out[2 * t] = lo;
  out[2 * t + 1] = hi | mark;
}

// Block -> tile mapping.  The row tiles that touch the slab's first and last row are
// given the lowest block indices so that they are dispatched first: their halo pushes
// leave early and the neighbours' next step never waits for them.
__device__ __forceinline__ void tile_of_block(const int tiles_x, const int tiles_y, int& tx, int& ty) {
  const unsigned b = blockIdx.x;
  const unsigned slot = b / (unsigned)tiles_x;
  tx = (int)(b - slot * (unsigned)tiles_x);
  if (slot == 0) ty = 0;
  else if (slot == 1) ty = tiles_y - 1;
  else ty = (int)slot - 1;
}

template <typename real, bool MULTI>
__device__ __forceinline__ void boundary_wait(const StepArgs<real>& a, const bool is_boundary) {
  if (MULTI && is_boundary) {
    if (threadIdx.x == 0 && threadIdx.y == 0) {
      while (ld_acquire_sys(a.flag_from_below) < a.step) { }
      while (ld_acquire_sys(a.flag_from_above) < a.step) { }
    }
    __syncthreads();
  }
}

template <typename real, bool MULTI>
__device__ __forceinline__ void boundary_signal(const StepArgs<real>& a, const bool is_boundary) {
  if (MULTI && is_boundary) {
    __threadfence_system();            // this thread's halo pushes are visible system-wide
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
      const unsigned long long nb = (unsigned long long)a.tiles_x * (a.tiles_y > 1 ? 2ULL : 1ULL);
      const unsigned long long old = atomicAdd(a.boundary_done, 1ULL);
      if (old + 1ULL == nb * (a.step + 1ULL)) {
        __threadfence_system();
        st_release_sys(a.up_flag, a.step + 1ULL);
        st_release_sys(a.dn_flag, a.step + 1ULL);
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// K1a  lbm_step_vec4: four cells per thread, 128-bit loads/stores straight from/to the
// SoA planes (any nx: rows are padded to the pitch, a multiple of 32 elements).  x-shifted planes come from the aligned vector plus one
// element of the neighbouring lane (warp shuffle); the two edge lanes of a warp fetch
// that element with a scalar load that also implements the periodic wrap in x.
// blockDim = (BX, BY), BX a multiple of 32 so that a warp never spans two rows.
// ------------------------------------------------------------------------------------
template <typename real, bool STRICT, bool CG>
__device__ __forceinline__ unsigned long long vec4_tile(const StepArgs<real>& a, const int tx, const int ty) {
  const int x0 = (tx * (int)blockDim.x + (int)threadIdx.x) * 4;
  const int r = ty * (int)blockDim.y + (int)threadIdx.y;        // local row
  const bool active = (x0 < a.nx) && (r < a.rows);
  const int lane = threadIdx.x & 31;
  unsigned long long q = 0ULL;

  // clamp so that inactive threads still form valid addresses (they take part in the
  // shuffles but never store)
  const int xc = active ? x0 : 0;
  const int rc = active ? r : 0;
  const long long PS = a.plane_stride;
  const long long oC = (long long)rc * a.pitch, oS = oC - a.pitch, oN = oC + a.pitch;

  // Row sources.  South/north neighbours of the slab's first/last row live in the halo
  // window; planes 1,3 of the centre row / 5,6 of the south row / 7,8 of the north row
  // come from the accelerated side row when that row is global row ny-2 (the window
  // already holds accelerated values).  All of this is warp-uniform.
  const bool first = (rc == 0), last = (rc == a.rows - 1);
  const bool cA = (rc == a.accel_row), sA = (rc - 1 == a.accel_row), nA = (rc + 1 == a.accel_row);
  const real* p0 = a.src + oC;
  const real* p1 = cA ? a.side_src + 0 * a.pitch : a.src + 1 * PS + oC;
  const real* p3 = cA ? a.side_src + 1 * a.pitch : a.src + 3 * PS + oC;
  const real* p2 = first ? a.halo_s + 0 * a.pitch : a.src + 2 * PS + oS;
  const real* p5 = first ? a.halo_s + 1 * a.pitch : (sA ? a.side_src + 2 * a.pitch : a.src + 5 * PS + oS);
  const real* p6 = first ? a.halo_s + 2 * a.pitch : (sA ? a.side_src + 3 * a.pitch : a.src + 6 * PS + oS);
  const real* p4 = last ? a.halo_n + 0 * a.pitch : a.src + 4 * PS + oN;
  const real* p7 = last ? a.halo_n + 1 * a.pitch : (nA ? a.side_src + 4 * a.pitch : a.src + 7 * PS + oN);
  const real* p8 = last ? a.halo_n + 2 * a.pitch : (nA ? a.side_src + 5 * a.pitch : a.src + 8 * PS + oN);

  // edge elements first (scalar, predicated), then the nine aligned vectors
  const bool need_w = (lane == 0) || (xc == 0);
  const bool need_e = (lane == 31) || (xc + 4 >= a.nx);
  const int xw = (xc == 0) ? a.nx - 1 : xc - 1;
  const int xe = (xc + 4 >= a.nx) ? 0 : xc + 4;
  real w1e = 0, w5e = 0, w8e = 0, e3e = 0, e6e = 0, e7e = 0;
  if (need_w) { w1e = load1<real, CG>(p1 + xw); w5e = load1<real, CG>(p5 + xw); w8e = load1<real, CG>(p8 + xw); }
  if (need_e) { e3e = load1<real, CG>(p3 + xe); e6e = load1<real, CG>(p6 + xe); e7e = load1<real, CG>(p7 + xe); }

  const Vec4<real> v0 = load4<real, CG>(p0 + xc), v1 = load4<real, CG>(p1 + xc), v2 = load4<real, CG>(p2 + xc),
                   v3 = load4<real, CG>(p3 + xc), v4 = load4<real, CG>(p4 + xc), v5 = load4<real, CG>(p5 + xc),
                   v6 = load4<real, CG>(p6 + xc), v7 = load4<real, CG>(p7 + xc), v8 = load4<real, CG>(p8 + xc);
  const uint32_t mword = a.mask[(long long)rc * a.mask_pitch + (xc >> 5)];
  const uint32_t mbits = (mword >> (xc & 31)) & 0xFu;

  // element x0-1 of planes 1,5,8 and x0+4 of planes 3,6,7 from the neighbouring lanes
  real l1 = __shfl_up_sync(0xffffffffu, v1.w, 1), l5 = __shfl_up_sync(0xffffffffu, v5.w, 1),
       l8 = __shfl_up_sync(0xffffffffu, v8.w, 1);
  real r3 = __shfl_down_sync(0xffffffffu, v3.x, 1), r6 = __shfl_down_sync(0xffffffffu, v6.x, 1),
       r7 = __shfl_down_sync(0xffffffffu, v7.x, 1);
  if (need_w) { l1 = w1e; l5 = w5e; l8 = w8e; }
  if (need_e) { r3 = e3e; r6 = e6e; r7 = e7e; }

  real in[4][9], out[4][9];
  in[0][0] = v0.x; in[1][0] = v0.y; in[2][0] = v0.z; in[3][0] = v0.w;
  in[0][1] = l1;   in[1][1] = v1.x; in[2][1] = v1.y; in[3][1] = v1.z;
  in[0][2] = v2.x; in[1][2] = v2.y; in[2][2] = v2.z; in[3][2] = v2.w;
  in[0][3] = v3.y; in[1][3] = v3.z; in[2][3] = v3.w; in[3][3] = r3;
  in[0][4] = v4.x; in[1][4] = v4.y; in[2][4] = v4.z; in[3][4] = v4.w;
  in[0][5] = l5;   in[1][5] = v5.x; in[2][5] = v5.y; in[3][5] = v5.z;
  in[0][6] = v6.y; in[1][6] = v6.z; in[2][6] = v6.w; in[3][6] = r6;
  in[0][7] = v7.y; in[1][7] = v7.z; in[2][7] = v7.w; in[3][7] = r7;
  in[0][8] = l8;   in[1][8] = v8.x; in[2][8] = v8.y; in[3][8] = v8.z;

  // Widths that are not a multiple of 4: the row's last thread holds 1-3 valid cells, the
  // rest of its vector is row padding (loaded and stored, never used).  The east neighbour
  // of the last valid cell is column 0 -- the wrap element r3/r6/r7 already holds -- and the
  // padding cells are treated as obstacles so that they add nothing to the average.
  uint32_t obits = mbits;
  const int nvalid = a.nx - xc;
  if (nvalid < 4) {
    if (nvalid == 1) { in[0][3] = r3; in[0][6] = r6; in[0][7] = r7; }
    else if (nvalid == 2) { in[1][3] = r3; in[1][6] = r6; in[1][7] = r7; }
    else { in[2][3] = r3; in[2][6] = r6; in[2][7] = r7; }
    obits |= (0xFu << nvalid) & 0xFu;
  }

#pragma unroll
  for (int j = 0; j < 4; j++) {
    const real s = cell_update<real, STRICT>(in[j], (obits >> j) & 1u, a.omega, out[j]);
    q += to_fixed(s);
    if (active && !(s < (real)LBM_SPEED_LIMIT)) atomicOr(a.av + 1, LBM_NONFINITE_MARK);   // NaN / blow-up
  }

  if (active) {
    real* d = a.dst + oC + xc;
#pragma unroll
    for (int k = 0; k < 9; k++) {
      Vec4<real> v; v.x = out[0][k]; v.y = out[1][k]; v.z = out[2][k]; v.w = out[3][k];
      st4(d + k * PS, v);
    }
    if (first | last | cA) {          // warp-uniform: a warp never spans two rows
      if (cA) {
        // next step's accelerate_flow on the row just produced (d2q9-bgk.c:229-260)
#pragma unroll
        for (int j = 0; j < 4; j++)
          cell_accelerate<real, STRICT>(out[j][1], out[j][3], out[j][5], out[j][6], out[j][7], out[j][8],
                                        (mbits >> j) & 1u, a.aw1, a.aw2);
        const int ks[6] = {1, 3, 5, 6, 7, 8};
#pragma unroll
        for (int i = 0; i < 6; i++) {
          Vec4<real> v; v.x = out[0][ks[i]]; v.y = out[1][ks[i]]; v.z = out[2][ks[i]]; v.w = out[3][ks[i]];
          st4(a.side_dst + (long long)i * a.pitch + xc, v);
        }
      }
      if (first) {                    // speeds 4,7,8 are pulled by the row below
        const int ks[3] = {4, 7, 8};
#pragma unroll
        for (int i = 0; i < 3; i++) {
          Vec4<real> v; v.x = out[0][ks[i]]; v.y = out[1][ks[i]]; v.z = out[2][ks[i]]; v.w = out[3][ks[i]];
          st4(a.push_dn + (long long)i * a.pitch + xc, v);
        }
      }
      if (last) {                     // speeds 2,5,6 are pulled by the row above
        const int ks[3] = {2, 5, 6};
#pragma unroll
        for (int i = 0; i < 3; i++) {
          Vec4<real> v; v.x = out[0][ks[i]]; v.y = out[1][ks[i]]; v.z = out[2][ks[i]]; v.w = out[3][ks[i]];
          st4(a.push_up + (long long)i * a.pitch + xc, v);
        }
      }
    }
  } else {
    q = 0ULL;
  }
  return q;
}

template <typename real, bool STRICT, bool MULTI>
__global__ void __launch_bounds__(LBM_BLOCK_THREADS, (sizeof(real) == 4 ? LBM_MIN_BLOCKS : 1))
lbm_step_vec4(const __grid_constant__ StepArgs<real> a) {
  int tx, ty;
  tile_of_block(a.tiles_x, a.tiles_y, tx, ty);
  // Programmatic dependent launch: when the host launches the steps with
  // cudaLaunchAttributeProgrammaticStreamSerialization this grid may be scheduled while
  // the previous step drains; everything the previous step wrote is visible after this
  // call.  A no-op for an ordinary launch.
  cudaGridDependencySynchronize();
  const bool is_boundary = (ty == 0) || (ty == a.tiles_y - 1);
  boundary_wait<real, MULTI>(a, is_boundary);
  const unsigned long long q = vec4_tile<real, STRICT, false>(a, tx, ty);
  block_accumulate(q, a.av);
  boundary_signal<real, MULTI>(a, is_boundary);
}

// ------------------------------------------------------------------------------------
// K1c  lbm_step_tma (fp32, opt-in with LBM_GPU_KERNEL_TMA): the same step with the nine
// pulled rows of a 512-cell tile staged in shared memory by the Tensor Memory Accelerator
// (one 3-D tensor map per lattice buffer: x, row, plane).  The copy engine does the y part
// of the propagate shift (box row = r - e_y).  It cannot do the x part: the innermost box
// coordinate must be a multiple of 16 bytes -- x = 1 raises an illegal-instruction fault
// (tools/probe/tma_probe.cu) -- so the x-shifted planes get a 4-element halo box on the
// side they pull from and the threads read that one extra element from shared memory,
// where K1a uses a warp shuffle.  The two ends of a row (periodic wrap in x) and the few
// rows that read the halo window or the accelerated side row keep the direct-load path
// (vec4_tile), chosen per block.  Built to measure whether staging buys anything over
// K1a: it does not (profiles/r01_kernel_variants.md), every byte is used once either way.
// ------------------------------------------------------------------------------------
#define LBM_TMA_TILE (LBM_BLOCK_THREADS * 4)
#define LBM_TMA_BOX (LBM_TMA_TILE < 256 ? LBM_TMA_TILE : 256)
static_assert(9 * LBM_TMA_TILE * sizeof(float) + 9 * 32 * sizeof(float) + 16 <= 48 * 1024,
              "K1c stages its tile in static shared memory: LBM_BLOCK_THREADS must be <= 256");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool STRICT, bool MULTI>
__global__ void __launch_bounds__(LBM_BLOCK_THREADS, LBM_MIN_BLOCKS)
lbm_step_tma(const __grid_constant__ StepArgs<float> a, const __grid_constant__ CUtensorMap src_map,
             const __grid_constant__ CUtensorMap halo_map) {
  __shared__ alignas(128) float tile[9][LBM_TMA_TILE];
  __shared__ alignas(128) float halo[9][32];          // [k][0..3]: planes 1,5,8 cells xt-4..xt-1; 3,6,7 xt+T..xt+T+3
                                                      // (rows 128 B apart: TMA destinations are 128-byte aligned)
  __shared__ alignas(8) unsigned long long mbar;
  int tx, ty;
  tile_of_block(a.tiles_x, a.tiles_y, tx, ty);
  const bool is_boundary = (ty == 0) || (ty == a.tiles_y - 1);
  boundary_wait<float, MULTI>(a, is_boundary);

  const int r = ty;                                   // blockDim = (LBM_BLOCK_THREADS, 1): one row per tile
  const bool special = (r == 0) || (r == a.rows - 1) || (r == a.accel_row) || (r - 1 == a.accel_row) ||
                       (r + 1 == a.accel_row);
  unsigned long long q = 0ULL;
  if (special) {
    q = vec4_tile<float, STRICT, false>(a, tx, ty);
  } else {
    const int tid = threadIdx.x;
    const int xt = tx * LBM_TMA_TILE;                 // first column of the tile
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)),
                   "r"((uint32_t)((9 * LBM_TMA_TILE + 6 * 4) * sizeof(float))) : "memory");
      // e_k of d2q9-bgk.c:7-13: the value pulled into (x, r) for speed k sits at (x - ex, r - ey).
      // A TMA box is at most 256 elements per dimension: LBM_TMA_TILE / 256 boxes per plane,
      // plus a 4-element box left of the tile (ex = +1) or right of it (ex = -1).
#define LBM_TMA_LOAD(k, ex, ey)                                                                                  \
  _Pragma("unroll") for (int b = 0; b < LBM_TMA_TILE / LBM_TMA_BOX; b++)                                        \
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" \
                 ::"r"(smem_u32(&tile[k][b * LBM_TMA_BOX])), "l"(&src_map), "r"(smem_u32(&mbar)),                 \
                 "r"(xt + b * LBM_TMA_BOX), "r"(r - (ey)), "r"(k) : "memory");                                   \
  if ((ex) != 0)                                                                                                 \
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" \
                 ::"r"(smem_u32(&halo[k][0])), "l"(&halo_map), "r"(smem_u32(&mbar)),                              \
                 "r"((ex) > 0 ? xt - 4 : xt + LBM_TMA_TILE), "r"(r - (ey)), "r"(k) : "memory")
      LBM_TMA_LOAD(0, 0, 0);  LBM_TMA_LOAD(1, 1, 0);   LBM_TMA_LOAD(2, 0, 1);
      LBM_TMA_LOAD(3, -1, 0); LBM_TMA_LOAD(4, 0, -1);  LBM_TMA_LOAD(5, 1, 1);
      LBM_TMA_LOAD(6, -1, 1); LBM_TMA_LOAD(7, -1, -1); LBM_TMA_LOAD(8, 1, -1);
#undef LBM_TMA_LOAD
    }
    const int x0 = xt + tid * 4;
    const bool active = x0 < a.nx;
    const int xc = active ? x0 : 0;
    const long long PS = a.plane_stride;
    const long long oC = (long long)r * a.pitch, oS = oC - a.pitch, oN = oC + a.pitch;
    // the wrap elements at the two ends of the row come by ordinary loads (the tensor map
    // zero-fills x < 0 and reads row padding beyond nx)
    const bool need_w = active && (x0 == 0);
    const bool need_e = active && (x0 + 4 >= a.nx);
    float w1e = 0, w5e = 0, w8e = 0, e3e = 0, e6e = 0, e7e = 0;
    if (need_w) {
      w1e = a.src[1 * PS + oC + a.nx - 1]; w5e = a.src[5 * PS + oS + a.nx - 1]; w8e = a.src[8 * PS + oN + a.nx - 1];
    }
    if (need_e) { e3e = a.src[3 * PS + oC]; e6e = a.src[6 * PS + oS]; e7e = a.src[7 * PS + oN]; }
    const uint32_t mword = a.mask[(long long)r * a.mask_pitch + (xc >> 5)];
    uint32_t obits = (mword >> (xc & 31)) & 0xFu;

    {   // wait for the rows (phase 0 of a barrier used once)
      uint32_t done = 0;
      while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.b32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
    }
    float in[4][9], out[4][9];
#define LBM_ROW(k) (*reinterpret_cast<const float4*>(&tile[k][tid * 4]))
#define LBM_WEST(k) (tid == 0 ? halo[k][3] : tile[k][tid * 4 - 1])
#define LBM_EAST(k) (tid == LBM_BLOCK_THREADS - 1 ? halo[k][0] : tile[k][tid * 4 + 4])
    { const float4 v = LBM_ROW(0); in[0][0] = v.x; in[1][0] = v.y; in[2][0] = v.z; in[3][0] = v.w; }
    { const float4 v = LBM_ROW(2); in[0][2] = v.x; in[1][2] = v.y; in[2][2] = v.z; in[3][2] = v.w; }
    { const float4 v = LBM_ROW(4); in[0][4] = v.x; in[1][4] = v.y; in[2][4] = v.z; in[3][4] = v.w; }
    { const float4 v = LBM_ROW(1); in[0][1] = LBM_WEST(1); in[1][1] = v.x; in[2][1] = v.y; in[3][1] = v.z; }
    { const float4 v = LBM_ROW(5); in[0][5] = LBM_WEST(5); in[1][5] = v.x; in[2][5] = v.y; in[3][5] = v.z; }
    { const float4 v = LBM_ROW(8); in[0][8] = LBM_WEST(8); in[1][8] = v.x; in[2][8] = v.y; in[3][8] = v.z; }
    { const float4 v = LBM_ROW(3); in[0][3] = v.y; in[1][3] = v.z; in[2][3] = v.w; in[3][3] = LBM_EAST(3); }
    { const float4 v = LBM_ROW(6); in[0][6] = v.y; in[1][6] = v.z; in[2][6] = v.w; in[3][6] = LBM_EAST(6); }
    { const float4 v = LBM_ROW(7); in[0][7] = v.y; in[1][7] = v.z; in[2][7] = v.w; in[3][7] = LBM_EAST(7); }
#undef LBM_ROW
#undef LBM_WEST
#undef LBM_EAST
    if (need_w) { in[0][1] = w1e; in[0][5] = w5e; in[0][8] = w8e; }
    if (need_e) {
      const int nvalid = a.nx - x0;
      if (nvalid == 1) { in[0][3] = e3e; in[0][6] = e6e; in[0][7] = e7e; }
      else if (nvalid == 2) { in[1][3] = e3e; in[1][6] = e6e; in[1][7] = e7e; }
      else if (nvalid == 3) { in[2][3] = e3e; in[2][6] = e6e; in[2][7] = e7e; }
      else { in[3][3] = e3e; in[3][6] = e6e; in[3][7] = e7e; }
      if (nvalid < 4) obits |= (0xFu << nvalid) & 0xFu;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float s = cell_update<float, STRICT>(in[j], (obits >> j) & 1u, a.omega, out[j]);
      q += to_fixed(s);
      if (active && !(s < (float)LBM_SPEED_LIMIT)) atomicOr(a.av + 1, LBM_NONFINITE_MARK);
    }
    if (active) {
      float* d = a.dst + oC + x0;
#pragma unroll
      for (int k = 0; k < 9; k++) {
        Vec4<float> v; v.x = out[0][k]; v.y = out[1][k]; v.z = out[2][k]; v.w = out[3][k];
        st4(d + k * PS, v);
      }
    } else {
      q = 0ULL;
    }
  }
  block_accumulate(q, a.av);
  boundary_signal<float, MULTI>(a, is_boundary);
}

// ------------------------------------------------------------------------------------
// K1b  lbm_step_scalar: one cell per thread, any nx.  Same arithmetic; used for grids
// whose width is not a multiple of 4 and as an independent cross-check of K1a.
// ------------------------------------------------------------------------------------
template <typename real, bool STRICT, bool CG>
__device__ __forceinline__ unsigned long long scalar_tile(const StepArgs<real>& a, const int tx, const int ty) {
#define LD(ptr) load1<real, CG>(ptr)
  const int x = tx * (int)blockDim.x + (int)threadIdx.x;
  const int r = ty * (int)blockDim.y + (int)threadIdx.y;
  const bool active = (x < a.nx) && (r < a.rows);
  unsigned long long q = 0ULL;
  if (active) {
    const long long PS = a.plane_stride;
    const long long oC = (long long)r * a.pitch, oS = oC - a.pitch, oN = oC + a.pitch;
    const bool first = (r == 0), last = (r == a.rows - 1);
    const bool cA = (r == a.accel_row), sA = (r - 1 == a.accel_row), nA = (r + 1 == a.accel_row);
    const int xw = (x == 0) ? a.nx - 1 : x - 1;
    const int xe = (x + 1 == a.nx) ? 0 : x + 1;
    real p[9], o[9];
    p[0] = LD(a.src + oC + x);
    p[1] = cA ? LD(a.side_src + 0 * a.pitch + xw) : LD(a.src + 1 * PS + oC + xw);
    p[3] = cA ? LD(a.side_src + 1 * a.pitch + xe) : LD(a.src + 3 * PS + oC + xe);
    p[2] = first ? LD(a.halo_s + 0 * a.pitch + x) : LD(a.src + 2 * PS + oS + x);
    p[5] = first ? LD(a.halo_s + 1 * a.pitch + xw) : (sA ? LD(a.side_src + 2 * a.pitch + xw) : LD(a.src + 5 * PS + oS + xw));
    p[6] = first ? LD(a.halo_s + 2 * a.pitch + xe) : (sA ? LD(a.side_src + 3 * a.pitch + xe) : LD(a.src + 6 * PS + oS + xe));
    p[4] = last ? LD(a.halo_n + 0 * a.pitch + x) : LD(a.src + 4 * PS + oN + x);
    p[7] = last ? LD(a.halo_n + 1 * a.pitch + xe) : (nA ? LD(a.side_src + 4 * a.pitch + xe) : LD(a.src + 7 * PS + oN + xe));
    p[8] = last ? LD(a.halo_n + 2 * a.pitch + xw) : (nA ? LD(a.side_src + 5 * a.pitch + xw) : LD(a.src + 8 * PS + oN + xw));
    const bool obst = (a.mask[(long long)r * a.mask_pitch + (x >> 5)] >> (x & 31)) & 1u;
    const real s = cell_update<real, STRICT>(p, obst, a.omega, o);
    q = to_fixed(s);
    if (!(s < (real)LBM_SPEED_LIMIT)) atomicOr(a.av + 1, LBM_NONFINITE_MARK);
#pragma unroll
    for (int k = 0; k < 9; k++) a.dst[k * PS + oC + x] = o[k];
    if (cA) {
      cell_accelerate<real, STRICT>(o[1], o[3], o[5], o[6], o[7], o[8], obst, a.aw1, a.aw2);
      a.side_dst[0 * a.pitch + x] = o[1]; a.side_dst[1 * a.pitch + x] = o[3];
      a.side_dst[2 * a.pitch + x] = o[5]; a.side_dst[3 * a.pitch + x] = o[6];
      a.side_dst[4 * a.pitch + x] = o[7]; a.side_dst[5 * a.pitch + x] = o[8];
    }
    if (first) {
      a.push_dn[0 * a.pitch + x] = o[4];
      a.push_dn[1 * a.pitch + x] = o[7];
      a.push_dn[2 * a.pitch + x] = o[8];
    }
    if (last) {
      a.push_up[0 * a.pitch + x] = o[2];
      a.push_up[1 * a.pitch + x] = o[5];
      a.push_up[2 * a.pitch + x] = o[6];
    }
  }
#undef LD
  return q;
}

template <typename real, bool STRICT, bool MULTI>
__global__ void __launch_bounds__(LBM_BLOCK_THREADS)
lbm_step_scalar(const __grid_constant__ StepArgs<real> a) {
  int tx, ty;
  tile_of_block(a.tiles_x, a.tiles_y, tx, ty);
  const bool is_boundary = (ty == 0) || (ty == a.tiles_y - 1);
  boundary_wait<real, MULTI>(a, is_boundary);
  const unsigned long long q = scalar_tile<real, STRICT, false>(a, tx, ty);
  block_accumulate(q, a.av);
  boundary_signal<real, MULTI>(a, is_boundary);
}

// ------------------------------------------------------------------------------------
// K5  lbm_steps_persistent: ALL timesteps of a run in one cooperative launch, for grids
// small enough to live in L2 (the reference's shipped inputs: 0.6-38 MB per buffer).
// There the per-step cost of K1a is launch latency, not bandwidth.  Every block owns a
// fixed set of tiles, loops over the steps, and meets the other blocks at a grid barrier
// (one atomic per block) between steps; buffers swap roles inside the kernel.  Single
// slab only (the halo window is the slab's own).  Loads are L2-only (see ld4cg).
// VEC = 4 cells per thread (K1a's tile) or, for the smallest grids, 1 cell per thread
// (K1b's tile): four times as many threads share the step's dependent-latency chain.
// Must be launched with cudaLaunchCooperativeKernel so that all blocks are resident.
// ------------------------------------------------------------------------------------
template <typename real>
struct PersistArgs {
  StepArgs<real> s;            // geometry, constants and the step-0 pointers
  real* lattice[2];
  real* side[2];
  real* window;                // own window: section (parity b, direction d) at ((b*2+d)*3)*pitch
  unsigned long long* av;      // n_steps x LBM_AV_SLOTS x LBM_AV_STRIDE words
  unsigned long long* barrier; // zeroed before the launch
  int first_parity;            // buffer index read by the first step
  int n_steps;
  int n_tiles;
};

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void grid_barrier(unsigned long long* counter, const unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    __threadfence();
    atomicAdd(counter, 1ULL);
    while (ld_acquire_gpu(counter) < target) { }
    __threadfence();
  }
  __syncthreads();
}

template <typename real, bool STRICT, int VEC>
__global__ void __launch_bounds__(LBM_BLOCK_THREADS, (sizeof(real) == 4 ? LBM_PERSIST_MIN_BLOCKS : 1))
lbm_steps_persistent(const __grid_constant__ PersistArgs<real> pa) {
  StepArgs<real> a = pa.s;
  const int pitch = a.pitch;
  for (int t = 0; t < pa.n_steps; t++) {
    const int src = (pa.first_parity + t) & 1, dst = src ^ 1;
    a.src = pa.lattice[src];
    a.dst = pa.lattice[dst];
    a.side_src = pa.side[src];
    a.side_dst = pa.side[dst];
    a.halo_s = pa.window + (size_t)((src * 2 + 0) * 3) * pitch;
    a.halo_n = pa.window + (size_t)((src * 2 + 1) * 3) * pitch;
    a.push_up = pa.window + (size_t)((dst * 2 + 0) * 3) * pitch;
    a.push_dn = pa.window + (size_t)((dst * 2 + 1) * 3) * pitch;
    a.av = pa.av + (size_t)t * (LBM_AV_STRIDE * LBM_AV_SLOTS);
#if LBM_K5_EXPERIMENT != 2
    unsigned long long q = 0ULL;
    for (int tile = blockIdx.x; tile < pa.n_tiles; tile += gridDim.x) {
      const int ty = tile / a.tiles_x;
      const int tx = tile - ty * a.tiles_x;
      q += (VEC == 4) ? vec4_tile<real, STRICT, true>(a, tx, ty) : scalar_tile<real, STRICT, true>(a, tx, ty);
    }
    warp_accumulate(q, a.av);
#endif
#if LBM_K5_EXPERIMENT != 1
    // (a variant where the last arriver publishes an epoch in a separate word that the
    // others poll was measured SLOWER -- one more L2 round trip: 128^2 3.0 -> 3.6 us/step)
    grid_barrier(pa.barrier, (unsigned long long)gridDim.x * (unsigned long long)(t + 1));
#endif
  }
}

// ------------------------------------------------------------------------------------
// K3  lbm_prepare: run once after the lattice was created or uploaded.  Fills the side
// row (accelerate_flow applied to row ny-2 of the current buffer, d2q9-bgk.c:229-260)
// and pushes the slab's edge rows into the neighbours' ghost rows of the same buffer.
// One thread per column.
// ------------------------------------------------------------------------------------
template <typename real>
struct PrepareArgs {
  const real* cur;            // current lattice buffer
  real* side_cur;             // side row of the same parity
  const uint32_t* mask;
  real* push_up;              // neighbour above's window, current parity, "from below"
  real* push_dn;              // neighbour below's window, current parity, "from above"
  long long plane_stride;
  int nx, rows, pitch, mask_pitch, accel_row;
  real aw1, aw2;
};

template <typename real>
__global__ void lbm_prepare(const PrepareArgs<real> a) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= a.nx) return;
  const long long PS = a.plane_stride;
#pragma unroll 1
  for (int which = 0; which < 3; which++) {
    // 0: accelerate row -> side; 1: first row -> push down; 2: last row -> push up
    const int r = (which == 0) ? a.accel_row : (which == 1 ? 0 : a.rows - 1);
    if (r < 0) continue;
    const long long o = (long long)r * a.pitch + x;
    real f[9];
#pragma unroll
    for (int k = 0; k < 9; k++) f[k] = a.cur[k * PS + o];
    if (r == a.accel_row) {
      const bool obst = (a.mask[(long long)r * a.mask_pitch + (x >> 5)] >> (x & 31)) & 1u;
      cell_accelerate<real, true>(f[1], f[3], f[5], f[6], f[7], f[8], obst, a.aw1, a.aw2);
    }
    if (which == 0) {
      a.side_cur[0 * a.pitch + x] = f[1]; a.side_cur[1 * a.pitch + x] = f[3];
      a.side_cur[2 * a.pitch + x] = f[5]; a.side_cur[3 * a.pitch + x] = f[6];
      a.side_cur[4 * a.pitch + x] = f[7]; a.side_cur[5 * a.pitch + x] = f[8];
    } else if (which == 1) {
      a.push_dn[0 * a.pitch + x] = f[4];
      a.push_dn[1 * a.pitch + x] = f[7];
      a.push_dn[2 * a.pitch + x] = f[8];
    } else {
      a.push_up[0 * a.pitch + x] = f[2];
      a.push_up[1 * a.pitch + x] = f[5];
      a.push_up[2 * a.pitch + x] = f[6];
    }
  }
}

// ------------------------------------------------------------------------------------
// K4  setup / output kernels (not on the per-step path)
// ------------------------------------------------------------------------------------
// rest state, d2q9-bgk.c:2802-2823
template <typename real>
__global__ void lbm_init_rest(real* buf, long long plane_stride, int pitch, int nx, int rows,
                              real w0, real w1, real w2) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)rows * pitch;
  if (n >= total) return;
#pragma unroll
  for (int k = 0; k < 9; k++) buf[k * plane_stride + n] = (k == 0) ? w0 : (k < 5 ? w1 : w2);
}

// AoS rows (9 reals per cell, dense nx) -> SoA planes; `aos` holds nrows rows that go
// to local rows [r0, r0+nrows)
template <typename real>
__global__ void lbm_aos_to_soa(const real* __restrict__ aos, real* __restrict__ buf, long long plane_stride,
                               int pitch, int nx, int r0, long long ncells) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncells * 9) return;
  const long long cell = i / 9;
  const int k = (int)(i - cell * 9);
  const long long row = cell / nx;
  const int x = (int)(cell - row * nx);
  buf[k * plane_stride + (r0 + row) * pitch + x] = aos[i];
}

template <typename real>
__global__ void lbm_soa_to_aos(const real* __restrict__ buf, real* __restrict__ aos, long long plane_stride,
                               int pitch, int nx, int r0, long long ncells) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncells * 9) return;
  const long long cell = i / 9;
  const int k = (int)(i - cell * 9);
  const long long row = cell / nx;
  const int x = (int)(cell - row * nx);
  aos[i] = buf[k * plane_stride + (r0 + row) * pitch + x];
}

// int-per-cell obstacle rows -> bit mask rows; one warp packs 32 cells with a ballot.
// Also counts the blocked cells.
__global__ void lbm_pack_mask(const int* __restrict__ obst, uint32_t* __restrict__ mask, int mask_pitch,
                              int nx, int r0, int nrows, unsigned long long* blocked_count) {
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // global warp
  const int lane = threadIdx.x & 31;
  const int words_per_row = (nx + 31) >> 5;
  if (gw >= (long long)nrows * words_per_row) return;
  const int row = (int)(gw / words_per_row);
  const int wi = (int)(gw - (long long)row * words_per_row);
  const int x = wi * 32 + lane;
  const bool b = (x < nx) && (obst[(long long)row * nx + x] != 0);
  const uint32_t bits = __ballot_sync(0xffffffffu, b);
  if (lane == 0) {
    mask[(long long)(r0 + row) * mask_pitch + wi] = bits;
    if (bits) atomicAdd(blocked_count, (unsigned long long)__popc(bits));
  }
}

// dense bit-packed host rows ((nx+31)/32 words per row) -> pitched mask rows
__global__ void lbm_copy_mask_bits(const uint32_t* __restrict__ in, uint32_t* __restrict__ mask, int mask_pitch,
                                   int nx, int r0, int nrows, unsigned long long* blocked_count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int words_per_row = (nx + 31) >> 5;
  if (i >= (long long)nrows * words_per_row) return;
  const int row = (int)(i / words_per_row);
  const int wi = (int)(i - (long long)row * words_per_row);
  uint32_t bits = in[i];
  const int valid = nx - wi * 32;
  if (valid < 32) bits &= (1u << valid) - 1u;
  mask[(long long)(r0 + row) * mask_pitch + wi] = bits;
  if (bits) atomicAdd(blocked_count, (unsigned long long)__popc(bits));
}

// write_values' per-cell fields (d2q9-bgk.c:2937-2976) for local rows [r0, r0+nrows):
// dense nx-wide outputs.  Also usable as av_velocity (d2q9-bgk.c:2665-2714) through
// the |u| accumulator (warp_accumulate) when av != NULL.
template <typename real>
__global__ void lbm_fields(const real* __restrict__ buf, const uint32_t* __restrict__ mask,
                           long long plane_stride, int pitch, int mask_pitch, int nx, int r0, int nrows,
                           real density, real* __restrict__ ux_out, real* __restrict__ uy_out,
                           real* __restrict__ u_out, real* __restrict__ p_out,
                           unsigned long long* av) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long q = 0ULL;
  if (n < (long long)nrows * nx) {
    const int row = (int)(n / nx);
    const int x = (int)(n - (long long)row * nx);
    const int r = r0 + row;                      // local row
    const bool obst = (mask[(long long)r * mask_pitch + (x >> 5)] >> (x & 31)) & 1u;
    real f[9];
#pragma unroll
    for (int k = 0; k < 9; k++) f[k] = buf[k * plane_stride + (long long)r * pitch + x];
    real ux, uy, rho;
    real u = cell_macroscopic<real>(f, ux, uy, rho);
    const real c_sq = (real)1 / (real)3;
    real pr = Ops<real, true>::mul(rho, c_sq);
    if (obst) { ux = 0; uy = 0; u = 0; pr = Ops<real, true>::mul(density, c_sq); }
    if (ux_out) ux_out[n] = ux;
    if (uy_out) uy_out[n] = uy;
    if (u_out) u_out[n] = u;
    if (p_out) p_out[n] = pr;
    q = to_fixed(u);
  }
  if (av) warp_accumulate(q, av);
}

// Digest of the lattice rows [r0, r0+nrows): two wrapping 64-bit integer sums, so they
// are exact, order-independent and additive over slabs / ranks.
//   mass      sum over cells and speeds of round(f * 2^32): total_density of the reference
//             (d2q9-bgk.c:2900-2916) in fixed point -- stays constant step after step;
//   checksum  sum of bit_pattern(f_k(cell)) * odd_weight(global cell index, k): equal for
//             two lattices iff (up to 2^-64) every speed of every cell has the same bits,
//             whatever the decomposition.  Full-size stand-in for a lattice comparison.
__device__ __forceinline__ unsigned long long bits_of(float f) { return (unsigned long long)__float_as_uint(f); }
__device__ __forceinline__ unsigned long long bits_of(double f) { return (unsigned long long)__double_as_longlong(f); }

template <typename real>
__global__ void lbm_digest(const real* __restrict__ buf, long long plane_stride, int pitch, int nx, int r0,
                           int nrows, long long global_row0, unsigned long long* out /* [mass, checksum] */) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long mass = 0ULL, sum = 0ULL;
  if (n < (long long)nrows * nx) {
    const int row = (int)(n / nx);
    const int x = (int)(n - (long long)row * nx);
    const unsigned long long g = (unsigned long long)(global_row0 + row) * (unsigned long long)nx + (unsigned long long)x;
#pragma unroll
    for (int k = 0; k < 9; k++) {
      const real f = buf[k * plane_stride + (long long)(r0 + row) * pitch + x];
      mass += (unsigned long long)__double2ll_rn((double)f * 4294967296.0);
      const unsigned long long w = (g * 0x9E3779B97F4A7C15ULL + (unsigned long long)(k + 1) * 0xC2B2AE3D27D4EB4FULL) | 1ULL;
      sum += bits_of(f) * w;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    mass += __shfl_down_sync(0xffffffffu, mass, off);
    sum += __shfl_down_sync(0xffffffffu, sum, off);
  }
  if ((threadIdx.x & 31) == 0) {
    if (mass) atomicAdd(out + 0, mass);
    if (sum) atomicAdd(out + 1, sum);
  }
}

}  // namespace lbm
