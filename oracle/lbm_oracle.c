/*
 * lbm_oracle.c -- CPU restatement of the d2q9-bgk timestep loop.
 *
 * THIS IS TEST INFRASTRUCTURE.  It is the checker the CUDA path is compared with,
 * never the thing shipped or measured as the product.  Only tests/, the smoke test
 * in __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product (advanced-hpc-lbm_b200/csrc + the C host in
 * advanced-hpc-lbm_b200/host) never links, loads or calls anything in oracle/.
 *
 * What it restates: /root/reference/d2q9-bgk.c -- accelerate_flow (:1888-1918),
 * propagate (:2123-2152), rebound (:2199-2228), collision (:2554-2663),
 * av_velocity (:2665-2714), the fused live step timestep_new2 (:228-1813), the
 * step loop of main (:180-201), the rest-state initialisation (:2802-2823) and the
 * final_state fields of write_values (:2935-2976).  Each function in
 * lbm_oracle_impl.h cites the lines it follows.
 *
 * Parity is PINNED (tests/test_oracle_*.py, -m "not gpu"):
 *   - the f64 instantiation reproduces the reference's golden av_vels files
 *     (check/{128x128,128x256,256x256,1024x1024}.av_vels.dat) and golden
 *     final_state pressures, stored compactly under tests/golden/;
 *   - where oracle/_ref/ holds the reference compiled from /root/reference (see
 *     oracle/build_oracle.py), the f32 instantiation is compared bit-for-bit with
 *     the reference's own timestep_new2 / accelerate_flow / propagate / rebound /
 *     collision / av_velocity on random lattices.
 *
 * Build: oracle/build_oracle.py (gcc -O2 -ffp-contract=off -fopenmp; no fast-math,
 * so the operation order is the source order).
 */
#include <math.h>
#include <stdlib.h>
#include <stddef.h>

#define REAL float
#define SQRT sqrtf
#define NAME(x) x##_f32
#define R(x) x##f
#include "lbm_oracle_impl.h"
#undef REAL
#undef SQRT
#undef NAME
#undef R

#define REAL double
#define SQRT sqrt
#define NAME(x) x##_f64
#define R(x) x
#include "lbm_oracle_impl.h"
#undef REAL
#undef SQRT
#undef NAME
#undef R

int oracle_abi_version(void) { return 1; }
