/* Diagnostic entry points of the tcgen05 building blocks (csrc/tc.cuh): self-test, operand-layout probe and MMA pacing
 * microbenchmark. Built into libpgmorl_b200_diag.so (pgmorl_b200/build.py), NOT into the product library: nothing on the
 * MOPG / selection path calls them. Used by tests/test_gpu_tc.py and profiles/tc_*.py. */
#ifndef PGMORL_B200_DIAG_H
#define PGMORL_B200_DIAG_H
#include "pgmorl_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------
 * Self-test of the tcgen05 (5th-generation tensor core) building blocks of the K3 tensor-core path
 * (csrc/tc.cuh): TF32 UMMA through every shared-memory operand view K3 uses, TMEM load/store, the 3-way
 * TF32 split. No reference counterpart (the reference runs torch-CPU GEMMs, algo/ppo.py:62-107).
 * out [n_out >= 16] device floats: [0..4] max |error| of five exact integer GEMMs (must be 0),
 * [5] 1xTF32 / [6] 3xTF32 / [7] FP32-FMA relative error vs FP64, [8] 0 = inputs truncated, 1 = rounded,
 * [9],[10] cycles for 1 / 24 MMAs issue->complete, [11] TMEM load cycles, [12] TMEM store/load error. */
int pgm_tc_selftest(float *out, int n_out, void *stream);

/* Layout-discovery aid for the same building blocks: ONE tf32 MMA (K = 8) with caller-chosen descriptor strides;
 * out [128 * N] receives the raw TMEM accumulator (lane-major). fill 0 = operand words hold their word index,
 * fill 1 = K-major identity image with R rows; ltA/ltB = descriptor layout type; a_tmem: A operand from TMEM; kind 0 = tf32, 1 = f16 (K = 16,
 * operands filled per halfword); offA/offB = byte offsets added to the operand start addresses. Used by profiles/tc_layout_probe.py only. */
int pgm_tc_layout_probe(float *out, int M, int N, int a_mn, int b_mn, int fillA, int fillB, int RA, int RB,
                        int lboA, int sboA, int lboB, int sboB, int d_lane_off, int ltA, int ltB, int a_tmem,
                        int kind, int offA, int offB, void *stream);

/* MMA pacing microbenchmark (profiles/tc_mma_bench.py): nmma kind::f16 MMAs (M x N x 16) from zero-filled
 * SWIZZLE_128B images, rotating over nacc accumulators, operand start addresses advancing by a_step / b_step bytes.
 * out [6] floats: per repetition {cycles issue..complete, cycles spent issuing}. */
int pgm_tc_mma_bench(float *out, int M, int N, int a_mn, int b_mn, int nmma, int nacc, int a_step, int b_step,
                     void *stream);

/* FP32 FMA peak probe (csrc/diag/ffma_peak.cu): `ctas` CTAs x 512 threads, each thread `iters` x 64 packed FFMA2 on 16
 * independent accumulators; out [ctas * 512] floats. FLOPs of one launch = ctas * 512 * iters * 64 * 2 lanes * 2. bench.py
 * times it with CUDA events (2 CTAs per SM) and uses the result as the denominator of the FP32-FFMA roofline. */
int pgm_ffma2_burn(float *out, int ctas, int iters, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PGMORL_B200_DIAG_H */
