// FP32 FMA peak of this part, measured: every SM retires packed FFMA2 (fma.rn.f32x2, two FMAs per lane per instruction)
// from 16 independent accumulators per thread, 16 warps per CTA, 2 CTAs per SM. bench.py times the launch with CUDA events and
// uses the result as the denominator of the FP32-FFMA roofline (MEASURED_PEAKS.json carries only HBM and tensor peaks).
// Diagnostic library only (include/pgmorl_b200_diag.h).
#include "../../../include/pgmorl_b200_diag.h"
#include "../common.cuh"

namespace pgm {
typedef unsigned long long u64;
__device__ __forceinline__ u64 burn_ffma2(u64 a, u64 b, u64 c) {
    u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__global__ void __launch_bounds__(512, 2) ffma2_burn_kernel(float *out, int iters, float seed) {
    const float a0 = seed + threadIdx.x * 1e-6f, b0 = seed * 0.5f;
    u64 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (u64)__float_as_uint(1e-3f * i) << 32 | __float_as_uint(2e-3f * i);
    const u64 A0 = ((u64)__float_as_uint(a0) << 32) | __float_as_uint(a0 * 0.99f), A1 = ((u64)__float_as_uint(a0 * 0.98f) << 32) | __float_as_uint(a0 * 0.97f);
    const u64 B0 = ((u64)__float_as_uint(b0) << 32) | __float_as_uint(b0 * 0.99f), B1 = ((u64)__float_as_uint(b0 * 0.98f) << 32) | __float_as_uint(b0 * 0.97f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = burn_ffma2((i & 1) ? A1 : A0, (i & 2) ? B1 : B0, acc[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += __uint_as_float((unsigned)(acc[i] & 0xffffffffu)) + __uint_as_float((unsigned)(acc[i] >> 32));
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace pgm

// Launch the burn kernel on `ctas` CTAs of 512 threads; out needs ctas * 512 floats. FLOPs of the launch:
// ctas * 512 threads * iters * 64 FFMA2 * 2 lanes * 2.
extern "C" int pgm_ffma2_burn(float *out, int ctas, int iters, void *stream) {
    PGM_REQUIRE(out && ctas > 0 && iters > 0, "pgm_ffma2_burn: bad arguments");
    pgm::ffma2_burn_kernel<<<ctas, 512, 0, (cudaStream_t)stream>>>(out, iters, 1.0f);
    PGM_CUDA(cudaGetLastError());
    return PGM_OK;
}
