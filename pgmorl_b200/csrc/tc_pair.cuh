// FP16-pair operand helpers shared by the tensor-core kernels (K1 forward, K3 update): a = a1 + a2 with a1 = a rounded to
// 11 significant bits and a2 = fp16(a - a1); images are [rows][64 halfwords] with the SWIZZLE_128B XOR (tc.cuh).
#pragma once
#include <cuda_fp16.h>

#include "tc.cuh"

namespace pgm {

constexpr float TC_SH = 256.f, TC_SD = 4096.f, TC_SW = 256.f;   // scales of activations / backward signals / weights

__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }
// loads that must be ISSUED where they are written (prefetches): volatile asm keeps the compiler from sinking them
__device__ __forceinline__ float4 ld_nc_f4(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_cg_f4(const float4 *p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_nc_f32(const float *p) {
    float v; asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v;
}
__device__ __forceinline__ int ld_nc_s32(const int32_t *p) {
    int v; asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {   // two floats -> fp16x2 (a in the low half), saturating
    uint32_t r; asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r;
}
// Same without saturation, for the operands that come from OUTSIDE the kernel (observations, parameters): a value beyond
// fp16's range (|x| >= 65 520 for an observation; |w| >= 255.9 for a parameter, which is stored x 2^8) becomes inf and turns
// the outputs of that task into inf / NaN -- a loud failure instead of a silently clamped operand. Activations (|h| <= 1,
// stored x 2^8) and the backward signals (x 2^12) keep the saturating form.
__device__ __forceinline__ uint32_t pack_h2_ovf(float a, float b) {
    uint32_t r; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r;
}
// a1 = v rounded to 11 significant bits (round half away: integer add + mask, no conversion-pipe instruction);
// exactly representable in fp16 whenever v is in fp16's normal range
__device__ __forceinline__ float round11(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }
// store 8 consecutive features (logical 16-byte chunks ca / cb of the a1 / a2 destination rows; swz8 = row & 7 is the
// SWIZZLE_128B XOR) as an fp16 pair
__device__ __forceinline__ void store_pair8(unsigned char *rowa, uint32_t ca, unsigned char *rowb, uint32_t cb, uint32_t swz8, const float *v) {
    float h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = round11(v[i]);
    uint4 p1, p2;
    p1.x = pack_h2(h[0], h[1]); p1.y = pack_h2(h[2], h[3]); p1.z = pack_h2(h[4], h[5]); p1.w = pack_h2(h[6], h[7]);
    p2.x = pack_h2(v[0] - h[0], v[1] - h[1]); p2.y = pack_h2(v[2] - h[2], v[3] - h[3]);
    p2.z = pack_h2(v[4] - h[4], v[5] - h[5]); p2.w = pack_h2(v[6] - h[6], v[7] - h[7]);
    *reinterpret_cast<uint4 *>(rowa + ((ca ^ swz8) << 4)) = p1;
    *reinterpret_cast<uint4 *>(rowb + ((cb ^ swz8) << 4)) = p2;
}
// store_pair8 for externally supplied values (observations): overflow becomes inf, see pack_h2_ovf
__device__ __forceinline__ void store_pair8_ovf(unsigned char *rowa, uint32_t ca, unsigned char *rowb, uint32_t cb, uint32_t swz8, const float *v) {
    float h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = round11(v[i]);
    uint4 p1, p2;
    p1.x = pack_h2_ovf(h[0], h[1]); p1.y = pack_h2_ovf(h[2], h[3]); p1.z = pack_h2_ovf(h[4], h[5]); p1.w = pack_h2_ovf(h[6], h[7]);
    p2.x = pack_h2_ovf(v[0] - h[0], v[1] - h[1]); p2.y = pack_h2_ovf(v[2] - h[2], v[3] - h[3]);
    p2.z = pack_h2_ovf(v[4] - h[4], v[5] - h[5]); p2.w = pack_h2_ovf(v[6] - h[6], v[7] - h[7]);
    *reinterpret_cast<uint4 *>(rowa + ((ca ^ swz8) << 4)) = p1;
    *reinterpret_cast<uint4 *>(rowb + ((cb ^ swz8) << 4)) = p2;
}
// halfword index of (row, feature) inside a [rows][64 halfwords] SWIZZLE_128B image
__device__ __forceinline__ int sw128_hw(int row, int f) { return row * 64 + ((((f >> 3) ^ (row & 7))) << 3) + (f & 7); }
__device__ __forceinline__ void put_pair(__half *ia, int hwa, __half *ib, int hwb, float v) {
    const float h = round11(v);
    ia[hwa] = __float2half_rn(h); ib[hwb] = __float2half_rn(v - h);
}

}  // namespace pgm
