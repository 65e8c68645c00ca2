// Self-test of the tcgen05 building blocks in tc.cuh: one CTA runs small TF32 GEMMs through every operand
// view the K3 tensor-core kernel relies on (chunk images as K-major and MN-major A / B, N = 16 / 32 / 64,
// accumulate flag, 3-way TF32 split) and checks them on the device against plain FP32/FP64 loops.
// Exposed as pgm_tc_selftest by the DIAGNOSTIC library libpgmorl_b200_diag.so (include/pgmorl_b200_diag.h), not by the
// product library; tests/test_gpu_tc.py asserts on the result vector.
#include <cuda_fp16.h>

#include "../../../include/pgmorl_b200_diag.h"
#include "../common.cuh"
#include "../tc.cuh"

namespace pgm {

namespace {

__device__ __forceinline__ float ival(int a, int b, int sa, int sb, int mod) { return (float)(((a * sa + b * sb) % mod) - mod / 2); }
__device__ __forceinline__ float rnd(uint32_t i, uint32_t salt) {   // deterministic pseudo-random in (-1, 1)
    uint32_t x = i * 2654435761u + salt * 40503u + 12345u;
    x ^= x >> 16; x *= 2246822519u; x ^= x >> 13; x *= 3266489917u; x ^= x >> 16;
    return (float)(x & 0xFFFFFF) * (2.f / 16777216.f) - 1.f;
}
// element (r, f) of a chunk image with R rows
__device__ __forceinline__ int img(int r, int f, int R) { return (f >> 2) * R * 4 + r * 4 + (f & 3); }
// element (r, f) of a row image with R rows (tc.cuh): word index
__device__ __forceinline__ int rimg(int r, int f, int R) { return (int)(tc::row_image_offset((uint32_t)r, (uint32_t)f, (uint32_t)R) >> 2); }

__device__ float block_max(float v, float *red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float t = red[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) t = fmaxf(t, red[i]);
    __syncthreads();
    return t;
}

}  // namespace

__global__ void __launch_bounds__(128, 1) tc_probe_kernel(float *out) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float red[8];
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t phase = 0;

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's lane quadrant

    float *A = sm;                 // up to 128 x 128 floats (64 KB)
    float *B = sm + 128 * 128;     // up to 64 x 128 floats  (32 KB)
    float *B2 = B + 64 * 128;      // second B image (lo part)  (32 KB)
    float *Alo = B2 + 2 * 64 * 128; // lo part of A in T6, 128 x 64 floats (32 KB)

    auto sync_issue = [&]() { tc::fence_async_smem(); tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after(); };
    auto wait_done = [&]() { tc::mbar_wait(&bar, phase); phase ^= 1; tc::tc_fence_after(); };

    // ---------------- T1: K-major A [128 x 32], K-major B [64 x 32] ----------------
    for (int i = tid; i < 128 * 32; i += 128) { const int r = i / 32, k = i % 32; A[img(r, k, 128)] = ival(r, k, 3, 5, 9); }
    for (int i = tid; i < 64 * 32; i += 128) { const int n = i / 32, k = i % 32; B[img(n, k, 64)] = ival(n, k, 7, 2, 7); }
    sync_issue();
    if (tid == 0) {
        const uint64_t da = tc::desc_kmajor(tc::smem_addr(A), 128), db = tc::desc_kmajor(tc::smem_addr(B), 64);
        const uint32_t id = tc::idesc_tf32(128, 64, 0, 0);
        for (int ks = 0; ks < 4; ++ks)
            tc::mma_tf32(tmem, tc::desc_advance(da, ks * 2 * 128 * 16), tc::desc_advance(db, ks * 2 * 64 * 16), id, ks > 0);
        tc::mma_commit(&bar);
    }
    wait_done();
    {
        float d[64], e = 0.f;
        tc::tmem_ld32(tlane, d); tc::tmem_ld32(tlane + 32, d + 32); tc::tmem_ld_wait();
        const int r = tid;
        for (int n = 0; n < 64; ++n) {
            float ref = 0.f;
            for (int k = 0; k < 32; ++k) ref += ival(r, k, 3, 5, 9) * ival(n, k, 7, 2, 7);
            e = fmaxf(e, fabsf(ref - d[n]));
        }
        e = block_max(e, red);
        if (tid == 0) out[0] = e;
    }

    // ---------------- T2: MN-major A [K = 64 rows][128 features], MN-major B [64 rows][64 features] ----------------
    tc::tc_fence_before();
    __syncthreads();
    for (int i = tid; i < 64 * 128; i += 128) { const int k = i / 128, m = i % 128; A[rimg(k, m, 64)] = ival(k, m, 5, 3, 11); }
    for (int i = tid; i < 64 * 64; i += 128) { const int k = i / 64, n = i % 64; B[rimg(k, n, 64)] = ival(k, n, 2, 9, 5); }
    sync_issue();
    if (tid == 0) {
        const uint64_t da = tc::desc_mnmajor(tc::smem_addr(A), 64), db = tc::desc_mnmajor(tc::smem_addr(B), 64);
        const uint32_t id = tc::idesc_tf32(128, 64, 1, 1);
        for (int ks = 0; ks < 8; ++ks)
            tc::mma_tf32(tmem + 64, tc::desc_advance(da, ks * 1024), tc::desc_advance(db, ks * 1024), id, ks > 0);
        tc::mma_commit(&bar);
    }
    wait_done();
    {
        float d[64], e = 0.f;
        tc::tmem_ld32(tlane + 64, d); tc::tmem_ld32(tlane + 96, d + 32); tc::tmem_ld_wait();
        const int m = tid;
        for (int n = 0; n < 64; ++n) {
            float ref = 0.f;
            for (int k = 0; k < 64; ++k) ref += ival(k, m, 5, 3, 11) * ival(k, n, 2, 9, 5);
            e = fmaxf(e, fabsf(ref - d[n]));
        }
        e = block_max(e, red);
        if (tid == 0) out[1] = e;
    }

    // ---------------- T3: K-major A [128 x 16], MN-major B = W[j = 16][n = 64] (image rows = j) ----------------
    tc::tc_fence_before();
    __syncthreads();
    for (int i = tid; i < 128 * 16; i += 128) { const int r = i / 16, j = i % 16; A[img(r, j, 128)] = ival(r, j, 3, 7, 7); }
    for (int i = tid; i < 16 * 64; i += 128) { const int j = i / 64, n = i % 64; B[rimg(j, n, 16)] = ival(j, n, 5, 2, 9); }
    sync_issue();
    if (tid == 0) {
        const uint64_t da = tc::desc_kmajor(tc::smem_addr(A), 128), db = tc::desc_mnmajor(tc::smem_addr(B), 16);
        const uint32_t id = tc::idesc_tf32(128, 64, 0, 1);
        for (int ks = 0; ks < 2; ++ks)
            tc::mma_tf32(tmem, tc::desc_advance(da, ks * 2 * 128 * 16), tc::desc_advance(db, ks * 1024), id, ks > 0);
        tc::mma_commit(&bar);
    }
    wait_done();
    {
        float d[64], e = 0.f;
        tc::tmem_ld32(tlane, d); tc::tmem_ld32(tlane + 32, d + 32); tc::tmem_ld_wait();
        const int r = tid;
        for (int n = 0; n < 64; ++n) {
            float ref = 0.f;
            for (int j = 0; j < 16; ++j) ref += ival(r, j, 3, 7, 7) * ival(j, n, 5, 2, 9);
            e = fmaxf(e, fabsf(ref - d[n]));
        }
        e = block_max(e, red);
        if (tid == 0) out[2] = e;
    }

    // ---------------- T4: N = 16 (K-major A [128 x 32], K-major B [16 x 32]); T5: weight-gradient shape ----------------
    tc::tc_fence_before();
    __syncthreads();
    for (int i = tid; i < 128 * 32; i += 128) { const int r = i / 32, k = i % 32; A[img(r, k, 128)] = ival(r, k, 3, 5, 9); }
    for (int i = tid; i < 16 * 32; i += 128) { const int n = i / 32, k = i % 32; B[img(n, k, 16)] = ival(n, k, 7, 2, 7); }
    sync_issue();
    if (tid == 0) {
        const uint64_t da = tc::desc_kmajor(tc::smem_addr(A), 128), db = tc::desc_kmajor(tc::smem_addr(B), 16);
        const uint32_t id = tc::idesc_tf32(128, 16, 0, 0);
        for (int ks = 0; ks < 4; ++ks)
            tc::mma_tf32(tmem + 128, tc::desc_advance(da, ks * 2 * 128 * 16), tc::desc_advance(db, ks * 2 * 16 * 16), id, ks > 0);
        tc::mma_commit(&bar);
    }
    wait_done();
    {
        float d[16], e = 0.f;
        tc::tmem_ld16(tlane + 128, d); tc::tmem_ld_wait();
        const int r = tid;
        for (int n = 0; n < 16; ++n) {
            float ref = 0.f;
            for (int k = 0; k < 32; ++k) ref += ival(r, k, 3, 5, 9) * ival(n, k, 7, 2, 7);
            e = fmaxf(e, fabsf(ref - d[n]));
        }
        e = block_max(e, red);
        if (tid == 0) out[3] = e;
    }
    // T5: weight-gradient shape: A = activations [128 rows][128 features] MN-major (M = features, K = rows),
    //     B = x image [128 rows][32 features] MN-major with N = 16 taken from features 16..31 (chunk offset 4)
    tc::tc_fence_before();
    __syncthreads();
    for (int i = tid; i < 128 * 128; i += 128) { const int k = i / 128, m = i % 128; A[rimg(k, m, 128)] = ival(k, m, 5, 3, 11); }
    for (int i = tid; i < 128 * 32; i += 128) { const int k = i / 32, n = i % 32; B[rimg(k, n, 128)] = ival(k, n, 2, 9, 5); }
    sync_issue();
    if (tid == 0) {
        // B starts 64 bytes into the 128-byte rows (features 16..31): the swizzle acts on the final address
        const uint64_t da = tc::desc_mnmajor(tc::smem_addr(A), 128), db = tc::desc_mnmajor(tc::smem_addr(B) + 64, 128);
        const uint32_t id = tc::idesc_tf32(128, 16, 1, 1);
        for (int ks = 0; ks < 16; ++ks)
            tc::mma_tf32(tmem + 160, tc::desc_advance(da, ks * 1024), tc::desc_advance(db, ks * 1024), id, ks > 0);
        tc::mma_commit(&bar);
    }
    wait_done();
    {
        float d[16], e = 0.f;
        tc::tmem_ld16(tlane + 160, d); tc::tmem_ld_wait();
        const int m = tid;
        for (int n = 0; n < 16; ++n) {
            float ref = 0.f;
            for (int k = 0; k < 128; ++k) ref += ival(k, m, 5, 3, 11) * ival(k, 16 + n, 2, 9, 5);
            e = fmaxf(e, fabsf(ref - d[n]));
        }
        e = block_max(e, red);
        if (tid == 0) out[4] = e;
    }

    // ---------------- T6: precision. random A [128 x 64], B [64 x 64] K-major; 1 x TF32 vs 3 x TF32 split ----------------
    tc::tc_fence_before();
    __syncthreads();
    for (int i = tid; i < 128 * 64; i += 128) {
        const int r = i / 64, k = i % 64;
        const float v = rnd(i, 1), h = tc::tf32_hi(v);
        A[img(r, k, 128)] = h;
        Alo[img(r, k, 128)] = v - h;
    }
    for (int i = tid; i < 64 * 64; i += 128) {
        const int n = i / 64, k = i % 64;
        const float v = rnd(i, 2), h = tc::tf32_hi(v);
        B[img(n, k, 64)] = h;
        B2[img(n, k, 64)] = v - h;
    }
    sync_issue();
    if (tid == 0) {
        const uint64_t dah = tc::desc_kmajor(tc::smem_addr(A), 128), dal = tc::desc_kmajor(tc::smem_addr(Alo), 128);
        const uint64_t dbh = tc::desc_kmajor(tc::smem_addr(B), 64), dbl = tc::desc_kmajor(tc::smem_addr(B2), 64);
        const uint32_t id = tc::idesc_tf32(128, 64, 0, 0);
        for (int ks = 0; ks < 8; ++ks)     // cols 0..63: hi*hi only
            tc::mma_tf32(tmem, tc::desc_advance(dah, ks * 4096), tc::desc_advance(dbh, ks * 2048), id, ks > 0);
        for (int ks = 0; ks < 8; ++ks) {   // cols 64..127: 3-way split, small terms first
            tc::mma_tf32(tmem + 64, tc::desc_advance(dal, ks * 4096), tc::desc_advance(dbh, ks * 2048), id, ks > 0);
            tc::mma_tf32(tmem + 64, tc::desc_advance(dah, ks * 4096), tc::desc_advance(dbl, ks * 2048), id, 1);
        }
        for (int ks = 0; ks < 8; ++ks)
            tc::mma_tf32(tmem + 64, tc::desc_advance(dah, ks * 4096), tc::desc_advance(dbh, ks * 2048), id, 1);
        tc::mma_commit(&bar);
    }
    wait_done();
    {
        float d1[64], d3[64], e1 = 0.f, e3 = 0.f, ef = 0.f;
        tc::tmem_ld32(tlane, d1); tc::tmem_ld32(tlane + 32, d1 + 32);
        tc::tmem_ld32(tlane + 64, d3); tc::tmem_ld32(tlane + 96, d3 + 32); tc::tmem_ld_wait();
        const int r = tid;
        for (int n = 0; n < 64; ++n) {
            double ref = 0.0, mag = 0.0;
            float f32 = 0.f;
            for (int k = 0; k < 64; ++k) {
                const float a = rnd(r * 64 + k, 1), b = rnd(n * 64 + k, 2);
                ref += (double)a * (double)b; mag += fabs((double)a * (double)b);
                f32 = fmaf(a, b, f32);
            }
            e1 = fmaxf(e1, (float)(fabs(ref - (double)d1[n]) / mag));
            e3 = fmaxf(e3, (float)(fabs(ref - (double)d3[n]) / mag));
            ef = fmaxf(ef, (float)(fabs(ref - (double)f32) / mag));
        }
        e1 = block_max(e1, red); e3 = block_max(e3, red); ef = block_max(ef, red);
        if (tid == 0) { out[5] = e1; out[6] = e3; out[7] = ef; }
    }

    // ---------------- T7: does the tensor core truncate or round FP32 inputs to TF32? ----------------
    tc::tc_fence_before();
    __syncthreads();
    for (int i = tid; i < 128 * 8; i += 128) { const int r = i / 8, k = i % 8; A[img(r, k, 128)] = (k == 0) ? (1.f + 1.5f * 0.00048828125f) : 0.f; }
    for (int i = tid; i < 16 * 8; i += 128) { const int n = i / 8, k = i % 8; B[img(n, k, 16)] = (k == 0) ? 1.f : 0.f; }
    sync_issue();
    if (tid == 0) {
        tc::mma_tf32(tmem, tc::desc_kmajor(tc::smem_addr(A), 128), tc::desc_kmajor(tc::smem_addr(B), 16), tc::idesc_tf32(128, 16, 0, 0), 0);
        tc::mma_commit(&bar);
    }
    wait_done();
    {
        float d[8];
        tc::tmem_ld8(tlane, d); tc::tmem_ld_wait();
        if (tid == 0) out[8] = (d[0] - 1.f) * 1024.f;     // 0 = truncation, 1 = round to nearest
    }

    // ---------------- T8: latency of issue -> commit -> wait (cycles) for 1 and for 24 MMAs of 128x64x8 ----------------
    for (int rep = 0; rep < 2; ++rep) {
        const int nm = rep == 0 ? 1 : 24;
        tc::tc_fence_before();
        __syncthreads();
        long long t0 = clock64();
        if (tid == 0) {
            const uint64_t da = tc::desc_kmajor(tc::smem_addr(A), 128), db = tc::desc_kmajor(tc::smem_addr(B), 64);
            const uint32_t id = tc::idesc_tf32(128, 64, 0, 0);
            for (int i = 0; i < nm; ++i) tc::mma_tf32(tmem + 64, da, db, id, i > 0);
            tc::mma_commit(&bar);
        }
        wait_done();
        long long t1 = clock64();
        if (tid == 0) out[9 + rep] = (float)(t1 - t0);
    }
    // TMEM load round trip (64 columns) and a full store->load
    {
        tc::tc_fence_before();
        __syncthreads();
        float d[64];
        long long t0 = clock64();
        tc::tmem_ld32(tlane, d); tc::tmem_ld32(tlane + 32, d + 32); tc::tmem_ld_wait();
        long long t1 = clock64();
        float s = 0.f;
        for (int i = 0; i < 64; ++i) s += d[i];
        for (int i = 0; i < 32; ++i) d[i] = (float)(tid * 100 + i);
        tc::tmem_st32(tlane + 192, d); tc::tmem_st_wait();
        float q[32];
        tc::tmem_ld32(tlane + 192, q); tc::tmem_ld_wait();
        float e = 0.f;
        for (int i = 0; i < 32; ++i) e = fmaxf(e, fabsf(q[i] - (float)(tid * 100 + i)));
        e = block_max(e, red);
        if (tid == 0) { out[11] = (float)(t1 - t0); out[12] = e; out[13] = s * 0.f; }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

}  // namespace pgm

extern "C" int pgm_tc_selftest(float *out, int n_out, void *stream) {
    PGM_REQUIRE(out && n_out >= 16, "pgm_tc_selftest: need a device buffer of >= 16 floats");
    const size_t smem = (128 * 128 + 3 * 64 * 128 + 128 * 64) * sizeof(float);   // 192 KB
    PGM_CUDA(cudaFuncSetAttribute(pgm::tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PGM_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * n_out, (cudaStream_t)stream));
    pgm::tc_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(out);
    PGM_CUDA(cudaGetLastError());
    return PGM_OK;
}

// ------------------------------------------------------------------------------------------------------
// Layout discovery aid (profiles/tc_layout_probe.py): one MMA (K = 8) with caller-chosen descriptors.
// fill == 0: operand words hold their own word index (mod 2048, exact in TF32); fill == 1: K-major identity
// image with R rows (element (r, k) = (r == k)); a_tmem: A comes from TMEM (lane m, column k holds m*8+k).
// D (128 lanes x N columns, raw TMEM order) -> out.
namespace pgm {
__global__ void __launch_bounds__(128, 1) tc_layout_kernel(float *out, int M, int N, int a_mn, int b_mn, int fillA, int fillB,
                                                           int RA, int RB, uint32_t lboA, uint32_t sboA, uint32_t lboB,
                                                           uint32_t sboB, uint32_t d_lane_off, uint32_t ltA, uint32_t ltB,
                                                           int a_tmem, int kind, uint32_t offA, uint32_t offB) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
    float *A = sm, *B = sm + 8192;   // 32 KB each
    if (kind == 0) {
        for (int i = tid; i < 8192; i += 128) {
            A[i] = fillA == 0 ? (float)(i % 2048) : 0.f;
            B[i] = fillB == 0 ? (float)(i % 2048) : 0.f;
        }
        __syncthreads();
        if (fillA == 1) for (int i = tid; i < 8; i += 128) A[img(i, i, RA)] = 1.f;
        if (fillB == 1) for (int i = tid; i < 8; i += 128) B[img(i, i, RB)] = 1.f;
    } else {   // fp16: halfword h holds (h mod 2048); identity = K-major no-swizzle image, 8 halfwords per 16-byte chunk
        __half *Ah = reinterpret_cast<__half *>(A), *Bh = reinterpret_cast<__half *>(B);
        for (int i = tid; i < 16384; i += 128) {
            Ah[i] = __float2half(fillA == 0 ? (float)(i % 2048) : 0.f);
            Bh[i] = __float2half(fillB == 0 ? (float)(i % 2048) : 0.f);
        }
        __syncthreads();
        if (fillA == 1) for (int i = tid; i < 16; i += 128) Ah[(i >> 3) * RA * 8 + i * 8 + (i & 7)] = __float2half(1.f);
        if (fillB == 1) for (int i = tid; i < 16; i += 128) Bh[(i >> 3) * RB * 8 + i * 8 + (i & 7)] = __float2half(1.f);
    }
    {   // pre-fill the accumulator columns with a sentinel; A operand in TMEM columns 256..263
        float z[32];
        for (int i = 0; i < 32; ++i) z[i] = -7.f;
        for (int c = 0; c < 256; c += 32) tc::tmem_st32(tlane + c, z);
        for (int i = 0; i < 32; ++i) z[i] = (float)(tid * 8 + i);
        tc::tmem_st32(tlane + 256, z);
        tc::tmem_st_wait();
    }
    tc::fence_async_smem(); tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
    if (tid == 0) {
        const uint64_t db = tc::make_desc(tc::smem_addr(B) + offB, lboB, sboB, ltB);
        const uint64_t da = tc::make_desc(tc::smem_addr(A) + offA, lboA, sboA, ltA);
        if (kind == 1) tc::mma_f16(tmem + (d_lane_off << 16), da, db, tc::idesc_f16(M, N, a_mn, b_mn), 0);
        else if (a_tmem) tc::mma_tf32_ts(tmem + (d_lane_off << 16), tmem + 256, db, tc::idesc_tf32(M, N, a_mn, b_mn), 0);
        else tc::mma_tf32(tmem + (d_lane_off << 16), da, db, tc::idesc_tf32(M, N, a_mn, b_mn), 0);
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::tc_fence_after();
    for (int c = 0; c < N; c += 8) {
        float d[8];
        tc::tmem_ld8(tlane + c, d); tc::tmem_ld_wait();
        for (int i = 0; i < 8; ++i) out[tid * N + c + i] = d[i];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}
}  // namespace pgm

extern "C" int pgm_tc_layout_probe(float *out, int M, int N, int a_mn, int b_mn, int fillA, int fillB, int RA, int RB,
                                   int lboA, int sboA, int lboB, int sboB, int d_lane_off, int ltA, int ltB, int a_tmem,
                                   int kind, int offA, int offB, void *stream) {
    PGM_REQUIRE(out && N % 8 == 0 && N <= 256, "pgm_tc_layout_probe: bad arguments");
    const size_t smem = 2 * 8192 * sizeof(float);
    PGM_CUDA(cudaFuncSetAttribute(pgm::tc_layout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pgm::tc_layout_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(out, M, N, a_mn, b_mn, fillA, fillB, RA, RB, (uint32_t)lboA,
                                                                 (uint32_t)sboA, (uint32_t)lboB, (uint32_t)sboB,
                                                                 (uint32_t)d_lane_off, (uint32_t)ltA, (uint32_t)ltB, a_tmem, kind,
                                                                 (uint32_t)offA, (uint32_t)offB);
    PGM_CUDA(cudaGetLastError());
    return PGM_OK;
}

// ------------------------------------------------------------------------------------------------------
// MMA pacing microbenchmark (profiles/tc_mma_bench.py): 48 kind::f16 MMAs of shape M x N x 16 from SWIZZLE_128B
// images (zeros), rotating over NACC accumulators; every descriptor input is warp-uniform (see tc::elect_one).
// out[2*rep] = cycles from the first issue to completion, out[2*rep+1] = cycles spent issuing.
namespace pgm {
template <int NACC>
__global__ void __launch_bounds__(128, 1) tc_mma_bench_kernel(float *out, int M, int N, int a_mn, int b_mn, int a_step, int b_step) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tc::uniform_warp_idx();
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    for (int i = tid; i < 49152; i += 128) sm[i] = 0.f;      // 192 KB of zero operands
    tc::fence_async_smem(); tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
    const uint32_t tmem = tc::uniform_u32(tmem_base_s);
    uint32_t phase = 0;
    const uint32_t aA = tc::smem_addr(sm), aB = tc::smem_addr(sm) + 98304;
    const uint64_t da = a_mn ? tc::make_desc(aA, 16384, 1024, 2) : tc::make_desc(aA, 16, 1024, 2);
    const uint64_t db = b_mn ? tc::make_desc(aB, 16384, 1024, 2) : tc::make_desc(aB, 16, 1024, 2);
    const uint32_t id = tc::idesc_f16(M, N, a_mn, b_mn);
    for (int rep = 0; rep < 3; ++rep) {
        tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
        const long long t0 = clock64();
        if (warp == 0 && tc::elect_one()) {
#pragma unroll
            for (int i = 0; i < 48; ++i)
                tc::mma_f16(tmem + (uint32_t)(i % NACC) * 128u, tc::desc_advance(da, (uint32_t)((i & 7) * a_step)),
                            tc::desc_advance(db, (uint32_t)((i & 7) * b_step)), id, 1);
            tc::mma_commit(&bar);
        }
        const long long t1 = clock64();
        tc::mbar_wait(&bar, phase); phase ^= 1;
        tc::tc_fence_after();
        const long long t2 = clock64();
        if (tid == 0) { out[2 * rep] = (float)(t2 - t0); out[2 * rep + 1] = (float)(t1 - t0); }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}
}  // namespace pgm

extern "C" int pgm_tc_mma_bench(float *out, int M, int N, int a_mn, int b_mn, int nmma, int nacc, int a_step, int b_step, void *stream) {
    PGM_REQUIRE(out && nmma == 48 && nacc >= 1 && nacc <= 3 && N * nacc <= 512, "pgm_tc_mma_bench: nmma must be 48, nacc 1..3");
    const size_t smem = 49152 * sizeof(float);
    auto go = [&](auto kern) -> int {
        PGM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<1, 128, smem, (cudaStream_t)stream>>>(out, M, N, a_mn, b_mn, a_step, b_step);
        PGM_CUDA(cudaGetLastError());
        return PGM_OK;
    };
    if (nacc == 1) return go(pgm::tc_mma_bench_kernel<1>);
    if (nacc == 2) return go(pgm::tc_mma_bench_kernel<2>);
    return go(pgm::tc_mma_bench_kernel<3>);
}
