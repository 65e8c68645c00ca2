// K1 on tensor cores: population-batched actor-critic forward over whole trajectories with the FP16-pair UMMA scheme of
// the K3 tensor-core kernel (k3_tc.cuh: same operand images, same precision argument; forward only).
// Included by k1_forward.cu (shares K1Args).
//
// One CTA = one (task, network half) and a contiguous range of 128-row tiles; 256 threads; tiles go through the pipeline
// in pairs so that an epilogue of one tile runs under the MMAs of the other:
//   obs (coalesced flat load, staged in shared memory, prefetched one pair ahead) -> X image pair (+ ones column)
//   G1 -> E1 tanh -> H1 pair -> G2 -> E2 tanh (+b2) -> H2 pair -> G3 (head, N = 16) -> E3: value | action, log-prob
// The critic half covers all rows_v rows, the actor half the rows_a rows that carry an action.
#pragma once
#include "tc_pair.cuh"

namespace pgm {

constexpr int K1T_THREADS = 256;
constexpr uint32_t K1T_TILE_BYTES = 81920;   // H1a | H1b | H2a | H2b | X (16 KB each)

struct K1tSmem { uint32_t tile[2], W2a, W2b, W1, WhA1, WhA2, WhZ, stage, bias, misc, total; };
__host__ __device__ inline K1tSmem k1t_smem_layout(int O) {
    K1tSmem s; uint32_t o = 0;
    s.tile[0] = o; o += K1T_TILE_BYTES; s.tile[1] = o; o += K1T_TILE_BYTES;
    s.W2a = o; o += 8192; s.W2b = o; o += 8192; s.W1 = o; o += 8192;
    s.WhA1 = o; o += 1024; s.WhA2 = o; o += 1024; s.WhZ = o; o += 1024;
    s.stage = o; o += ((uint32_t)(2 * 128 * O * 4) + 1023u) / 1024u * 1024u;      // fp32 obs of the two tiles
    s.bias = o; o += 512;                                                          // b2[64] bh[8] ls[8] ...
    s.misc = o; o += 128; s.total = o;
    return s;
}

template <int O, int A, int M>
__global__ void __launch_bounds__(K1T_THREADS, 1) k1_tc_kernel(const K1Args a) {
    constexpr int NK1 = (O + 1 + 15) / 16, NXC = (O + 1 + 7) / 8;
    constexpr int NLD = (128 * O + K1T_THREADS - 1) / K1T_THREADS;     // obs floats per thread per tile
    static_assert(O + 1 <= 24 && A <= 8 && M <= 8, "k1_tc: dims outside the tensor-core path");
    extern __shared__ __align__(1024) unsigned char smem_raw[];

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = tc::uniform_warp_idx();
    const int q = warp & 3, hcol = warp >> 2, g = hcol, r = tid & 127;
    const int task = blockIdx.z, half = blockIdx.y;
    const bool actor = half == 0;
    const int KH = actor ? A : M;
    const NetLayout &L = a.L;
    const K1tSmem sl = k1t_smem_layout(O);
    constexpr int ob1 = H * O, oW2 = ob1 + H, ob2 = oW2 + H * H, oWh = ob2 + H;

    const int rows = actor ? a.rows_a : a.rows_v;
    const int ntiles = (rows + 127) >> 7;
    const int t0 = blockIdx.x * a.chunks_per_cta, t1 = min(ntiles, t0 + a.chunks_per_cta);
    if (t0 >= t1) return;                                         // uniform per CTA, before any barrier / TMEM allocation

    __half *W2a = (__half *)(smem_raw + sl.W2a), *W2b = (__half *)(smem_raw + sl.W2b), *W1i = (__half *)(smem_raw + sl.W1);
    __half *WhA1 = (__half *)(smem_raw + sl.WhA1), *WhA2 = (__half *)(smem_raw + sl.WhA2);
    float *stage = (float *)(smem_raw + sl.stage);
    float *b2s = (float *)(smem_raw + sl.bias), *bhs = b2s + 64, *lss = b2s + 72;
    uint64_t *mbars = (uint64_t *)(smem_raw + sl.misc);
    uint32_t *tmem_ptr_s = (uint32_t *)(smem_raw + sl.misc + 32);

    for (int i = tid; i < (int)(sl.total / 16); i += K1T_THREADS) reinterpret_cast<float4 *>(smem_raw)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    if (warp == 0) tc::tmem_alloc(tmem_ptr_s, 128);
    if (tid == 0) { tc::mbar_init(mbars, 1); tc::mbar_init(mbars + 1, 1); tc::fence_mbar_init(); }
    {   // operand images of the half's weights (k3_tc.cuh layout)
        const float *gpar = a.params + (size_t)task * L.n_par;
        const int nW = oWh + KH * H;
        for (int e = tid; e < nW + KH + (actor ? A : 0); e += K1T_THREADS) {
            const float p = __ldg(gpar + L.to_global(half, e));
            if (e < ob1) { const int j = e / O, c = e - j * O; put_pair(W1i, sw128_hw(j, c), W1i, sw128_hw(j, 32 + c), p * TC_SW); }
            else if (e < oW2) { const int j = e - ob1; put_pair(W1i, sw128_hw(j, O), W1i, sw128_hw(j, 32 + O), p * TC_SW); }
            else if (e < ob2) { const int j = (e - oW2) >> 6, k = (e - oW2) & 63; put_pair(W2a, sw128_hw(j, k), W2b, sw128_hw(j, k), p * TC_SW); }
            else if (e < oWh) b2s[e - ob2] = p;
            else if (e < nW) { const int aa = (e - oWh) >> 6, k = (e - oWh) & 63; put_pair(WhA1, sw128_hw(aa, k), WhA2, sw128_hw(aa, k), p * TC_SW); }
            else if (e < nW + KH) bhs[e - nW] = p;
            else lss[e - nW - KH] = p;
        }
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tc::uniform_u32(*tmem_ptr_s);
    const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);

    const uint32_t aH1 = tc::smem_addr(smem_raw + sl.tile[0]), aH2 = aH1 + 32768, aX = aH1 + 65536;
    const uint32_t aW1 = tc::smem_addr(W1i), aW2a = tc::smem_addr(W2a), aW2b = tc::smem_addr(W2b);
    auto dK = [](uint32_t addr) { return tc::make_desc(addr, 16, 1024, 2); };
    const uint64_t dXa_k = dK(aX), dXb_k = dK(aX + 64), dW1a_k = dK(aW1), dW1b_k = dK(aW1 + 64);
    const uint64_t dH1a_k = dK(aH1), dH1b_k = dK(aH1 + 16384), dH2a_k = dK(aH2), dH2b_k = dK(aH2 + 16384);
    const uint64_t dW2a_k = dK(aW2a), dW2b_k = dK(aW2b);
    const uint64_t dWh1_k = tc::make_desc(tc::smem_addr(WhA1), 16, 2048, 2), dWh2_k = tc::make_desc(tc::smem_addr(WhA2), 16, 1024, 2);
    constexpr uint32_t ID_KK = tc::idesc_f16(128, 64, 0, 0), ID_HEAD = tc::idesc_f16(128, 16, 0, 0);
    const uint32_t swz = (uint32_t)(r & 7);
    uint32_t ph[2] = {0u, 0u};
    auto sync_all = [&]() { tc::tmem_st_wait(); tc::tmem_ld_wait(); tc::fence_async_smem(); tc::tc_fence_before(); __syncthreads(); };
    auto mma3 = [&](uint32_t d, uint64_t a1, uint64_t a2, uint64_t b1, uint64_t b2, uint32_t id, uint32_t acc) {
        tc::mma_f16(d, a2, b1, id, acc); tc::mma_f16(d, a1, b2, id, 1); tc::mma_f16(d, a1, b1, id, 1);
    };

    const float *obs = a.obs + (size_t)task * a.rows_v * O;
    const float *eps = a.eps ? a.eps + (a.eps_shared ? 0 : (size_t)task * a.rows_a * A) : nullptr;
    float *action = a.action ? a.action + (size_t)task * a.rows_a * A : nullptr;
    float *value = a.value + (size_t)task * a.rows_v * M;
    float *logp = a.logp ? a.logp + (size_t)task * a.rows_a : nullptr;

    // obs of a tile = 128 * O contiguous floats: flat coalesced loads, one pair of tiles ahead (clamped, masked at use)
    float xin[2][NLD];
    auto load_obs = [&](int tp) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int k = 0; k < NLD; ++k) {
                const long long e = (long long)(tp + i) * 128 * O + tid + K1T_THREADS * k;
                const long long lim = (long long)rows * O;
                xin[i][k] = ld_nc_f32(obs + (e < lim ? e : 0));
            }
    };
    load_obs(t0);

    for (int tp = t0; tp < t1; tp += 2) {
        const int nt = tp + 1 < t1 ? 2 : 1;
        const bool mine = g < nt;
        const int myrow = (tp + g) * 128 + r;
        const bool row_valid = mine && myrow < rows;
        // ---- stage the pair's observations, prefetch the next pair's ----
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int k = 0; k < NLD; ++k) {
                const int f = tid + K1T_THREADS * k;
                if (i < nt && f < 128 * O) stage[i * 128 * O + f] = ((long long)(tp + i) * 128 * O + f < (long long)rows * O) ? xin[i][k] : 0.f;
            }
        if (tp + 2 < t1) load_obs(tp + 2);
        // my row's noise / given action (actor)
        float r_in[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) r_in[d] = 0.f;
        if (actor && row_valid && a.mode != PGM_ACT_DETERMINISTIC) {
            const float *src = (a.mode == PGM_ACT_SAMPLE ? eps : action) + (size_t)myrow * A;
#pragma unroll
            for (int d = 0; d < A; ++d) r_in[d] = __ldg(src + d);
        }
        __syncthreads();
        // ---- X image pair: item = (row, 8-feature chunk) ----
        for (int it = tid; it < nt * 128 * NXC; it += K1T_THREADS) {
            const int i = it / (128 * NXC), rem = it - i * 128 * NXC, c = rem / 128, row = rem - c * 128;
            const bool ok = (tp + i) * 128 + row < rows;
            float xv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int f = 8 * c + e;
                xv[e] = f < O ? stage[i * 128 * O + row * O + f] : ((f == O && ok) ? 1.f : 0.f);
            }
            unsigned char *xr = smem_raw + sl.tile[i] + 65536 + row * 128;
            store_pair8_ovf(xr, (uint32_t)c, xr, (uint32_t)c + 4u, (uint32_t)(row & 7), xv);
        }
        sync_all();
        if (warp == 0 && tc::elect_one()) {
            tc::tc_fence_after();
            for (int i = 0; i < nt; ++i) {
                const uint32_t so = (uint32_t)i * K1T_TILE_BYTES, acc = tmem + (uint32_t)i * 64u;
#pragma unroll
                for (int ks = 0; ks < NK1; ++ks)
                    mma3(acc, tc::desc_advance(dXa_k, so + 32 * ks), tc::desc_advance(dXb_k, so + 32 * ks),
                         tc::desc_advance(dW1a_k, 32 * ks), tc::desc_advance(dW1b_k, 32 * ks), ID_KK, ks > 0);
                tc::mma_commit(mbars + i);
            }
        }
        // ---- E1 ----
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (i >= nt) break;
            unsigned char *rowh1 = smem_raw + sl.tile[i] + r * 128;
            tc::mbar_wait(mbars + i, ph[i]); ph[i] ^= 1; tc::tc_fence_after();
            float z[32];
            tc::tmem_ld32(tq + (uint32_t)i * 64u + 32 * hcol, z);
            tc::tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) z[k] = fast_tanh(z[k] * (1.f / TC_SW)) * TC_SH;
#pragma unroll
            for (int c = 0; c < 4; ++c) store_pair8(rowh1, (uint32_t)(4 * hcol + c), rowh1 + 16384, (uint32_t)(4 * hcol + c), swz, z + 8 * c);
            sync_all();
            if (warp == 0 && tc::elect_one()) {
                tc::tc_fence_after();
                const uint32_t so = (uint32_t)i * K1T_TILE_BYTES, acc = tmem + (uint32_t)i * 64u;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    mma3(acc, tc::desc_advance(dH1a_k, so + 32 * ks), tc::desc_advance(dH1b_k, so + 32 * ks),
                         tc::desc_advance(dW2a_k, 32 * ks), tc::desc_advance(dW2b_k, 32 * ks), ID_KK, ks > 0);
                tc::mma_commit(mbars + i);
            }
        }
        // ---- E2 ----
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (i >= nt) break;
            unsigned char *rowh2 = smem_raw + sl.tile[i] + 32768 + r * 128;
            tc::mbar_wait(mbars + i, ph[i]); ph[i] ^= 1; tc::tc_fence_after();
            float z[32];
            tc::tmem_ld32(tq + (uint32_t)i * 64u + 32 * hcol, z);
            tc::tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
                const float4 bv = *reinterpret_cast<const float4 *>(b2s + 32 * hcol + k);
                z[k] = fmaf(z[k], 1.f / (TC_SH * TC_SW), bv.x); z[k + 1] = fmaf(z[k + 1], 1.f / (TC_SH * TC_SW), bv.y);
                z[k + 2] = fmaf(z[k + 2], 1.f / (TC_SH * TC_SW), bv.z); z[k + 3] = fmaf(z[k + 3], 1.f / (TC_SH * TC_SW), bv.w);
            }
#pragma unroll
            for (int k = 0; k < 32; ++k) z[k] = fast_tanh(z[k]) * TC_SH;
#pragma unroll
            for (int c = 0; c < 4; ++c) store_pair8(rowh2, (uint32_t)(4 * hcol + c), rowh2 + 16384, (uint32_t)(4 * hcol + c), swz, z + 8 * c);
            sync_all();
            if (warp == 0 && tc::elect_one()) {
                tc::tc_fence_after();
                const uint32_t so = (uint32_t)i * K1T_TILE_BYTES, acc = tmem + (uint32_t)i * 64u;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    mma3(acc, tc::desc_advance(dH2a_k, so + 32 * ks), tc::desc_advance(dH2b_k, so + 32 * ks),
                         tc::desc_advance(dWh1_k, 32 * ks), tc::desc_advance(dWh2_k, 32 * ks), ID_HEAD, ks > 0);
                tc::mma_commit(mbars + i);
            }
        }
        // ---- E3: head output of my row (group g owns tile g) ----
        if (mine) {
            tc::mbar_wait(mbars + g, g == 0 ? ph[0] : ph[1]); tc::tc_fence_after();
            float ho[8];
            tc::tmem_ld8(tq + (uint32_t)g * 64u, ho);
            tc::tmem_ld_wait();
            if (row_valid) {
                if (actor) {
                    float lp = 0.f;
#pragma unroll
                    for (int d = 0; d < 8; ++d) {
                        if (d < A) {
                            const float mean = fmaf(ho[d], 1.f / (TC_SH * TC_SW), bhs[d]);
                            const float ls = lss[d], sd = expf(ls);
                            float act;
                            if (a.mode == PGM_ACT_SAMPLE) act = fmaf(sd, r_in[d], mean);
                            else if (a.mode == PGM_ACT_DETERMINISTIC) act = mean;
                            else act = r_in[d];
                            if (a.mode != PGM_ACT_EVALUATE) action[(size_t)myrow * A + d] = act;
                            const float diff = act - mean;
                            lp += -(diff * diff) / (2.f * sd * sd) - ls - 0.91893853320467274178f;
                        }
                    }
                    logp[myrow] = lp;
                } else {
#pragma unroll
                    for (int m = 0; m < 8; ++m)
                        if (m < M) value[(size_t)myrow * M + m] = fmaf(ho[m], 1.f / (TC_SH * TC_SW), bhs[m]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) if (i < nt) ph[i] ^= 1;      // both head GEMMs are complete once the next barrier is passed
        tc::tmem_ld_wait();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
    }
    if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

}  // namespace pgm
