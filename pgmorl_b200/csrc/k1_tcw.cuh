// K1 on tensor cores for WIDE observations (Humanoid, O = 376): bulk rollout inference with the FP16-pair UMMA scheme
// (k3_tc.cuh / k3_tcw.cuh), forward only. Included by k1_forward.cu after k1_tc.cuh (shares K1Args).
//
// One CTA = one (task, network half) and a contiguous range of 128-row tiles. The half's W1 stays resident as six
// 64-feature FP16-pair block images (96 KB); the observations of a tile are read as FP32 rows (16 consecutive threads
// read 256 contiguous bytes), split into pairs in registers and written into two alternating 32 KB block stages while
// the MMAs of the previous block run:
//   per tile:  6 x { x block -> stage, 12 MMAs }  ->  E1 tanh -> H pair -> G2 -> E2 tanh (+b2) -> H pair (same buffer)
//              -> G3 (head, N = 32) -> E3: value | action, log-prob (warps 0..3, thread = row)
// The critic half covers all rows_v rows, the actor half the rows_a rows that carry an action.
#pragma once
#include "tc_pair.cuh"

namespace pgm {

struct K1wSmem { uint32_t W1, W2a, W2b, Wh1, Wh2, Hb, XS, bias, misc, total; };
__host__ __device__ inline K1wSmem k1w_smem_layout() {
    K1wSmem s; uint32_t o = 0;
    s.W1 = o; o += 6 * 16384;                           // [6 blocks][a1 8 KB | a2 8 KB], [64 rows j][64 halfwords]
    s.W2a = o; o += 8192; s.W2b = o; o += 8192;
    s.Wh1 = o; o += 4096; s.Wh2 = o; o += 4096;         // [32 rows a][64 halfwords k]
    s.Hb = o; o += 32768;                               // h1, then h2: a1 | a2
    s.XS = o; o += 2 * 32768;                           // two x block stages: a1 | a2
    s.bias = o; o += 512;                               // b2[64] | bh[32] | logstd[32]
    s.misc = o; o += 128; s.total = o;
    return s;
}

template <int O, int A, int M>
__global__ void __launch_bounds__(K1T_THREADS, 1) k1_tcw_kernel(const K1Args a) {
    constexpr int NB = (O + 1 + 63) / 64;
    static_assert(O % 4 == 0 && NB == 6 && A <= 24 && M <= 8, "k1_tcw: dims outside the wide tensor-core path");
    extern __shared__ __align__(1024) unsigned char smem_raw[];

    const int tid = threadIdx.x;
    const int warp = tc::uniform_warp_idx();
    const int q = warp & 3, hcol = warp >> 2, r = tid & 127;
    const int task = blockIdx.z, half = blockIdx.y;
    const bool actor = half == 0;
    const int KH = actor ? A : M;
    const NetLayout &L = a.L;
    const K1wSmem sl = k1w_smem_layout();
    constexpr int ob1 = H * O, oW2 = ob1 + H, ob2 = oW2 + H * H, oWh = ob2 + H;

    const int rows = actor ? a.rows_a : a.rows_v;
    const int ntiles = (rows + 127) >> 7;
    const int t0 = blockIdx.x * a.chunks_per_cta, t1 = min(ntiles, t0 + a.chunks_per_cta);
    if (t0 >= t1) return;                                         // uniform per CTA, before any barrier / TMEM allocation

    __half *W1i = (__half *)(smem_raw + sl.W1);
    __half *W2a = (__half *)(smem_raw + sl.W2a), *W2b = (__half *)(smem_raw + sl.W2b);
    __half *Wh1 = (__half *)(smem_raw + sl.Wh1), *Wh2 = (__half *)(smem_raw + sl.Wh2);
    unsigned char *S_h = smem_raw + sl.Hb;
    float *b2s = (float *)(smem_raw + sl.bias), *bhs = b2s + 64, *lss = b2s + 96;
    uint64_t *mbars = (uint64_t *)(smem_raw + sl.misc);           // 0 chain, 1-2 x stages
    uint32_t *tmem_ptr_s = (uint32_t *)(smem_raw + sl.misc + 32);

    for (int i = tid; i < (int)(sl.total / 16); i += K1T_THREADS) reinterpret_cast<float4 *>(smem_raw)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    if (warp == 0) tc::tmem_alloc(tmem_ptr_s, 64);
    if (tid == 0) { for (int i = 0; i < 3; ++i) tc::mbar_init(mbars + i, 1); tc::fence_mbar_init(); }
    {   // operand images of the half's weights
        const float *gpar = a.params + (size_t)task * L.n_par;
        const int nW = oWh + KH * H;
        for (int e = tid; e < nW + KH + (actor ? A : 0); e += K1T_THREADS) {
            const float p = __ldg(gpar + L.to_global(half, e));
            if (e < oW2) {                                        // W1[j][c], b1[j] = column O
                const int j = e < ob1 ? e / O : e - ob1, c = e < ob1 ? e - j * O : O;
                __half *img = W1i + (c >> 6) * 8192;
                put_pair(img, sw128_hw(j, c & 63), img + 4096, sw128_hw(j, c & 63), p * TC_SW);
            }
            else if (e < ob2) { const int j = (e - oW2) >> 6, k = (e - oW2) & 63; put_pair(W2a, sw128_hw(j, k), W2b, sw128_hw(j, k), p * TC_SW); }
            else if (e < oWh) b2s[e - ob2] = p;
            else if (e < nW) { const int aa = (e - oWh) >> 6, k = (e - oWh) & 63; put_pair(Wh1, sw128_hw(aa, k), Wh2, sw128_hw(aa, k), p * TC_SW); }
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

    const uint32_t aW1 = tc::smem_addr(W1i), aH = tc::smem_addr(S_h), aXS = tc::smem_addr(smem_raw + sl.XS);
    auto dK = [](uint32_t addr) { return tc::make_desc(addr, 16, 1024, 2); };
    const uint64_t dHa_k = dK(aH), dHb_k = dK(aH + 16384);
    const uint64_t dW2a_k = dK(tc::smem_addr(W2a)), dW2b_k = dK(tc::smem_addr(W2b));
    const uint64_t dWh1_k = dK(tc::smem_addr(Wh1)), dWh2_k = dK(tc::smem_addr(Wh2));
    constexpr uint32_t ID_KK = tc::idesc_f16(128, 64, 0, 0), ID_HEAD = tc::idesc_f16(128, 32, 0, 0);
    const uint32_t swz = (uint32_t)(r & 7);
    uint32_t ph[3] = {0u, 0u, 0u};
    auto sync_all = [&]() { tc::tmem_st_wait(); tc::tmem_ld_wait(); tc::fence_async_smem(); tc::tc_fence_before(); __syncthreads(); };
    auto mma3 = [&](uint32_t d, uint64_t a1, uint64_t a2, uint64_t b1, uint64_t b2, uint32_t id, uint32_t acc) {
        tc::mma_f16(d, a2, b1, id, acc); tc::mma_f16(d, a1, b2, id, 1); tc::mma_f16(d, a1, b1, id, 1);
    };

    const float *obs = a.obs + (size_t)task * a.rows_v * O;
    const float *eps = a.eps ? a.eps + (a.eps_shared ? 0 : (size_t)task * a.rows_a * A) : nullptr;
    float *action = a.action ? a.action + (size_t)task * a.rows_a * A : nullptr;
    float *value = a.value + (size_t)task * a.rows_v * M;
    float *logp = a.logp ? a.logp + (size_t)task * a.rows_a : nullptr;

    // x block (tile t, block b): item k of a thread = (row (tid >> 4) + 16 k, features 64 b + 4 (tid & 15) .. + 3)
    float4 xin[8];
    const int pc = tid & 15;
    auto load_x = [&](int t, int b) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int row = t * 128 + (tid >> 4) + 16 * k, f = 64 * b + 4 * pc;
            const bool ok = row < rows && f < O;                 // O % 4 == 0: a float4 never straddles the end of a row
            xin[k] = ld_nc_f4(reinterpret_cast<const float4 *>(obs + (ok ? (size_t)row * O + f : 0)));
            if (!ok) xin[k] = make_float4((f == O && row < rows) ? 1.f : 0.f, 0.f, 0.f, 0.f);      // ones column, zero padding
        }
    };
    auto store_x = [&](int st) {
        unsigned char *img = smem_raw + sl.XS + st * 32768;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int row = (tid >> 4) + 16 * k;
            const float4 v = xin[k];
            const float h0 = round11(v.x), h1 = round11(v.y), h2 = round11(v.z), h3 = round11(v.w);
            unsigned char *d = img + row * 128 + ((uint32_t)((pc >> 1) ^ (row & 7)) << 4) + (pc & 1) * 8;
            *reinterpret_cast<uint2 *>(d) = make_uint2(pack_h2_ovf(h0, h1), pack_h2_ovf(h2, h3));
            *reinterpret_cast<uint2 *>(d + 16384) = make_uint2(pack_h2_ovf(v.x - h0, v.y - h1), pack_h2_ovf(v.z - h2, v.w - h3));
        }
    };
    load_x(t0, 0);
    int nblk = 0;                                                 // x blocks staged so far by this CTA

    for (int t = t0; t < t1; ++t) {
        const int myrow = t * 128 + r;
        const bool row_valid = hcol == 0 && myrow < rows;
        // ---------------- G1 over six blocks ----------------
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int st = b & 1;
            if (nblk >= 2) { tc::mbar_wait(mbars + 1 + st, ph[1 + st]); ph[1 + st] ^= 1u; }     // MMAs of the block two back are done
            ++nblk;
            store_x(st);
            if (b + 1 < NB) load_x(t, b + 1); else if (t + 1 < t1) load_x(t + 1, 0);
            tc::tmem_ld_wait();                                   // (b == 0) the previous tile's head output has been read
            tc::fence_async_smem();
            tc::tc_fence_before();
            __syncthreads();
            if (warp == 0 && tc::elect_one()) {
                tc::tc_fence_after();
                const uint64_t xa = dK(aXS + st * 32768), xb = dK(aXS + st * 32768 + 16384);
                const uint64_t wa = dK(aW1 + b * 16384), wb = dK(aW1 + b * 16384 + 8192);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    mma3(tmem, tc::desc_advance(xa, 32 * ks), tc::desc_advance(xb, 32 * ks),
                         tc::desc_advance(wa, 32 * ks), tc::desc_advance(wb, 32 * ks), ID_KK, (b > 0 || ks > 0) ? 1u : 0u);
                tc::mma_commit(mbars + 1 + st);
                if (b == NB - 1) tc::mma_commit(mbars + 0);
            }
        }
        // ---------------- E1 ----------------
        {
            unsigned char *rowh = S_h + r * 128;
            tc::mbar_wait(mbars + 0, ph[0]); ph[0] ^= 1u; tc::tc_fence_after();
            float z[32];
            tc::tmem_ld32(tq + 32 * hcol, z);
            tc::tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) z[k] = fast_tanh(z[k] * (1.f / TC_SW)) * TC_SH;
#pragma unroll
            for (int c = 0; c < 4; ++c) store_pair8(rowh, (uint32_t)(4 * hcol + c), rowh + 16384, (uint32_t)(4 * hcol + c), swz, z + 8 * c);
            sync_all();
            if (warp == 0 && tc::elect_one()) {
                tc::tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    mma3(tmem, tc::desc_advance(dHa_k, 32 * ks), tc::desc_advance(dHb_k, 32 * ks),
                         tc::desc_advance(dW2a_k, 32 * ks), tc::desc_advance(dW2b_k, 32 * ks), ID_KK, ks > 0);
                tc::mma_commit(mbars + 0);
            }
        }
        // ---------------- E2 (h2 overwrites h1: G2 is complete) ----------------
        {
            unsigned char *rowh = S_h + r * 128;
            tc::mbar_wait(mbars + 0, ph[0]); ph[0] ^= 1u; tc::tc_fence_after();
            float z[32];
            tc::tmem_ld32(tq + 32 * hcol, z);
            tc::tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) z[k] = fast_tanh(fmaf(z[k], 1.f / (TC_SH * TC_SW), b2s[32 * hcol + k])) * TC_SH;
#pragma unroll
            for (int c = 0; c < 4; ++c) store_pair8(rowh, (uint32_t)(4 * hcol + c), rowh + 16384, (uint32_t)(4 * hcol + c), swz, z + 8 * c);
            sync_all();
            if (warp == 0 && tc::elect_one()) {
                tc::tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    mma3(tmem, tc::desc_advance(dHa_k, 32 * ks), tc::desc_advance(dHb_k, 32 * ks),
                         tc::desc_advance(dWh1_k, 32 * ks), tc::desc_advance(dWh2_k, 32 * ks), ID_HEAD, ks > 0);
                tc::mma_commit(mbars + 0);
            }
        }
        // my row's noise / given action (actor), in flight under the head GEMM
        float r_in[24];
#pragma unroll
        for (int d = 0; d < 24; ++d) r_in[d] = 0.f;
        if (actor && row_valid && a.mode != PGM_ACT_DETERMINISTIC) {
            const float *src = (a.mode == PGM_ACT_SAMPLE ? eps : action) + (size_t)myrow * A;
#pragma unroll
            for (int d = 0; d < A; ++d) r_in[d] = __ldg(src + d);
        }
        // ---------------- E3: head output of my row ----------------
        tc::mbar_wait(mbars + 0, ph[0]); ph[0] ^= 1u; tc::tc_fence_after();
        if (hcol == 0) {
            float ho[32];
            tc::tmem_ld32(tq, ho);
            tc::tmem_ld_wait();
            if (row_valid) {
                if (actor) {
                    float lp = 0.f;
#pragma unroll
                    for (int d = 0; d < 24; ++d) {
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
    }
    tc::tmem_ld_wait();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 64);
}

}  // namespace pgm
