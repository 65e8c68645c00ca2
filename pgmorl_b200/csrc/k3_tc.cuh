// K3 tensor-core path: the PPO minibatch loop (a2c/algo/ppo.py:62-107) on tcgen05 (UMMA, FP32 accumulate in TMEM).
// Included by k3_ppo.cu (shares K3Args, k3_pack_kernel, fast_tanh, the DSMEM helpers).
//
// One cluster of two CTAs per task: rank 0 = actor, rank 1 = critic (they share only the global gradient norm).
// A CTA keeps its half's weights resident in shared memory as UMMA B operands and runs TWO independent 128-thread
// pipelines ("groups"); group g owns the 128-row tiles t = g, g+2, ... of every minibatch, thread = row = TMEM lane.
// While one group waits for its MMAs the other runs its epilogue, which hides the issue->complete latency of the
// dependent GEMM chain.  Per tile:
//   gather record -> x (tf32 hi|lo) to TMEM, x (fp16 a1|a2) to smem
//   G1  Z1 = x W1^T (+b1 through a ones column)   kind::tf32, A: TMEM hi/lo, B: smem hi/lo, 3 MMAs per K step
//   E1  h1 = tanh(Z1) -> TMEM hi/lo (operand of G2) and smem fp16 pair (operand of the weight-gradient GEMMs)
//   G2/E2 likewise for layer 2;  G3  head = h2 Wh^T (N = 16)
//   E3  per-row loss, d loss / d head -> smem            (thread owns its row: no cross-thread reduction)
//   G4  dz2pre = dOut Wh;  GWh  dWh^T += h2^T dOut       kind::f16 from here on, 3 MMAs per K step
//   E4  dz2 = dz2pre (1 - h2^2), h2 = hi + lo exact from TMEM
//   G5  dz1pre = dz2 W2;   GW2  dW2 += dz2^T h1
//   E5  dz1 = dz1pre (1 - h1^2)
//   G1X [dW1 | db1 ; db2] += [dz1 | dz2]^T [x | 1]
// Precision. Forward: 3-way TF32 split (hi*hi + lo*hi + hi*lo), measured 1.5e-7 like an FP32 FMA chain. Backward:
// every operand is an FP16 PAIR a = a1 + a2 (a1 = fp16(a), a2 = fp16(a - a1): 22 significant bits) held in
// power-of-two scaled form so that neither part underflows (activations x 2^8, backward signals x 2^12, backward
// weights x 2^8; saturating conversions); a1*b1 + a1*b2 + a2*b1 reproduces the FP32 product to ~2^-21.  A pair
// costs 4 bytes per element -- the footprint of ONE tf32 copy -- and, unlike tf32 (whose MN-major operands exist
// only in a special 32-byte-swizzle layout), one [row][64 halfwords] SWIZZLE_128B image serves both as the
// K-major A operand of G4/G5 and as the MN-major operand of the weight-gradient GEMMs (contraction over rows).
// Step tail: gradients TMEM -> registers of their owner threads, squared-norm exchange with the peer CTA through
// DSMEM (one cluster barrier), clip + Adam (moments in an L2-resident workspace, thread-owned float4 slots), new
// weights re-split into the operand images.
#pragma once
#include <cuda_fp16.h>

#include "tc.cuh"

namespace pgm {

constexpr int TC_SLOT4 = 15;                 // float4 parameter slots per thread
constexpr int TC_THREADS = 256;
constexpr uint32_t TC_GROUP_BYTES = 81920;   // H1a | H1b | H2a | H2b | X   (16 KB each, [128 rows][64 halfwords])
constexpr uint32_t TC_MISC_BYTES = 1280;
constexpr float TC_SH = 256.f, TC_SD = 4096.f, TC_SW = 256.f;   // scales of activations / backward signals / backward weights

struct TcSmem {   // byte offsets inside dynamic shared memory
    uint32_t grp[2], W2h, W2l, W2Ta, W2Tb, Whh, Whl, WhTz, WhTa, WhTb, W1h, W1l, misc, total;
};
__host__ __device__ inline int tc_nch1(int O) { return (O + 1 + 3) / 4; }   // allocated 16-byte K chunks per W1 image
__host__ __device__ inline TcSmem tc_smem_layout(int O) {
    TcSmem s; uint32_t o = 0;
    s.grp[0] = o; o += TC_GROUP_BYTES; s.grp[1] = o; o += TC_GROUP_BYTES;
    s.W2h = o; o += 16384; s.W2l = o; o += 16384;
    s.W2Ta = o; o += 8192; s.W2Tb = o; o += 8192;          // [64 rows k][64 halfwords j], SWIZZLE_128B
    s.Whh = o; o += 2048; s.Whl = o; o += 2048;
    s.WhTz = o; o += 1024; s.WhTa = o; o += 1024; s.WhTb = o; o += 1024;   // K chunk of zeros | a1 | a2, [64 rows k][8 halfwords a]
    // The last K step of G1 may read one chunk past each W1 image (multiplied by the zero padding of x): what
    // follows must hold finite floats -> W1l follows W1h, the float part of `misc` follows W1l.
    s.W1h = o; o += tc_nch1(O) * 1024; s.W1l = o; o += tc_nch1(O) * 1024;
    s.misc = o; o += TC_MISC_BYTES; s.total = o;
    return s;
}
// misc (floats): b2[64] bh[8] ls[8] red[40] part[8][16] ssq[4] | at byte 1024: double sh_d[4], mbarriers, tmem ptr
constexpr int TCM_B2 = 0, TCM_BH = 64, TCM_LS = 72, TCM_RED = 80, TCM_PART = 120, TCM_SSQ = 248;

// TMEM columns: group g at g*192: ACT [0,128) A operands of the forward GEMMs, ACC [128,192) accumulators;
// weight-gradient accumulators shared by both groups
constexpr uint32_t TC_ACT = 0, TC_ACC = 128, TC_GSTRIDE = 192, TC_GW2 = 384, TC_G1X = 448, TC_GWH = 480;

__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {   // two floats -> fp16x2 (a in the low half), saturating
    uint32_t r; asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r;
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) {
    float2 r;
    asm("{ .reg .b16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h; }" : "=f"(r.x), "=f"(r.y) : "r"(v));
    return r;
}

// fp16 pair of (v0, v1): p1 = fp16x2 of the values, p2 = fp16x2 of the residuals
__device__ __forceinline__ void split_h2(float v0, float v1, uint32_t &p1, uint32_t &p2) {
    p1 = pack_h2(v0, v1);
    const float2 f = unpack_h2(p1);
    p2 = pack_h2(v0 - f.x, v1 - f.y);
}
// store 8 consecutive features (chunk c of 16 bytes) of row r into the a1 / a2 images of a [row][64 halfwords] buffer
// (logical chunks ca / cb; swz8 = row & 7 is the SWIZZLE_128B XOR)
__device__ __forceinline__ void store_pair8(unsigned char *rowa, uint32_t ca, unsigned char *rowb, uint32_t cb, uint32_t swz8, const float *v) {
    uint4 p1, p2;
    split_h2(v[0], v[1], p1.x, p2.x); split_h2(v[2], v[3], p1.y, p2.y);
    split_h2(v[4], v[5], p1.z, p2.z); split_h2(v[6], v[7], p1.w, p2.w);
    *reinterpret_cast<uint4 *>(rowa + ((ca ^ swz8) << 4)) = p1;
    *reinterpret_cast<uint4 *>(rowb + ((cb ^ swz8) << 4)) = p2;
}
// halfword index of (row, feature) inside a [rows][64 halfwords] SWIZZLE_128B image
__device__ __forceinline__ int sw128_hw(int row, int f) { return row * 64 + ((((f >> 3) ^ (row & 7))) << 3) + (f & 7); }

// half-local parameter index (reference order inside the half: W1, b1, W2, b2, head W, head b, logstd) owned by
// slot s of thread (q = lane quadrant, h = column half, lane), or -1.
__device__ __forceinline__ int tc_own(int s, int q, int h, int lane, int tid, int O, int KH, int A, bool actor) {
    const int oW2 = H * O + H, ob2 = oW2 + H * H, oWh = ob2 + H, obh = oWh + KH * H, ols = obh + KH;
    if (s < 32) { if (lane >= 16) return -1; return oW2 + (16 * q + lane) * H + 32 * h + s; }
    if (s < 48) {
        const int c = 16 * h + (s - 32);
        if (q < 2) { const int j = 32 * q + lane; return c < O ? j * O + c : (c == O ? H * O + j : -1); }
        return c == O ? ob2 + 32 * (q - 2) + lane : -1;
    }
    if (s < 56) { if (h != 0 || lane >= 16) return -1; const int a = s - 48; return a < KH ? oWh + a * H + 16 * q + lane : -1; }
    if (s == 56) return tid < KH ? obh + tid : -1;
    if (s == 57) return (actor && tid < A) ? ols + tid : -1;
    return -1;
}

template <int O, int A, int M>
__global__ void __launch_bounds__(TC_THREADS, 1) k3_tc_kernel(const K3Args a) {
    constexpr int OP = (O + 3) / 4 * 4;
    constexpr int KX = (O + 1 + 7) / 8 * 8;          // layer-1 contraction: x, ones column at index O, zero padding
    constexpr int NKX = KX / 8;
    constexpr int RSG = (OP + A + 2 * M + 2 + 3) / 4 * 4;
    constexpr int NCH1 = (O + 1 + 3) / 4;
    static_assert(KX <= 24 && A <= 8 && M <= 8 && RSG <= 32, "k3_tc: dims outside the tensor-core path");
    extern __shared__ __align__(1024) unsigned char smem_raw[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = tid >> 7, r = tid & 127;           // group, row inside the tile (= TMEM lane)
    const int q = warp & 3, hcol = warp >> 2;        // lane quadrant, column half used in the step tail
    const int task = blockIdx.x >> 1;
    const unsigned rank = group_rank<2>();
    const bool actor = rank == 0;
    const int half = (int)rank;
    const int KH = actor ? A : M;
    const NetLayout &L = a.L;
    const TcSmem sl = tc_smem_layout(O);

    unsigned char *sg = smem_raw + sl.grp[g];
    unsigned char *S_h1 = sg, *S_h2 = sg + 32768, *S_x = sg + 65536;     // each: a1 image | a2 image (X: one image)
    float *W2h = (float *)(smem_raw + sl.W2h), *W2l = (float *)(smem_raw + sl.W2l);
    __half *W2Ta = (__half *)(smem_raw + sl.W2Ta), *W2Tb = (__half *)(smem_raw + sl.W2Tb);
    float *Whh = (float *)(smem_raw + sl.Whh), *Whl = (float *)(smem_raw + sl.Whl);
    __half *WhTa = (__half *)(smem_raw + sl.WhTa), *WhTb = (__half *)(smem_raw + sl.WhTb);
    float *W1h = (float *)(smem_raw + sl.W1h), *W1l = (float *)(smem_raw + sl.W1l);
    float *misc = (float *)(smem_raw + sl.misc);
    float *b2s = misc + TCM_B2, *bhs = misc + TCM_BH, *lss = misc + TCM_LS, *red = misc + TCM_RED;
    float *part = misc + TCM_PART, *ssqS = misc + TCM_SSQ;
    double *sh_d = (double *)(smem_raw + sl.misc + 1024);
    uint64_t *mbars = (uint64_t *)(smem_raw + sl.misc + 1024 + 32);     // [0..1] chain (A), [2..3] weight grads (B)
    uint32_t *tmem_ptr_s = (uint32_t *)(smem_raw + sl.misc + 1024 + 64);
    uint64_t *mbA = mbars + g, *mbB = mbars + 2 + g;

    // ---------------- one-time setup ----------------
    for (int i = tid; i < (int)(TC_MISC_BYTES / 4); i += TC_THREADS) misc[i] = 0.f;
    for (int i = tid; i < (int)(2 * TC_GROUP_BYTES / 16); i += TC_THREADS)
        reinterpret_cast<float4 *>(smem_raw)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < 1024 / 16; i += TC_THREADS)
        reinterpret_cast<float4 *>(smem_raw + sl.WhTz)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    if (warp == 0) tc::tmem_alloc(tmem_ptr_s, 512);
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) tc::mbar_init(mbars + i, 1);
        tc::fence_mbar_init();
    }
    const float *gpar = a.params + (size_t)task * L.n_par;
    const float *pbase = gpar + L.half_base(half), *phead = gpar + L.half_head(half);
    {   // operand images of the half's weights
        const float *gW1 = pbase, *gb1 = pbase + H * O, *gW2 = gb1 + H, *gb2 = gW2 + H * H;
        for (int i = tid; i < H * NCH1 * 4; i += TC_THREADS) {            // W1 image: row j, feature f (f == O: bias)
            const int j = i / (NCH1 * 4), f = i - j * (NCH1 * 4);
            const float v = f < O ? __ldg(gW1 + j * O + f) : (f == O ? __ldg(gb1 + j) : 0.f);
            const float hi = tc::tf32_hi(v);
            const int w = (f >> 2) * 256 + j * 4 + (f & 3);
            W1h[w] = hi; W1l[w] = v - hi;
        }
        for (int i = tid; i < H * H; i += TC_THREADS) {
            const int j = i >> 6, k = i & 63;
            const float v = __ldg(gW2 + i), hi = tc::tf32_hi(v);
            const int w = (k >> 2) * 256 + j * 4 + (k & 3);
            W2h[w] = hi; W2l[w] = v - hi;
            const __half t1 = __float2half_rn(v * TC_SW);                   // backward copy: row k, feature j, fp16 pair
            W2Ta[sw128_hw(k, j)] = t1; W2Tb[sw128_hw(k, j)] = __float2half_rn(v * TC_SW - __half2float(t1));
        }
        for (int i = tid; i < 8 * H; i += TC_THREADS) {                    // head: rows a < 8 (zero beyond KH)
            const int aa = i >> 6, k = i & 63;
            const float v = aa < KH ? __ldg(phead + aa * H + k) : 0.f, hi = tc::tf32_hi(v);
            const int w = (k >> 2) * 32 + aa * 4 + (k & 3);
            Whh[w] = hi; Whl[w] = v - hi;
            const __half t1 = __float2half_rn(v * TC_SW);                   // backward copy: row k, 8 halfwords a
            WhTa[k * 8 + aa] = t1; WhTb[k * 8 + aa] = __float2half_rn(v * TC_SW - __half2float(t1));
        }
        for (int i = tid; i < H; i += TC_THREADS) b2s[i] = __ldg(gb2 + i);
        for (int i = tid; i < 8; i += TC_THREADS) {
            bhs[i] = i < KH ? __ldg(phead + KH * H + i) : 0.f;
            lss[i] = (actor && i < A) ? __ldg(phead + KH * H + KH + i) : 0.f;
        }
    }
    // Adam moments: reference order -> thread-owned slots in the workspace; validity mask of my slots
    float4 *mv4 = reinterpret_cast<float4 *>(a.mv) + ((size_t)(task * 2 + half) * 2) * TC_SLOT4 * TC_THREADS;
    float4 *mM = mv4, *mV = mv4 + TC_SLOT4 * TC_THREADS;
    unsigned long long vmask = 0ull;
    for (int s4 = 0; s4 < TC_SLOT4; ++s4) {
        float mm[4], vv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int e = tc_own(4 * s4 + c, q, hcol, lane, tid, O, KH, A, actor);
            mm[c] = 0.f; vv[c] = 0.f;
            if (e >= 0) {
                vmask |= 1ull << (4 * s4 + c);
                if (!a.grad_only) {
                    const size_t gi = (size_t)task * L.n_par + L.to_global(half, e);
                    mm[c] = a.adam_m[gi]; vv[c] = a.adam_v[gi];
                }
            }
        }
        if (!a.grad_only) {
            mM[s4 * TC_THREADS + tid] = make_float4(mm[0], mm[1], mm[2], mm[3]);
            mV[s4 * TC_THREADS + tid] = make_float4(vv[0], vv[1], vv[2], vv[3]);
        }
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    sync_group<2>();                 // barrier inits + TMEM address visible; peer CTA is alive before any DSMEM store
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_ptr_s;
    const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);          // my lane quadrant
    const uint32_t tg = tq + (uint32_t)g * TC_GSTRIDE;              // + my group's column window

    {   // zero the weight-gradient accumulators (each thread: the cells it will read in the step tail)
        float z[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) z[i] = 0.f;
        tc::tmem_st32(tq + TC_GW2 + 32 * hcol, z);
        tc::tmem_st16(tq + TC_G1X + 16 * hcol, z);
        if (hcol == 0) tc::tmem_st8(tq + TC_GWH, z);
        tc::tmem_st_wait();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();

    // ---------------- constants ----------------
    const float clip = (float)a.hy.clip_param;
    const float inv_mb = 1.f / (float)a.mb;
    const float vscale = (float)(a.hy.value_loss_coef * 0.5 / ((double)a.mb * M));
    const float omb1 = (float)(1.0 - a.hy.beta1);
    const float b2f = (float)a.hy.beta2, omb2 = (float)(1.0 - a.hy.beta2);
    const float aeps = (float)a.hy.adam_eps;
    const float ecoef = (float)a.hy.entropy_coef;
    const int step0 = a.grad_only ? 0 : a.adam_step[task];
    double b1pow = pow(a.hy.beta1, (double)step0), b2pow = pow(a.hy.beta2, (double)step0);   // used by thread 0
    const double lr = a.grad_only ? 0.0 : a.lr[task];
    float loss_act = 0.f, loss_val = 0.f, loss_ent = 0.f;

    const int32_t *perm = a.perm + ((a.perm_shared || a.grad_only) ? 0 : (size_t)task * a.E * a.S);
    const float *recg = a.rec + (size_t)task * a.S * RSG;
    const int ntiles = (a.mb + 127) >> 7;
    uint32_t phA = 0, phB = 0;
    bool pendB = false;
    const bool issuer = r == 0;

    // descriptors (constant for the whole launch)
    const uint64_t dW1h = tc::desc_kmajor(tc::smem_addr(W1h), 64), dW1l = tc::desc_kmajor(tc::smem_addr(W1l), 64);
    const uint64_t dW2h = tc::desc_kmajor(tc::smem_addr(W2h), 64), dW2l = tc::desc_kmajor(tc::smem_addr(W2l), 64);
    // backward B operands (fp16): W2^T as a K-major SWIZZLE_128B image; Wh^T as [zero chunk | data chunk], K = 16
    const uint64_t dW2Ta = tc::make_desc(tc::smem_addr(W2Ta), 16, 1024, 2), dW2Tb = tc::make_desc(tc::smem_addr(W2Tb), 16, 1024, 2);
    const uint64_t dWhTa = tc::make_desc(tc::smem_addr(smem_raw + sl.WhTz), 1024, 128, 0);
    const uint64_t dWhTb = tc::make_desc(tc::smem_addr(smem_raw + sl.WhTz), 2048, 128, 0);
    const uint64_t dWhh = tc::make_desc(tc::smem_addr(Whh), 128, 0), dWhl = tc::make_desc(tc::smem_addr(Whl), 128, 0);
    // activation images [128 rows][64 halfwords], SWIZZLE_128B.  MN-major view (M/N = feature, K = row): LBO = stride to
    // the next 64-feature block (H1 -> H2 = 32 KB, used by the stacked G1X operand), SBO = 1024, 2048 B per K step.
    // K-major view (M = row, K = feature): SBO = 1024, 32 B per K step.
    const uint32_t aH1 = tc::smem_addr(S_h1), aH2 = tc::smem_addr(S_h2), aX = tc::smem_addr(S_x);
    const uint64_t dH1a_mn = tc::make_desc(aH1, 32768, 1024, 2), dH1b_mn = tc::make_desc(aH1 + 16384, 32768, 1024, 2);
    const uint64_t dH2a_mn = tc::make_desc(aH2, 32768, 1024, 2), dH2b_mn = tc::make_desc(aH2 + 16384, 32768, 1024, 2);
    const uint64_t dH2a_k = tc::make_desc(aH2, 16, 1024, 2), dH2b_k = tc::make_desc(aH2 + 16384, 16, 1024, 2);
    const uint64_t dXa_mn = tc::make_desc(aX, 16384, 1024, 2), dXb_mn = tc::make_desc(aX + 64, 16384, 1024, 2);       // features 0..31 | 32..63
    const uint64_t dXa_do = tc::make_desc(aX + 48, 16384, 1024, 2), dXb_do = tc::make_desc(aX + 112, 16384, 1024, 2);  // dOut: 24..31 | 56..63
    const uint64_t dXa_k = tc::make_desc(aX + 32, 16, 1024, 2), dXb_k = tc::make_desc(aX + 96, 16, 1024, 2);           // K window 16..31 | 48..63
    constexpr uint32_t ID_FWD = tc::idesc_tf32(128, 64, 0, 0), ID_HEAD = tc::idesc_tf32(128, 16, 0, 0);
    constexpr uint32_t ID_BWD = tc::idesc_f16(128, 64, 0, 0);
    constexpr uint32_t ID_GWH = tc::idesc_f16(64, 8, 1, 1), ID_GW2 = tc::idesc_f16(64, 64, 1, 1), ID_G1X = tc::idesc_f16(128, 32, 1, 1);
    const uint32_t tACT = tmem + (uint32_t)g * TC_GSTRIDE + TC_ACT, tACC = tmem + (uint32_t)g * TC_GSTRIDE + TC_ACC;
    const uint32_t swz = (uint32_t)(r & 7);
    unsigned char *rowx = S_x + r * 128, *rowh1 = S_h1 + r * 128, *rowh2 = S_h2 + r * 128;   // a2 images at +16384

    auto pre_issue = [&]() {      // my TMEM / smem writes are done and ordered before the group's MMA issue
        tc::tmem_st_wait(); tc::tmem_ld_wait(); tc::fence_async_smem(); tc::tc_fence_before(); group_bar(g);
    };
    auto waitA = [&]() { tc::mbar_wait(mbA, phA); phA ^= 1; tc::tc_fence_after(); };
    auto waitB = [&]() { tc::mbar_wait(mbB, phB); phB ^= 1; tc::tc_fence_after(); };

    for (int s = 0; s < a.nsteps; ++s) {
        const int ep = s / a.B, bb = s - ep * a.B;
        const int32_t *pb = perm + (size_t)ep * a.S + (size_t)bb * a.mb;
        float gbh[8], gls[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { gbh[i] = 0.f; gls[i] = 0.f; }
        if (actor && tid == 0) {   // entropy with the parameters this step starts from
            float ent = 0.f;
            for (int d = 0; d < A; ++d) ent += 0.5f + 0.91893853320467274178f + lss[d];
            loss_ent += ent;
        }

        for (int t = g; t < ntiles; t += 2) {
            const int rowi = t * 128 + r;
            const bool valid = rowi < a.mb;
            // ---------------- gather ----------------
            float rec[RSG];
            {
                const int idx = valid ? __ldg(pb + rowi) : 0;
                const float4 *rp = reinterpret_cast<const float4 *>(recg + (size_t)idx * RSG);
#pragma unroll
                for (int i = 0; i < RSG / 4; ++i) {
                    const float4 v = valid ? __ldg(rp + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                    rec[4 * i] = v.x; rec[4 * i + 1] = v.y; rec[4 * i + 2] = v.z; rec[4 * i + 3] = v.w;
                }
            }
            if (pendB) { waitB(); pendB = false; }      // previous tile's weight-gradient MMAs released S_x / S_h*
            {
                float xh[KX], xl[KX];
#pragma unroll
                for (int f = 0; f < KX; ++f) {
                    const float v = f < O ? rec[f] : ((f == O && valid) ? 1.f : 0.f);
                    xh[f] = tc::tf32_hi(v); xl[f] = v - xh[f];
                }
#pragma unroll
                for (int c = 0; c < NKX; ++c) {
                    tc::tmem_st8(tg + TC_ACT + 8 * c, xh + 8 * c);
                    tc::tmem_st8(tg + TC_ACT + 32 + 8 * c, xl + 8 * c);
                    float xv[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) xv[i] = xh[8 * c + i] + xl[8 * c + i];
                    store_pair8(rowx, (uint32_t)c, rowx, (uint32_t)c + 4u, swz, xv);      // a1 in features 8c.., a2 in features 32+8c..
                }
            }
            pre_issue();
            if (issuer) {
                tc::tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < NKX; ++ks) {       // small terms first
                    tc::mma_tf32_ts(tACC, tACT + 32 + 8 * ks, tc::desc_advance(dW1h, ks * 2048), ID_FWD, ks > 0);
                    tc::mma_tf32_ts(tACC, tACT + 8 * ks, tc::desc_advance(dW1l, ks * 2048), ID_FWD, 1);
                }
#pragma unroll
                for (int ks = 0; ks < NKX; ++ks) tc::mma_tf32_ts(tACC, tACT + 8 * ks, tc::desc_advance(dW1h, ks * 2048), ID_FWD, 1);
                tc::mma_commit(mbA);
            }
            // ---------------- E1: h1 = tanh(Z1) ----------------
            waitA();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float z[32], hi[32];
                tc::tmem_ld32(tg + TC_ACC + 32 * hh, z);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) { z[i] = fast_tanh(z[i]); hi[i] = tc::tf32_hi(z[i]); }
                tc::tmem_st32(tg + TC_ACT + 32 * hh, hi);
#pragma unroll
                for (int i = 0; i < 32; ++i) hi[i] = z[i] - hi[i];
                tc::tmem_st32(tg + TC_ACT + 64 + 32 * hh, hi);
#pragma unroll
                for (int i = 0; i < 32; ++i) z[i] *= TC_SH;
#pragma unroll
                for (int c = 0; c < 4; ++c) store_pair8(rowh1, (uint32_t)(4 * hh + c), rowh1 + 16384, (uint32_t)(4 * hh + c), swz, z + 8 * c);
            }
            pre_issue();
            if (issuer) {
                tc::tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    tc::mma_tf32_ts(tACC, tACT + 64 + 8 * ks, tc::desc_advance(dW2h, ks * 2048), ID_FWD, ks > 0);
                    tc::mma_tf32_ts(tACC, tACT + 8 * ks, tc::desc_advance(dW2l, ks * 2048), ID_FWD, 1);
                }
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) tc::mma_tf32_ts(tACC, tACT + 8 * ks, tc::desc_advance(dW2h, ks * 2048), ID_FWD, 1);
                tc::mma_commit(mbA);
            }
            // ---------------- E2: h2 = tanh(Z2 + b2) ----------------
            waitA();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float z[32], lo[32];
                tc::tmem_ld32(tg + TC_ACC + 32 * hh, z);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 bv = *reinterpret_cast<const float4 *>(b2s + 32 * hh + i);
                    z[i] += bv.x; z[i + 1] += bv.y; z[i + 2] += bv.z; z[i + 3] += bv.w;
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) { z[i] = fast_tanh(z[i]); lo[i] = tc::tf32_hi(z[i]); }
                tc::tmem_st32(tg + TC_ACT + 32 * hh, lo);
#pragma unroll
                for (int i = 0; i < 32; ++i) lo[i] = z[i] - lo[i];
                tc::tmem_st32(tg + TC_ACT + 64 + 32 * hh, lo);
#pragma unroll
                for (int i = 0; i < 32; ++i) z[i] *= TC_SH;
#pragma unroll
                for (int c = 0; c < 4; ++c) store_pair8(rowh2, (uint32_t)(4 * hh + c), rowh2 + 16384, (uint32_t)(4 * hh + c), swz, z + 8 * c);
            }
            pre_issue();
            if (issuer) {
                tc::tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    tc::mma_tf32_ts(tACC, tACT + 64 + 8 * ks, tc::desc_advance(dWhh, ks * 256), ID_HEAD, ks > 0);
                    tc::mma_tf32_ts(tACC, tACT + 8 * ks, tc::desc_advance(dWhl, ks * 256), ID_HEAD, 1);
                }
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) tc::mma_tf32_ts(tACC, tACT + 8 * ks, tc::desc_advance(dWhh, ks * 256), ID_HEAD, 1);
                tc::mma_commit(mbA);
            }
            // ---------------- E3: per-row loss and d loss / d head ----------------
            waitA();
            {
                float ho[8], dq[8];
                tc::tmem_ld8(tg + TC_ACC, ho);
                tc::tmem_ld_wait();
                if (actor) {
                    float lp = 0.f, diffv[8], ivv[8];
#pragma unroll
                    for (int d = 0; d < 8; ++d) {
                        diffv[d] = 0.f; ivv[d] = 0.f;
                        if (d < A) {
                            const float ls = lss[d];
                            const float iv = expf(-2.f * ls);
                            const float diff = rec[OP + d] - (ho[d] + bhs[d]);
                            lp += -0.5f * (diff * diff * iv) - ls - 0.91893853320467274178f;
                            diffv[d] = diff; ivv[d] = iv;
                        }
                    }
                    const float ratio = expf(lp - rec[OP + A]);
                    const float adv = rec[OP + A + 1 + 2 * M];
                    const float surr1 = ratio * adv;
                    const float rcl = fminf(fmaxf(ratio, 1.f - clip), 1.f + clip);
                    const float surr2 = rcl * adv;
                    const float w1 = surr1 < surr2 ? 1.f : (surr1 == surr2 ? 0.5f : 0.f);
                    const float inr = (ratio >= 1.f - clip && ratio <= 1.f + clip) ? 1.f : 0.f;
                    const float dmin = w1 * adv + (1.f - w1) * adv * inr;
                    const float dlp = valid ? -inv_mb * dmin * ratio : 0.f;
                    if (valid) loss_act -= fminf(surr1, surr2);
#pragma unroll
                    for (int d = 0; d < 8; ++d) {
                        const float go = dlp * diffv[d] * ivv[d];                 // d loss / d mean
                        gbh[d] += go;
                        gls[d] = fmaf(dlp, diffv[d] * diffv[d] * ivv[d] - (d < A ? 1.f : 0.f), gls[d]);
                        dq[d] = go * TC_SD;
                    }
                } else {
#pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        float go = 0.f;
                        if (m < M) {
                            const float V = ho[m] + bhs[m], vo = rec[OP + A + 1 + m], R = rec[OP + A + 1 + M + m];
                            const float dlt = V - vo;
                            const float vcl = vo + fminf(fmaxf(dlt, -clip), clip);
                            const float ea = V - R, eb = vcl - R;
                            const float la = ea * ea, lb = eb * eb;
                            const float wa = la > lb ? 1.f : (la == lb ? 0.5f : 0.f);
                            const float pas = (dlt >= -clip && dlt <= clip) ? 1.f : 0.f;
                            if (valid) { loss_val += fmaxf(la, lb); go = vscale * (wa * 2.f * ea + (1.f - wa) * 2.f * eb * pas); }
                        }
                        gbh[m] += go;
                        dq[m] = go * TC_SD;
                    }
                }
                store_pair8(rowx, 3u, rowx, 7u, swz, dq);          // a1 in features 24..31, a2 in features 56..63
            }
            pre_issue();
            if (issuer) {
                tc::tc_fence_after();
                // dz2pre = dOut Wh: K window = X features 16..31 (x16.., ones, padding | dOut) against [zeros | Wh^T]
                tc::mma_f16(tACC, dXb_k, dWhTa, ID_BWD, 0);
                tc::mma_f16(tACC, dXa_k, dWhTb, ID_BWD, 1);
                tc::mma_f16(tACC, dXa_k, dWhTa, ID_BWD, 1);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {                                               // dWh^T += h2^T dOut
                    tc::mma_f16(tmem + TC_GWH, tc::desc_advance(dH2b_mn, ks * 2048), tc::desc_advance(dXa_do, ks * 2048), ID_GWH, 1);
                    tc::mma_f16(tmem + TC_GWH, tc::desc_advance(dH2a_mn, ks * 2048), tc::desc_advance(dXb_do, ks * 2048), ID_GWH, 1);
                    tc::mma_f16(tmem + TC_GWH, tc::desc_advance(dH2a_mn, ks * 2048), tc::desc_advance(dXa_do, ks * 2048), ID_GWH, 1);
                }
                tc::mma_commit(mbA);
            }
            // ---------------- E4: dz2 = dz2pre (1 - h2^2) ----------------
            waitA();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float z[32], hi[32], lo[32];
                tc::tmem_ld32(tg + TC_ACC + 32 * hh, z);
                tc::tmem_ld32(tg + TC_ACT + 32 * hh, hi);
                tc::tmem_ld32(tg + TC_ACT + 64 + 32 * hh, lo);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float hv = hi[i] + lo[i];
                    z[i] = z[i] * (1.f / TC_SW) * fmaf(-hv, hv, 1.f);      // dz2 (still x 2^12)
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) store_pair8(rowh2, (uint32_t)(4 * hh + c), rowh2 + 16384, (uint32_t)(4 * hh + c), swz, z + 8 * c);
            }
            pre_issue();
            if (issuer) {
                tc::tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {                                               // dz1pre = dz2 W2
                    tc::mma_f16(tACC, tc::desc_advance(dH2b_k, ks * 32), tc::desc_advance(dW2Ta, ks * 32), ID_BWD, ks > 0);
                    tc::mma_f16(tACC, tc::desc_advance(dH2a_k, ks * 32), tc::desc_advance(dW2Tb, ks * 32), ID_BWD, 1);
                }
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) tc::mma_f16(tACC, tc::desc_advance(dH2a_k, ks * 32), tc::desc_advance(dW2Ta, ks * 32), ID_BWD, 1);
                tc::mma_commit(mbA);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {                                               // dW2 += dz2^T h1
                    tc::mma_f16(tmem + TC_GW2, tc::desc_advance(dH2b_mn, ks * 2048), tc::desc_advance(dH1a_mn, ks * 2048), ID_GW2, 1);
                    tc::mma_f16(tmem + TC_GW2, tc::desc_advance(dH2a_mn, ks * 2048), tc::desc_advance(dH1b_mn, ks * 2048), ID_GW2, 1);
                    tc::mma_f16(tmem + TC_GW2, tc::desc_advance(dH2a_mn, ks * 2048), tc::desc_advance(dH1a_mn, ks * 2048), ID_GW2, 1);
                }
                tc::mma_commit(mbB);
            }
            // ---------------- E5: dz1 = dz1pre (1 - h1^2), h1 = (a1 + a2) / 2^8 from its smem pair ----------------
            waitA();
            {
                float dz[64];
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    float z[32];
                    tc::tmem_ld32(tg + TC_ACC + 32 * hh, z);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t off = (((uint32_t)(4 * hh + c)) ^ swz) << 4;
                        const uint4 u1 = *reinterpret_cast<const uint4 *>(rowh1 + off);
                        const uint4 u2 = *reinterpret_cast<const uint4 *>(rowh1 + 16384 + off);
                        const uint32_t w1[4] = {u1.x, u1.y, u1.z, u1.w}, w2[4] = {u2.x, u2.y, u2.z, u2.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float2 fa = unpack_h2(w1[i]), fb = unpack_h2(w2[i]);
                            const float ha = (fa.x + fb.x) * (1.f / TC_SH), hb = (fa.y + fb.y) * (1.f / TC_SH);
                            dz[32 * hh + 8 * c + 2 * i] = z[8 * c + 2 * i] * (1.f / TC_SW) * fmaf(-ha, ha, 1.f);
                            dz[32 * hh + 8 * c + 2 * i + 1] = z[8 * c + 2 * i + 1] * (1.f / TC_SW) * fmaf(-hb, hb, 1.f);
                        }
                    }
                }
                waitB();                                   // dW2 MMAs are done reading h1
#pragma unroll
                for (int c = 0; c < 8; ++c) store_pair8(rowh1, (uint32_t)c, rowh1 + 16384, (uint32_t)c, swz, dz + 8 * c);
            }
            pre_issue();
            if (issuer) {
                tc::tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {            // [dW1 | db1 ; db2] += [dz1 | dz2]^T [x | 1 | dOut]
                    tc::mma_f16(tmem + TC_G1X, tc::desc_advance(dH1b_mn, ks * 2048), tc::desc_advance(dXa_mn, ks * 2048), ID_G1X, 1);
                    tc::mma_f16(tmem + TC_G1X, tc::desc_advance(dH1a_mn, ks * 2048), tc::desc_advance(dXb_mn, ks * 2048), ID_G1X, 1);
                    tc::mma_f16(tmem + TC_G1X, tc::desc_advance(dH1a_mn, ks * 2048), tc::desc_advance(dXa_mn, ks * 2048), ID_G1X, 1);
                }
                tc::mma_commit(mbB);
            }
            pendB = true;
        }   // tiles

        // ================= step tail =================
        if (pendB) { waitB(); pendB = false; }
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();

        float gr[60];
        tc::tmem_ld32(tq + TC_GW2 + 32 * hcol, gr);
        tc::tmem_ld16(tq + TC_G1X + 16 * hcol, gr + 32);
        if (hcol == 0) tc::tmem_ld8(tq + TC_GWH, gr + 48);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) gr[i] *= 1.f / (TC_SD * TC_SH);            // dW2   = (dz2 2^12)^T (h1 2^8)
#pragma unroll
        for (int i = 32; i < 48; ++i) gr[i] *= 1.f / TC_SD;                     // dW1.. = (dz 2^12)^T x
#pragma unroll
        for (int i = 48; i < 56; ++i) gr[i] *= 1.f / (TC_SD * TC_SH);           // dWh^T = (h2 2^8)^T (dOut 2^12)
        {   // hand the accumulator cells I own back zeroed for the next step
            float z[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) z[i] = 0.f;
            tc::tmem_st32(tq + TC_GW2 + 32 * hcol, z);
            tc::tmem_st16(tq + TC_G1X + 16 * hcol, z);
            if (hcol == 0) tc::tmem_st8(tq + TC_GWH, z);
        }
        // head bias / logstd gradients: per-thread row sums -> warp -> CTA
#pragma unroll
        for (int i = 0; i < 8; ++i) { gbh[i] = warp_sum(gbh[i]); gls[i] = warp_sum(gls[i]); }
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { part[warp * 16 + i] = gbh[i]; part[warp * 16 + 8 + i] = gls[i]; }
        }
        if (tid == 0 && !a.grad_only) {   // Adam scalars of step k = step0 + s + 1, in double
            b1pow *= a.hy.beta1; b2pow *= a.hy.beta2;
            sh_d[0] = lr / (1.0 - b1pow);
            sh_d[1] = 1.0 / sqrt(1.0 - b2pow);
        }
        __syncthreads();
        gr[56] = 0.f; gr[57] = 0.f; gr[58] = 0.f; gr[59] = 0.f;
        if (tid < 8) {
            float sb = 0.f, sl_ = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) { sb += part[w * 16 + tid]; sl_ += part[w * 16 + 8 + tid]; }
            gr[56] = sb; gr[57] = sl_ - ecoef;          // d(-ecoef * entropy) / d logstd = -ecoef
        }
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < 58; ++i) {
            if (!((vmask >> i) & 1ull)) gr[i] = 0.f;
            sq = fmaf(gr[i], gr[i], sq);
        }
        sq = block_sum(sq, red);

        if (a.grad_only) {
#pragma unroll
            for (int i = 0; i < 58; ++i) {
                const int e = tc_own(i, q, hcol, lane, tid, O, KH, A, actor);
                if (e >= 0) a.grad_out[(size_t)task * L.n_par + L.to_global(half, e)] = gr[i];
            }
            break;
        }

        float *ssq2 = ssqS + 2 * (s & 1);     // slots alternate by step parity: the peer may still be reading the last ones
        if (tid == 0) {
            ssq2[rank] = sq;
            st_dsmem1(mapa_u32(smem_u32(ssq2 + rank), rank ^ 1u), sq);
        }
        // moments of my slots: in flight across the cluster barrier
        float4 m4[TC_SLOT4], v4[TC_SLOT4];
#pragma unroll
        for (int s4 = 0; s4 < TC_SLOT4; ++s4) {
            if ((vmask >> (4 * s4)) & 0xFull) { m4[s4] = __ldcg(mM + s4 * TC_THREADS + tid); v4[s4] = __ldcg(mV + s4 * TC_THREADS + tid); }
            else { m4[s4] = make_float4(0.f, 0.f, 0.f, 0.f); v4[s4] = m4[s4]; }
        }
        sync_group<2>();
        const float tot = ssq2[0] + ssq2[1];
        const float coef = fminf(1.f, (float)a.hy.max_grad_norm / (sqrtf(tot) + 1e-6f));
        const float step_size = (float)sh_d[0], ibc2 = (float)sh_d[1];

#define TC_ADAM(P_, G_, M_, V_)                                                         \
        {                                                                               \
            const float gq = (G_) * coef;                                               \
            M_ = fmaf(gq - M_, omb1, M_);                                               \
            V_ = fmaf(omb2 * gq, gq, V_ * b2f);                                         \
            const float denom = fmaf(fast_sqrt(V_), ibc2, aeps);                        \
            P_ -= step_size * __fdividef(M_, denom);                                    \
        }
        // ---- W2 rows: slots 0..31 (lanes < 16): row j, features 32h .. 32h+31 ----
        if (lane < 16) {
            const int j = 16 * q + lane;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int w = (8 * hcol + c) * 256 + j * 4;
                const float4 ph = *reinterpret_cast<const float4 *>(W2h + w), pl = *reinterpret_cast<const float4 *>(W2l + w);
                float p0 = ph.x + pl.x, p1 = ph.y + pl.y, p2 = ph.z + pl.z, p3 = ph.w + pl.w;
                TC_ADAM(p0, gr[4 * c], m4[c].x, v4[c].x) TC_ADAM(p1, gr[4 * c + 1], m4[c].y, v4[c].y)
                TC_ADAM(p2, gr[4 * c + 2], m4[c].z, v4[c].z) TC_ADAM(p3, gr[4 * c + 3], m4[c].w, v4[c].w)
                const float4 nh = make_float4(tc::tf32_hi(p0), tc::tf32_hi(p1), tc::tf32_hi(p2), tc::tf32_hi(p3));
                *reinterpret_cast<float4 *>(W2h + w) = nh;
                *reinterpret_cast<float4 *>(W2l + w) = make_float4(p0 - nh.x, p1 - nh.y, p2 - nh.z, p3 - nh.w);
                const int k0 = 32 * hcol + 4 * c;
                const float pv[4] = {p0, p1, p2, p3};
#pragma unroll
                for (int e = 0; e < 4; ++e) {                 // backward copy: row k, feature j, fp16 pair x 2^8
                    const __half t1 = __float2half_rn(pv[e] * TC_SW);
                    const int hw = sw128_hw(k0 + e, j);
                    W2Ta[hw] = t1; W2Tb[hw] = __float2half_rn(pv[e] * TC_SW - __half2float(t1));
                }
            }
        }
        // ---- slots 32..47: W1 | b1 rows (quadrants 0,1) or b2 (quadrants 2,3), columns 16h .. 16h+15 ----
        if (q < 2) {
            const int j = 32 * q + lane;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int ch = 4 * hcol + c;
                if (ch < NCH1) {
                    const int w = ch * 256 + j * 4;
                    const float4 ph = *reinterpret_cast<const float4 *>(W1h + w), pl = *reinterpret_cast<const float4 *>(W1l + w);
                    float p0 = ph.x + pl.x, p1 = ph.y + pl.y, p2 = ph.z + pl.z, p3 = ph.w + pl.w;
                    TC_ADAM(p0, gr[32 + 4 * c], m4[8 + c].x, v4[8 + c].x) TC_ADAM(p1, gr[33 + 4 * c], m4[8 + c].y, v4[8 + c].y)
                    TC_ADAM(p2, gr[34 + 4 * c], m4[8 + c].z, v4[8 + c].z) TC_ADAM(p3, gr[35 + 4 * c], m4[8 + c].w, v4[8 + c].w)
                    // columns past the bias keep gradient 0 and moments 0: the update leaves their zeros in place
                    const float4 nh = make_float4(tc::tf32_hi(p0), tc::tf32_hi(p1), tc::tf32_hi(p2), tc::tf32_hi(p3));
                    *reinterpret_cast<float4 *>(W1h + w) = nh;
                    *reinterpret_cast<float4 *>(W1l + w) = make_float4(p0 - nh.x, p1 - nh.y, p2 - nh.z, p3 - nh.w);
                }
            }
        } else if (hcol == O / 16) {
            const int j = 32 * (q - 2) + lane;
            constexpr int c = (O % 16) / 4, e = O % 4;
            float p = b2s[j];
            float gg = e == 0 ? gr[32 + 4 * c] : (e == 1 ? gr[33 + 4 * c] : (e == 2 ? gr[34 + 4 * c] : gr[35 + 4 * c]));
            float mm = e == 0 ? m4[8 + c].x : (e == 1 ? m4[8 + c].y : (e == 2 ? m4[8 + c].z : m4[8 + c].w));
            float vv = e == 0 ? v4[8 + c].x : (e == 1 ? v4[8 + c].y : (e == 2 ? v4[8 + c].z : v4[8 + c].w));
            TC_ADAM(p, gg, mm, vv)
            b2s[j] = p;
            if (e == 0) { m4[8 + c].x = mm; v4[8 + c].x = vv; } else if (e == 1) { m4[8 + c].y = mm; v4[8 + c].y = vv; }
            else if (e == 2) { m4[8 + c].z = mm; v4[8 + c].z = vv; } else { m4[8 + c].w = mm; v4[8 + c].w = vv; }
        }
        // ---- slots 48..55: head weights Wh[a][k], k = 16q + lane (column half 0, lanes < 16) ----
        if (hcol == 0 && lane < 16) {
            const int k = 16 * q + lane;
            float pn[8];
#pragma unroll
            for (int aa = 0; aa < 8; ++aa) {
                const int w = (k >> 2) * 32 + aa * 4 + (k & 3);
                float p = Whh[w] + Whl[w];
                float mm = aa < 4 ? (&m4[12].x)[aa] : (&m4[13].x)[aa - 4];
                float vv = aa < 4 ? (&v4[12].x)[aa] : (&v4[13].x)[aa - 4];
                TC_ADAM(p, gr[48 + aa], mm, vv)
                if (aa < 4) { (&m4[12].x)[aa] = mm; (&v4[12].x)[aa] = vv; } else { (&m4[13].x)[aa - 4] = mm; (&v4[13].x)[aa - 4] = vv; }
                const float nh = tc::tf32_hi(p);
                Whh[w] = nh; Whl[w] = p - nh;
                pn[aa] = p * TC_SW;
            }
            uint4 q1, q2;                                     // backward copy: row k, 8 halfwords a, fp16 pair x 2^8
            split_h2(pn[0], pn[1], q1.x, q2.x); split_h2(pn[2], pn[3], q1.y, q2.y);
            split_h2(pn[4], pn[5], q1.z, q2.z); split_h2(pn[6], pn[7], q1.w, q2.w);
            *reinterpret_cast<uint4 *>(WhTa + k * 8) = q1;
            *reinterpret_cast<uint4 *>(WhTb + k * 8) = q2;
        }
        // ---- slots 56, 57: head bias and logstd ----
        if (tid < 8) {
            if (tid < KH) { float p = bhs[tid]; TC_ADAM(p, gr[56], m4[14].x, v4[14].x) bhs[tid] = p; }
            if (actor && tid < A) { float p = lss[tid]; TC_ADAM(p, gr[57], m4[14].y, v4[14].y) lss[tid] = p; }
        }
#undef TC_ADAM
#pragma unroll
        for (int s4 = 0; s4 < TC_SLOT4; ++s4)
            if ((vmask >> (4 * s4)) & 0xFull) { __stcg(mM + s4 * TC_THREADS + tid, m4[s4]); __stcg(mV + s4 * TC_THREADS + tid, v4[s4]); }
        tc::tmem_st_wait();
        tc::fence_async_smem();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
    }   // steps

    // ---------------- write back ----------------
    if (!a.grad_only) {
        __threadfence_block();
        for (int s4 = 0; s4 < TC_SLOT4; ++s4) {
            if (!((vmask >> (4 * s4)) & 0xFull)) continue;
            const float4 mq = __ldcg(mM + s4 * TC_THREADS + tid), vq = __ldcg(mV + s4 * TC_THREADS + tid);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int sidx = 4 * s4 + c;
                const int e = tc_own(sidx, q, hcol, lane, tid, O, KH, A, actor);
                if (e < 0) continue;
                const size_t gi = (size_t)task * L.n_par + L.to_global(half, e);
                a.adam_m[gi] = (&mq.x)[c]; a.adam_v[gi] = (&vq.x)[c];
                float p;
                if (sidx < 32) { const int j = 16 * q + lane, k = 32 * hcol + sidx, w = (k >> 2) * 256 + j * 4 + (k & 3); p = W2h[w] + W2l[w]; }
                else if (sidx < 48) {
                    const int cc = 16 * hcol + sidx - 32;
                    if (q < 2) { const int j = 32 * q + lane, w = (cc >> 2) * 256 + j * 4 + (cc & 3); p = W1h[w] + W1l[w]; }
                    else p = b2s[32 * (q - 2) + lane];
                } else if (sidx < 56) { const int k = 16 * q + lane, aa = sidx - 48, w = (k >> 2) * 32 + aa * 4 + (k & 3); p = Whh[w] + Whl[w]; }
                else if (sidx == 56) p = bhs[tid];
                else p = lss[tid];
                a.params[gi] = p;
            }
        }
    }
    {   // losses: the critic CTA reports the value loss, the actor CTA the action loss and the entropy
        const float la = block_sum(loss_act, red), lv = block_sum(loss_val, red), le = block_sum(loss_ent, red);
        if (tid == 0) {
            const float ns = (float)a.nsteps;
            if (actor) {
                a.losses[task * 3 + 1] = la * inv_mb / ns;
                a.losses[task * 3 + 2] = le / ns;
                if (!a.grad_only) a.adam_step[task] = step0 + a.nsteps;
            } else {
                a.losses[task * 3 + 0] = lv * 0.5f / ((float)a.mb * M) / ns;
            }
        }
    }
    tc::tc_fence_before();
    sync_group<2>();                 // no DSMEM traffic may target an exited CTA
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace pgm
