// Shared-memory resident 64-64 tanh MLP half (actor or critic) and the register-tiled
// FP32 FFMA building blocks used by K1 (forward) and K3 (forward + backward).
//
// Thread tiling: a CTA of 256 threads is a 16x16 grid (tr = tid & 15, tc = tid >> 4).
// For row-tile ops a thread owns rows {tr + 16*i, i < TM} and columns {4*tc .. 4*tc+3} of a
// [RC = 16*TM] x 64 tile. Operands are read from shared memory with 128-bit loads along the
// contraction dimension; row strides are 4*odd floats so that the 8 rows touched by a quarter
// warp fall into distinct 16-byte bank groups.
#pragma once
#include "common.cuh"

namespace pgm {

struct HalfNet {           // pointers into shared memory
    float *W1;             // [64][ldw1]   cols >= O are zero
    float *b1;             // [64]
    float *W2;             // [64][LDH]
    float *b2;             // [64]
    float *Wh;             // [KH][LDH]    head: fc_mean (actor) or critic_linear (critic)
    float *bh;             // [KH]
    float *ls;             // [A] logstd (actor only)
    int KH;
};

__host__ __device__ inline int halfnet_smem_floats(const NetLayout &L, int half) {
    int KH = L.head_dim(half);
    return H * L.ldw1 + H + H * LDH + H + KH * LDH + round_up(KH, 4) + (half == 0 ? round_up(L.A, 4) : 0);
}

#ifdef __CUDACC__
__device__ inline float *halfnet_carve(HalfNet &n, float *p, const NetLayout &L, int half) {
    n.KH = L.head_dim(half);
    n.W1 = p; p += H * L.ldw1;
    n.b1 = p; p += H;
    n.W2 = p; p += H * LDH;
    n.b2 = p; p += H;
    n.Wh = p; p += n.KH * LDH;
    n.bh = p; p += round_up(n.KH, 4);
    n.ls = p; if (half == 0) p += round_up(L.A, 4);
    return p;
}

template <bool CG>
__device__ __forceinline__ float ldp(const float *p) { return CG ? __ldcg(p) : __ldg(p); }

// Copy one half's parameters from the flat global vector `g` (reference order) into the padded
// shared-memory image. CG=true reads through L2 only (other CTAs of the cluster just wrote them).
template <bool CG>
__device__ inline void halfnet_load(const HalfNet &n, const float *__restrict__ g, const NetLayout &L, int half) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const float *base = g + L.half_base(half);
    const float *head = g + L.half_head(half);
    const int O = L.O;
    for (int i = tid; i < H * L.ldw1; i += nt) {
        int j = i / L.ldw1, k = i - j * L.ldw1;
        n.W1[i] = k < O ? ldp<CG>(base + j * O + k) : 0.f;
    }
    const float *b1 = base + H * O, *W2 = b1 + H, *b2 = W2 + H * H;
    for (int i = tid; i < H; i += nt) { n.b1[i] = ldp<CG>(b1 + i); n.b2[i] = ldp<CG>(b2 + i); }
    for (int i = tid; i < H * H; i += nt) n.W2[(i >> 6) * LDH + (i & 63)] = ldp<CG>(W2 + i);
    for (int i = tid; i < n.KH * H; i += nt) n.Wh[(i >> 6) * LDH + (i & 63)] = ldp<CG>(head + i);
    const float *bh = head + n.KH * H;
    for (int i = tid; i < n.KH; i += nt) n.bh[i] = ldp<CG>(bh + i);
    if (half == 0)
        for (int i = tid; i < L.A; i += nt) n.ls[i] = ldp<CG>(bh + n.KH + i);
}

__device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void sts4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }

#define PGM_DOT4(acc, a, b) \
    acc = fmaf((a).x, (b).x, acc); acc = fmaf((a).y, (b).y, acc); acc = fmaf((a).z, (b).z, acc); acc = fmaf((a).w, (b).w, acc)

// out[r][j] = tanh(bias[j] + sum_{k<K} in[r][k] * W[j][k]);  r = tr+16i, j = 4tc+c.  K % 4 == 0.
template <int TM>
__device__ __forceinline__ void layer_fwd_tanh(const float *__restrict__ in, int ldi, const float *__restrict__ W,
                                               int ldw, const float *__restrict__ bias, int K,
                                               float *__restrict__ out, int tr, int tc) {
    float acc[TM][4];
    const float4 bv = lds4(bias + 4 * tc);
#pragma unroll
    for (int i = 0; i < TM; ++i) { acc[i][0] = bv.x; acc[i][1] = bv.y; acc[i][2] = bv.z; acc[i][3] = bv.w; }
    const float *ip = in + tr * ldi;
    const float *wp = W + 4 * tc * ldw;
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        float4 a[TM], b[4];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = lds4(ip + i * 16 * ldi + k);
#pragma unroll
        for (int c = 0; c < 4; ++c) b[c] = lds4(wp + c * ldw + k);
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) { PGM_DOT4(acc[i][c], a[i], b[c]); }
    }
#pragma unroll
    for (int i = 0; i < TM; ++i)
        sts4(out + (tr + 16 * i) * LDH + 4 * tc,
             make_float4(tanhf(acc[i][0]), tanhf(acc[i][1]), tanhf(acc[i][2]), tanhf(acc[i][3])));
}

// Head: ho[r][a] = bh[a] + sum_k h[r][k] * Wh[a][k]   (a < KH, ho row stride ldo)
template <int TM>
__device__ __forceinline__ void head_fwd(const float *__restrict__ h, const HalfNet &n, float *__restrict__ ho,
                                         int ldo, int tr, int tc) {
    for (int a = tc; a < n.KH; a += 16) {
        float acc[TM];
        const float b = n.bh[a];
#pragma unroll
        for (int i = 0; i < TM; ++i) acc[i] = b;
        const float *wp = n.Wh + a * LDH;
#pragma unroll 4
        for (int k = 0; k < H; k += 4) {
            const float4 w = lds4(wp + k);
#pragma unroll
            for (int i = 0; i < TM; ++i) { const float4 x = lds4(h + (tr + 16 * i) * LDH + k); PGM_DOT4(acc[i], x, w); }
        }
#pragma unroll
        for (int i = 0; i < TM; ++i) ho[(tr + 16 * i) * ldo + a] = acc[i];
    }
}

// Forward of one half for a chunk of RC = 16*TM rows: x [RC][ldx] -> h1, h2 [RC][LDH], ho [RC][ldo].
// Ends with a __syncthreads() so that ho/h1/h2 are visible to every thread.
template <int TM>
__device__ __forceinline__ void half_forward(const float *x, int ldx, const HalfNet &n, const NetLayout &L,
                                             float *h1, float *h2, float *ho, int ldo, int tr, int tc) {
    layer_fwd_tanh<TM>(x, ldx, n.W1, L.ldw1, n.b1, L.OP, h1, tr, tc);
    __syncthreads();
    layer_fwd_tanh<TM>(h1, LDH, n.W2, LDH, n.b2, H, h2, tr, tc);
    __syncthreads();
    head_fwd<TM>(h2, n, ho, ldo, tr, tc);
    __syncthreads();
}
#endif  // __CUDACC__

}  // namespace pgm
