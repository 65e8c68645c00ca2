// K2 -- vector-reward GAE + weight-scalarised, normalised advantage.
//
// Replaces RolloutStorage.compute_returns (a2c/storage.py:83-94: use_gae and
// use_proper_time_limits) and the PPO.update preamble (a2c/algo/ppo.py:43-56) with
// WeightedSumScalarization.evaluate (morl/scalarization_methods.py:28-29).
//
// The reference's recurrence, per env column n and objective m, going backwards in t:
//     delta_t = r_t + gamma * V_{t+1} * mask_{t+1} - V_t
//     g_t     = (delta_t + gamma*lam*mask_{t+1} * g_{t+1}) * bad_{t+1}
//     ret_t   = g_t + V_t
// is the affine map g_t = a_t * g_{t+1} + d_t with a_t = gamma*lam*mask*bad, d_t = delta_t*bad.
// One CTA per task, one warp per env column: the warp walks T in blocks of 32 steps (lane = step,
// so global loads of a block are contiguous across the CTA's warps), composes the 32 affine maps
// with a shuffle scan and chains blocks through a carried g. The scalarised raw advantage
// sum_m w_m * sqrt(var_m + 1e-8) * g_{t,m} is written to `adv`, then the CTA normalises it with
// the mean and unbiased std over all T*N samples (two-pass, block reductions by warp shuffles).
//
// Streaming kernel: 36 B per env-step at M = 2 (read r, V, masks; write ret, adv) -> HBM bound.
#include "common.cuh"

namespace pgm {

constexpr int K2_MAXM = 8;

template <int M>
__global__ void __launch_bounds__(1024) k2_gae_kernel(const float *__restrict__ rewards, const float *__restrict__ value,
                                                      const float *__restrict__ masks, const float *__restrict__ bad_masks,
                                                      const float *__restrict__ weights, const float *__restrict__ obj_var,
                                                      float gamma, float lam, float *__restrict__ returns,
                                                      float *__restrict__ adv, int T, int N) {
    __shared__ float red[34];
    const int task = blockIdx.x;
    const int lane = threadIdx.x & 31, n = threadIdx.x >> 5;   // warp n <-> env column n
    rewards += (size_t)task * T * N * M;
    value += (size_t)task * (T + 1) * N * M;
    masks += (size_t)task * (T + 1) * N;
    bad_masks += (size_t)task * (T + 1) * N;
    returns += (size_t)task * T * N * M;
    const bool do_adv = (weights != nullptr) && (adv != nullptr);
    if (do_adv) adv += (size_t)task * T * N;

    float ws[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        float s = obj_var ? sqrtf(__ldg(obj_var + task * M + m) + 1e-8f) : 1.f;
        ws[m] = do_adv ? __ldg(weights + task * M + m) * s : 0.f;
    }
    const float gl = gamma * lam;
    float carry[M];
#pragma unroll
    for (int m = 0; m < M; ++m) carry[m] = 0.f;
    float psum = 0.f;

    const int nblk = (T + 31) / 32;
    for (int blk = nblk - 1; blk >= 0; --blk) {
        const int t = blk * 32 + lane;
        const bool ok = t < T;
        float a = 1.f, d[M], v[M];
#pragma unroll
        for (int m = 0; m < M; ++m) { d[m] = 0.f; v[m] = 0.f; }
        if (ok) {
            const float mk = __ldg(masks + (size_t)(t + 1) * N + n);
            const float bd = __ldg(bad_masks + (size_t)(t + 1) * N + n);
            a = gl * mk * bd;
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const float r = __ldg(rewards + ((size_t)t * N + n) * M + m);
                v[m] = __ldg(value + ((size_t)t * N + n) * M + m);
                const float vn = __ldg(value + ((size_t)(t + 1) * N + n) * M + m);
                d[m] = (r + gamma * vn * mk - v[m]) * bd;
            }
        }
        // suffix composition over lanes: after the scan, g_t = d + a * carry where (a, d) is the
        // composition of the maps of steps t, t+1, ..., end of block
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const float an = __shfl_down_sync(0xffffffffu, a, off);
            float dn[M];
#pragma unroll
            for (int m = 0; m < M; ++m) dn[m] = __shfl_down_sync(0xffffffffu, d[m], off);
            if (lane + off < 32) {
#pragma unroll
                for (int m = 0; m < M; ++m) d[m] = fmaf(a, dn[m], d[m]);
                a *= an;
            }
        }
        float g[M], s = 0.f;
#pragma unroll
        for (int m = 0; m < M; ++m) {
            g[m] = fmaf(a, carry[m], d[m]);
            s = fmaf(ws[m], g[m], s);
        }
        if (ok) {
#pragma unroll
            for (int m = 0; m < M; ++m) returns[((size_t)t * N + n) * M + m] = g[m] + v[m];
            if (do_adv) { adv[(size_t)t * N + n] = s; psum += s; }
        }
#pragma unroll
        for (int m = 0; m < M; ++m) carry[m] = __shfl_sync(0xffffffffu, g[m], 0);
    }
    if (!do_adv) return;

    // normalise: (adv - mean) / (unbiased std + 1e-5) over the task's T*N samples
    const int S = T * N;
    const float mean = block_sum(psum, red) / (float)S;   // block_sum syncs -> raw adv visible CTA-wide
    float q = 0.f;
    for (int i = threadIdx.x; i < S; i += blockDim.x) { const float c = adv[i] - mean; q = fmaf(c, c, q); }
    const float var = block_sum(q, red) / (float)(S - 1);
    const float inv = 1.f / (sqrtf(var) + 1e-5f);
    for (int i = threadIdx.x; i < S; i += blockDim.x) adv[i] = (adv[i] - mean) * inv;
}

}  // namespace pgm

using namespace pgm;

extern "C" int pgm_gae_adv_f32(const float *rewards, const float *value, const float *masks, const float *bad_masks,
                               const float *weights, const float *obj_var, float gamma, float lam, float *returns,
                               float *adv, int P, int T, int N, int M, void *stream) {
    PGM_REQUIRE(rewards && value && masks && bad_masks && returns, "pgm_gae_adv_f32: null pointer");
    PGM_REQUIRE(P > 0 && T > 0 && N > 0 && N <= 32, "pgm_gae_adv_f32: need 1 <= N <= 32 env columns (got %d)", N);
    PGM_REQUIRE(M >= 1 && M <= K2_MAXM, "pgm_gae_adv_f32: obj_num %d unsupported", M);
    PGM_REQUIRE(!(weights && adv) || (long long)T * N >= 2, "pgm_gae_adv_f32: need >= 2 samples to normalise");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(P), block(32 * N);
#define PGM_K2(MM) case MM: k2_gae_kernel<MM><<<grid, block, 0, st>>>(rewards, value, masks, bad_masks, weights, obj_var, gamma, lam, returns, adv, T, N); break;
    switch (M) { PGM_K2(1) PGM_K2(2) PGM_K2(3) PGM_K2(4) PGM_K2(5) PGM_K2(6) PGM_K2(7) PGM_K2(8) }
#undef PGM_K2
    PGM_CUDA(cudaGetLastError());
    return PGM_OK;
}
