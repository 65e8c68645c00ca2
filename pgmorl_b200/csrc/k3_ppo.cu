// K3 -- PPO update for the whole population shard in ONE launch.
//
// Replaces PPO.update's minibatch loop (a2c/algo/ppo.py:62-107): evaluate_actions,
// clipped surrogate, clipped vector value loss, autograd backward, clip_grad_norm_ and
// Adam.step, plus the sampler of a2c/storage.py:118-154 (consumes the host permutation).
//
// Parallel decomposition. A task is a chain of E*B dependent optimiser steps, so the only
// parallelism inside a step is over the minibatch rows and over the two independent networks
// (actor / critic share nothing but the global gradient norm). One thread-block CLUSTER of C
// CTAs owns one task for the whole chain:
//     C == 1 : one CTA runs both halves, all mb rows                     (large populations)
//     C >= 2 : CTAs [0, C/2) = actor, [C/2, C) = critic; each of the G = C/2 CTAs of a half
//              runs mb/G rows                                            (small populations)
// Per step every CTA: gathers its rows' packed records (cp.async, double buffered), runs
// forward + hand-derived backward on 16*TM-row chunks with weights resident in shared memory,
// and takes part in an all-reduce of the half's gradient: reduce-scatter to the G slice owners
// (fixed summation order -> deterministic), norm exchange over all C CTAs, clip + Adam on the
// slice, all-gather of the new parameters. The generic kernel in this file does it through
// L2 scratch slots and three cluster barriers per step (any dims, incl. C == 1); the fast
// kernel for small networks (k3_fast.cuh) keeps everything in (distributed) shared memory and
// runs the exchange as barrier-free dataflow (bulk copies + st.async into mbarriers, TMA
// multicast for the parameters).
//
// FP32 FFMA throughout; Adam bias corrections in double.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "net.cuh"

namespace pgm {

struct K3Args {
    float *params, *adam_m, *adam_v;
    int32_t *adam_step;
    const double *lr;
    const float *rec;       // packed records [P][S][RSG]
    const int32_t *perm;    // [P or 1][E][S]  (update)  or  [mb] (grad mode)
    float *losses;          // [P][3]
    float *gpart;           // scratch: partial / reduced gradients [P][2][G][NHP]
    float *pstage;          // scratch: updated parameters on their way to the TMA multicast [P][2][NHP] (fast path, TAIL >= 2)
    float *ssq;             // scratch: squared-norm partials       [P][16]
    float *lpart;           // scratch: loss partial sums           [P][16][4]
    float *grad_out;        // grad mode only: [P][n_par]
    float *mv;              // tensor-core path: Adam moments per half in reference order [P][2 halves][m|v][TC_NHP]
    long long *trace;       // PGM_K3_TRACE builds only: [CTA][4 steps][16 marks] clock64 / globaltimer
    int perm_shared, E, B, mb, S, nsteps, grad_only;
    int Rg;                 // rows per CTA per step (multiple of the chunk size)
    int RSG, RSS, NHP;      // record stride in global / shared memory; padded half size
    int stage_floats;       // fast path: staging floats for row-split partials
    int rs;                 // tensor-core path: CTAs per network half (row split), 1 or 2
    pgm_ppo_hyper hy;
    NetLayout L;
};

// record field offsets (floats): x[0,OP) act[OP,OP+A) logp_old v_old[M] ret[M] adv
__host__ __device__ inline int rec_stride(const NetLayout &L) { return round_up(L.OP + L.A + 2 * L.M + 2, 4); }

__global__ void k3_pack_kernel(const float *__restrict__ obs, size_t obs_ts, const float *__restrict__ action,
                               const float *__restrict__ logp, const float *__restrict__ vold, size_t v_ts,
                               const float *__restrict__ ret, const float *__restrict__ adv, float *__restrict__ rec,
                               int P, int S, NetLayout L, int RSG) {
    const size_t total = (size_t)P * S * RSG;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int f = (int)(i % RSG);
        const size_t ts = i / RSG;
        const int s = (int)(ts % S);
        const int task = (int)(ts / S);
        const int O = L.O, OP = L.OP, A = L.A, M = L.M;
        float v = 0.f;
        if (f < O) v = obs[task * obs_ts + (size_t)s * O + f];
        else if (f < OP) v = 0.f;
        else if (f < OP + A) v = action[((size_t)task * S + s) * A + (f - OP)];
        else if (f == OP + A) v = logp[(size_t)task * S + s];
        else if (f < OP + A + 1 + M) v = vold[task * v_ts + (size_t)s * M + (f - OP - A - 1)];
        else if (f < OP + A + 1 + 2 * M) v = ret[((size_t)task * S + s) * M + (f - OP - A - 1 - M)];
        else if (f == OP + A + 1 + 2 * M) v = adv[(size_t)task * S + s];
        rec[i] = v;
    }
}

#ifdef PGM_K3_TRACE
#define PGM_TR(ph)                                                                                   \
    if (a.trace && tid == 0 && s >= 8 && s < 12) {                                                   \
        long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                       \
        a.trace[(((size_t)blockIdx.x * 4 + (s - 8)) * 16 + (ph)) * 2] = clock64();                   \
        a.trace[(((size_t)blockIdx.x * 4 + (s - 8)) * 16 + (ph)) * 2 + 1] = gt_;                     \
    }
#else
#define PGM_TR(ph)
#endif

template <int C>
__device__ __forceinline__ void sync_group() {
    if (C == 1) {
        __syncthreads();
    } else {
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    }
}

template <int C>
__device__ __forceinline__ unsigned group_rank() {
    if (C == 1) return 0;
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}

// half-local flat parameter index -> offset inside the padded shared-memory image of the half
__device__ __forceinline__ int half_img_off(const HalfNet &n, const NetLayout &L, int e) {
    const int O = L.O;
    if (e < H * O) { const int j = e / O; return j * L.ldw1 + (e - j * O); }
    e -= H * O;
    if (e < H) return (int)(n.b1 - n.W1) + e;
    e -= H;
    if (e < H * H) return (int)(n.W2 - n.W1) + (e >> 6) * LDH + (e & 63);
    e -= H * H;
    if (e < H) return (int)(n.b2 - n.W1) + e;
    e -= H;
    if (e < n.KH * H) return (int)(n.Wh - n.W1) + (e >> 6) * LDH + (e & 63);
    e -= n.KH * H;
    if (e < n.KH) return (int)(n.bh - n.W1) + e;
    return (int)(n.ls - n.W1) + (e - n.KH);
}

__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// C: CTAs per task; TM: rows per thread per chunk (chunk = 16*TM rows);
// KG1: 4-column groups of dW1 per thread (OP <= 64*KG1); NA: head rows per thread (A,M <= 16*NA);
// DB: double-buffered record gather.
// Generic path (any supported dims, incl. wide observations and C == 1): each CTA updates a 1/G
// slice of its half in global memory and all CTAs reload the parameters (three barriers per step).
// Small networks on C >= 2 take the fast path in k3_fast.cuh instead.
template <int C, int TM, int KG1, int NA, bool DB>
__global__ void __launch_bounds__(NTHREADS, 1) k3_ppo_kernel(const K3Args a) {
    constexpr int RC = 16 * TM;
    constexpr int NHALF = (C == 1) ? 2 : 1;
    constexpr int G = (C == 1) ? 1 : C / 2;
    static_assert(!(DB && NHALF == 2), "double-buffered gather needs one half per CTA");
    extern __shared__ __align__(16) float smem[];
    __shared__ float red[34];
    __shared__ double sh_d[4];

    const NetLayout &L = a.L;
    const int tid = threadIdx.x, tr = tid & 15, tc = tid >> 4;
    const int task = blockIdx.x / C;
    const unsigned rank = group_rank<C>();
    const int half0 = (C == 1) ? 0 : (int)(rank / G);
    const int g = (C == 1) ? 0 : (int)(rank % G);
    const int O = L.O, OP = L.OP, A = L.A, M = L.M;
    const int ldo = ((A > M ? A : M) | 1);
    const int RSS = a.RSS, RSG = a.RSG;
    const int nchunk = a.Rg / RC;

    // ---- shared memory carve ----
    HalfNet net[NHALF];
    float *p = smem;
#pragma unroll
    for (int hh = 0; hh < NHALF; ++hh) p = halfnet_carve(net[hh], p, L, half0 + hh);
    float *h1 = p; p += RC * LDH;
    float *h2 = p; p += RC * LDH;
    float *dz = p; p += RC * LDH;
    float *ho = p; p += round_up(RC * ldo, 4);
    float *els = p; p += round_up(RC * ldo, 4);
    float *recb = p;   // [DB ? 2 : 1][RC][RSS]

    float *gparams = a.params + (size_t)task * L.n_par;
#pragma unroll
    for (int hh = 0; hh < NHALF; ++hh) halfnet_load<false>(net[hh], gparams, L, half0 + hh);

    const float clip = (float)a.hy.clip_param;
    const float inv_mb = 1.f / (float)a.mb;
    const float vscale = (float)(a.hy.value_loss_coef * 0.5 / ((double)a.mb * M));
    const float omb1 = (float)(1.0 - a.hy.beta1);
    const float b2f = (float)a.hy.beta2, omb2 = (float)(1.0 - a.hy.beta2);
    const float aeps = (float)a.hy.adam_eps;
    const float ecoef = (float)a.hy.entropy_coef;
    const int step0 = a.grad_only ? 0 : a.adam_step[task];
    double b1pow = pow(a.hy.beta1, (double)step0), b2pow = pow(a.hy.beta2, (double)step0);   // thread 0 uses them
    const double lr = a.grad_only ? 0.0 : a.lr[task];

    float loss_act = 0.f, loss_val = 0.f, loss_ent = 0.f;   // running sums (per thread)

    const int32_t *perm = a.perm + ((a.perm_shared || a.grad_only) ? 0 : (size_t)task * a.E * a.S);
    const float *recg = a.rec + (size_t)task * a.S * RSG;
    const int q4 = RSG / 4;

    // gather chunk `ci` (flat index over steps x chunks) into buffer `buf`
    auto gather = [&](int ci, int buf) {
        const int s = ci / nchunk, c = ci - s * nchunk;
        const int e = s / a.B, b = s - e * a.B;
        const int row0 = g * a.Rg + c * RC;                       // first row of this chunk inside the minibatch
        const int32_t *pb = perm + (size_t)e * a.S + (size_t)b * a.mb;
        float *dst = recb + buf * RC * RSS;
        for (int i = tid; i < RC * q4; i += NTHREADS) {
            const int r = i / q4, q = i - r * q4;
            float *d = dst + r * RSS + 4 * q;
            if (row0 + r < a.mb) {
                const int idx = __ldg(pb + row0 + r);
                cp_async16(d, recg + (size_t)idx * RSG + 4 * q);
            } else {
                sts4(d, make_float4(0.f, 0.f, 0.f, 0.f));
            }
        }
        cp_async_commit();
    };

    const int total_chunks = a.nsteps * nchunk;
    if (DB) gather(0, 0);
    __syncthreads();

    for (int s = 0; s < a.nsteps; ++s) {
        PGM_TR(0)
#pragma unroll
        for (int hh = 0; hh < NHALF; ++hh) {
            const int half = half0 + hh;
            const HalfNet &n = net[hh];
            const int KH = n.KH;
            // gradient accumulators (registers, persist over the chunks of this step)
            float gW2[4][4], gb2[4], gW1[KG1][4][4], gb1[4], gWh[NA][4], gbh[NA], gls[NA];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                gb2[i] = 0.f; gb1[i] = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    gW2[i][j] = 0.f;
#pragma unroll
                    for (int q = 0; q < KG1; ++q) gW1[q][i][j] = 0.f;
                }
            }
#pragma unroll
            for (int ia = 0; ia < NA; ++ia) { gbh[ia] = 0.f; gls[ia] = 0.f; gWh[ia][0] = gWh[ia][1] = gWh[ia][2] = gWh[ia][3] = 0.f; }

            for (int c = 0; c < nchunk; ++c) {
                const int ci = s * nchunk + c;
                const float *x;
                if (DB) {
                    cp_async_wait<0>();
                    __syncthreads();            // chunk ci landed; every thread is done with chunk ci-1
                    if (ci + 1 < total_chunks) gather(ci + 1, (ci + 1) & 1);
                    x = recb + (ci & 1) * RC * RSS;
                } else {
                    __syncthreads();            // previous users of recb are done
                    gather(ci, 0);
                    cp_async_wait<0>();
                    __syncthreads();
                    x = recb;
                }
                const int row0 = g * a.Rg + c * RC;

                PGM_TR(1)
                // ---------------- forward ----------------
                half_forward<TM>(x, RSS, n, L, h1, h2, ho, ldo, tr, tc);

                PGM_TR(2)
                // ---------------- loss + d(loss)/d(head output), thread per row ----------------
                if (tid < RC) {
                    const int r = tid;
                    const bool valid = row0 + r < a.mb;
                    const float *rp = x + r * RSS;
                    if (half == 0) {
                        float lp = 0.f;
                        for (int d = 0; d < A; ++d) {
                            const float ls = n.ls[d], sd = expf(ls);
                            const float diff = rp[OP + d] - ho[r * ldo + d];
                            lp += -(diff * diff) / (2.f * sd * sd) - ls - 0.91893853320467274178f;
                        }
                        const float ratio = expf(lp - rp[OP + A]);
                        const float adv = rp[OP + A + 1 + 2 * M];
                        const float surr1 = ratio * adv;
                        const float rcl = fminf(fmaxf(ratio, 1.f - clip), 1.f + clip);
                        const float surr2 = rcl * adv;
                        const float w1 = surr1 < surr2 ? 1.f : (surr1 == surr2 ? 0.5f : 0.f);
                        const float inr = (ratio >= 1.f - clip && ratio <= 1.f + clip) ? 1.f : 0.f;
                        const float dmin = w1 * adv + (1.f - w1) * adv * inr;
                        const float dlp = valid ? -inv_mb * dmin * ratio : 0.f;
                        if (valid) loss_act -= fminf(surr1, surr2);
                        for (int d = 0; d < A; ++d) {
                            const float sd = expf(n.ls[d]);
                            const float iv = 1.f / (sd * sd);
                            const float diff = rp[OP + d] - ho[r * ldo + d];
                            ho[r * ldo + d] = dlp * diff * iv;
                            els[r * ldo + d] = dlp * (diff * diff * iv - 1.f);
                        }
                    } else {
                        for (int m = 0; m < M; ++m) {
                            const float V = ho[r * ldo + m];
                            const float vo = rp[OP + A + 1 + m];
                            const float R = rp[OP + A + 1 + M + m];
                            const float d = V - vo;
                            const float vcl = vo + fminf(fmaxf(d, -clip), clip);
                            const float ea = V - R, eb = vcl - R;
                            const float la = ea * ea, lb = eb * eb;
                            const float wa = la > lb ? 1.f : (la == lb ? 0.5f : 0.f);
                            const float pas = (d >= -clip && d <= clip) ? 1.f : 0.f;
                            if (valid) loss_val += fmaxf(la, lb);
                            ho[r * ldo + m] = valid ? vscale * (wa * 2.f * ea + (1.f - wa) * 2.f * eb * pas) : 0.f;
                        }
                    }
                }
                __syncthreads();

                PGM_TR(3)
                // ---------------- phase A: head weight grads, dz2 ----------------
                {
                    const int kg = tid & 15, a0 = tid >> 4;
#pragma unroll
                    for (int ia = 0; ia < NA; ++ia) {
                        const int aa = a0 + 16 * ia;
                        if (aa < KH) {
#pragma unroll 4
                            for (int r = 0; r < RC; ++r) {
                                const float sv = ho[r * ldo + aa];
                                const float4 hv = lds4(h2 + r * LDH + 4 * kg);
                                gWh[ia][0] = fmaf(sv, hv.x, gWh[ia][0]); gWh[ia][1] = fmaf(sv, hv.y, gWh[ia][1]);
                                gWh[ia][2] = fmaf(sv, hv.z, gWh[ia][2]); gWh[ia][3] = fmaf(sv, hv.w, gWh[ia][3]);
                                if (kg == 0) gbh[ia] += sv;
                                if (half == 0 && kg == 1) gls[ia] += els[r * ldo + aa];
                            }
                        }
                    }
                    float acc[TM][4];
#pragma unroll
                    for (int i = 0; i < TM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
                    for (int aa = 0; aa < KH; ++aa) {
                        const float4 w = lds4(n.Wh + aa * LDH + 4 * tc);
#pragma unroll
                        for (int i = 0; i < TM; ++i) {
                            const float sv = ho[(tr + 16 * i) * ldo + aa];
                            acc[i][0] = fmaf(sv, w.x, acc[i][0]); acc[i][1] = fmaf(sv, w.y, acc[i][1]);
                            acc[i][2] = fmaf(sv, w.z, acc[i][2]); acc[i][3] = fmaf(sv, w.w, acc[i][3]);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < TM; ++i) {
                        const float4 hv = lds4(h2 + (tr + 16 * i) * LDH + 4 * tc);
                        sts4(dz + (tr + 16 * i) * LDH + 4 * tc,
                             make_float4(acc[i][0] * (1.f - hv.x * hv.x), acc[i][1] * (1.f - hv.y * hv.y),
                                         acc[i][2] * (1.f - hv.z * hv.z), acc[i][3] * (1.f - hv.w * hv.w)));
                    }
                }
                __syncthreads();

                PGM_TR(4)
                // ---------------- phase B: dW2, db2, dz1 (into the h2 buffer) ----------------
                {
                    const int tj = tid & 15, tk = tid >> 4;
#pragma unroll 4
                    for (int r = 0; r < RC; ++r) {
                        const float4 dv = lds4(dz + r * LDH + 4 * tj);
                        const float4 hv = lds4(h1 + r * LDH + 4 * tk);
                        const float dd[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            gW2[jj][0] = fmaf(dd[jj], hv.x, gW2[jj][0]); gW2[jj][1] = fmaf(dd[jj], hv.y, gW2[jj][1]);
                            gW2[jj][2] = fmaf(dd[jj], hv.z, gW2[jj][2]); gW2[jj][3] = fmaf(dd[jj], hv.w, gW2[jj][3]);
                        }
                        if (tk == 0) { gb2[0] += dv.x; gb2[1] += dv.y; gb2[2] += dv.z; gb2[3] += dv.w; }
                    }
                    float acc[TM][4];
#pragma unroll
                    for (int i = 0; i < TM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll 2
                    for (int j = 0; j < H; j += 4) {
                        float4 av[TM], bv[4];
#pragma unroll
                        for (int i = 0; i < TM; ++i) av[i] = lds4(dz + (tr + 16 * i) * LDH + j);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) bv[jj] = lds4(n.W2 + (j + jj) * LDH + 4 * tc);
#pragma unroll
                        for (int i = 0; i < TM; ++i) {
                            const float aa[4] = {av[i].x, av[i].y, av[i].z, av[i].w};
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                acc[i][0] = fmaf(aa[jj], bv[jj].x, acc[i][0]); acc[i][1] = fmaf(aa[jj], bv[jj].y, acc[i][1]);
                                acc[i][2] = fmaf(aa[jj], bv[jj].z, acc[i][2]); acc[i][3] = fmaf(aa[jj], bv[jj].w, acc[i][3]);
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < TM; ++i) {
                        const float4 hv = lds4(h1 + (tr + 16 * i) * LDH + 4 * tc);
                        sts4(h2 + (tr + 16 * i) * LDH + 4 * tc,
                             make_float4(acc[i][0] * (1.f - hv.x * hv.x), acc[i][1] * (1.f - hv.y * hv.y),
                                         acc[i][2] * (1.f - hv.z * hv.z), acc[i][3] * (1.f - hv.w * hv.w)));
                    }
                }
                __syncthreads();

                PGM_TR(5)
                // ---------------- phase C: dW1, db1 ----------------
                {
                    const int tj = tid & 15, tk = tid >> 4;
#pragma unroll
                    for (int q = 0; q < KG1; ++q) {
                        const int k0 = 4 * (tk + 16 * q);
                        if (k0 < OP) {
#pragma unroll 4
                            for (int r = 0; r < RC; ++r) {
                                const float4 dv = lds4(h2 + r * LDH + 4 * tj);
                                const float4 xv = lds4(x + r * RSS + k0);
                                const float dd[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj) {
                                    gW1[q][jj][0] = fmaf(dd[jj], xv.x, gW1[q][jj][0]); gW1[q][jj][1] = fmaf(dd[jj], xv.y, gW1[q][jj][1]);
                                    gW1[q][jj][2] = fmaf(dd[jj], xv.z, gW1[q][jj][2]); gW1[q][jj][3] = fmaf(dd[jj], xv.w, gW1[q][jj][3]);
                                }
                                if (q == 0 && tk == 0) { gb1[0] += dv.x; gb1[1] += dv.y; gb1[2] += dv.z; gb1[3] += dv.w; }
                            }
                        }
                    }
                }
                // no barrier needed here: the next chunk's forward (or the barrier below) orders
                // the reuse of h2 / x behind every thread's phase C
            }

            PGM_TR(6)
            // ---- write this CTA's partial gradient of `half` to its scratch slot ----
            {
                float *gp = a.gpart + ((size_t)(task * 2 + half) * G + g) * a.NHP;
                const int tj = tid & 15, tk = tid >> 4;
                const int lb1 = H * O, lW2 = lb1 + H, lb2 = lW2 + H * H, lWh = lb2 + H, lbh = lWh + KH * H, lls = lbh + KH;
#pragma unroll
                for (int q = 0; q < KG1; ++q)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const int k = 4 * (tk + 16 * q) + kk;
                            if (k < O) gp[(4 * tj + jj) * O + k] = gW1[q][jj][kk];
                        }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    sts4(gp + lW2 + (4 * tj + jj) * H + 4 * tk, make_float4(gW2[jj][0], gW2[jj][1], gW2[jj][2], gW2[jj][3]));
                if (tk == 0) {
                    sts4(gp + lb1 + 4 * tj, make_float4(gb1[0], gb1[1], gb1[2], gb1[3]));
                    sts4(gp + lb2 + 4 * tj, make_float4(gb2[0], gb2[1], gb2[2], gb2[3]));
                }
                const int kg = tid & 15, a0 = tid >> 4;
#pragma unroll
                for (int ia = 0; ia < NA; ++ia) {
                    const int aa = a0 + 16 * ia;
                    if (aa < KH) {
                        sts4(gp + lWh + aa * H + 4 * kg, make_float4(gWh[ia][0], gWh[ia][1], gWh[ia][2], gWh[ia][3]));
                        if (kg == 0) gp[lbh + aa] = gbh[ia];
                        if (half == 0 && kg == 1) gp[lls + aa] = gls[ia];
                    }
                }
            }
            if (half == 0 && g == 0 && tid == 0) {   // entropy with the parameters this step started from
                float ent = 0.f;
                for (int d = 0; d < A; ++d) ent += 0.5f + 0.91893853320467274178f + n.ls[d];
                loss_ent += ent;
            }
        }   // halves

        PGM_TR(7)
        sync_group<C>();   // (1) all partial gradients of this task are in L2
        PGM_TR(8)

        // ---- reduce my slice over the G partials, squared-norm partial ----
        float sq = 0.f;
#pragma unroll
        for (int hh = 0; hh < NHALF; ++hh) {
            const int half = half0 + hh;
            const int nH = L.half_size(half);
            const int per = round_up((nH + G - 1) / G, 4);
            const int s0 = g * per, s1 = min(nH, s0 + per);
            float *slot0 = a.gpart + (size_t)(task * 2 + half) * G * a.NHP;
            const int lls = L.n_base + L.head_dim(half) * H + L.head_dim(half);
            for (int e = s0 + tid; e < s1; e += NTHREADS) {
                float gs = 0.f;
#pragma unroll
                for (int gg = 0; gg < G; ++gg) gs += __ldcg(slot0 + (size_t)gg * a.NHP + e);
                if (half == 0 && e >= lls) gs -= ecoef;      // d(-ecoef * entropy)/d logstd
                sq = fmaf(gs, gs, sq);
                if (a.grad_only) a.grad_out[(size_t)task * L.n_par + L.to_global(half, e)] = gs;
                else slot0[(size_t)g * a.NHP + e] = gs;   // keep the reduced value in my own slot (only I read it)
            }
        }
        sq = block_sum(sq, red);
        if (tid == 0) a.ssq[task * 16 + rank] = sq;
        if (a.grad_only) break;

        if (tid == 0) {   // Adam scalars of step k = step0 + s + 1, in double
            b1pow *= a.hy.beta1; b2pow *= a.hy.beta2;
            sh_d[0] = lr / (1.0 - b1pow);            // step_size
            sh_d[1] = 1.0 / sqrt(1.0 - b2pow);       // 1 / bias_correction2_sqrt
        }
        sync_group<C>();   // (2) squared-norm partials visible

        float tot = 0.f;
#pragma unroll
        for (int rr = 0; rr < C; ++rr) tot += __ldcg(a.ssq + task * 16 + rr);
        const float coef = fminf(1.f, (float)a.hy.max_grad_norm / (sqrtf(tot) + 1e-6f));
        const float step_size = (float)sh_d[0], ibc2 = (float)sh_d[1];
#pragma unroll
        for (int hh = 0; hh < NHALF; ++hh) {
            const int half = half0 + hh;
            const int nH = L.half_size(half);
            const int per = round_up((nH + G - 1) / G, 4);
            const int s0 = g * per, s1 = min(nH, s0 + per);
            const float *mine = a.gpart + ((size_t)(task * 2 + half) * G + g) * a.NHP;
            for (int e = s0 + tid; e < s1; e += NTHREADS) {
                const size_t gi = (size_t)task * L.n_par + L.to_global(half, e);
                const float gr = __ldcg(mine + e) * coef;
                float m = a.adam_m[gi], v = a.adam_v[gi], pw = a.params[gi];
                m = fmaf(gr - m, omb1, m);                 // exp_avg.lerp_(grad, 1 - beta1)
                v = fmaf(omb2 * gr, gr, v * b2f);          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
                const float denom = sqrtf(v) * ibc2 + aeps;
                pw -= step_size * (m / denom);
                a.adam_m[gi] = m; a.adam_v[gi] = v; a.params[gi] = pw;
            }
        }
        sync_group<C>();   // (3) new parameters of this task are in L2
#pragma unroll
        for (int hh = 0; hh < NHALF; ++hh) halfnet_load<true>(net[hh], gparams, L, half0 + hh);
        __syncthreads();
    }

    // ---- losses: per-CTA partial sums -> rank 0 combines in fixed order ----
    {
        const float la = block_sum(loss_act, red), lv = block_sum(loss_val, red), le = block_sum(loss_ent, red);
        if (tid == 0) {
            a.lpart[(task * 16 + rank) * 4 + 0] = lv;
            a.lpart[(task * 16 + rank) * 4 + 1] = la;
            a.lpart[(task * 16 + rank) * 4 + 2] = le;
        }
        sync_group<C>();
        if (rank == 0 && tid == 0) {
            float sv = 0.f, sa = 0.f, se = 0.f;
            for (int rr = 0; rr < C; ++rr) {
                sv += __ldcg(a.lpart + (task * 16 + rr) * 4 + 0);
                sa += __ldcg(a.lpart + (task * 16 + rr) * 4 + 1);
                se += __ldcg(a.lpart + (task * 16 + rr) * 4 + 2);
            }
            const float ns = (float)a.nsteps;
            a.losses[task * 3 + 0] = sv * 0.5f / ((float)a.mb * M) / ns;
            a.losses[task * 3 + 1] = sa * inv_mb / ns;
            a.losses[task * 3 + 2] = se / ns;
            if (!a.grad_only) a.adam_step[task] = step0 + a.nsteps;
        }
    }
}

}  // namespace pgm
#include "k3_fast.cuh"
#include "k3_tc.cuh"
#include "k3_tcw.cuh"
namespace pgm {

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct K3Plan {
    int C, G, TM, KG1, NA, RC, Rg, RSG, RSS, NHP;
    bool DB, fast, tc, tcw;
    int tail;     // fast path step tail (k3_fast.cuh): 0 three cluster barriers, 1 redundant Adam, 2 TMA multicast broadcast
    int rs;
    int stage_floats;
    size_t smem;
    size_t off_rec, off_gpart, off_ssq, off_lpart, off_trace, off_mv, total;
    size_t off_xp, off_scr, off_w1img, off_pmv;      // wide tensor-core path (k3_tcw.cuh)
};

static size_t k3_smem_bytes(const NetLayout &L, int C, int TM, bool DB, int RSS) {
    const int RC = 16 * TM;
    const int ldo = ((L.A > L.M ? L.A : L.M) | 1);
    size_t f = 0;
    if (C == 1) f += halfnet_smem_floats(L, 0) + halfnet_smem_floats(L, 1);
    else f += halfnet_smem_floats(L, 0) > halfnet_smem_floats(L, 1) ? halfnet_smem_floats(L, 0) : halfnet_smem_floats(L, 1);
    f += 3 * (size_t)RC * LDH + 2 * (size_t)round_up(RC * ldo, 4);
    f += (size_t)(DB ? 2 : 1) * RC * RSS;
    return f * sizeof(float);
}

// shapes the tensor-core kernel is instantiated for (Walker2d / HalfCheetah and Hopper-v3)
static bool k3_tc_dims(int O, int A, int M) { return (O == 17 && A == 6 && M == 2) || (O == 11 && A == 3 && M == 3); }
static size_t k3_tc_mv_bytes(int P) { return (size_t)P * 2 * 2 * 2 * TC_NHP * sizeof(float); }   // [P][half][rs <= 2][m|v]
constexpr int K3_CLUSTER_TC = 32;   // `cluster` values that select the tensor-core path explicitly: 32 = 2 CTAs per task,
constexpr int K3_CLUSTER_TC2 = 64;  // 64 = 4 CTAs per task (row tiles split over two CTAs per network half)
constexpr int K3_CLUSTER_TC4 = 128; // 128 = 8 CTAs per task (wide-observation kernel only)
// shape the wide-observation tensor-core kernel (k3_tcw.cuh) is instantiated for: Humanoid
static bool k3_tcw_dims(int O, int A, int M) { return O == 376 && A == 17 && M == 2; }
static size_t k3_tcw_bytes(int P, int S, int O, int A, size_t *oxp, size_t *oscr, size_t *ow1, size_t *opmv) {
    size_t off = 0;
    auto seg = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t a0 = seg((size_t)P * S * TW_XROW_HW * sizeof(__half));
    const size_t a1 = seg((size_t)P * S * TW_SCF * sizeof(float));
    const size_t a2 = seg((size_t)P * 2 * TW_NB * 8192 * sizeof(__half));
    const size_t a3 = seg((size_t)P * 2 * 3 * tw_nhp(O, A) * sizeof(float));
    if (oxp) { *oxp = a0; *oscr = a1; *ow1 = a2; *opmv = a3; }
    return off;
}

// Clusters of 2*rs CTAs of the wide kernel that the device keeps resident at once (a cluster has to fit one GPC, so this
// is less than SMs / cluster size); -1 when it cannot be queried (no device: workspace sizing on a build machine).
static int k3_tcw_resident_clusters(int rs) {
    static int cache[5] = {-2, -2, -2, -2, -2};
    if (cache[rs] != -2) return cache[rs];
    const void *kern = rs == 4 ? (const void *)k3_tcw_kernel<376, 17, 2, 4>
                               : (rs == 2 ? (const void *)k3_tcw_kernel<376, 17, 2, 2> : (const void *)k3_tcw_kernel<376, 17, 2, 1>);
    const size_t smem = tw_smem_layout().total;
    int n = -1;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * rs * 64); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2 * rs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) n = -1;
    }
    if (n < 0) (void)cudaGetLastError();
    return cache[rs] = n;
}

// Step tails of the FFMA cluster kernel (k3_fast.cuh, template parameter TAIL); every variant performs the same arithmetic
// in the same order, so results are bit-identical (tests/test_gpu_kernels.py::test_k3_step_tails_bit_identical). Same-box
// A/B at the headline configuration (6 tasks x 16 CTAs, profiles/r02/README.md), ms per MOPG iteration:
//   tail 0  3.876  three cluster barriers: partial images pulled through ld.shared::cluster, parameters pushed (round 1)
//   tail 1  4.04   two barriers: register tiles pushed to the slice owner, whole-half Adam in every CTA
//   tail 2  3.734  parameter broadcast by TMA bulk copy with cluster multicast, third barrier gone
//   tail 3  3.839  tail 2 + gradient exchange through L2 scratch slots
//   tail 4  3.712  tail 2 + gradient reduce-scatter pushed by cp.async.bulk shared::cta -> shared::cluster, first barrier gone
//   tail 5  3.575  tail 4 + norm partials by st.async with mbarrier completion: no cluster barrier left in the step (default)
//   tail 6  3.557* tail 5 + the W2-only gradient slices pushed right after the {dW2, dz1} phase (*same box: tail 5 3.525;
//                  the mid-phase proxy fence costs more than the early transfer saves)
// A tail can be forced per call by OR-ing a flag into `cluster` (tests, A/B) or per process with PGM_K3_TAIL=0..5.
constexpr int K3_DEFAULT_TAIL = 5;
constexpr int K3_CLUSTER_TAIL2 = 0x100;    // tail 1
constexpr int K3_CLUSTER_TAILMC = 0x200;   // tail 2
constexpr int K3_CLUSTER_TAILGL = 0x400;   // tail 3
constexpr int K3_CLUSTER_TAIL0 = 0x800;    // tail 0
constexpr int K3_CLUSTER_TAILBP = 0x1000;  // tail 4
constexpr int K3_CLUSTER_TAILNB = 0x2000;  // tail 5
constexpr int K3_CLUSTER_TAILEP = 0x4000;  // tail 6
constexpr int K3_CLUSTER_FLAGS = K3_CLUSTER_TAIL2 | K3_CLUSTER_TAILMC | K3_CLUSTER_TAILGL | K3_CLUSTER_TAIL0 | K3_CLUSTER_TAILBP |
                                 K3_CLUSTER_TAILNB | K3_CLUSTER_TAILEP;

static int k3_plan(K3Plan &pl, int P, int S, int mb, int O, int A, int M, int cluster, int sms) {
    NetLayout L(O, A, M);
    const bool tail2 = (cluster & K3_CLUSTER_TAIL2) != 0, tailmc = (cluster & K3_CLUSTER_TAILMC) != 0;
    const bool tailgl = (cluster & K3_CLUSTER_TAILGL) != 0, tail0 = (cluster & K3_CLUSTER_TAIL0) != 0;
    const bool tailbp = (cluster & K3_CLUSTER_TAILBP) != 0, tailnb = (cluster & K3_CLUSTER_TAILNB) != 0;
    const bool tailep = (cluster & K3_CLUSTER_TAILEP) != 0;
    cluster &= ~K3_CLUSTER_FLAGS;
    pl.tc = false; pl.tcw = false; pl.tail = 0; pl.off_mv = 0;
    // wide observations (Humanoid): the streamed tensor-core kernel; row split over as many CTAs per half as fill the SMs
    if (k3_tcw_dims(O, A, M) && (cluster == 0 || cluster == K3_CLUSTER_TC || cluster == K3_CLUSTER_TC2 || cluster == K3_CLUSTER_TC4)) {
        const int tiles = (mb + 127) / 128;
        if (cluster == 0) {
            // most CTAs per task whose clusters are all resident at once (a second wave of clusters doubles the step)
            auto fits = [&](int rs) {
                const int res = k3_tcw_resident_clusters(rs);
                return tiles >= rs && 2 * rs * P <= sms && (res < 0 || P <= res);
            };
            pl.rs = fits(4) ? 4 : (fits(2) ? 2 : 1);
        }
        else pl.rs = cluster == K3_CLUSTER_TC4 ? 4 : (cluster == K3_CLUSTER_TC2 ? 2 : 1);
        pl.tc = true; pl.tcw = true; pl.C = 2 * pl.rs; pl.G = 1; pl.TM = 0; pl.KG1 = 0; pl.NA = 0; pl.RC = 128; pl.Rg = 0;
        pl.RSG = 0; pl.RSS = 0; pl.NHP = 64; pl.DB = false; pl.fast = false; pl.stage_floats = 0;
        pl.smem = tw_smem_layout().total;
        size_t off = 0;
        auto seg = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
        pl.off_rec = 0;
        pl.off_gpart = seg((size_t)P * 2 * pl.G * pl.NHP * sizeof(float));
        pl.off_ssq = seg((size_t)P * 16 * sizeof(float));
        pl.off_lpart = seg((size_t)P * 16 * 4 * sizeof(float));
#ifdef PGM_K3_TRACE
        pl.off_trace = seg((size_t)P * 8 * 2 * 48 * sizeof(long long));
#else
        pl.off_trace = 0;
#endif
        const size_t base = off;
        off += k3_tcw_bytes(P, S, O, A, &pl.off_xp, &pl.off_scr, &pl.off_w1img, &pl.off_pmv);
        pl.off_xp += base; pl.off_scr += base; pl.off_w1img += base; pl.off_pmv += base;
        pl.total = off;
        return PGM_OK;
    }
    // auto: the tensor-core path wins as soon as the FFMA path can no longer give every task a 16-CTA cluster (measured on
    // a B200, profiles/k3_sweep.py: 1.2x at P = 8, 2.4x at 16, 3.3x from 37 tasks on; below 8 tasks FFMA is 6 % faster)
    if (cluster == 0 && k3_tc_dims(O, A, M) && P >= 8) {
        // 4 CTAs per task (row tiles split over two CTAs per half) while they fit in one wave: 4-CTA clusters tile the 148
        // SMs less densely than pairs (measured: 32 clusters run in one wave, 37 need two)
        cluster = (mb > 128 && 4 * P <= sms - 20) ? K3_CLUSTER_TC2 : K3_CLUSTER_TC;
    }
    pl.rs = 1;
    if (cluster == K3_CLUSTER_TC || cluster == K3_CLUSTER_TC2) {
        pl.rs = cluster == K3_CLUSTER_TC2 ? 2 : 1;
        PGM_REQUIRE(k3_tc_dims(O, A, M), "ppo: the tensor-core path is built for (O,A,M) = (17,6,2) and (11,3,3), got (%d,%d,%d)", O, A, M);
        pl.tc = true; pl.C = 2 * pl.rs; pl.G = 1; pl.TM = 0; pl.KG1 = 0; pl.NA = 0; pl.RC = 128; pl.Rg = 0;
        pl.RSG = rec_stride(L); pl.RSS = 0; pl.NHP = 64; pl.DB = false; pl.fast = false; pl.stage_floats = 0;
        pl.smem = tc_smem_layout().total;
        size_t off = 0;
        auto seg = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
        pl.off_rec = seg((size_t)P * S * pl.RSG * sizeof(float));
        pl.off_gpart = seg((size_t)P * 2 * pl.G * pl.NHP * sizeof(float));
        pl.off_ssq = seg((size_t)P * 16 * sizeof(float));
        pl.off_lpart = seg((size_t)P * 16 * 4 * sizeof(float));
        pl.off_mv = seg(k3_tc_mv_bytes(P));
#ifdef PGM_K3_TRACE
        pl.off_trace = seg((size_t)P * 16 * 4 * 16 * 2 * sizeof(long long));
#else
        pl.off_trace = 0;
#endif
        pl.total = off;
        return PGM_OK;
    }
    PGM_REQUIRE(O >= 1 && O <= 384 && A >= 1 && A <= 32 && M >= 1 && M <= 16,
                "ppo: unsupported dims O=%d A=%d M=%d (O<=384, A<=32, M<=16)", O, A, M);
    PGM_REQUIRE(cluster == 0 || cluster == 1 || cluster == 2 || cluster == 4 || cluster == 8 || cluster == 16,
                "ppo: cluster must be 0 (auto), 1, 2, 4, 8, 16 (FP32 FFMA paths), 32, 64 (tensor-core paths) or 128 (wide tensor-core path) (got %d)", cluster);
    int C = cluster;
    if (C == 0) {   // fill the SMs: double the cluster while every task still gets its CTAs resident at once
        C = 2;      // one CTA per network half is the throughput configuration (large populations)
        // 16-CTA clusters are non-portable and at most one fits a GPC (ncu: launch__cluster_max_active = 7 with
        // this kernel's shared-memory footprint): use them only when every task's cluster is resident at once
        const int cmax = (P <= 7) ? 16 : 8;
        while (C < cmax && (long long)P * C * 2 <= sms && mb / C >= 32) C *= 2;
    }
    const bool big = (L.OP > 64) || (A > 16) || (M > 16);
    pl.C = C; pl.G = C == 1 ? 1 : C / 2;
    pl.KG1 = big ? 6 : 1; pl.NA = big ? 2 : 1;
    const int rows = (mb + pl.G - 1) / pl.G;          // rows per CTA per step
    pl.TM = big ? 2 : (rows <= 32 ? 2 : 4);
    pl.RC = 16 * pl.TM;
    pl.Rg = round_up(rows, pl.RC);
    pl.RSG = rec_stride(L);
    pl.RSS = stride4odd(pl.RSG);
    {
        const int i0 = halfnet_smem_floats(L, 0), i1 = halfnet_smem_floats(L, 1);
        pl.NHP = round_up(i0 > i1 ? i0 : i1, 64);   // covers both the reference-order half and its padded image
    }
    pl.DB = (C > 1) && !big;
    pl.fast = (C > 1) && !big && L.OP <= 48 && pl.RC * (pl.RSG / 4) <= 4 * NTHREADS;
    pl.stage_floats = 0;
    if (pl.fast) {
        const int s0 = k3_fast_stage_floats(L.OP, A), s1 = k3_fast_stage_floats(L.OP, M);
        pl.stage_floats = s0 > s1 ? s0 : s1;
        // step tail: default from K3_DEFAULT_TAIL; others on request (cluster flags, or PGM_K3_TAIL=0|1|2 in the
        // environment for whole-program A/B runs)
        static const int env_tail = [] { const char *e = getenv("PGM_K3_TAIL"); return (e && e[0] >= '0' && e[0] <= '6') ? e[0] - '0' : -1; }();
        pl.tail = tail0 ? 0 : (tail2 ? 1 : (tailmc ? 2 : (tailgl ? 3 : (tailbp ? 4 : (tailnb ? 5 : (tailep ? 6 : (env_tail >= 0 ? env_tail : K3_DEFAULT_TAIL)))))));
        if (pl.tail == 1 && !(pl.G >= 2 && k3_fast_smem_bytes(L, pl.TM, pl.RSS, pl.NHP, pl.stage_floats, pl.G, true) <= 227 * 1024))
            pl.tail = 0;
        if (pl.tail >= 2 && pl.G < 2) pl.tail = 0;
        if (pl.tail >= 4 && k3_fast_smem_bytes(L, pl.TM, pl.RSS, pl.NHP, pl.stage_floats, pl.G, false, false, true) > 227 * 1024)
            pl.tail = 2;
        pl.smem = k3_fast_smem_bytes(L, pl.TM, pl.RSS, pl.NHP, pl.stage_floats, pl.G, pl.tail == 1, pl.tail == 3, pl.tail >= 4);
    } else {
        pl.smem = k3_smem_bytes(L, C, pl.TM, pl.DB, pl.RSS);
    }
    PGM_REQUIRE(pl.smem <= 227 * 1024, "ppo: configuration needs %zu B of shared memory (O=%d, cluster=%d)", pl.smem, O, C);
    size_t off = 0;
    auto seg = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    pl.off_rec = seg((size_t)P * S * pl.RSG * sizeof(float));
    pl.off_gpart = seg((size_t)P * 2 * (pl.G + 1) * pl.NHP * sizeof(float));      // + one staging row per half (pstage)
    pl.off_ssq = seg((size_t)P * 16 * sizeof(float));
    pl.off_lpart = seg((size_t)P * 16 * 4 * sizeof(float));
#ifdef PGM_K3_TRACE
    pl.off_trace = seg((size_t)P * 16 * 4 * 16 * 2 * sizeof(long long));
#else
    pl.off_trace = 0;
#endif
    pl.total = off;
    return PGM_OK;
}

template <typename Kern>
static int k3_launch_k(Kern kern, int C, const K3Args &a, const K3Plan &pl, int P, cudaStream_t st) {
    PGM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    if (C > 8) PGM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(P * C); cfg.blockDim = dim3(pl.tc ? TC_THREADS : NTHREADS); cfg.dynamicSmemBytes = pl.smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = C > 1 ? 1 : 0;
    PGM_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    return PGM_OK;
}

template <int C>
static int k3_launch_c(const K3Args &a, const K3Plan &pl, int P, cudaStream_t st) {
    if constexpr (C > 1) {
        if (pl.fast) {
            if (pl.tail == 1)
                return pl.TM == 2 ? k3_launch_k(k3_ppo_fast_kernel<C, 2, 1>, C, a, pl, P, st)
                                  : k3_launch_k(k3_ppo_fast_kernel<C, 4, 1>, C, a, pl, P, st);
            if (pl.tail == 2)
                return pl.TM == 2 ? k3_launch_k(k3_ppo_fast_kernel<C, 2, 2>, C, a, pl, P, st)
                                  : k3_launch_k(k3_ppo_fast_kernel<C, 4, 2>, C, a, pl, P, st);
            if (pl.tail == 3)
                return pl.TM == 2 ? k3_launch_k(k3_ppo_fast_kernel<C, 2, 3>, C, a, pl, P, st)
                                  : k3_launch_k(k3_ppo_fast_kernel<C, 4, 3>, C, a, pl, P, st);
            if (pl.tail == 4)
                return pl.TM == 2 ? k3_launch_k(k3_ppo_fast_kernel<C, 2, 4>, C, a, pl, P, st)
                                  : k3_launch_k(k3_ppo_fast_kernel<C, 4, 4>, C, a, pl, P, st);
            if (pl.tail == 5)
                return pl.TM == 2 ? k3_launch_k(k3_ppo_fast_kernel<C, 2, 5>, C, a, pl, P, st)
                                  : k3_launch_k(k3_ppo_fast_kernel<C, 4, 5>, C, a, pl, P, st);
            if (pl.tail == 6)
                return pl.TM == 2 ? k3_launch_k(k3_ppo_fast_kernel<C, 2, 6>, C, a, pl, P, st)
                                  : k3_launch_k(k3_ppo_fast_kernel<C, 4, 6>, C, a, pl, P, st);

            return pl.TM == 2 ? k3_launch_k(k3_ppo_fast_kernel<C, 2, 0>, C, a, pl, P, st)
                              : k3_launch_k(k3_ppo_fast_kernel<C, 4, 0>, C, a, pl, P, st);
        }
        if (pl.KG1 == 6) return k3_launch_k(k3_ppo_kernel<C, 2, 6, 2, false>, C, a, pl, P, st);
        return pl.TM == 2 ? k3_launch_k(k3_ppo_kernel<C, 2, 1, 1, true>, C, a, pl, P, st)
                          : k3_launch_k(k3_ppo_kernel<C, 4, 1, 1, true>, C, a, pl, P, st);
    } else {
        if (pl.KG1 == 6) return k3_launch_k(k3_ppo_kernel<1, 2, 6, 2, false>, 1, a, pl, P, st);
        return pl.TM == 2 ? k3_launch_k(k3_ppo_kernel<1, 2, 1, 1, false>, 1, a, pl, P, st)
                          : k3_launch_k(k3_ppo_kernel<1, 4, 1, 1, false>, 1, a, pl, P, st);
    }
}

template <typename Kern>
static int k3_launch_w(Kern kern, const K3Args &a, const TwExtra &x, const K3Plan &pl, int P, cudaStream_t st) {
    PGM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(P * pl.C); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = pl.smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = pl.C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    PGM_CUDA(cudaLaunchKernelEx(&cfg, kern, a, x));
    return PGM_OK;
}

static int k3_launch(const K3Args &a, const K3Plan &pl, int P, cudaStream_t st) {
    if (pl.tc) {
        if (a.L.O == 17) return pl.rs == 2 ? k3_launch_k(k3_tc_kernel<17, 6, 2, 2>, pl.C, a, pl, P, st)
                                           : k3_launch_k(k3_tc_kernel<17, 6, 2, 1>, pl.C, a, pl, P, st);
        return pl.rs == 2 ? k3_launch_k(k3_tc_kernel<11, 3, 3, 2>, pl.C, a, pl, P, st)
                          : k3_launch_k(k3_tc_kernel<11, 3, 3, 1>, pl.C, a, pl, P, st);
    }
    switch (pl.C) {
        case 1: return k3_launch_c<1>(a, pl, P, st);
        case 2: return k3_launch_c<2>(a, pl, P, st);
        case 4: return k3_launch_c<4>(a, pl, P, st);
        case 8: return k3_launch_c<8>(a, pl, P, st);
        case 16: return k3_launch_c<16>(a, pl, P, st);
    }
    set_error("ppo: bad cluster %d", pl.C);
    return PGM_ERR_ARG;
}

static int sm_count(int &sms) {
    int dev = 0;
    PGM_CUDA(cudaGetDevice(&dev));
    PGM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    return PGM_OK;
}

}  // namespace pgm

using namespace pgm;

#ifdef PGM_K3_TRACE
static size_t g_last_trace_offset = 0;
// instrumented builds only (libpgmorl_b200_trace.so): byte offset of the clock marks inside the workspace of the last launch
extern "C" size_t pgm_ppo_last_trace_offset() { return g_last_trace_offset; }
#endif

extern "C" size_t pgm_ppo_workspace_bytes(int P, int S, int O, int A, int M, int cluster) {
    // upper bound over every plan the launcher may choose for these dims (cluster 0 = auto)
    NetLayout L(O, A, M);
    cluster &= ~K3_CLUSTER_FLAGS;
    const int G = cluster == 0 ? 8 : (cluster == 1 ? 1 : cluster / 2);
    const int i0 = halfnet_smem_floats(L, 0), i1 = halfnet_smem_floats(L, 1);
    const int NHP = round_up(i0 > i1 ? i0 : i1, 64);
    size_t t = 0;
    auto seg = [&](size_t b) { t += (b + 255) / 256 * 256; };
    seg((size_t)P * S * rec_stride(L) * sizeof(float));
    seg((size_t)P * 2 * (G + 1) * NHP * sizeof(float));
    seg((size_t)P * 16 * sizeof(float));
    seg((size_t)P * 64 * sizeof(float));
    if (k3_tc_dims(O, A, M)) seg(k3_tc_mv_bytes(P));
    if (k3_tcw_dims(O, A, M)) seg(k3_tcw_bytes(P, S, O, A, nullptr, nullptr, nullptr, nullptr) + 1024);
#ifdef PGM_K3_TRACE
    seg((size_t)P * 16 * 4 * 16 * 2 * sizeof(long long));
#endif
    return t;
}

static int ppo_common(float *params, float *adam_m, float *adam_v, int32_t *adam_step, const double *lr,
                      const float *obs, size_t obs_ts, const float *action, const float *logp_old,
                      const float *value_old, size_t v_ts, const float *returns, const float *adv,
                      const int32_t *perm, int perm_shared, int E, int B, int mb, const pgm_ppo_hyper *hy,
                      float *losses, float *grad_out, void *workspace, size_t wbytes, int cluster, int P, int S,
                      int O, int A, int M, cudaStream_t st) {
    PGM_REQUIRE(params && obs && action && logp_old && value_old && returns && adv && perm && hy && losses && workspace,
                "ppo: null pointer argument");
    PGM_REQUIRE(P > 0 && S > 0 && mb > 0 && mb <= S, "ppo: bad sizes P=%d S=%d mb=%d", P, S, mb);
    PGM_REQUIRE(((uintptr_t)workspace & 255) == 0, "ppo: workspace must be 256-byte aligned");
    int sms = 148;
    if (int rc = sm_count(sms)) return rc;
    K3Plan pl;
    if (int rc = k3_plan(pl, P, S, mb, O, A, M, cluster, sms)) return rc;
    if (pl.total > wbytes) {
        set_error("ppo: workspace too small: need %zu bytes, got %zu", pl.total, wbytes);
        return PGM_ERR_WORKSPACE;
    }
    char *ws = (char *)workspace;
    K3Args a;
    a.params = params; a.adam_m = adam_m; a.adam_v = adam_v; a.adam_step = adam_step; a.lr = lr;
    a.rec = (const float *)(ws + pl.off_rec); a.perm = perm; a.losses = losses;
    a.gpart = (float *)(ws + pl.off_gpart); a.pstage = a.gpart + (size_t)P * 2 * pl.G * pl.NHP; a.ssq = (float *)(ws + pl.off_ssq); a.lpart = (float *)(ws + pl.off_lpart);
#ifdef PGM_K3_TRACE
    a.trace = (long long *)(ws + pl.off_trace);
    g_last_trace_offset = pl.off_trace;
#else
    a.trace = nullptr;
#endif
    a.mv = (float *)(ws + pl.off_mv);
    a.grad_out = grad_out; a.perm_shared = perm_shared; a.E = E; a.B = B; a.mb = mb; a.S = S;
    a.grad_only = grad_out != nullptr; a.nsteps = a.grad_only ? 1 : E * B;
    a.Rg = pl.Rg; a.RSG = pl.RSG; a.RSS = pl.RSS; a.NHP = pl.NHP; a.stage_floats = pl.stage_floats; a.rs = pl.rs; a.hy = *hy; a.L = NetLayout(O, A, M);
    if (pl.tcw) {
        TwExtra x;
        x.xp = (const __half *)(ws + pl.off_xp); x.scr = (const float *)(ws + pl.off_scr);
        x.w1img = (__half *)(ws + pl.off_w1img); x.pmv = (float *)(ws + pl.off_pmv);
        const size_t total = (size_t)P * S * (TW_NB * 64 + TW_SCF);
        int blocks = (int)((total + 255) / 256);
        if (blocks > sms * 16) blocks = sms * 16;
        k3w_pack_kernel<376, 17, 2><<<blocks, 256, 0, st>>>(obs, obs_ts, action, logp_old, value_old, v_ts, returns, adv,
                                                            (__half *)(ws + pl.off_xp), (float *)(ws + pl.off_scr), P, S);
        PGM_CUDA(cudaGetLastError());
        // features O+1 .. 383 of the last W1 block image are never written by the kernel: they must be zero, not stale bytes
        PGM_CUDA(cudaMemsetAsync(ws + pl.off_w1img, 0, (size_t)P * 2 * TW_NB * 8192 * sizeof(__half), st));
        if (pl.rs == 4) return k3_launch_w(k3_tcw_kernel<376, 17, 2, 4>, a, x, pl, P, st);
        if (pl.rs == 2) return k3_launch_w(k3_tcw_kernel<376, 17, 2, 2>, a, x, pl, P, st);
        return k3_launch_w(k3_tcw_kernel<376, 17, 2, 1>, a, x, pl, P, st);
    }
    {
        const size_t total = (size_t)P * S * pl.RSG;
        int blocks = (int)((total + 255) / 256);
        if (blocks > sms * 16) blocks = sms * 16;
        k3_pack_kernel<<<blocks, 256, 0, st>>>(obs, obs_ts, action, logp_old, value_old, v_ts, returns, adv,
                                               (float *)(ws + pl.off_rec), P, S, a.L, pl.RSG);
        PGM_CUDA(cudaGetLastError());
    }
    // padding entries of the gradient slots are never written by the kernel: keep them zero
    PGM_CUDA(cudaMemsetAsync(ws + pl.off_gpart, 0, (size_t)P * 2 * (pl.G + 1) * pl.NHP * sizeof(float), st));
    return k3_launch(a, pl, P, st);
}

extern "C" int pgm_ppo_update_f32(float *params, float *adam_m, float *adam_v, int32_t *adam_step, const double *lr,
                                  const float *obs, size_t obs_task_stride, const float *action,
                                  const float *logp_old, const float *value_old, size_t value_task_stride,
                                  const float *returns, const float *adv, const int32_t *perm, int perm_shared,
                                  int E, int B, const pgm_ppo_hyper *hyper_host, float *losses, void *workspace,
                                  size_t workspace_bytes, int cluster, int P, int S, int O, int A, int M,
                                  void *stream) {
    PGM_REQUIRE(adam_m && adam_v && adam_step && lr, "pgm_ppo_update_f32: null optimizer state");
    PGM_REQUIRE(E > 0 && B > 0 && S / B > 0, "pgm_ppo_update_f32: bad E=%d B=%d for S=%d", E, B, S);
    // the reference's BatchSampler(drop_last=True) yields S // (S // B) minibatches per epoch (a2c/storage.py:133-137), which
    // is B only when the remainder S % B is smaller than one minibatch; otherwise it would take more than E*B steps
    PGM_REQUIRE(S / (S / B) == B, "pgm_ppo_update_f32: S=%d samples in B=%d minibatches of %d leave a remainder of %d >= one "
                "minibatch: the reference would run %d minibatches per epoch; choose S, B with S %% B < S / B", S, B, S / B, S % B, S / (S / B));
    return ppo_common(params, adam_m, adam_v, adam_step, lr, obs, obs_task_stride, action, logp_old, value_old,
                      value_task_stride, returns, adv, perm, perm_shared, E, B, S / B, hyper_host, losses, nullptr,
                      workspace, workspace_bytes, cluster, P, S, O, A, M, (cudaStream_t)stream);
}

extern "C" int pgm_ppo_grad_f32(const float *params, const float *obs, size_t obs_task_stride, const float *action,
                                const float *logp_old, const float *value_old, size_t value_task_stride,
                                const float *returns, const float *adv, const int32_t *idx, int mb,
                                const pgm_ppo_hyper *hyper_host, float *grad, float *losses, void *workspace,
                                size_t workspace_bytes, int cluster, int P, int S, int O, int A, int M, void *stream) {
    PGM_REQUIRE(grad, "pgm_ppo_grad_f32: null grad");
    return ppo_common(const_cast<float *>(params), nullptr, nullptr, nullptr, nullptr, obs, obs_task_stride, action,
                      logp_old, value_old, value_task_stride, returns, adv, idx, 1, 1, 1, mb, hyper_host, losses, grad,
                      workspace, workspace_bytes, cluster, P, S, O, A, M, (cudaStream_t)stream);
}
