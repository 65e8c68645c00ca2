// K4 front-end -- training data of the hyperbolic prediction models, on the device (FLOAT64 compares, bit-identical).
//
// Replaces the neighbourhood search of collect_nearest_data and the widening loop around it
// (morl/population_2d.py:12-21,37-54, morl/population_3d.py:13-21,33-49): for a population member with opt-graph
// node k, the edges (i -> s) whose source node i satisfies |objs_k - objs_i| < |objs_k| * threshold in every objective,
// with (threshold, sigma) doubled from (0.1, 0.03) until the selected edges carry more than 3 pairwise-distinct
// successor weights (L2 distance >= 1e-5, first-occurrence scan) -- or, in the 3-objective variant, until
// threshold >= 1. Like the host restatement it replaces (pgmorl_b200/prediction.py) the loop also ends once the
// threshold has overflowed to +inf, where the reference would spin forever.
//
// One CTA per member. Because the threshold only ever doubles, the neighbourhoods are nested: each node gets the
// first step s at which it becomes near (its comparisons use the same products |objs_k| * (0.1 * 2^s) numpy forms),
// and the loop over steps jumps from one such event to the next instead of re-testing every node per doubling.
// The second kernel gathers (x, y) = (successor weight, objective gain) per objective into the layout K4 reads and
// forms the per-fit upper bounds; the Gaussian point weights need numpy's exp bit for bit and stay on the host.
#include <cmath>

#include "common.cuh"

namespace pgm {

constexpr int K4I_THREADS = 256;
constexpr int K4I_NEVER = 0xFFFF;

__device__ __forceinline__ int block_sum(int v, int *scratch) {
    // sum over the CTA, result in every thread (scratch: 8 ints + 1)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int w = 0; w < K4I_THREADS / 32; ++w) t += scratch[w];
    return t;
}

__device__ __forceinline__ int block_min(int v, int *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    int t = scratch[0];
#pragma unroll
    for (int w = 1; w < K4I_THREADS / 32; ++w) t = min(t, scratch[w]);
    return t;
}

// GS = false: the per-member scratch (successor weights and ids of the listed edges, first-near step of every node) lives
// in dynamic shared memory; GS = true (opt-graphs beyond ~200 KB of scratch): in a global workspace slice per CTA.
template <int M, bool GS>
__global__ void __launch_bounds__(K4I_THREADS) k4_neighbours_kernel(const double *__restrict__ objs, int n_nodes,
                                                                    const int32_t *__restrict__ parent, int E,
                                                                    const double *__restrict__ edge_w,
                                                                    const int32_t *__restrict__ node_ids, int cap_threshold,
                                                                    int s_inf, int s_cap,
                                                                    int32_t *__restrict__ klen, int32_t *__restrict__ steps,
                                                                    int32_t *__restrict__ edge_idx, unsigned char *gscratch,
                                                                    size_t scratch_per_cta) {
    extern __shared__ __align__(16) unsigned char k4i_smem[];
    unsigned char *member_scratch = GS ? gscratch + (size_t)blockIdx.x * scratch_per_cta : k4i_smem;
    double *wl = reinterpret_cast<double *>(member_scratch);                        // [E][M] successor weights of the listed edges
    int *e_list = reinterpret_cast<int *>(wl + (size_t)E * M);               // [E]
    unsigned short *s_node = reinterpret_cast<unsigned short *>(e_list + E); // [n_nodes]
    __shared__ int scratch[K4I_THREADS / 32 + 1];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x, k = node_ids[b];

    double ok[M], aok[M];
#pragma unroll
    for (int m = 0; m < M; ++m) { ok[m] = objs[(size_t)k * M + m]; aok[m] = fabs(ok[m]); }

    // s_inf: number of doublings after which 0.1 * 2^s is +inf; s_cap: first step with threshold >= 1 (3-objective stop)
    // ---- first step at which each node is inside the neighbourhood ----
    for (int i = tid; i < n_nodes; i += K4I_THREADS) {
        double rel[M];
#pragma unroll
        for (int m = 0; m < M; ++m) rel[m] = fabs(ok[m] - objs[(size_t)i * M + m]);
        double thr = 0.1;
        int s = 0;
        while (true) {
            bool c = true;
#pragma unroll
            for (int m = 0; m < M; ++m) c = c && (rel[m] < __dmul_rn(aok[m], thr));
            if (c) break;
            if (isinf(thr)) { s = K4I_NEVER; break; }
            thr *= 2.0;
            ++s;
        }
        s_node[i] = (unsigned short)s;
    }
    __syncthreads();

    int s = 0, K = 0;
    while (true) {
        // ---- ordered list of the edges whose source node is near at step s ----
        if (tid == 0) s_base = 0;
        __syncthreads();
        for (int e0 = 0; e0 < E; e0 += K4I_THREADS) {
            const int e = e0 + tid;
            const bool in = e < E && (int)s_node[parent[e]] <= s;
            const unsigned bal = __ballot_sync(0xffffffffu, in);
            if (lane == 0) scratch[warp] = __popc(bal);
            __syncthreads();
            int off = s_base;
            for (int w = 0; w < warp; ++w) off += scratch[w];
            if (in) e_list[off + __popc(bal & ((1u << lane) - 1u))] = e;
            __syncthreads();
            if (tid == 0) {
                int t = 0;
                for (int w = 0; w < K4I_THREADS / 32; ++w) t += scratch[w];
                s_base += t;
            }
            __syncthreads();
        }
        K = s_base;
        for (int i = tid; i < K * M; i += K4I_THREADS) wl[i] = edge_w[(size_t)e_list[i / M] * M + i % M];
        __syncthreads();
        // ---- more than 3 pairwise-distinct successor weights? (weight i counts when no earlier one is within 1e-5) ----
        int mine = 0;
        for (int i = tid; i < K; i += K4I_THREADS) {
            double wi[M];
#pragma unroll
            for (int m = 0; m < M; ++m) wi[m] = wl[i * M + m];
            bool distinct = true;
            for (int j = 0; j < i && distinct; ++j) {
                const double *wj = wl + j * M;
                double d = wi[0] - wj[0];
                double sq = __dmul_rn(d, d);
#pragma unroll
                for (int m = 1; m < M; ++m) { d = wi[m] - wj[m]; sq = fma(d, d, sq); }
                if (sqrt(sq) < 1e-5) distinct = false;
            }
            mine += distinct ? 1 : 0;
        }
        const int n_distinct = block_sum(mine, scratch);
        if (n_distinct > 3 || (cap_threshold && s >= s_cap) || s >= s_inf) break;      // threshold(s) >= 1  |  = +inf
        // ---- next step at which the neighbourhood grows (or one of the two stops applies) ----
        int nxt = K4I_NEVER;
        for (int i = tid; i < n_nodes; i += K4I_THREADS) {
            const int si = s_node[i];
            if (si > s && si < nxt) nxt = si;
        }
        nxt = block_min(nxt, scratch);
        nxt = min(nxt, s_inf);
        if (cap_threshold) nxt = min(nxt, s_cap);
        s = nxt;
        __syncthreads();
    }
    if (tid == 0) { klen[b] = K; steps[b] = s; }
    for (int r = tid; r < K; r += K4I_THREADS) edge_idx[(size_t)b * E + r] = e_list[r];
}

// one warp per fit f = member * M + objective: x, y rows of the K4 input pack, the upper bounds
// [clip(max y - min y, 1, 500), 20, 5, 500] (population_2d.py:100-104) and, once per member, the source node of every edge
__global__ void k4_gather_kernel(const int32_t *__restrict__ edge_idx, const int32_t *__restrict__ klen, int E, int M,
                                 const int32_t *__restrict__ parent, const double *__restrict__ edge_w,
                                 const double *__restrict__ edge_dy, int F, int Kmax, double *__restrict__ x,
                                 double *__restrict__ y, double *__restrict__ ub, int32_t *__restrict__ klen_f,
                                 int32_t *__restrict__ source) {
    const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (f >= F) return;
    const int b = f / M, dim = f - b * M, K = klen[b];
    double hi = -INFINITY, lo = INFINITY;
    for (int r = lane; r < K; r += 32) {
        const int e = edge_idx[(size_t)b * E + r];
        const double yv = edge_dy[(size_t)e * M + dim];
        x[(size_t)f * Kmax + r] = edge_w[(size_t)e * M + dim];
        y[(size_t)f * Kmax + r] = yv;
        hi = fmax(hi, yv); lo = fmin(lo, yv);
        if (dim == 0) source[(size_t)b * Kmax + r] = parent[e];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    }
    if (lane == 0) {
        const double span = K > 0 ? hi - lo : 1.0;
        ub[4 * f + 0] = fmin(fmax(span, 1.0), 500.0);
        ub[4 * f + 1] = 20.0; ub[4 * f + 2] = 5.0; ub[4 * f + 3] = 500.0;
        klen_f[f] = K;
    }
}

constexpr size_t K4I_SMEM_LIMIT = 200 * 1024;
static size_t k4i_scratch_bytes(int n_nodes, int M, int E) {
    const size_t b = (size_t)E * M * sizeof(double) + (size_t)E * sizeof(int) + (size_t)n_nodes * sizeof(unsigned short);
    return (b + 15) & ~(size_t)15;
}

}  // namespace pgm

using namespace pgm;

extern "C" size_t pgm_fit_neighbours_workspace_bytes(int n_nodes, int M, int E, int n) {
    const size_t per = k4i_scratch_bytes(n_nodes, M, E);
    return per <= K4I_SMEM_LIMIT ? 0 : per * (size_t)(n > 0 ? n : 0);
}

extern "C" int pgm_fit_neighbours_f64(const double *objs, int n_nodes, int M, const int32_t *parent, int E,
                                      const double *edge_w, const int32_t *node_ids, int n, int cap_threshold,
                                      int32_t *klen, int32_t *steps, int32_t *edge_idx, void *workspace,
                                      size_t workspace_bytes, void *stream) {
    PGM_REQUIRE(objs && node_ids && klen && steps, "pgm_fit_neighbours_f64: null pointer");
    PGM_REQUIRE(E == 0 || (parent && edge_w && edge_idx), "pgm_fit_neighbours_f64: null edge arrays with E=%d", E);
    PGM_REQUIRE(M >= 2 && M <= 4, "pgm_fit_neighbours_f64: M=%d objectives not in 2..4", M);
    PGM_REQUIRE(n >= 0 && n_nodes >= 1 && E >= 0, "pgm_fit_neighbours_f64: bad sizes n=%d nodes=%d E=%d", n, n_nodes, E);
    if (n == 0) return PGM_OK;
    const size_t per = k4i_scratch_bytes(n_nodes, M, E);
    const bool gs = per > K4I_SMEM_LIMIT || workspace != nullptr;      // a workspace, when given, is used
    if (gs) {
        PGM_REQUIRE(workspace && ((uintptr_t)workspace & 15) == 0, "pgm_fit_neighbours_f64: this opt-graph (%d nodes, %d edges) needs a 16-byte aligned workspace", n_nodes, E);
        if (workspace_bytes < per * (size_t)n) {
            set_error("pgm_fit_neighbours_f64: workspace too small: need %zu, got %zu", per * (size_t)n, workspace_bytes);
            return PGM_ERR_WORKSPACE;
        }
    }
    int s_inf = 0, s_cap = 0;
    for (double thr = 0.1; !std::isinf(thr); thr *= 2.0) ++s_inf;
    for (double thr = 0.1; !(thr >= 1.0); thr *= 2.0) ++s_cap;
    auto launch = [&](auto kern, size_t smem) -> int {
        if (smem) PGM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<n, K4I_THREADS, smem, (cudaStream_t)stream>>>(objs, n_nodes, parent, E, edge_w, node_ids, cap_threshold, s_inf, s_cap,
                                                              klen, steps, edge_idx, (unsigned char *)workspace, per);
        PGM_CUDA(cudaGetLastError());
        return PGM_OK;
    };
    if (gs) {
        if (M == 2) return launch(k4_neighbours_kernel<2, true>, 0);
        if (M == 3) return launch(k4_neighbours_kernel<3, true>, 0);
        return launch(k4_neighbours_kernel<4, true>, 0);
    }
    if (M == 2) return launch(k4_neighbours_kernel<2, false>, per);
    if (M == 3) return launch(k4_neighbours_kernel<3, false>, per);
    return launch(k4_neighbours_kernel<4, false>, per);
}

extern "C" int pgm_fit_gather_f64(const int32_t *edge_idx, const int32_t *klen, int n, int E, int M, const int32_t *parent,
                                  const double *edge_w, const double *edge_dy, int Kmax, double *x, double *y, double *ub,
                                  int32_t *klen_f, int32_t *source, void *stream) {
    PGM_REQUIRE(klen && x && y && ub && klen_f && source, "pgm_fit_gather_f64: null pointer");
    PGM_REQUIRE(E == 0 || (edge_idx && parent && edge_w && edge_dy), "pgm_fit_gather_f64: null edge arrays with E=%d", E);
    PGM_REQUIRE(n >= 0 && M >= 2 && M <= 4 && Kmax >= 1 && E >= 0, "pgm_fit_gather_f64: bad sizes n=%d M=%d Kmax=%d E=%d", n, M, Kmax, E);
    if (n == 0) return PGM_OK;
    const int F = n * M, warps = 4;
    k4_gather_kernel<<<(F + warps - 1) / warps, 32 * warps, 0, (cudaStream_t)stream>>>(edge_idx, klen, E, M, parent, edge_w,
                                                                                       edge_dy, F, Kmax, x, y, ub, klen_f, source);
    PGM_CUDA(cudaGetLastError());
    return PGM_OK;
}
