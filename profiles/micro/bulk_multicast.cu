// Functional + timing probe of the non-tensor TMA bulk copy with cluster multicast (the parameter broadcast of the K3 FFMA
// cluster kernel, csrc/k3_fast.cuh): every CTA of a 16-CTA cluster publishes ONE slice from global memory to the same
// shared-memory offset of 8 CTAs (its "half"), each destination's mbarrier collects complete_tx bytes from 8 sources.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_multicast bulk_multicast.cu && ./bulk_multicast
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int C = 16, G = 8, SL = 3104;            // slice bytes (194 float4)
template <int MODE, bool BAR>
__global__ void __launch_bounds__(256, 1) k(const float *g, float *out, long long *cyc, int iters) {
    extern __shared__ __align__(16) float buf[];    // G * SL bytes
    __shared__ __align__(8) uint64_t mbar;
    const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&mbar), d = (uint32_t)__cvta_generic_to_shared(buf);
    unsigned rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int half = rank / G, gg = rank % G;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const float *src = g + ((size_t)(blockIdx.x / C) * C + rank) * (SL / 4);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(G * SL) : "memory");
            asm volatile("fence.proxy.async.global;" ::: "memory");
            const uint16_t mask = (uint16_t)(0xFFu << (half * G));
            if (MODE == 0)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                             ::"r"(d + gg * SL), "l"(src), "r"(SL), "r"(mb), "h"(mask) : "memory");
            else          // unicast: every CTA fetches the whole image of its half itself
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(d), "l"(g + ((size_t)(blockIdx.x / C) * C + half * G) * (SL / 4)), "r"(G * SL), "r"(mb) : "memory");
        }
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(mb), "r"(it & 1) : "memory");
        // everyone must have consumed the data before the next round overwrites it (timing variant without: benign race)
        if (BAR) asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = (t1 - t0) / iters;
    for (int i = threadIdx.x; i < G * SL / 4; i += blockDim.x) out[(size_t)blockIdx.x * (G * SL / 4) + i] = buf[i];
}

template <int MODE, bool BAR>
int run(const char *what) {
    const int clusters = 6, n = clusters * C, iters = 200;
    float *g, *out; long long *cyc;
    cudaMallocManaged(&g, (size_t)n * SL); cudaMallocManaged(&out, (size_t)n * G * SL); cudaMallocManaged(&cyc, n * 8);
    for (int i = 0; i < n * SL / 4; ++i) g[i] = (float)i;
    auto kern = k<MODE, BAR>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G * SL);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = G * SL;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {C, 1, 1};
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, (const float *)g, out, cyc, iters);
    e = e ? e : cudaDeviceSynchronize();
    if (e) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    long bad = 0;
    for (int b = 0; b < n; ++b) {
        const int cl = b / C, rank = b % C, half = rank / G;
        for (int s = 0; s < G; ++s)
            for (int i = 0; i < SL / 4; ++i) {
                const float want = (float)(((cl * C + half * G + s) * (SL / 4)) + i);
                if (out[(size_t)b * (G * SL / 4) + s * (SL / 4) + i] != want) ++bad;
            }
    }
    printf("%-44s: %ld mismatches; cycles per round (8 x %d B into every CTA): CTA0 %lld, CTA15 %lld\n", what, bad, SL, cyc[0], cyc[15]);
    cudaFree(g); cudaFree(out); cudaFree(cyc);
    return bad != 0;
}

int main() {
    int rc = 0;
    rc |= run<0, true>("multicast slices + cluster barrier");
    rc |= run<0, false>("multicast slices, no barrier");
    rc |= run<1, true>("unicast whole image + cluster barrier");
    rc |= run<1, false>("unicast whole image, no barrier");
    return rc;
}
