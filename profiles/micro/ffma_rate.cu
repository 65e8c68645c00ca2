// Microbenchmark: issue rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on one SM partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_rate ffma_rate.cu && ./ffma_rate
// Prints FMA lanes retired per clock per SM for 1..16 warps per SM (grid = #SMs).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, float seed) {
    float a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    float b0 = seed * 0.5f, b1 = b0 + 1, b2 = b0 + 2, b3 = b0 + 3;
    float acc[16];
    u64 acc2[16];
    for (int i = 0; i < 16; ++i) { acc[i] = i; acc2[i] = (u64)i; }
    u64 A0 = ((u64)__float_as_uint(a0) << 32) | __float_as_uint(a1), A1 = ((u64)__float_as_uint(a2) << 32) | __float_as_uint(a3);
    u64 B0 = ((u64)__float_as_uint(b0) << 32) | __float_as_uint(b1), B1 = ((u64)__float_as_uint(b2) << 32) | __float_as_uint(b3);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {   // 16 independent scalar FFMA, outer-product style operand reuse
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                acc[0] = fmaf(a0, b0, acc[0]); acc[1] = fmaf(a0, b1, acc[1]); acc[2] = fmaf(a0, b2, acc[2]); acc[3] = fmaf(a0, b3, acc[3]);
                acc[4] = fmaf(a1, b0, acc[4]); acc[5] = fmaf(a1, b1, acc[5]); acc[6] = fmaf(a1, b2, acc[6]); acc[7] = fmaf(a1, b3, acc[7]);
                acc[8] = fmaf(a2, b0, acc[8]); acc[9] = fmaf(a2, b1, acc[9]); acc[10] = fmaf(a2, b2, acc[10]); acc[11] = fmaf(a2, b3, acc[11]);
                acc[12] = fmaf(a3, b0, acc[12]); acc[13] = fmaf(a3, b1, acc[13]); acc[14] = fmaf(a3, b2, acc[14]); acc[15] = fmaf(a3, b3, acc[15]);
            }
        } else if (MODE == 1) {   // 16 independent packed FFMA2 (vector x vector)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int i = 0; i < 16; ++i) acc2[i] = ffma2((i & 1) ? A1 : A0, (i & 2) ? B1 : B0, acc2[i]);
            }
        } else {   // packed FFMA2 with a scalar-broadcast operand
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    u64 d; float s = (i & 1) ? a1 : a0;
                    asm volatile("{ .reg .b64 t; mov.b64 t, {%1, %1}; fma.rn.f32x2 %0, t, %2, %3; }" : "=l"(d) : "f"(s), "l"((i & 2) ? B1 : B0), "l"(acc2[i]));
                    acc2[i] = d;
                }
            }
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 16; ++i) s += acc[i] + (float)(acc2[i] & 0xffff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int lanes_per_inst) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    int iters = 4096;
    printf("%s:", name);
    for (int warps = 1; warps <= 16; warps *= 2) {
        k<MODE><<<148, warps * 32>>>(out, cyc, iters, 1.0f); cudaDeviceSynchronize();
        k<MODE><<<148, warps * 32>>>(out, cyc, iters, 1.0f); cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
        double fma_per_clk = (double)iters * 64 * lanes_per_inst * warps * 32 / c;
        printf("  %2dw: %6.1f FMA/clk/SM", warps, fma_per_clk);
    }
    printf("\n");
}
int main() {
    run<0>("scalar FFMA        ", 1);
    run<1>("FFMA2 vec x vec    ", 2);
    run<2>("FFMA2 scalar x vec ", 2);
    return 0;
}
