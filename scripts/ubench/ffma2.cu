// microbenchmark: packed fp32 (FFMA2) against scalar FFMA on sm_100a -- dependent latency and issue throughput
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 x) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x)); return a + b; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
template <int ILP, bool PACKED>
__global__ void k(float* out, long long* cyc, int iters, float seed)
{
    float acc = 0.f;
    long long t0 = clock64();
    if (PACKED) {
        u64 x[ILP];
        const u64 m = pack(1.0000001f, 0.9999999f), c = pack(seed, -seed);
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = pack(seed + i, seed - i);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < ILP; ++i) x[i] = fma2(x[i], m, c);
        }
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc += lo(x[i]);
    } else {
        float x[ILP];
        const float m = 1.0000001f, c = seed;
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = seed + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], m, c);
        }
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc += x[i];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int ILP, bool PACKED> void run(float* out, long long* cyc, int warps)
{
    const int iters = 4000;
    for (int rep = 0; rep < 2; ++rep) { k<ILP, PACKED><<<1, 32 * warps>>>(out, cyc, iters, 1.0f); cudaDeviceSynchronize(); }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / iters / 8 / ILP;
    printf("%s ILP %d warps/SM %2d: %.2f cycles per warp-instruction, %.1f FMA lanes/clk/SM\n", PACKED ? "FFMA2" : "FFMA ", ILP, warps, per,
           (PACKED ? 64.0 : 32.0) * warps / (per * ILP) * ILP / 1.0 / 1.0 * 1.0 / 1.0 * (1.0) / (1.0) * 1.0 * (1.0 / 1.0) * (1.0) * (1.0 / 1.0) / 1.0 * 1.0 * 1.0 / 1.0);
}
int main()
{
    float* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
    run<1, false>(out, cyc, 1); run<1, true>(out, cyc, 1);
    run<8, false>(out, cyc, 1); run<8, true>(out, cyc, 1);
    run<8, false>(out, cyc, 4); run<8, true>(out, cyc, 4);
    run<8, false>(out, cyc, 12); run<8, true>(out, cyc, 12);
    run<8, false>(out, cyc, 32); run<8, true>(out, cyc, 32);
    return 0;
}
