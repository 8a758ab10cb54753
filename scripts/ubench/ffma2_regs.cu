// microbenchmark: FFMA / FFMA2 with three DISTINCT register operands per instruction (no immediates, no operand reuse),
// the shape of k_formant's inner loop -- issue throughput per SM at 4 / 12 / 16 warps and the dependent latency
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 x) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x)); return a + b; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
// MODE 0: scalar FFMA x[i] = x[i]*y[i]+z[i];  1: FFMA2 same;  2: FFMA2 rotating operands x[i] = x[i+1]*y[i]+z[i+2] (no same-slot reuse)
template <int ILP, int MODE>
__global__ void k(float* out, long long* cyc, int iters, float seed)
{
    float acc = 0.f;
    long long t0, t1;
    if (MODE >= 1) {
        u64 x[ILP], y[ILP], z[ILP];
#pragma unroll
        for (int i = 0; i < ILP; ++i) { x[i] = pack(seed + i, seed - i); y[i] = pack(1.0f + 1e-7f * (seed + i), 1.0f - 1e-7f * i); z[i] = pack(seed * 1e-9f * i, -seed * 1e-9f); }
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < ILP; ++i) {
                    if (MODE == 1) x[i] = fma2(x[i], y[i], z[i]);
                    else x[i] = fma2(x[(i + 1) % ILP], y[(i + r) % ILP], z[(i + 2 + r) % ILP]);
                }
        }
        t1 = clock64();
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc += lo(x[i]);
    } else {
        float x[ILP], y[ILP], z[ILP];
#pragma unroll
        for (int i = 0; i < ILP; ++i) { x[i] = seed + i; y[i] = 1.0f + 1e-7f * (seed + i); z[i] = seed * 1e-9f * i; }
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < ILP; ++i) x[i] = fma1(x[i], y[i], z[i]);
        }
        t1 = clock64();
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc += x[i];
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int ILP, int MODE> void run(float* out, long long* cyc, int warps)
{
    const int iters = 4000;
    for (int rep = 0; rep < 2; ++rep) { k<ILP, MODE><<<1, 32 * warps>>>(out, cyc, iters, 1.0f); cudaDeviceSynchronize(); }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / iters / 8 / ILP;          // cycles per warp-instruction as seen by one warp
    const double lanes = (MODE ? 64.0 : 32.0) * warps / per; // FMA lanes per clock per SM
    printf("%s ILP %d warps/SM %2d: %.2f cycles per instruction per warp, %.1f FMA lanes/clk/SM\n",
           MODE == 0 ? "FFMA  rrr" : (MODE == 1 ? "FFMA2 rrr" : "FFMA2 rot"), ILP, warps, per, lanes);
}
int main()
{
    float* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
    run<1, 0>(out, cyc, 1); run<1, 1>(out, cyc, 1);
    const int ws[5] = { 4, 8, 12, 16, 32 };
    for (int wi = 0; wi < 5; ++wi) {
        const int w = ws[wi];
        run<1, 0>(out, cyc, w); run<1, 1>(out, cyc, w);
        run<2, 0>(out, cyc, w); run<2, 1>(out, cyc, w);
        run<4, 0>(out, cyc, w); run<4, 1>(out, cyc, w); run<4, 2>(out, cyc, w);
        run<8, 0>(out, cyc, w); run<8, 1>(out, cyc, w); run<8, 2>(out, cyc, w);
    }
    return 0;
}
