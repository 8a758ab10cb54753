// microbenchmark: the straight-line quad loop of k_phase_pair's chain warp, piece by piece, in a lone warp.
// variant bits: 1 = broadcast LDS.128 of the increments (else registers), 2 = lane-0 STS.128 of block-start phases,
// 4 = "p < 1" test with an out-of-line redo call (never taken here), 8 = a second idle-ish warp in the CTA
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float4 lds128(unsigned a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(unsigned a, float x, float y, float z, float w)
{
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __noinline__ float redo(float p, unsigned a)
{
    for (int i = 0; i < 8; ++i) { p = __fadd_rn(p, lds128(a + i * 16).x); if (p >= 1.0f) p -= 1.0f; }
    return p;
}
template <int V>
__global__ void k_quad(float* out, long long* cyc, int tiles)
{
    __shared__ __align__(16) float sF[2048];
    __shared__ __align__(16) float sP[64];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sF[i] = 1e-9f * (1 + (i & 3));
    __syncthreads();
    if (threadIdx.x >= 32) { if (V & 8) { float x = 0; for (int i = 0; i < tiles * 64; ++i) x = __fadd_rn(x, 1.0f); out[1] = x; } return; }
    const unsigned sF_a = (unsigned)__cvta_generic_to_shared(sF), sP_a = (unsigned)__cvta_generic_to_shared(sP);
    const bool lane0 = threadIdx.x == 0;
    float phase = 0.f;
    long long t0 = clock64();
    for (int tile = 0; tile < tiles; ++tile) {
        const unsigned fa_ = sF_a + (tile & 7) * 1024, pa_ = sP_a + (tile & 1) * 128;
        float4 fv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) fv[i] = (V & 1) ? lds128(fa_ + i * 16) : make_float4(1e-9f, 2e-9f, 3e-9f, 4e-9f);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float4 fn[8];
            if (q < 7) {
#pragma unroll
                for (int i = 0; i < 8; ++i) fn[i] = (V & 1) ? lds128(fa_ + (q + 1) * 128 + i * 16) : fv[i];
            }
            float p = phase, ps[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if ((i & 1) == 0) ps[i >> 1] = p;
                p = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(p, fv[i].x), fv[i].y), fv[i].z), fv[i].w);
            }
            if ((V & 2) && lane0) sts128(pa_ + q * 16, ps[0], ps[1], ps[2], ps[3]);
            if (V & 4) {
                if (__builtin_expect(p < 1.0f, 1)) phase = p; else phase = redo(ps[0], fa_ + q * 128);
            } else phase = p;
            if (q < 7) {
#pragma unroll
                for (int i = 0; i < 8; ++i) fv[i] = fn[i];
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = phase; cyc[0] = t1 - t0; }
}
template <int V> void run(float* out, long long* cyc)
{
    const int tiles = 4000;
    for (int rep = 0; rep < 2; ++rep) { k_quad<V><<<1, (V & 8) ? 64 : 32>>>(out, cyc, tiles); cudaDeviceSynchronize(); }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("variant %2d (lds %d sts %d check %d warp2 %d): %.2f cycles/sample, %.1f per quad\n", V, V & 1, (V >> 1) & 1, (V >> 2) & 1, (V >> 3) & 1,
           (double)c / tiles / 256, (double)c / tiles / 8);
}
int main()
{
    float* out; long long* cyc;
    cudaMalloc(&out, 1024); cudaMalloc(&cyc, 1024);
    run<0>(out, cyc); run<1>(out, cyc); run<2>(out, cyc); run<3>(out, cyc); run<4>(out, cyc); run<5>(out, cyc); run<7>(out, cyc); run<15>(out, cyc);
    return 0;
}
