// microbenchmark: latency of a dependent f32 add chain in a lone warp, with and without broadcast LDS feeding it
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_chain(float* out, long long* cyc, int iters, int mode)
{
    __shared__ __align__(16) float sF[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sF[i] = 1e-4f * (1 + (i & 3));
    __syncthreads();
    float p = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (mode == 0) {
#pragma unroll
            for (int k = 0; k < 32; ++k) p = __fadd_rn(p, 1e-7f);
        } else if (mode == 1) {
            const float* f = sF + (it & 7) * 32;
            float4 fv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) fv[i] = *reinterpret_cast<const float4*>(f + i * 4);
#pragma unroll
            for (int i = 0; i < 8; ++i) p = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(p, fv[i].x), fv[i].y), fv[i].z), fv[i].w);
        } else if (mode == 2) {   // loads for the NEXT iteration issued before the chain (explicit double buffer)
            static_assert(true, "");
            const float* f = sF + (it & 7) * 32;
            float4 fv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) fv[i] = *reinterpret_cast<const float4*>(f + i * 4);
#pragma unroll
            for (int i = 0; i < 8; ++i) p = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(p, fv[i].x), fv[i].y), fv[i].z), fv[i].w);
            if (p >= 1.0f) p = p - 1.0f;
        } else {   // integer add chain for comparison
            unsigned q = __float_as_uint(p);
#pragma unroll
            for (int k = 0; k < 32; ++k) q = q * 3u + 1u;
            p = __uint_as_float(q & 0x3fffffff);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x] = p; cyc[blockIdx.x] = t1 - t0; }
}
int main()
{
    float* out; long long* cyc;
    cudaMalloc(&out, 1024); cudaMalloc(&cyc, 1024);
    for (int mode = 0; mode < 4; ++mode)
        for (int warps = 1; warps <= 2; ++warps) {
            k_chain<<<1, 32 * warps>>>(out, cyc, 20000, mode);
            cudaDeviceSynchronize();
            k_chain<<<1, 32 * warps>>>(out, cyc, 20000, mode);
            cudaDeviceSynchronize();
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("mode %d warps %d: %.2f cycles per add (32 adds/iter)\n", mode, warps, (double)c / 20000 / 32);
        }
    return 0;
}
