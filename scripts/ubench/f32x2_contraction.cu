typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pk(float lo, float hi){ f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c){ f2_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2_t add2(f2_t a, f2_t b){ f2_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2_t mul2(f2_t a, f2_t b){ f2_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__global__ void ka(const f2_t* x, f2_t* y) { f2_t a=x[threadIdx.x], b=x[threadIdx.x+32], c=x[threadIdx.x+64], d=x[threadIdx.x+96]; y[threadIdx.x] = add2(mul2(a,b), mul2(c,d)); }
__global__ void kb(const f2_t* x, f2_t* y) { f2_t a=x[threadIdx.x], b=x[threadIdx.x+32], c=x[threadIdx.x+64], d=x[threadIdx.x+96]; const f2_t nz = pk(-0.0f,-0.0f); y[threadIdx.x] = add2(fma2(a,b,nz), fma2(c,d,nz)); }
__global__ void kc(const float* x, float* y) { float a=x[threadIdx.x], b=x[threadIdx.x+32], c=x[threadIdx.x+64], d=x[threadIdx.x+96]; y[threadIdx.x] = __fadd_rn(__fmul_rn(a,b), __fmul_rn(c,d)); }
__device__ __forceinline__ void unpk(f2_t x, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x)); }
__global__ void kd(const f2_t* x, f2_t* y) { f2_t a=x[threadIdx.x], b=x[threadIdx.x+32], c=x[threadIdx.x+64], d=x[threadIdx.x+96]; f2_t p=mul2(a,b), q=mul2(c,d); float p0,p1,q0,q1; unpk(p,p0,p1); unpk(q,q0,q1); y[threadIdx.x] = pk(__fadd_rn(p0,q0), __fadd_rn(p1,q1)); }
