// Throughput of the exact-order (unfused) multiply-accumulate the correlation kernels need, scalar vs packed f32x2.
// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even under -fmad=false, so the packed unfused forms
// are written as fma.rn.f32x2 with a -0.0 addend (exact product) and a 1.0 multiplier (exact sum).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o f32x2_bench f32x2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { return ((u64)__float_as_uint(hi) << 32) | __float_as_uint(lo); }
// the constants must be opaque to ptxas (it folds fma(fma(a,b,-0),1,c) into fma(a,b,c) when it can see them)
__device__ __forceinline__ u64 mul2(u64 a, u64 b, u64 negzero) {
    u64 d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(negzero));
    return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b, u64 one) {
    u64 d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(one), "l"(b));
    return d;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// MODE 0: 16 FMUL + 16 FADD   1: 8 mul2 + 8 add2 (unfused, packed)   2: 8 fused FFMA2   3: 16 scalar FFMA (intrinsic)
template <int MODE>
__global__ void k(float *out, int iters, float seed, u64 negzero, u64 one) {
    float a[16], x = seed + threadIdx.x, y = seed * 0.5f;
    u64 p[8];
    float w[16];
    u64 w2[8];
    for (int i = 0; i < 16; i++) { a[i] = i; w[i] = seed * (i + 1); }
    for (int i = 0; i < 8; i++) { p[i] = i; w2[i] = pk(w[2 * i], w[2 * i + 1]); }
    u64 xx = pk(x, y), yy = pk(y, x);
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = a[i] + x * w[i];
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = add2(p[i], mul2(xx, w2[i], negzero), one);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = fma2(xx, w2[i], p[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = __fmaf_rn(x, w[i], a[i]);
        }
        x += 1e-9f;
        y += 1e-9f;
        xx += 1;
        yy += 3;
    }
    float s = 0;
    for (int i = 0; i < 16; i++) s += a[i];
    for (int i = 0; i < 8; i++) s += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *d;
    cudaMalloc(&d, 148 * 8 * 1024 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 100000;
    const u64 NZ = 0x8000000080000000ull, ONE = 0x3f8000003f800000ull;
    const char *names[4] = {"scalar FMUL+FADD", "packed unfused (2 FFMA2 per mul+add pair)", "packed fused FFMA2", "scalar FFMA"};
    double best_unfused = 0.0;
    for (int mode = 0; mode < 4; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters, 1.0f, NZ, ONE);
            else if (mode == 1) k<1><<<148 * 8, 256>>>(d, iters, 1.0f, NZ, ONE);
            else if (mode == 2) k<2><<<148 * 8, 256>>>(d, iters, 1.0f, NZ, ONE);
            else k<3><<<148 * 8, 256>>>(d, iters, 1.0f, NZ, ONE);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            double pairs = 148.0 * 8 * 256 * (double)iters * 16;       // multiply-accumulate lane operations
            printf("mode %d (%s): %.2f ms, %.2f T mul+acc pairs/s\n", mode, names[mode], ms, pairs / ms / 1e9);
            if (mode == 1 && 2.0 * pairs / ms / 1e9 > best_unfused) best_unfused = 2.0 * pairs / ms / 1e9;
        }
    }
    // the ceiling the correlation / low-pass kernels are measured against: one rounded multiply and one rounded add per pair
    printf("{\"unfused_tflops\": %.3f, \"how\": \"tools/microbench/f32x2_bench.cu mode 1: nothing but exact-product / exact-sum FFMA2 pairs, 148 x 8 CTAs of 256 threads, 8 independent chains per thread\"}\n",
           best_unfused);
    return 0;
}
