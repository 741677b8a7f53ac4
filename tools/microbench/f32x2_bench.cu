// Throughput of scalar FMUL+FADD vs packed mul/add.f32x2 (unfused), to decide whether the exact-order correlation kernels
// should use the packed forms.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o f32x2_bench f32x2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
template <int MODE>
__global__ void k(float *out, int iters, float seed) {
    float a[8], x = seed + threadIdx.x, y = seed * 0.5f;
    unsigned long long p[4];
    for (int i = 0; i < 8; i++) a[i] = i;
    for (int i = 0; i < 4; i++) p[i] = i;
    unsigned long long xx = ((unsigned long long)__float_as_uint(x) << 32) | __float_as_uint(y);
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = a[i] + x * y;      // 8 FMUL + 8 FADD (fmad=false)
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) p[i] = add2(p[i], mul2(xx, xx));   // 4 MUL2 + 4 ADD2 = same flops
        }
        x += 1e-9f;
        xx += 1;
    }
    float s = 0;
    for (int i = 0; i < 8; i++) s += a[i];
    for (int i = 0; i < 4; i++) s += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *d;
    cudaMalloc(&d, 148 * 8 * 1024 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 200000;
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters, 1.0f);
            else k<1><<<148 * 8, 256>>>(d, iters, 1.0f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            double flops = 148.0 * 8 * 256 * (double)iters * 16;
            printf("mode %d (%s): %.2f ms, %.2f TFLOP/s (mul and add counted separately)\n", mode, mode ? "f32x2" : "scalar", ms, flops / ms / 1e9);
        }
    }
    return 0;
}
