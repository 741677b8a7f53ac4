// Experiment: what does ONE long-lived background warp cost a K4-shaped foreground kernel that shares its SM, by the kind of
// instructions the background warp issues?  (The Fano worker warps slow the K4 warps of their SM by 25-50 %, far more than
// their share of the issue slots -- profiles/r2_interference_warp_times.txt; this isolates the ingredient.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o interference interference.cu && ./interference
// Foreground: 608 threads, 47 KB of shared memory, two CTAs per SM; per step one 8-byte shared load per lane, two broadcast
// 16-byte loads and 16 packed fma.rn.f32x2 in four dependent chains (the instruction mix of k_sync_lags).
// Background: one-warp CTAs holding 31 KB of shared memory, one (or two) per SM, running for the whole measurement:
//   0 imad   dependent integer multiply-add chain (FMA pipe, ~0.25 instructions per clock)
//   1 alu    eight independent chains of LOP3/IADD3/SHF (ALU pipe, as many instructions per clock as the warp can issue)
//   2 alu1   ONE dependent chain of LOP3/IADD3/SHF (ALU pipe, ~0.25 per clock)
//   3 sel    predicate/select heavy integer code with 4 chains (the shape of the Fano loop)
//   4 lds    alu1 + a dependent shared-memory load/store pair every 16 instructions
//   5 sleep  __nanosleep only (presence without issue)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CKR(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

typedef unsigned long long pk2;
__device__ __forceinline__ pk2 pk_fma(pk2 a, pk2 b, pk2 c) {
    pk2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

__global__ void __launch_bounds__(608, 2) k_fg(float *out, int iters, pk2 one) {
    __shared__ float2 win[8 * 609];
    __shared__ ulonglong2 tab[512];
    for (int i = threadIdx.x; i < 8 * 609; i += blockDim.x) win[i] = make_float2(1e-3f * i, 2e-3f * i);
    for (int i = threadIdx.x; i < 512; i += blockDim.x) tab[i] = make_ulonglong2(one, one + i);
    __syncthreads();
    pk2 a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    const float2 *wp = win + (threadIdx.x % 600);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const float2 v = wp[r * 609 / 8];
            pk2 x, y;
            asm("mov.b64 %0, {%1, %1};" : "=l"(x) : "f"(v.x));
            asm("mov.b64 %0, {%1, %1};" : "=l"(y) : "f"(v.y));
            const ulonglong2 w01 = tab[(it * 8 + r) & 255], w23 = tab[256 + ((it * 8 + r) & 255)];
            a0 = pk_fma(pk_fma(x, w01.x, a0), one, pk_fma(y, w01.y, a0));
            a1 = pk_fma(pk_fma(x, w01.y, a1), one, pk_fma(y, w01.x, a1));
            a2 = pk_fma(pk_fma(x, w23.x, a2), one, pk_fma(y, w23.y, a2));
            a3 = pk_fma(pk_fma(x, w23.y, a3), one, pk_fma(y, w23.x, a3));
            a0 = pk_fma(a0, one, a1);
            a1 = pk_fma(a1, one, a2);
            a2 = pk_fma(a2, one, a3);
            a3 = pk_fma(a3, one, a0);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(a0 ^ a1 ^ a2 ^ a3));
}

__global__ void __launch_bounds__(128) k_bg(int mode, long long clocks, unsigned *sink) {
    extern __shared__ unsigned bg_all[];
    unsigned *bg = bg_all + (threadIdx.x >> 5) * 4096;
    const long long t0 = clock64();
    unsigned x = threadIdx.x, y = x * 3u + 1u, z = x ^ 0x55u, w = x + 7u, p = x * 5u, q = x + 11u, r = x ^ 0x33u, s = x + 19u;
    bg[threadIdx.x & 31] = x;
    while (clock64() - t0 < clocks) {
        if (mode == 0) {
#pragma unroll
            for (int k = 0; k < 64; k++) x = x * 1664525u + 1013904223u;
        } else if (mode == 1) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                x = (x ^ (x >> 3)) + 0x9e3779b9u; y = (y ^ (y << 5)) + 0x7f4a7c15u; z = (z & 0xfffffff0u) ^ (z >> 7); w = (w | 1u) + (w >> 2);
                p = (p ^ (p >> 11)) + 3u; q = (q ^ (q << 7)) + 5u; r = (r & 0x0fffffffu) ^ (r >> 9); s = (s | 2u) + (s >> 4);
            }
        } else if (mode == 2) {
#pragma unroll
            for (int k = 0; k < 64; k++) x = (x ^ (x >> 3)) + 0x9e3779b9u;
        } else if (mode == 3) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const bool a = (int)(x + y) >= (int)z, b = !a && (w < p), c = !a && !b;
                x = a ? x + 1u : (c ? x - 1u : x);
                y = a ? (y ^ (z >> 9)) : (c ? q : y);
                z = a ? z + 60u : z - ((b || c) ? 60u : 0u);
                w = a ? p : (c ? r : w);
                p = a ? s : (c ? (p >> 1) : p);
                q = (q << 1) | (a ? 1u : 0u);
                r = b ? (r ^ 1u) : r + 3u;
                s = c ? s + x : s ^ y;
            }
        } else if (mode == 4) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
#pragma unroll
                for (int j = 0; j < 16; j++) x = (x ^ (x >> 3)) + 0x9e3779b9u;
                bg[((threadIdx.x & 31) + (x & 31u) * 32u) & 4095u] = x;
                x ^= bg[((threadIdx.x & 31) + ((x >> 5) & 31u) * 32u) & 4095u];
            }
        } else if (mode == 6) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                x = x * 1664525u + 1013904223u; y = y * 22695477u + 1u; z = z * 69069u + 5u; w = w * 1103515245u + 12345u;
                p = p * 134775813u + 1u; q = q * 214013u + 2531011u; r = r * 16807u + 7u; s = s * 48271u + 11u;
            }
        } else {
            __nanosleep(2000);
        }
    }
    if ((x ^ y ^ z ^ w ^ p ^ q ^ r ^ s) == 0xdeadbeefu) sink[0] = x + bg[0];
}

static float time_fg(cudaStream_t st, float *out, int ctas) {
    cudaEvent_t e0, e1;
    CKR(cudaEventCreate(&e0));
    CKR(cudaEventCreate(&e1));
    CKR(cudaEventRecord(e0, st));
    for (int r = 0; r < 3; r++) k_fg<<<ctas, 608, 0, st>>>(out, 4000, 0x3f8000003f800000ull);
    CKR(cudaEventRecord(e1, st));
    CKR(cudaEventSynchronize(e1));
    float ms = 0;
    CKR(cudaEventElapsedTime(&ms, e0, e1));
    return ms / 3;
}

int main() {
    CKR(cudaSetDevice(0));
    cudaDeviceProp prop;
    CKR(cudaGetDeviceProperties(&prop, 0));
    const int nsm = prop.multiProcessorCount;
    float *d_out;
    unsigned *d_sink;
    CKR(cudaMalloc(&d_out, (size_t)16 * nsm * 608 * sizeof(float)));
    CKR(cudaMalloc(&d_sink, 64));
    cudaStream_t s_fg, s_bg;
    CKR(cudaStreamCreateWithFlags(&s_fg, cudaStreamNonBlocking));
    CKR(cudaStreamCreateWithFlags(&s_bg, cudaStreamNonBlocking));
    // one common carve-out, as the library sets it
    CKR(cudaFuncSetAttribute(k_fg, cudaFuncAttributePreferredSharedMemoryCarveout, 72));
    CKR(cudaFuncSetAttribute(k_bg, cudaFuncAttributePreferredSharedMemoryCarveout, 72));
    time_fg(s_fg, d_out, 16 * nsm);
    const float base = time_fg(s_fg, d_out, 16 * nsm);
    printf("foreground alone: %.3f ms\n", base);
    const char *names[] = {"imad chain (fma pipe)", "alu x8 chains", "alu x1 chain", "select-heavy x4", "alu x1 + lds/sts", "sleep", "imad x8 chains"};
    CKR(cudaFuncSetAttribute(k_bg, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16 * 1024));
    // cta_warps: background warps per CTA (one CTA per SM; the warps of a CTA sit on different schedulers); nctas: CTAs launched
    const int shapes[][2] = {{1, 1}, {1, 2}, {2, 1}, {4, 1}, {1, 0}};   // {warps per CTA, CTAs per SM}; {1,0}: on every other SM only
    for (auto &sh : shapes)
        for (int mode : {1, 3, 6, 0, 2}) {
            const int cta_warps = sh[0], nctas = sh[1] ? nsm * sh[1] : nsm / 2;
            k_bg<<<nctas, 32 * cta_warps, cta_warps * 16 * 1024, s_bg>>>(mode, 400000000LL, d_sink);   // ~0.2 s: outlasts the measurement
            // give the background a moment to spread over the (empty) SMs before the foreground arrives
            cudaEvent_t go;
            CKR(cudaEventCreate(&go));
            k_bg<<<1, 32, 1024, s_fg>>>(5, 2000000LL, d_sink);
            const float ms = time_fg(s_fg, d_out, 16 * nsm);
            CKR(cudaDeviceSynchronize());
            printf("background %d CTAs of %d warps, %-24s: foreground %.3f ms (x%.3f)\n", nctas, cta_warps, names[mode], ms, ms / base);
            CKR(cudaEventDestroy(go));
        }
    return 0;
}
