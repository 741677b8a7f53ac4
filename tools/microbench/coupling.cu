// Experiment: is the slow-down that ONE background warp causes an FFMA2-bound foreground kernel (interference.cu: x1.2-1.5
// whether 1, 2 or 4 such warps share the SM) a property of the scheduler the warp sits on, amplified by the foreground's CTA
// shape?  A CTA's shared memory and slots are released when its LAST warp ends, and its warps are dealt evenly to the four
// schedulers: if one scheduler is slow, the other three wait for it.  Three foreground shapes, same total work:
//   static19   19-warp CTAs, one task (32 cells x 256 steps) per warp            -- the shape of k_sync_lags
//   single     1-warp CTAs, one task each                                        -- the hardware rebalances
//   dynamic8   8-warp CTAs, 19 tasks per CTA claimed through a shared counter    -- the CTA rebalances
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o coupling coupling.cu && ./coupling
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

constexpr int TASK_STEPS = 256;
constexpr int TAB = 128;

__device__ __forceinline__ unsigned run_task(const float2 *wp, const ulonglong2 *tab, pk2 one, int pitch) {
    pk2 a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int it = 0; it < TASK_STEPS / 8; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const float2 v = wp[r * pitch + it];
            pk2 x, y;
            asm("mov.b64 %0, {%1, %1};" : "=l"(x) : "f"(v.x));
            asm("mov.b64 %0, {%1, %1};" : "=l"(y) : "f"(v.y));
            const ulonglong2 w01 = tab[(it * 8 + r) & (TAB / 2 - 1)], w23 = tab[TAB / 2 + ((it * 8 + r) & (TAB / 2 - 1))];
            a0 = pk_fma(pk_fma(pk_fma(x, w01.x, one), one, a0), one, pk_fma(y, w01.y, one));
            a1 = pk_fma(pk_fma(pk_fma(x, w01.y, one), one, a1), one, pk_fma(y, w01.x, one));
            a2 = pk_fma(pk_fma(pk_fma(x, w23.x, one), one, a2), one, pk_fma(y, w23.y, one));
            a3 = pk_fma(pk_fma(pk_fma(x, w23.y, one), one, a3), one, pk_fma(y, w23.x, one));
        }
    }
    return (unsigned)(a0 ^ a1 ^ a2 ^ a3);
}

// tasks_per_cta tasks; the CTA's warps take them statically (dynamic == 0: task = warp id, needs warps == tasks) or from a counter
__global__ void __launch_bounds__(608, 2) k_fg(float *out, int tasks_per_cta, int dynamic, pk2 one, int reps) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int next_task;
    const int pitch = tasks_per_cta * 32 + TASK_STEPS / 8 + 1;
    ulonglong2 *tab = reinterpret_cast<ulonglong2 *>(smem);
    float2 *win = reinterpret_cast<float2 *>(smem + TAB * sizeof(ulonglong2));
    for (int i = threadIdx.x; i < 8 * pitch; i += blockDim.x) win[i] = make_float2(1e-3f * i, 2e-3f * i);
    for (int i = threadIdx.x; i < TAB; i += blockDim.x) tab[i] = make_ulonglong2(one, one + i);
    if (threadIdx.x == 0) next_task = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned acc = 0;
    if (!dynamic) {
        for (int rep = 0; rep < reps; rep++) acc ^= run_task(win + warp * 32 + lane, tab, one, pitch);
    } else {
        const int total = tasks_per_cta * reps;
        for (;;) {
            int task = 0;
            if (lane == 0) task = atomicAdd(&next_task, 1);
            task = __shfl_sync(0xffffffffu, task, 0);
            if (task >= total) break;
            acc ^= run_task(win + (task % tasks_per_cta) * 32 + lane, tab, one, pitch);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
}

__global__ void __launch_bounds__(128) k_bg(int mode, long long clocks, unsigned *sink) {
    const long long t0 = clock64();
    unsigned x = threadIdx.x, y = x * 3u + 1u, z = x ^ 0x55u, w = x + 7u, p = x * 5u, q = x + 11u, r = x ^ 0x33u, s = x + 19u;
    while (clock64() - t0 < clocks) {
        if (mode == 1) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                x = (x ^ (x >> 3)) + 0x9e3779b9u; y = (y ^ (y << 5)) + 0x7f4a7c15u; z = (z & 0xfffffff0u) ^ (z >> 7); w = (w | 1u) + (w >> 2);
                p = (p ^ (p >> 11)) + 3u; q = (q ^ (q << 7)) + 5u; r = (r & 0x0fffffffu) ^ (r >> 9); s = (s | 2u) + (s >> 4);
            }
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
        } else {
            __nanosleep(2000);
        }
    }
    if ((x ^ y ^ z ^ w ^ p ^ q ^ r ^ s) == 0xdeadbeefu) sink[0] = x;
}

struct Shape {
    const char *name;
    int warps, tasks, dynamic, reps, ctas_per_sm_total;
};

static float time_fg(cudaStream_t st, float *out, const Shape &sh, int nsm) {
    const int pitch = sh.tasks * 32 + TASK_STEPS / 8 + 1;
    const size_t smem = TAB * sizeof(ulonglong2) + (size_t)8 * pitch * sizeof(float2);
    cudaEvent_t e0, e1;
    CKR(cudaEventCreate(&e0));
    CKR(cudaEventCreate(&e1));
    CKR(cudaEventRecord(e0, st));
    for (int r = 0; r < 3; r++)
        k_fg<<<sh.ctas_per_sm_total * nsm, 32 * sh.warps, smem, st>>>(out, sh.tasks, sh.dynamic, 0x3f8000003f800000ull, sh.reps);
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
    CKR(cudaMalloc(&d_out, (size_t)64 * 19 * nsm * 608 * sizeof(float)));
    CKR(cudaMalloc(&d_sink, 64));
    cudaStream_t s_fg, s_bg;
    CKR(cudaStreamCreateWithFlags(&s_fg, cudaStreamNonBlocking));
    CKR(cudaStreamCreateWithFlags(&s_bg, cudaStreamNonBlocking));
    CKR(cudaFuncSetAttribute(k_fg, cudaFuncAttributePreferredSharedMemoryCarveout, 72));
    CKR(cudaFuncSetAttribute(k_bg, cudaFuncAttributePreferredSharedMemoryCarveout, 72));
    CKR(cudaFuncSetAttribute(k_fg, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    // every shape does 64 x 19 tasks of 256 steps per SM; reps = tasks a CTA works through in sequence / its width
    const Shape shapes[] = {
        {"static19 (19-warp CTAs, 1 task per warp)", 19, 19, 0, 4, 16},
        {"single (1-warp CTAs)", 1, 1, 0, 4, 16 * 19},
        {"dynamic8 (8-warp CTAs, 19 tasks from a counter)", 8, 19, 1, 4, 16},
        {"dynamic12 (12-warp CTAs, 19 tasks from a counter)", 12, 19, 1, 4, 16},
        {"dynamic19 (19-warp CTAs, 38 tasks from a counter)", 19, 38, 1, 2, 16},
    };
    const char *names[] = {"", "alu x8 chains", "", "select-heavy x4"};
    for (const Shape &sh : shapes) {
        time_fg(s_fg, d_out, sh, nsm);
        const float base = time_fg(s_fg, d_out, sh, nsm);
        printf("%-52s alone: %.3f ms\n", sh.name, base);
        for (int per_sm : {1, 2, 4})
            for (int mode : {1, 3}) {
                k_bg<<<nsm * per_sm, 32, 0, s_bg>>>(mode, 800000000LL, d_sink);     // ~0.4 s: outlasts the measurement
                k_bg<<<1, 32, 0, s_fg>>>(5, 2000000LL, d_sink);                       // let the background spread over the empty SMs first
                const float ms = time_fg(s_fg, d_out, sh, nsm);
                CKR(cudaDeviceSynchronize());
                printf("    background %d one-warp CTAs per SM, %-16s: %.3f ms (x%.3f)\n", per_sm, names[mode], ms, ms / base);
            }
    }
    return 0;
}
