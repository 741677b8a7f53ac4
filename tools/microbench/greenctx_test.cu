// Experiment: can the long Fano runs be confined to a private SM partition (CUDA green contexts, driver API through
// cudaGetDriverEntryPoint -- libcuda is not linked) while runtime-API kernels run on streams of both partitions?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o greenctx_test greenctx_test.cu && ./greenctx_test
// Prints: the partitions the driver offers for several minimum counts (with / without IGNORE_SM_COSCHEDULING), the SM ids
// kernels really ran on, and what a long-resident background kernel costs a shared-memory-heavy FP32 kernel when the two
// share SMs versus when they sit on disjoint partitions.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <set>
#include <vector>

#define CKR(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
#define CKD(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { printf("%s: CUresult %d\n", #x, (int)r_); return 1; } } while (0)

template <class F>
static F entry(const char *name) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        printf("driver entry point %s not available\n", name);
        return nullptr;
    }
    return (F)fn;
}

__global__ void k_smid(unsigned *out) {
    unsigned id;
    asm("mov.u32 %0, %%smid;" : "=r"(id));
    if (threadIdx.x == 0) out[blockIdx.x] = id;
    // stay a little so that the grid spreads over every SM of the partition
    long long t0 = clock64();
    while (clock64() - t0 < 20000) {}
}

// background: one warp per CTA, dependent integer loop for `clocks` cycles, `smem` bytes of dynamic shared memory held
__global__ void k_background(long long clocks, unsigned *sink) {
    extern __shared__ unsigned bg[];
    const long long t0 = clock64();
    unsigned x = threadIdx.x;
    while (clock64() - t0 < clocks) {
#pragma unroll
        for (int k = 0; k < 64; k++) x = x * 1664525u + 1013904223u;
    }
    if (x == 0xdeadbeefu) sink[0] = x + bg[0];
}

// foreground: FP32-bound, 47 KB of shared memory per 608-thread CTA (the shape of k_sync_lags)
__global__ void __launch_bounds__(608, 2) k_foreground(float *out, int iters) {
    __shared__ float tab[47 * 256];
    for (int i = threadIdx.x; i < 47 * 256; i += blockDim.x) tab[i] = (float)i * 1e-3f;
    __syncthreads();
    float a0 = 0, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
    for (int i = 0; i < iters; i++) {
        float w = tab[(threadIdx.x + i) & 8191];
        a0 = a0 * w + a1; a1 = a1 * w + a2; a2 = a2 * w + a3; a3 = a3 * w + a4;
        a4 = a4 * w + a5; a5 = a5 * w + a6; a6 = a6 * w + a7; a7 = a7 * w + a0;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

static float time_fg(cudaStream_t st, float *out, int ctas) {
    cudaEvent_t e0, e1;
    CKR(cudaEventCreate(&e0));
    CKR(cudaEventCreate(&e1));
    k_foreground<<<ctas, 608, 0, st>>>(out, 2000);
    CKR(cudaStreamSynchronize(st));
    CKR(cudaEventRecord(e0, st));
    for (int r = 0; r < 5; r++) k_foreground<<<ctas, 608, 0, st>>>(out, 20000);
    CKR(cudaEventRecord(e1, st));
    CKR(cudaEventSynchronize(e1));
    float ms = 0;
    CKR(cudaEventElapsedTime(&ms, e0, e1));
    return ms / 5;
}

static std::set<unsigned> smids(cudaStream_t st, unsigned *d, int n) {
    CKR(cudaMemsetAsync(d, 0xff, n * sizeof(unsigned), st));
    k_smid<<<n, 32, 0, st>>>(d);
    std::vector<unsigned> h(n);
    CKR(cudaMemcpyAsync(h.data(), d, n * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CKR(cudaStreamSynchronize(st));
    return std::set<unsigned>(h.begin(), h.end());
}

int main(int argc, char **argv) {
    const int want = argc > 1 ? atoi(argv[1]) : 16;
    CKR(cudaSetDevice(0));
    CKR(cudaFree(0));
    auto pDeviceGet = entry<CUresult (*)(CUdevice *, int)>("cuDeviceGet");
    auto pGetRes = entry<CUresult (*)(CUdevice, CUdevResource *, CUdevResourceType)>("cuDeviceGetDevResource");
    auto pSplit = entry<CUresult (*)(CUdevResource *, unsigned *, const CUdevResource *, CUdevResource *, unsigned, unsigned)>("cuDevSmResourceSplitByCount");
    auto pDesc = entry<CUresult (*)(CUdevResourceDesc *, CUdevResource *, unsigned)>("cuDevResourceGenerateDesc");
    auto pCreate = entry<CUresult (*)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned)>("cuGreenCtxCreate");
    auto pStream = entry<CUresult (*)(CUstream *, CUgreenCtx, unsigned, int)>("cuGreenCtxStreamCreate");
    if (!pDeviceGet || !pGetRes || !pSplit || !pDesc || !pCreate || !pStream) return 2;
    CUdevice dev;
    CKD(pDeviceGet(&dev, 0));
    CUdevResource all;
    CKD(pGetRes(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
    printf("device SMs: %u\n", all.sm.smCount);
    for (unsigned flags = 0; flags <= 1; flags++)
        for (unsigned mc : {2u, 4u, 6u, 8u, 10u, 12u, 16u, 20u, 24u, 32u}) {
            CUdevResource grp[1], rem;
            unsigned ng = 1;
            CUresult r = pSplit(grp, &ng, &all, &rem, flags, mc);
            if (r == CUDA_SUCCESS) printf("split flags=%u minCount=%2u -> groups %u, first group %u SMs, remainder %u SMs\n", flags, mc, ng, grp[0].sm.smCount, rem.sm.smCount);
            else printf("split flags=%u minCount=%2u -> CUresult %d\n", flags, mc, (int)r);
        }
    // the partition under test: `want` SMs (flag 1 when the default granularity cannot give it) + the remainder
    CUdevResource small, rest;
    unsigned ng = 1;
    unsigned flags = 0;
    CKD(pSplit(&small, &ng, &all, &rest, flags, (unsigned)want));
    if ((int)small.sm.smCount != want) {
        flags = CU_DEV_SM_RESOURCE_SPLIT_IGNORE_SM_COSCHEDULING;
        ng = 1;
        CKD(pSplit(&small, &ng, &all, &rest, flags, (unsigned)want));
    }
    printf("using flags=%u: small %u SMs, rest %u SMs\n", flags, small.sm.smCount, rest.sm.smCount);
    CUdevResourceDesc dsmall, drest;
    CKD(pDesc(&dsmall, &small, 1));
    CKD(pDesc(&drest, &rest, 1));
    CUgreenCtx gsmall, grest;
    CKD(pCreate(&gsmall, dsmall, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    CKD(pCreate(&grest, drest, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    CUstream s_small, s_rest, s_rest2;
    CKD(pStream(&s_small, gsmall, CU_STREAM_NON_BLOCKING, 0));
    CKD(pStream(&s_rest, grest, CU_STREAM_NON_BLOCKING, 0));
    CKD(pStream(&s_rest2, grest, CU_STREAM_NON_BLOCKING, 0));
    cudaStream_t s_plain;
    CKR(cudaStreamCreateWithFlags(&s_plain, cudaStreamNonBlocking));

    unsigned *d_ids, *d_sink;
    float *d_out;
    CKR(cudaMalloc(&d_ids, 4096 * sizeof(unsigned)));
    CKR(cudaMalloc(&d_sink, 64));
    CKR(cudaMalloc(&d_out, (size_t)4096 * 608 * sizeof(float)));
    auto a = smids((cudaStream_t)s_small, d_ids, 2048), b = smids((cudaStream_t)s_rest, d_ids, 2048), c = smids(s_plain, d_ids, 2048);
    int overlap = 0;
    for (unsigned x : a) overlap += (int)b.count(x);
    printf("SM ids used: small stream %zu, rest stream %zu (overlap %d), plain stream %zu\n", a.size(), b.size(), overlap, c.size());
    printf("small partition SMs:");
    for (unsigned x : a) printf(" %u", x);
    printf("\n");

    // dynamic shared memory above 48 KB on a green-context stream (function attribute set through the runtime)
    CKR(cudaFuncSetAttribute(k_background, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    k_background<<<8, 32, 100 * 1024, (cudaStream_t)s_small>>>(1000, d_sink);
    CKR(cudaStreamSynchronize((cudaStream_t)s_small));
    printf("100 KB dynamic shared memory launch on the small partition: ok\n");
    // cross-partition event dependency
    cudaEvent_t ev;
    CKR(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    k_background<<<8, 32, 0, (cudaStream_t)s_rest>>>(100000, d_sink);
    CKR(cudaEventRecord(ev, (cudaStream_t)s_rest));
    CKR(cudaStreamWaitEvent((cudaStream_t)s_small, ev, 0));
    k_background<<<8, 32, 0, (cudaStream_t)s_small>>>(1000, d_sink);
    CKR(cudaStreamSynchronize((cudaStream_t)s_small));
    printf("event recorded on one partition, waited on by the other: ok\n");

    const int nsm = (int)all.sm.smCount;
    const long long bgclk = 2000000000LL;    // ~1 s: outlasts every foreground measurement
    float base_plain = time_fg(s_plain, d_out, 2 * nsm * 8);
    float base_rest = time_fg((cudaStream_t)s_rest, d_out, 2 * nsm * 8);
    printf("foreground alone: plain stream %.3f ms, rest partition %.3f ms (x%.3f, SM ratio %.3f)\n", base_plain, base_rest,
           base_rest / base_plain, (double)nsm / rest.sm.smCount);
    for (int smem : {0, 21 * 1024, 84 * 1024}) {
        for (int per_sm : {1, 4, 8}) {
            // (a) background on the plain stream sharing SMs with the foreground: `want * per_sm` one-warp CTAs
            const int nbg = want * per_sm;
            k_background<<<nbg, 32, smem, s_plain>>>(bgclk / 4, d_sink);
            cudaStream_t s2;
            CKR(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
            float shared_ms = time_fg(s2, d_out, 2 * nsm * 8);
            CKR(cudaDeviceSynchronize());
            // (b) the same background confined to the small partition, foreground on the rest
            k_background<<<nbg, 32, smem, (cudaStream_t)s_small>>>(bgclk / 4, d_sink);
            float part_ms = time_fg((cudaStream_t)s_rest, d_out, 2 * nsm * 8);
            CKR(cudaDeviceSynchronize());
            CKR(cudaStreamDestroy(s2));
            printf("background %4d one-warp CTAs holding %2d KB: foreground %.3f ms sharing SMs (x%.3f), %.3f ms partitioned (x%.3f)\n",
                   nbg, smem / 1024, shared_ms, shared_ms / base_plain, part_ms, part_ms / base_plain);
        }
    }
    return 0;
}
