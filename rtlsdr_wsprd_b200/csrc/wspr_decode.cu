// Host side of libwsprd_b200.so: device-buffer ownership, the wave scheduler that drives the kernels in the
// reference's control flow (wsprd/wsprd.c:416-855), and the C ABI declared in include/wspr_b200.h.
//
// Scheduling: the reference walks the candidates of one capture serially because, in pass 0, every successful
// decode is subtracted from the samples before the next candidate is examined (wsprd.c:781-789).  Captures are
// independent, so every capture runs its own state machine (wspr_kernels.cuh, enum Phase) and the host drives
// *rounds*: each round advances every ready capture by one candidate (sync search -> soft symbols -> budgeted Fano
// -> unpack/resolve -> subtraction) and runs the pass set-up (spectrogram, candidate search, coarse sync) for the
// captures that enter a new pass.  The rare candidates that need a long Fano run or the 42-attempt jitter search
// are finished by a per-device pool of Fano worker warps while the rounds go on, so a straggler delays only its own capture.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <sched.h>
#include <unistd.h>
#include <fcntl.h>
#include <sys/file.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/wspr_b200.h"
#include "wspr_kernels.cuh"
#include "wspr_math.cuh"

using namespace wspr;

static_assert(sizeof(Spot) == sizeof(decoder_results) && sizeof(Spot) == 80, "decoder_results layout");
static_assert(offsetof(Spot, message) == 28 && offsetof(Spot, call) == 51 && offsetof(Spot, loc) == 64 &&
                  offsetof(Spot, pwr) == 71 && offsetof(Spot, cycles) == 76, "decoder_results layout");
static_assert(sizeof(Cand) == sizeof(cand) && sizeof(Cand) == 20, "cand layout");
static_assert(sizeof(decoder_options) == 40, "decoder_options layout");

static thread_local std::string g_err;
static int fail(int code, const char *what, cudaError_t e = cudaSuccess) {
    g_err = what;
    if (e != cudaSuccess) {
        g_err += ": ";
        g_err += cudaGetErrorString(e);
    }
    return code;
}
#define CK(call)                                                 \
    do {                                                         \
        cudaError_t e_ = (call);                                 \
        if (e_ != cudaSuccess) return fail(WSPR_ERR_CUDA, #call, e_); \
    } while (0)

extern "C" const char *wspr_last_error(void) { return g_err.c_str(); }
extern "C" unsigned long long wspr_kernel_launches(void) { return kernel_launch_count(); }

// ---- constant tables: the values the reference computes with the host libm at every call ----
static void host_tables(HostTables &t) {
    for (int i = 0; i < NFFT; i++) t.window[i] = sinf(0.006147931 * i);          // wsprd.c:510-513
    float norm = 0;
    for (int i = 0; i < NFILT; i++) {                                             // wsprd.c:359-368
        t.lpf_w[i] = sinf(M_PI * (float)i / (float)(NFILT - 1));
        norm = norm + t.lpf_w[i];
    }
    for (int i = 0; i < NFILT; i++) t.lpf_w[i] = t.lpf_w[i] / norm;
    t.lpf_psum[0] = 0.0f;
    for (int i = 1; i < NFILT; i++) t.lpf_psum[i] = t.lpf_psum[i - 1] + t.lpf_w[i];
    t.min_snr = powf(10.0, -8.0 / 10.0);                                          // wsprd.c:590
    t.floor_snr = 0.1 * t.min_snr;                                                // wsprd.c:595
}

// ---- per-device Fano service --------------------------------------------------------------------------------------
// The long Fano runs of parked candidates (wsprd.c:741-766 when the decoder does not converge at once) are served by ONE pool
// of worker warps per GPU, fed from a device-side queue that every context of the process pushes to (wspr_kernels.cu,
// k_fano_workers).  Optionally the pool runs on its own SM partition (CUDA green contexts through the driver API; libcuda is
// not linked, the entry points come from cudaGetDriverEntryPoint): the workers then never share an SM -- its issue slots, its
// shared-memory carve-out -- with the bulk kernels, which get the remaining SMs.  Every placement other than the default
// (two one-warp workers per SM, no partition, one common 164 KB carve-out) measured slower (profiles/r2_bench_variants.txt),
// so only two sizing knobs are read by the product library, once, when the first context on a device is created:
//   WSPR_FANO_POOL    worker warps alive at any time (default 2 per SM)
//   WSPR_FANO_PER_SM  worker warps allowed on one SM (0 = no limit; default 2)
// and the alternatives can be selected at run time only in the experiment build (make exp):
//   WSPR_FANO_SMS   SMs set aside for the pool (even; 0 = no partition, workers and bulk kernels share every SM)
//   WSPR_FANO_SHARE 1: the pool is confined to its WSPR_FANO_SMS SMs but the other kernels run on ALL SMs (they use whatever
//                   the workers leave of those SMs); 0: the other kernels keep off the pool's SMs
//   WSPR_FANO_OVERFLOW  with a partition: this many extra worker warps may run on the other kernels' SMs (at most
//                   WSPR_FANO_PER_SM per SM) while at least WSPR_FANO_OVERFLOW_BACKLOG candidates wait for lanes
//   WSPR_FANO_CTA_WARPS  worker warps per CTA (1, 2 or 4)
//   WSPR_CARVEOUT_KB  common shared-memory carve-out of every decode kernel (0 = the driver's per-kernel choice)
#ifndef WSPR_DEFAULT_FANO_SMS
#define WSPR_DEFAULT_FANO_SMS 0
#endif
#ifndef WSPR_DEFAULT_CARVEOUT_KB
#define WSPR_DEFAULT_CARVEOUT_KB 164              // K4's two 47 KB CTAs + two worker warps; measured in profiles/r2_bench_variants.txt
#endif
#ifndef WSPR_DEFAULT_FANO_OVERFLOW
#define WSPR_DEFAULT_FANO_OVERFLOW 0
#endif
#ifndef WSPR_DEFAULT_FANO_SHARE
#define WSPR_DEFAULT_FANO_SHARE 0
#endif
#ifndef WSPR_DEFAULT_FANO_CTA_WARPS
#define WSPR_DEFAULT_FANO_CTA_WARPS 1
#endif
#ifndef WSPR_DEFAULT_FANO_PER_SM
#define WSPR_DEFAULT_FANO_PER_SM 2
#endif
#ifndef WSPR_DEFAULT_FANO_POOL_PER_SM
#define WSPR_DEFAULT_FANO_POOL_PER_SM 2
#endif
constexpr int EXACT_SIZING_MIN_JOBS = 64;
constexpr int FANO_RING = 1 << 18;                 // candidates the queue can hold (64 contexts of 4096 captures, all parked)
constexpr int NFANO_STREAMS = 4;

struct FanoService {
    int device = -1;
    int fano_sms = 0, pool = 0, total_sms = 0, cta_warps = 1;
    int pool2 = 0;                                 // overflow worker warps on the other kernels' SMs (partition only)
    bool partitioned = false;
    bool shared_bulk = false;                      // the other kernels may use the pool's SMs as well (WSPR_FANO_SHARE)
    CUgreenCtx g_fano = nullptr, g_bulk = nullptr;
    FanoQueue *queue = nullptr;                    // device
    FanoQueueEntry *ring = nullptr;                // device
    std::mutex mu;
    // (records, memory) returned by destroyed contexts: [0] records armed with all 43 attempts, [1] quick-mode records
    std::vector<std::pair<size_t, ChainScratch *>> free_scratch[2];
    std::string note;
};
static std::mutex g_svc_mu;
static FanoService *g_svc[64] = {nullptr};

template <class F>
static F driver_entry(const char *name) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return (F)fn;
}

static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}
// placement alternatives that were measured and rejected (profiles/r2_bench_variants.txt): selectable at run time only in
// the experiment build (make exp), the product library runs its compiled-in defaults
static int exp_int(const char *name, int dflt) {
#ifdef WSPR_EXPERIMENTS
    return env_int(name, dflt);
#else
    (void)name;
    return dflt;
#endif
}

// SM partition: `want` SMs for the Fano pool, the rest for everything else
static bool fano_partition(FanoService *s, int want) {
    auto pDeviceGet = driver_entry<CUresult (*)(CUdevice *, int)>("cuDeviceGet");
    auto pGetRes = driver_entry<CUresult (*)(CUdevice, CUdevResource *, CUdevResourceType)>("cuDeviceGetDevResource");
    auto pSplit = driver_entry<CUresult (*)(CUdevResource *, unsigned *, const CUdevResource *, CUdevResource *, unsigned, unsigned)>(
        "cuDevSmResourceSplitByCount");
    auto pDesc = driver_entry<CUresult (*)(CUdevResourceDesc *, CUdevResource *, unsigned)>("cuDevResourceGenerateDesc");
    auto pCreate = driver_entry<CUresult (*)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned)>("cuGreenCtxCreate");
    if (!pDeviceGet || !pGetRes || !pSplit || !pDesc || !pCreate) return false;
    CUdevice dev;
    CUdevResource all, small, rest;
    if (pDeviceGet(&dev, s->device) != CUDA_SUCCESS || pGetRes(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
    unsigned ng = 1;
    // (the default granularity on sm_100 is 8 SMs; IGNORE_SM_COSCHEDULING gives any even count -- no clusters are used here)
    if (pSplit(&small, &ng, &all, &rest, CU_DEV_SM_RESOURCE_SPLIT_IGNORE_SM_COSCHEDULING, (unsigned)want) != CUDA_SUCCESS || ng != 1) return false;
    CUdevResourceDesc dsmall, drest;
    if (pDesc(&dsmall, &small, 1) != CUDA_SUCCESS || pDesc(&drest, &rest, 1) != CUDA_SUCCESS) return false;
    if (pCreate(&s->g_fano, dsmall, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
    if (pCreate(&s->g_bulk, drest, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
    s->fano_sms = (int)small.sm.smCount;
    return true;
}

// (the caller has made `device` current)
static FanoService *fano_service(int device) {
    std::lock_guard<std::mutex> lock(g_svc_mu);
    if (device < 0 || device >= 64) return nullptr;
    if (g_svc[device]) return g_svc[device];
    FanoService *s = new FanoService();
    s->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete s; return nullptr; }
    s->total_sms = prop.multiProcessorCount;
    init_kernel_attributes(exp_int("WSPR_CARVEOUT_KB", WSPR_DEFAULT_CARVEOUT_KB));
    const int want = exp_int("WSPR_FANO_SMS", WSPR_DEFAULT_FANO_SMS);
    if (want > 0 && want < s->total_sms) {
        s->partitioned = fano_partition(s, want);
        if (!s->partitioned) s->note = "SM partition unavailable (green contexts): Fano workers share the SMs";
    }
    if (!s->partitioned) s->fano_sms = 0;
    s->shared_bulk = s->partitioned && exp_int("WSPR_FANO_SHARE", WSPR_DEFAULT_FANO_SHARE) != 0;
    const int per_sm = (228 * 1024) / (fano_warp_smem_bytes() + 1024);
    s->pool = env_int("WSPR_FANO_POOL", s->partitioned ? s->fano_sms * per_sm : s->total_sms * WSPR_DEFAULT_FANO_POOL_PER_SM);
    if (s->pool < 1) s->pool = 1;
    s->cta_warps = exp_int("WSPR_FANO_CTA_WARPS", WSPR_DEFAULT_FANO_CTA_WARPS);
    FanoQueue h;
    memset(&h, 0, sizeof h);
    h.pool = s->pool;
    // with a partition: the per-SM limit is for the overflow workers (the partition's SMs hold as many as fit)
    s->pool2 = s->partitioned && !s->shared_bulk ? std::max(0, exp_int("WSPR_FANO_OVERFLOW", WSPR_DEFAULT_FANO_OVERFLOW)) : 0;
    h.pool2 = s->pool2;
    h.ovf_backlog = std::max(1, exp_int("WSPR_FANO_OVERFLOW_BACKLOG", 8));
    h.per_sm = env_int("WSPR_FANO_PER_SM", (s->partitioned && s->pool2 == 0) ? 0 : WSPR_DEFAULT_FANO_PER_SM);
    h.mask = FANO_RING - 1;
    if (cudaMalloc((void **)&s->ring, (size_t)FANO_RING * sizeof(FanoQueueEntry)) != cudaSuccess ||
        cudaMemset(s->ring, 0, (size_t)FANO_RING * sizeof(FanoQueueEntry)) != cudaSuccess ||
        cudaMalloc((void **)&s->queue, sizeof(FanoQueue)) != cudaSuccess) { delete s; return nullptr; }
    h.ring = s->ring;
    if (cudaMemcpy(s->queue, &h, sizeof h, cudaMemcpyHostToDevice) != cudaSuccess) { delete s; return nullptr; }
#ifdef WSPR_EXPERIMENTS
    exp_set_queue(s->queue);
#endif
    g_svc[device] = s;
    return s;
}

static cudaError_t service_stream(FanoService *s, bool fano, cudaStream_t *out) {
    if (s->partitioned && (fano || !s->shared_bulk)) {
        static auto pStream = driver_entry<CUresult (*)(CUstream *, CUgreenCtx, unsigned, int)>("cuGreenCtxStreamCreate");
        CUstream st = nullptr;
        if (!pStream || pStream(&st, fano ? s->g_fano : s->g_bulk, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return cudaErrorUnknown;
        *out = (cudaStream_t)st;
        return cudaSuccess;
    }
    return cudaStreamCreateWithFlags(out, cudaStreamNonBlocking);
}

// ChainScratch records are leased from the service and never freed while the process runs (see wspr_kernels.cuh).
// A worker may reach a record through an OLD ring entry (the head1 cursor lags behind while older candidates still have
// attempts to hand out), so an idle record must never look claimable: next >= nattempts.  That holds after every hand-back
// as long as a record is always armed with the same number of attempts -- a quick-mode candidate (attempt 0 only) leaves
// next = 1, which is "nothing left" for 1 attempt but "42 left" the moment the record is re-armed for 43.  Quick-mode
// decodes therefore park their candidates on records of their own (`quick` pool), leased when first needed.
static ChainScratch *lease_scratch(FanoService *s, size_t n, int quick = 0) {
    {
        std::lock_guard<std::mutex> lock(s->mu);
        auto &pool = s->free_scratch[quick ? 1 : 0];
        for (size_t i = 0; i < pool.size(); i++)
            if (pool[i].first >= n) {
                ChainScratch *p = pool[i].second;
                pool.erase(pool.begin() + i);
                return p;
            }
    }
    ChainScratch *p = nullptr;
    if (cudaMalloc((void **)&p, n * sizeof(ChainScratch)) != cudaSuccess) return nullptr;
    // next >= nattempts in every record: nothing to claim until a candidate is parked there
    if (cudaMemset(p, 0, n * sizeof(ChainScratch)) != cudaSuccess) { cudaFree(p); return nullptr; }
    return p;
}

struct wspr_ctx {
    int device = 0, maxcap = 0, np = 0, stride = 0, blocks = 0;
    int ncap = 0;
    FanoService *svc = nullptr;
    cudaStream_t st = nullptr;
    cudaStream_t fano_st[NFANO_STREAMS] = {nullptr};   // worker warps are launched here (the pool's SM partition, if any)
    cudaStream_t ovf_st[NFANO_STREAMS] = {nullptr};    // overflow worker warps (the other kernels' SMs)
    int fano_rr = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_wait = nullptr;   // blocking-sync event: host threads sleep instead of spinning while the GPU works
    cudaEvent_t ev_fano = nullptr;   // the round's candidates are in the queue
    float *I = nullptr, *Q = nullptr, *psT = nullptr, *smspec = nullptr;
    Cand *cands = nullptr;
    CapState *caps = nullptr;
    Spot *spots = nullptr;
    int *nres = nullptr;
    Job *jobs = nullptr;           // [maxcap] the candidate each capture is working on
    Attempt *att0 = nullptr;       // [maxcap] its jitter-0 attempt
    int *ident = nullptr, *setup_list = nullptr, *job_list = nullptr, *res_list = nullptr, *sub_list = nullptr, *defer_list = nullptr;
    float4 *P0 = nullptr, *P1 = nullptr, *tabs = nullptr;
    float *phi0 = nullptr;
    float2 *ref = nullptr, *cprod = nullptr;
    Counters *cnt = nullptr;       // device
    Counters *h_cnt = nullptr;     // pinned host mirror
    uint4 *moments = nullptr;      // front-end scratch of wspr_ctx_decimate: [streams][blocks] block moments
    size_t moments_cap = 0;
    ChainScratch *scratch = nullptr;   // [maxcap], leased from the service
    ChainScratch *scratch_quick = nullptr;   // [maxcap], the records of quick-mode decodes (leased by the first one)
    int *h_done = nullptr;         // pinned, device-visible: parked captures handed back by the Fano workers so far
    char *preload = nullptr;       // device [32768][13]: hashtable.txt calls (allocated on the first -H decode)
    int *stats = nullptr;          // device [8]: how deferred candidates were settled
    int h_stats[8] = {0};
    unsigned fano_budget = 4096;   // WSPR_FANO_BUDGET (read when the context is created)
    int park_linger_us = 300;      // WSPR_PARK_LINGER_US: how long to wait for more parked captures once one has come back
    bool trace = false;            // WSPR_TRACE
    float last_ms = 0.0f, sync_ms = 0.0f;
    int sync_launches = 0, rounds = 0, deferred = 0;
    double sync_cells = 0.0;
    bool time_kernels = false;
    std::vector<cudaEvent_t> kev;  // event pairs around the mode-0 sync kernel, one pair per round
    std::vector<int> kev_jobs;
};

template <class T>
static cudaError_t dalloc(T **p, size_t n) { return cudaMalloc((void **)p, n * sizeof(T)); }

extern "C" void wspr_ctx_destroy(wspr_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    void *ptrs[] = {c->I, c->Q, c->psT, c->smspec, c->cands, c->caps, c->spots, c->nres, c->jobs, c->att0, c->ident,
                    c->setup_list, c->job_list, c->res_list, c->sub_list, c->defer_list, c->P0, c->P1, c->tabs, c->phi0, c->ref, c->cprod,
                    c->cnt, c->stats, c->preload, c->moments};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (c->svc) {
        std::lock_guard<std::mutex> lock(c->svc->mu);
        if (c->scratch) c->svc->free_scratch[0].push_back({(size_t)c->maxcap, c->scratch});
        if (c->scratch_quick) c->svc->free_scratch[1].push_back({(size_t)c->maxcap, c->scratch_quick});
    }
    for (cudaStream_t s : c->fano_st)
        if (s) cudaStreamDestroy(s);
    for (cudaStream_t s : c->ovf_st)
        if (s) cudaStreamDestroy(s);
    for (cudaEvent_t e : c->kev) cudaEventDestroy(e);
    if (c->h_cnt) cudaFreeHost(c->h_cnt);
    if (c->h_done) cudaFreeHost(c->h_done);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_wait) cudaEventDestroy(c->ev_wait);
    if (c->ev_fano) cudaEventDestroy(c->ev_fano);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
}

static int ctx_init(wspr_ctx *c, int device, int maxcap, int samples) {
    if (maxcap <= 0 || samples < NFFT) return fail(WSPR_ERR_ARG, "wspr_ctx_create: bad sizes");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) return fail(WSPR_ERR_CUDA, "no CUDA device");
    if (device < 0) CK(cudaGetDevice(&device));
    CK(cudaSetDevice(device));
    c->device = device;
    c->maxcap = maxcap;
    c->np = samples;
    // a row holds the capture and, zeroed, whatever the spectrogram reads past its end: block i covers samples 128 i .. 128 i + 511
    // and there are 4 floor(samples / 512) - 1 blocks, so the last one ends at 512 floor(samples / 512) + 255 -- beyond a
    // capture whose length is below 256 modulo 512 (wsprd.c:516,536-541; the reference's buffers are full-size and zero-tailed)
    c->stride = (std::max(samples, NFFT * (samples / NFFT) + NFFT / 2) + 127) / 128 * 128;
    c->blocks = 4 * (samples / NFFT) - 1;                    // wsprd.c:516
    c->svc = fano_service(device);
    if (!c->svc) return fail(WSPR_ERR_CUDA, "Fano service set-up failed", cudaGetLastError());
    CK(service_stream(c->svc, false, &c->st));
    for (cudaStream_t &s : c->fano_st) CK(service_stream(c->svc, true, &s));
    if (c->svc->pool2 > 0)
        for (cudaStream_t &s : c->ovf_st) CK(service_stream(c->svc, false, &s));
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    CK(cudaEventCreateWithFlags(&c->ev_wait, cudaEventBlockingSync | cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_fano, cudaEventDisableTiming));
    c->fano_budget = (unsigned)std::max(256, env_int("WSPR_FANO_BUDGET", 4096));   // tuning knob, read once
    c->trace = getenv("WSPR_TRACE") != nullptr;
    c->park_linger_us = env_int("WSPR_PARK_LINGER_US", 300);
    size_t B = (size_t)maxcap;
    CK(dalloc(&c->I, B * c->stride));
    CK(dalloc(&c->Q, B * c->stride));
    CK(cudaMemset(c->I, 0, B * c->stride * sizeof(float)));
    CK(cudaMemset(c->Q, 0, B * c->stride * sizeof(float)));
    CK(dalloc(&c->psT, B * c->blocks * NFFT));
    CK(dalloc(&c->smspec, B * NSMOOTH));
    CK(dalloc(&c->cands, B * MAXCAND));
    CK(dalloc(&c->caps, B));
    CK(dalloc(&c->spots, B * MAXUNIQ));
    CK(cudaMemset(c->spots, 0, B * MAXUNIQ * sizeof(Spot)));
    CK(dalloc(&c->nres, B));
    CK(dalloc(&c->jobs, B));
    CK(dalloc(&c->att0, B));
    CK(dalloc(&c->ident, B));
    CK(dalloc(&c->setup_list, B));
    CK(dalloc(&c->job_list, B));
    CK(dalloc(&c->res_list, B));
    CK(dalloc(&c->sub_list, B));
    CK(dalloc(&c->defer_list, B));
    CK(dalloc(&c->P0, B * MAXLAGS * NSYM));
    CK(dalloc(&c->P1, B * NFREQ1 * NSYM));
    CK(dalloc(&c->tabs, B * NFREQ1 * 2 * SPS));
    CK(dalloc(&c->phi0, B * (NSIG / PHI_SEG)));            // running phase of every PHI_SEG-th sample of a subtraction
    CK(dalloc(&c->ref, B * NSIG));
    CK(dalloc(&c->cprod, B * CPAD));
    CK(dalloc(&c->cnt, 1));
    CK(dalloc(&c->stats, 8));
    CK(cudaMallocHost((void **)&c->h_cnt, sizeof(Counters)));
    {
        std::vector<int> id(B);
        for (size_t i = 0; i < B; i++) id[i] = (int)i;
        CK(cudaMemcpy(c->ident, id.data(), B * sizeof(int), cudaMemcpyHostToDevice));
    }
    c->scratch = lease_scratch(c->svc, B);
    if (!c->scratch) return fail(WSPR_ERR_CUDA, "ChainScratch allocation", cudaGetLastError());
    CK(cudaMallocHost((void **)&c->h_done, sizeof(int)));
    *c->h_done = 0;
    HostTables t;
    host_tables(t);
    upload_tables(t);
    CK(cudaGetLastError());
    return WSPR_OK;
}

extern "C" wspr_ctx *wspr_ctx_create(int device, int max_captures, int samples) {
    wspr_ctx *c = new wspr_ctx();
    if (ctx_init(c, device, max_captures, samples) != WSPR_OK) {
        std::string keep = g_err;
        wspr_ctx_destroy(c);
        g_err = keep;
        return nullptr;
    }
    return c;
}

extern "C" int wspr_ctx_upload(wspr_ctx *c, const float *I, const float *Q, int ncap) {
    if (!c || ncap < 0 || ncap > c->maxcap) return fail(WSPR_ERR_ARG, "wspr_ctx_upload: bad arguments");
    CK(cudaSetDevice(c->device));
    c->ncap = ncap;
    if (ncap == 0) return WSPR_OK;
    size_t w = (size_t)c->np * sizeof(float);
    CK(cudaMemcpy2DAsync(c->I, (size_t)c->stride * sizeof(float), I, w, w, ncap, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpy2DAsync(c->Q, (size_t)c->stride * sizeof(float), Q, w, w, ncap, cudaMemcpyHostToDevice, c->st));
    return WSPR_OK;
}

extern "C" int wspr_ctx_upload_device(wspr_ctx *c, const float *dI, const float *dQ, int ncap, int row_stride) {
    if (!c || ncap < 0 || ncap > c->maxcap || row_stride < c->np) return fail(WSPR_ERR_ARG, "wspr_ctx_upload_device");
    CK(cudaSetDevice(c->device));
    c->ncap = ncap;
    if (ncap == 0) return WSPR_OK;
    size_t w = (size_t)c->np * sizeof(float);
    CK(cudaMemcpy2DAsync(c->I, (size_t)c->stride * sizeof(float), dI, (size_t)row_stride * sizeof(float), w, ncap,
                         cudaMemcpyDeviceToDevice, c->st));
    CK(cudaMemcpy2DAsync(c->Q, (size_t)c->stride * sizeof(float), dQ, (size_t)row_stride * sizeof(float), w, ncap,
                         cudaMemcpyDeviceToDevice, c->st));
    return WSPR_OK;
}

// Raw streams straight into the context (BASELINE config 4): rtlsdr_callback (rtlsdr_wsprd.c:126-244) for `nstreams` whole
// streams resident on the device, written into the context's sample planes with the tail zeroed like the daemon's hand-off
// (rtlsdr_wsprd.c:285-288); follow with wspr_ctx_normalise (:291-305) and wspr_ctx_decode (:316).  Asynchronous on the
// context's stream.  Returns the samples every stream produced, or a negative error.
extern "C" int wspr_ctx_decimate(wspr_ctx *c, const uint8_t *d_raw, int nstreams, size_t n_iq, size_t stream_stride_bytes) {
    if (!c || !d_raw || nstreams < 0 || nstreams > c->maxcap) return fail(WSPR_ERR_ARG, "wspr_ctx_decimate: bad arguments");
    if (((uintptr_t)d_raw & 15) || (stream_stride_bytes & 15) || stream_stride_bytes < 2 * n_iq)
        return fail(WSPR_ERR_ARG, "wspr_ctx_decimate: raw streams must be 16-byte aligned and strided");
    CK(cudaSetDevice(c->device));
    c->ncap = nstreams;
    if (nstreams == 0) return 0;
    const int nblk = decimate_outputs(n_iq);
    const size_t need = (size_t)nstreams * (size_t)std::max(nblk, 1);
    if (c->moments_cap < need) {
        if (c->moments) CK(cudaFree(c->moments));
        c->moments = nullptr;
        c->moments_cap = 0;
        CK(dalloc(&c->moments, need));
        c->moments_cap = need;
    }
    launch_decimate(d_raw, n_iq, nstreams, stream_stride_bytes, c->moments, c->I, c->Q, c->stride, c->np, c->st);
    CK(cudaGetLastError());
    return std::min(nblk, c->np);
}

extern "C" int wspr_ctx_normalise(wspr_ctx *c) {
    if (!c) return fail(WSPR_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    launch_normalise(c->I, c->Q, c->ncap, c->np, c->stride, c->st);
    CK(cudaGetLastError());
    return WSPR_OK;
}

static DecodeParams make_params(const wspr_ctx *c, const decoder_options &o) {
    DecodeParams p;
    p.np = c->np;
    p.stride = c->stride;
    p.blocks = c->blocks;
    p.dialfreq = o.freq;
    p.quickmode = o.quickmode;
    p.subtraction = o.subtraction;
    p.npasses = o.npasses;
    p.minsync1 = 0.10;                                       // wsprd.c:424-433
    p.symfac = 50;
    p.minrms = 52.0 * (p.symfac / 64.0);
    p.delta = 60;
    p.maxcycles = 10000;
    p.lagstep = o.quickmode ? 16 : 8;                        // wsprd.c:715-717
    p.nlags = 256 / p.lagstep + 1;
    p.preload = nullptr;
    p.fano_budget = c->fano_budget;
    return p;
}

// wait for everything issued on the context's stream so far, yielding the CPU (several contexts per GPU and several
// GPUs per host are each driven by a host thread; spinning in cudaStreamSynchronize would oversubscribe the cores)
static bool wait_blocks() {                                    // WSPR_WAIT=block: sleep in the driver instead of polling
    static const bool b = [] { const char *e = getenv("WSPR_WAIT"); return e && e[0] == 'b'; }();
    return b;
}
static int wait_event(cudaEvent_t ev) {
    if (!wait_blocks()) {                                      // poll, giving the core away whenever someone else wants it
        for (;;) {
            cudaError_t q = cudaEventQuery(ev);
            if (q == cudaSuccess) return WSPR_OK;
            if (q != cudaErrorNotReady) return fail(WSPR_ERR_CUDA, "cudaEventQuery", q);
            sched_yield();
        }
    }
    CK(cudaEventSynchronize(ev));
    return WSPR_OK;
}
static int wait_stream(wspr_ctx *c) {
    CK(cudaEventRecord(c->ev_wait, c->st));
    return wait_event(c->ev_wait);
}

static int read_counters(wspr_ctx *c) {
    CK(cudaMemcpyAsync(c->h_cnt, c->cnt, sizeof(Counters), cudaMemcpyDeviceToHost, c->st));
    return wait_stream(c);
}

// (`seen`: the count read before the round was planned, so a hand-back in between is not missed)
// give the pool up to park_linger_us to hand `want` captures back (counted from `seen`): a round costs two host
// synchronisations and some twenty launches whatever its size
static void linger_parked(wspr_ctx *c, int seen, int want) {
    if (c->park_linger_us <= 0) return;
    const volatile int *done = c->h_done;
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    while (*done - seen < want) {
        clock_gettime(CLOCK_MONOTONIC, &t1);
        if ((t1.tv_sec - t0.tv_sec) * 1000000L + (t1.tv_nsec - t0.tv_nsec) / 1000L >= c->park_linger_us) break;
        if (wait_blocks()) usleep(50);
        else sched_yield();
    }
}
// every open capture is parked: sleep until at least one has been handed back, then linger for more
static int wait_parked(wspr_ctx *c, int seen, int want, const DecodeParams &p) {
    const volatile int *done = c->h_done;
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (unsigned spin = 0; *done == seen; spin++) {
        if ((spin & 1023u) == 1023u) {             // (a failed kernel would otherwise leave us here for good)
            cudaError_t e = cudaStreamQuery(c->fano_st[0]);
            if (e != cudaSuccess && e != cudaErrorNotReady) return fail(WSPR_ERR_CUDA, "Fano workers", e);
            // belt and braces: nothing has come back for a quarter of a second (a hopeless attempt alone takes 0.1-0.2 s, so
            // this is rare but legitimate) -- start a few workers.  They leave at once if the pool is complete or the queue
            // empty; if candidates were ever left in the queue with nobody alive to work on them (the window described
            // above k_fano_workers), this is what picks them up.
            clock_gettime(CLOCK_MONOTONIC, &t1);
            if ((t1.tv_sec - t0.tv_sec) * 1000L + (t1.tv_nsec - t0.tv_nsec) / 1000000L >= 250) {
                // (a ring that has overflowed has lost candidates: their captures would never come back)
                int overflow = 0;
                CK(cudaMemcpy(&overflow, &c->svc->queue->overflow, sizeof(int), cudaMemcpyDeviceToHost));
                if (overflow) return fail(WSPR_ERR_CUDA, "Fano queue overflow: parked candidates were lost");
                launch_fano_workers(c->svc->queue, 4, c->svc->cta_warps, false, p, c->fano_st[c->fano_rr++ % NFANO_STREAMS]);
                CK(cudaGetLastError());
                t0 = t1;
            }
        }
        if (wait_blocks()) usleep(50);
        else sched_yield();
    }
    if (want > 1) linger_parked(c, seen, want);
    return WSPR_OK;
}

// ---- options.usehashtable (reference -H): hashtable.txt in the CWD, wsprd.c:481-494 and :842-852 ----
// Several contexts (host threads) and several ranks may decode with -H in the same CWD: the file is read when a decode
// starts and rewritten when it ends, each under a process mutex plus an advisory lock on hashtable.txt.lock, and the rewrite
// goes through a temporary file and rename() so that a reader never sees a half-written table.  Between its read and its
// write-back a decode does not hold the lock; the write-back therefore re-reads the file and merges this decode's
// additions into it, so entries other decodes added in the meantime are kept.
constexpr int HT_SIZE = 32768;                                // HASHTAB_SIZE, wsprd/wsprd.h
static std::mutex g_ht_mu;
struct HtFileLock {
    int fd;
    HtFileLock() {
        g_ht_mu.lock();
        fd = open("hashtable.txt.lock", O_CREAT | O_RDWR, 0644);
        if (fd >= 0) flock(fd, LOCK_EX);
    }
    ~HtFileLock() {
        if (fd >= 0) {
            flock(fd, LOCK_UN);
            close(fd);
        }
        g_ht_mu.unlock();
    }
};
struct HostHashTables {
    std::vector<char> calls, locs;                            // [32768][13], [32768][5]
    HostHashTables() : calls((size_t)HT_SIZE * CALL_LEN, 0), locs((size_t)HT_SIZE * LOC_LEN, 0) {}
};
static void hashtable_read(HostHashTables &t) {
    FILE *f = fopen("hashtable.txt", "r+");
    if (!f) return;
    char line[80], hcall[80] = "", hgrid[80];
    int nh = -1;                                              // (like the reference, a line that does not parse repeats the previous entry)
    while (fgets(line, sizeof line, f) != NULL) {
        hgrid[0] = 0;
        sscanf(line, "%d %s %s", &nh, hcall, hgrid);
        if (nh >= 0 && nh < HT_SIZE) {
            snprintf(t.calls.data() + (size_t)nh * CALL_LEN, CALL_LEN, "%s", hcall);
            if (strlen(hgrid) > 0) snprintf(t.locs.data() + (size_t)nh * LOC_LEN, LOC_LEN, "%s", hgrid);
        }
    }
    fclose(f);
}
static void hashtable_write(const HostHashTables &t) {
    char tmp[64];
    snprintf(tmp, sizeof tmp, "hashtable.txt.%ld.tmp", (long)getpid());
    FILE *f = fopen(tmp, "w");
    if (!f) return;
    for (int i = 0; i < HT_SIZE; i++)
        if (t.calls[(size_t)i * CALL_LEN] != 0)
            fprintf(f, "%5d %s %s\n", i, t.calls.data() + (size_t)i * CALL_LEN, t.locs.data() + (size_t)i * LOC_LEN);
    fclose(f);
    rename(tmp, "hashtable.txt");
}
// entries the captures added during the decode, merged in capture order (one capture = the reference's semantics; with
// several captures in a batch every capture saw the table as it was on entry, and later captures win on write-back)
static int hashtable_merge(wspr_ctx *c, HostHashTables &t) {
    std::vector<CapState> caps(c->ncap);
    CK(cudaMemcpyAsync(caps.data(), c->caps, (size_t)c->ncap * sizeof(CapState), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    HtFileLock lock;
    hashtable_read(t);                                        // (what other decodes wrote since this one started)
    for (const CapState &cs : caps)
        for (int i = 0; i < cs.nhash && i < HASH_CAP; i++) {
            const HashEntry &e = cs.hash[i];
            if (e.h < 0 || e.h >= HT_SIZE) continue;
            snprintf(t.calls.data() + (size_t)e.h * CALL_LEN, CALL_LEN, "%s", e.call);
            if (e.loc[0]) snprintf(t.locs.data() + (size_t)e.h * LOC_LEN, LOC_LEN, "%s", e.loc);
        }
    hashtable_write(t);
    return WSPR_OK;
}

extern "C" int wspr_ctx_decode(wspr_ctx *c, decoder_options o) {
    if (!c) return fail(WSPR_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    const int ncap = c->ncap;
    DecodeParams p = make_params(c, o);
    HostHashTables *ht = nullptr;
    if (o.usehashtable && ncap > 0) {
        ht = new HostHashTables();
        {
            HtFileLock lock;
            hashtable_read(*ht);
        }
        cudaError_t e = c->preload ? cudaSuccess : cudaMalloc((void **)&c->preload, ht->calls.size());
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->preload, ht->calls.data(), ht->calls.size(), cudaMemcpyHostToDevice, c->st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->st);
        if (e != cudaSuccess) {
            delete ht;
            return fail(WSPR_ERR_CUDA, "hashtable upload", e);
        }
        p.preload = c->preload;
    }
    struct HtGuard {
        HostHashTables *h;
        ~HtGuard() { delete h; }
    } ht_guard{ht};
    c->sync_ms = 0.0f;
    c->sync_launches = 0;
    c->sync_cells = 0.0;
    c->rounds = 0;
    c->deferred = 0;
    c->kev_jobs.clear();
    size_t kev_used = 0;
    const bool trace = c->trace;
    CK(cudaEventRecord(c->ev0, c->st));
    CK(cudaMemsetAsync(c->stats, 0, 8 * sizeof(int), c->st));
    launch_reset_caps(c->caps, ncap, o.npasses, c->st);
    // quick-mode candidates (attempt 0 only) are parked on records of their own, see lease_scratch
    ChainScratch *scratch = c->scratch;
    if (p.quickmode) {
        if (!c->scratch_quick) {
            c->scratch_quick = lease_scratch(c->svc, (size_t)c->maxcap, 1);
            if (!c->scratch_quick) return fail(WSPR_ERR_CUDA, "ChainScratch allocation", cudaGetLastError());
            // (ordered before this stream's kernels; idle records, so nothing can be claimed in them meanwhile)
            CK(cudaMemsetAsync(c->scratch_quick, 0, (size_t)c->maxcap * sizeof(ChainScratch), c->st));
        }
        scratch = c->scratch_quick;
    }
    bool lingered = false;
    while (ncap > 0) {
        const int seen = *(volatile int *)c->h_done;
        launch_plan(c->caps, c->cands, c->jobs, c->setup_list, c->job_list, c->res_list, c->cnt, ncap, o.npasses, c->st);
        if (read_counters(c)) return WSPR_ERR_CUDA;
        const Counters h = *c->h_cnt;
        if (trace) {
            float ms = 0;
            cudaEventRecord(c->ev1, c->st);
            cudaEventSynchronize(c->ev1);
            cudaEventElapsedTime(&ms, c->ev0, c->ev1);
            fprintf(stderr, "[wspr] t=%8.2f ms round %3d: setup %5d jobs %5d resolve %5d waiting %5d done %5d\n", ms, c->rounds,
                    h.nsetup, h.njobs, h.nres, h.nwait, h.ndone);
        }
        if (h.ndone == ncap) break;
        if (h.nsetup == 0 && h.njobs == 0 && h.nres == 0) {   // everything still open is parked with the Fano workers
            if (wait_parked(c, seen, std::min(h.nwait, 64), p)) return WSPR_ERR_CUDA;
            continue;
        }
        // a handful of captures ready while many more are with the pool: give those a moment to come back and plan again
        // (once), rather than running a round of near-empty grids for every capture that trickles in
        const int ready = h.nsetup + h.njobs + h.nres;
        if (!lingered && h.nwait > 0 && ready < 8 && ready * 4 < h.nwait) {
            lingered = true;
            linger_parked(c, seen, std::min(h.nwait, 16));
            continue;
        }
        lingered = false;
        c->rounds++;
        // captures entering a pass: spectrogram, candidate search, coarse sync (wsprd.c:536-678)
        launch_spectrogram(c->I, c->Q, c->psT, c->setup_list, h.nsetup, p, c->st);
        launch_candidates(c->psT, c->cands, c->caps, c->smspec, c->setup_list, h.nsetup, p, c->st);
        launch_coarse(c->psT, c->cands, c->caps, c->setup_list, h.nsetup, p, c->st);
        // one candidate of every ready capture (wsprd.c:697-766)
        int nres_max = h.nres;
        if (h.njobs > 0) {
            if (c->time_kernels) {
                while (c->kev.size() < kev_used + 2) {
                    cudaEvent_t e;
                    CK(cudaEventCreate(&e));
                    c->kev.push_back(e);
                }
                CK(cudaEventRecord(c->kev[kev_used], c->st));
            }
            launch_sync_lags(c->I, c->Q, c->jobs, c->job_list, h.njobs, c->P0, c->tabs, p, c->st);
            if (c->time_kernels) {
                CK(cudaEventRecord(c->kev[kev_used + 1], c->st));
                kev_used += 2;
                c->kev_jobs.push_back(h.njobs);
            }
            launch_sync_freqs(c->I, c->Q, c->jobs, c->job_list, h.njobs, c->P0, c->P1, c->tabs, c->att0, p, c->st);
            launch_fano_round(c->att0, c->job_list, h.njobs, p, c->st);
            launch_collect(c->jobs, c->att0, c->caps, c->job_list, h.njobs, c->res_list, c->defer_list, c->cnt, p, c->st);
            // big rounds: read the round's counters back so that the grids below are sized exactly; small ones (the
            // straggler rounds, nine tenths of all rounds) size them from the job count and let the kernels read the
            // counts on the device -- one host synchronisation per round instead of two
            int ndefer_max = h.njobs;
            nres_max = h.nres + h.njobs;
            if (h.njobs >= EXACT_SIZING_MIN_JOBS) {
                if (read_counters(c)) return WSPR_ERR_CUDA;
                ndefer_max = c->h_cnt->ndefer;
                nres_max = c->h_cnt->nres;
            }
            if (ndefer_max > 0) {                             // finish them off the critical path
                launch_deferred(c->I, c->Q, c->jobs, c->att0, c->caps, c->defer_list, ndefer_max, c->cnt, scratch, c->tabs, c->stats,
                                c->h_done, c->svc->queue, p, c->st);
                CK(cudaEventRecord(c->ev_fano, c->st));
                cudaStream_t fs = c->fano_st[c->fano_rr++ % NFANO_STREAMS];
                CK(cudaStreamWaitEvent(fs, c->ev_fano, 0));
                const int attempts = ndefer_max * (p.quickmode ? 1 : NJIT);
                launch_fano_workers(c->svc->queue, std::min(c->svc->pool, (attempts + 31) / 32), c->svc->cta_warps, false, p, fs);
                if (c->svc->pool2 > 0) {
                    cudaStream_t os = c->ovf_st[c->fano_rr % NFANO_STREAMS];
                    CK(cudaStreamWaitEvent(os, c->ev_fano, 0));
                    launch_fano_workers(c->svc->queue, std::min(c->svc->pool2, (attempts + 31) / 32), 1, true, p, os);
                }
            }
        }
        // in-order tail of the candidate loop for everything that finished, then the subtractions (wsprd.c:768-822)
        launch_resolve(c->jobs, c->caps, c->spots, c->res_list, nres_max, c->sub_list, c->cnt, p, c->st);
        if (p.subtraction)
            launch_subtract(c->I, c->Q, c->caps, c->sub_list, nres_max, c->cnt, c->phi0, c->ref, c->cprod, p, c->st);
        CK(cudaGetLastError());
    }
    launch_finish(c->caps, c->spots, c->nres, c->stats, ncap, c->st);
    CK(cudaMemcpyAsync(c->h_stats, c->stats, 8 * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CK(cudaEventRecord(c->ev1, c->st));
    if (wait_stream(c)) return WSPR_ERR_CUDA;
    CK(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
    c->deferred = c->h_stats[4];
    {
        int overflow = 0;
        CK(cudaMemcpy(&overflow, &c->svc->queue->overflow, sizeof(int), cudaMemcpyDeviceToHost));
        if (overflow) return fail(WSPR_ERR_CUDA, "Fano queue overflow: parked candidates were lost");
    }
    if (c->h_stats[3] > 0)
        return fail(WSPR_ERR_LIMIT, "callsign hash list overflow: a capture produced more distinct hashed callsigns than HASH_CAP");
    if (ht && hashtable_merge(c, *ht)) return WSPR_ERR_CUDA;
    for (size_t k = 0; k + 1 < kev_used; k += 2) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, c->kev[k], c->kev[k + 1]));
        c->sync_ms += ms;
        c->sync_launches += 1;
        c->sync_cells += (double)c->kev_jobs[k / 2] * p.nlags * NSYM;
    }
    CK(cudaGetLastError());
    return WSPR_OK;
}

extern "C" float wspr_ctx_last_decode_ms(wspr_ctx *c) { return c ? c->last_ms : 0.0f; }
extern "C" void *wspr_ctx_stream(wspr_ctx *c) { return c ? (void *)c->st : nullptr; }
extern "C" float wspr_ctx_last_sync_ms(wspr_ctx *c) { return c ? c->sync_ms : 0.0f; }
extern "C" int wspr_ctx_last_sync_launches(wspr_ctx *c) { return c ? c->sync_launches : 0; }
extern "C" double wspr_ctx_last_sync_cells(wspr_ctx *c) { return c ? c->sync_cells : 0.0; }
extern "C" int wspr_ctx_last_rounds(wspr_ctx *c) { return c ? c->rounds : 0; }
extern "C" int wspr_ctx_last_deferred(wspr_ctx *c) { return c ? c->deferred : 0; }
extern "C" int wspr_ctx_last_stats(wspr_ctx *c, int *out8) {
    if (!c || !out8) return WSPR_ERR_ARG;
    for (int i = 0; i < 8; i++) out8[i] = c->h_stats[i];
    return WSPR_OK;
}
// Fano worker pool of `device` (-1: the current one): out[0] worker warps allowed, [1] SMs set aside for them (0: shared),
// [2] housekeeping periods (256 loop trips) the worker warps were alive for, [3] lane-periods with an attempt in the lane
// (utilisation = [3] / (32 [2])), [4] attempts decoded to their end, [5] attempts skipped or abandoned, [6] worker warps
// started; reset != 0 clears the counters.  Returns 0, or a negative error when no context exists on the device yet.
extern "C" int wspr_fano_stats(int device, unsigned long long *out8, int reset) {
    if (!out8) return fail(WSPR_ERR_ARG, "wspr_fano_stats");
    if (device < 0) CK(cudaGetDevice(&device));
    FanoService *s = (device >= 0 && device < 64) ? g_svc[device] : nullptr;
    if (!s) return fail(WSPR_ERR_ARG, "wspr_fano_stats: no context on this device yet");
    CK(cudaSetDevice(device));
    FanoQueue h;
    CK(cudaMemcpy(&h, s->queue, sizeof h, cudaMemcpyDeviceToHost));
    out8[0] = (unsigned long long)s->pool;
    out8[1] = (unsigned long long)s->fano_sms;
    out8[2] = h.st_warp_periods;
    out8[3] = h.st_lane_periods;
    out8[4] = h.st_attempts;
    out8[5] = h.st_dropped;
    out8[6] = h.st_warps;
    out8[7] = h.st_ovf_warp_periods;
    if (reset) CK(cudaMemset(&s->queue->st_warp_periods, 0, 6 * sizeof(unsigned long long)));
    return WSPR_OK;
}

#ifdef WSPR_EXPERIMENTS
// experiment builds only: out16 = [kernel 0: K4, 1: LPF][worker warps on the SM: 0,1,2,3+][sum of clocks, warps]
extern "C" int wspr_debug_hist(unsigned long long *out16, int reset) {
    exp_read_hist(out16, reset);
    return WSPR_OK;
}
#endif

extern "C" int wspr_ctx_time_kernels(wspr_ctx *c, int on) {
    if (!c) return WSPR_ERR_ARG;
    c->time_kernels = on != 0;
    return WSPR_OK;
}

extern "C" int wspr_ctx_download(wspr_ctx *c, decoder_results *out, int *n_results, float *I_out, float *Q_out) {
    if (!c) return fail(WSPR_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    const int ncap = c->ncap;
    if (ncap == 0) return WSPR_OK;
    if (out) CK(cudaMemcpyAsync(out, c->spots, (size_t)ncap * MAXUNIQ * sizeof(Spot), cudaMemcpyDeviceToHost, c->st));
    if (n_results) CK(cudaMemcpyAsync(n_results, c->nres, (size_t)ncap * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    size_t w = (size_t)c->np * sizeof(float);
    if (I_out) CK(cudaMemcpy2DAsync(I_out, w, c->I, (size_t)c->stride * sizeof(float), w, ncap, cudaMemcpyDeviceToHost, c->st));
    if (Q_out) CK(cudaMemcpy2DAsync(Q_out, w, c->Q, (size_t)c->stride * sizeof(float), w, ncap, cudaMemcpyDeviceToHost, c->st));
    return wait_stream(c);
}

// ---- stage-level access ----
__global__ void k_untranspose_ps(const float *__restrict__ psT, float *__restrict__ ps, int blocks) {
    // psT[cap][b][bin] -> ps[cap][bin][b]
    int cap = blockIdx.y;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= blocks * NFFT) return;
    int bin = i / blocks, b = i - bin * blocks;
    ps[(size_t)cap * blocks * NFFT + i] = psT[((size_t)cap * blocks + b) * NFFT + bin];
}

extern "C" int wspr_ctx_spectrogram(wspr_ctx *c, float *ps_out) {
    if (!c || !ps_out) return fail(WSPR_ERR_ARG, "wspr_ctx_spectrogram");
    CK(cudaSetDevice(c->device));
    decoder_options o;
    memset(&o, 0, sizeof o);
    DecodeParams p = make_params(c, o);
    launch_spectrogram(c->I, c->Q, c->psT, c->ident, c->ncap, p, c->st);
    float *tmp = nullptr;
    size_t n = (size_t)c->ncap * c->blocks * NFFT;
    CK(dalloc(&tmp, n));
    k_untranspose_ps<<<dim3((c->blocks * NFFT + 255) / 256, c->ncap), 256, 0, c->st>>>(c->psT, tmp, c->blocks);
    CK(cudaMemcpyAsync(ps_out, tmp, n * sizeof(float), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    CK(cudaFree(tmp));
    return WSPR_OK;
}

__global__ void k_set_pass(CapState *caps, int ncap, int ipass) {
    int cap = blockIdx.x * blockDim.x + threadIdx.x;
    if (cap < ncap) caps[cap].ipass = ipass;
}
__global__ void k_gather_npk(const CapState *caps, int *npk, int ncap) {
    int cap = blockIdx.x * blockDim.x + threadIdx.x;
    if (cap < ncap) npk[cap] = caps[cap].npk;
}

extern "C" int wspr_ctx_candidates(wspr_ctx *c, int maxdrift, cand *cands, int *npk, float *smspec) {
    if (!c || !cands || !npk) return fail(WSPR_ERR_ARG, "wspr_ctx_candidates");
    CK(cudaSetDevice(c->device));
    decoder_options o;
    memset(&o, 0, sizeof o);
    DecodeParams p = make_params(c, o);
    launch_reset_caps(c->caps, c->ncap, 1, c->st);
    // the coarse search takes its drift range from the pass number: 4 in passes 0/1, 0 in pass 2 (wsprd.c:524-531)
    k_set_pass<<<(c->ncap + 127) / 128, 128, 0, c->st>>>(c->caps, c->ncap, maxdrift == 0 ? 2 : 0);
    launch_spectrogram(c->I, c->Q, c->psT, c->ident, c->ncap, p, c->st);
    CK(cudaMemsetAsync(c->cands, 0, (size_t)c->ncap * MAXCAND * sizeof(Cand), c->st));
    launch_candidates(c->psT, c->cands, c->caps, c->smspec, c->ident, c->ncap, p, c->st);
    launch_coarse(c->psT, c->cands, c->caps, c->ident, c->ncap, p, c->st);
    k_gather_npk<<<(c->ncap + 127) / 128, 128, 0, c->st>>>(c->caps, c->nres, c->ncap);
    CK(cudaMemcpyAsync(npk, c->nres, (size_t)c->ncap * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(cands, c->cands, (size_t)c->ncap * MAXCAND * sizeof(Cand), cudaMemcpyDeviceToHost, c->st));
    if (smspec)
        CK(cudaMemcpyAsync(smspec, c->smspec, (size_t)c->ncap * NSMOOTH * sizeof(float), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return WSPR_OK;
}

// ---- one-shot entry points ----
extern "C" int wspr_decode_batch(const float *I, const float *Q, int ncaptures, int samples, decoder_options options,
                                 decoder_results *out, int *n_results, int device) {
    if (ncaptures < 0 || !I || !Q) return fail(WSPR_ERR_ARG, "wspr_decode_batch: bad arguments");
    if (ncaptures == 0) return WSPR_OK;
    wspr_ctx *c = wspr_ctx_create(device, ncaptures, samples);
    if (!c) return WSPR_ERR_CUDA;
    int rc = wspr_ctx_upload(c, I, Q, ncaptures);
    if (!rc) rc = wspr_ctx_decode(c, options);
    if (!rc) rc = wspr_ctx_download(c, out, n_results, nullptr, nullptr);
    std::string keep = g_err;
    wspr_ctx_destroy(c);
    g_err = keep;
    return rc;
}

// The reference is called from one thread at a time; keep one single-capture context per process.
static wspr_ctx *g_single = nullptr;
static int g_single_np = 0;
static wspr_ctx *single_ctx(int samples) {
    if (g_single && g_single_np != samples) {
        wspr_ctx_destroy(g_single);
        g_single = nullptr;
    }
    if (!g_single) {
        g_single = wspr_ctx_create(-1, 1, samples);
        g_single_np = samples;
    }
    return g_single;
}

extern "C" int wspr_decode(float *idat, float *qdat, int samples, decoder_options options, decoder_results *decodes,
                           int *n_results) {
    if (n_results) *n_results = 0;
    g_err.clear();
    wspr_ctx *c = single_ctx(samples);
    std::vector<decoder_results> tmp(MAXUNIQ);
    int n = 0;
    int rc = c ? WSPR_OK : WSPR_ERR_CUDA;
    if (!rc) rc = wspr_ctx_upload(c, idat, qdat, 1);
    if (!rc) rc = wspr_ctx_decode(c, options);
    if (!rc) rc = wspr_ctx_download(c, tmp.data(), &n, idat, qdat);
    if (rc) {
        fprintf(stderr, "wspr_decode (libwsprd_b200): %s\n", g_err.c_str());
        return 0;                                            // the reference has no error channel (wsprd.c:854)
    }
    for (int i = 0; i < n; i++) decodes[i] = tmp[i];
    if (n_results) *n_results = n;
    return 0;
}

// Fano decoder kernel on caller-supplied soft symbols (n vectors of 162 deinterleaved bytes): the device counterpart
// of fano() (wsprd/fano.h:14-28) for batches.  solo & 1: one attempt per warp (the latency of a lone attempt); solo & 4: the
// instantiation the decode kernels run (time-out test every 256 trips, maxnp not tracked) instead of the exact one.
// Outputs are host arrays of n entries (data: n x 12 bytes); returns 0 or a negative error.
extern "C" int wspr_fano_batch(const unsigned char *symbols, int n, int delta, unsigned maxcycles, unsigned stop_after, int solo,
                               int *rc, unsigned *metric, unsigned *cycles, unsigned *maxnp, unsigned char *data,
                               unsigned long long *clocks) {
    if (n < 0 || !symbols || !rc || !metric || !cycles || !maxnp || !data) return fail(WSPR_ERR_ARG, "wspr_fano_batch");
    if (n == 0) return WSPR_OK;
    unsigned char *d_sym = nullptr, *d_data = nullptr;
    int *d_rc = nullptr;
    unsigned *d_m = nullptr, *d_c = nullptr, *d_x = nullptr;
    unsigned long long *d_k = nullptr;
    int ret = WSPR_OK;
    cudaError_t e = cudaMalloc((void **)&d_sym, (size_t)n * NSYM);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_data, (size_t)n * 12);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_rc, (size_t)n * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_m, (size_t)n * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_c, (size_t)n * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_x, (size_t)n * sizeof(unsigned));
    if (e == cudaSuccess && clocks) e = cudaMalloc((void **)&d_k, (size_t)n * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemcpy(d_sym, symbols, (size_t)n * NSYM, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        launch_fano_test(d_sym, n, delta, maxcycles, stop_after, solo, d_rc, d_m, d_c, d_x, d_data, d_k, 0);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(rc, d_rc, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(metric, d_m, (size_t)n * sizeof(unsigned), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(cycles, d_c, (size_t)n * sizeof(unsigned), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(maxnp, d_x, (size_t)n * sizeof(unsigned), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(data, d_data, (size_t)n * 12, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && clocks) e = cudaMemcpy(clocks, d_k, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) ret = fail(WSPR_ERR_CUDA, "wspr_fano_batch", e);
    cudaFree(d_sym); cudaFree(d_data); cudaFree(d_rc); cudaFree(d_m); cudaFree(d_c); cudaFree(d_x); cudaFree(d_k);
    return ret;
}

// sync_and_demodulate: correlation grid on the GPU, the handful of scalar reductions on the host in the
// reference's order (wsprd.c:216-256)
extern "C" void sync_and_demodulate(float *id, float *qd, long np, unsigned char *symbols, float *freq, int ifmin,
                                    int ifmax, float fstep, int *shift, int lagmin, int lagmax, int lagstep, float *drift,
                                    int symfac, float *sync, int mode) {
    if (mode == 0) { ifmin = 0; ifmax = 0; fstep = 0.0; }
    else if (mode == 1) { lagmin = *shift; lagmax = *shift; }
    else if (mode == 2) { lagmin = *shift; lagmax = *shift; ifmin = 0; ifmax = 0; }
    else return;
    if (lagstep <= 0) lagstep = 1;
    const int nf = ifmax - ifmin + 1, nl = (lagmax - lagmin) / lagstep + 1;
    if (nf <= 0 || nl <= 0 || np <= 0) return;
    float *dI = nullptr, *dQ = nullptr;
    float4 *dP = nullptr;
    size_t npad = ((size_t)np + 3) / 4 * 4;
    std::vector<float4> P((size_t)nf * nl * NSYM);
    cudaError_t e = cudaMalloc((void **)&dI, npad * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void **)&dQ, npad * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void **)&dP, P.size() * sizeof(float4));
    if (e == cudaSuccess) e = cudaMemcpy(dI, id, np * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dQ, qd, np * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        launch_sync_generic(dI, dQ, (int)np, *freq, ifmin, ifmax, fstep, lagmin, lagmax, lagstep, *drift, dP, 0);
        e = cudaMemcpy(P.data(), dP, P.size() * sizeof(float4), cudaMemcpyDeviceToHost);
    }
    cudaFree(dI);
    cudaFree(dQ);
    cudaFree(dP);
    if (e != cudaSuccess) {
        fprintf(stderr, "sync_and_demodulate (libwsprd_b200): %s\n", cudaGetErrorString(e));
        return;
    }
    float syncmax = -1e30, fbest = 0.0;
    int best_shift = 0;
    float fsymb[NSYM];
    for (int fi = 0; fi < nf; fi++) {
        float f0 = *freq + (ifmin + fi) * fstep;
        for (int l = 0; l < nl; l++) {
            const float4 *p = &P[((size_t)fi * nl + l) * NSYM];
            float ss = 0.0, totp = 0.0;
            for (int i = 0; i < NSYM; i++) {
                totp = totp + p[i].x + p[i].y + p[i].z + p[i].w;
                float cmet = (p[i].y + p[i].w) - (p[i].x + p[i].z);
                ss = sync_bit(i) ? ss + cmet : ss - cmet;
                if (mode == 2) fsymb[i] = sync_bit(i) ? p[i].w - p[i].y : p[i].z - p[i].x;
            }
            ss = ss / totp;
            if (ss > syncmax) {
                syncmax = ss;
                best_shift = lagmin + l * lagstep;
                fbest = f0;
            }
        }
    }
    *sync = syncmax;
    if (mode <= 1) {
        *shift = best_shift;
        *freq = fbest;
        return;
    }
    float fsum = 0.0, f2sum = 0.0;
    for (int i = 0; i < NSYM; i++) {
        fsum += fsymb[i] / NSYM;
        f2sum += fsymb[i] * fsymb[i] / NSYM;
    }
    float fac = sqrt(f2sum - fsum * fsum);
    for (int i = 0; i < NSYM; i++) {
        float v = symfac * fsymb[i] / fac;
        if (v > 127) v = 127.0;
        if (v < -128) v = -128.0;
        symbols[i] = v + 128;
    }
}

// wsprd/wsprd.h:92-98 (exported by the reference, never called by it): the per-symbol subtraction
extern "C" void subtract_signal(float *id, float *qd, long np, float f0, int shift, float drift, const unsigned char *channel_symbols) {
    wspr_ctx *c = single_ctx((int)np);
    int rc = c ? WSPR_OK : WSPR_ERR_CUDA;
    if (!rc) rc = wspr_ctx_upload(c, id, qd, 1);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(c->caps, channel_symbols, NSYM, cudaMemcpyHostToDevice, c->st);   // (scratch use of the state buffer)
        if (e == cudaSuccess) {
            launch_subtract_symbolwise(c->I, c->Q, (int)np, f0, shift, drift, reinterpret_cast<const unsigned char *>(c->caps), c->st);
            e = cudaGetLastError();
        }
        if (e != cudaSuccess) rc = fail(WSPR_ERR_CUDA, "subtract_signal", e);
    }
    if (!rc) rc = wspr_ctx_download(c, nullptr, nullptr, id, qd);
    if (rc) fprintf(stderr, "subtract_signal (libwsprd_b200): %s\n", g_err.c_str());
}

extern "C" void subtract_signal2(float *id, float *qd, long np, float f0, int shift, float drift,
                                 const unsigned char *channel_symbols) {
    wspr_ctx *c = single_ctx((int)np);
    int rc = c ? WSPR_OK : WSPR_ERR_CUDA;
    if (!rc) rc = wspr_ctx_upload(c, id, qd, 1);
    if (!rc) {
        CapState cs;
        memset(&cs, 0, sizeof cs);
        cs.sub_pending = 1;
        cs.sub_f0 = f0;
        cs.sub_shift = shift;
        cs.sub_drift = drift;
        memcpy(cs.chan, channel_symbols, NSYM);
        int zero = 0;
        decoder_options o;
        memset(&o, 0, sizeof o);
        DecodeParams p = make_params(c, o);
        Counters hc;
        memset(&hc, 0, sizeof hc);
        hc.nsub = 1;
        cudaError_t e = cudaMemcpyAsync(c->caps, &cs, sizeof cs, cudaMemcpyHostToDevice, c->st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->sub_list, &zero, sizeof(int), cudaMemcpyHostToDevice, c->st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->cnt, &hc, sizeof hc, cudaMemcpyHostToDevice, c->st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->st);
        if (e == cudaSuccess) {
            launch_subtract(c->I, c->Q, c->caps, c->sub_list, 1, c->cnt, c->phi0, c->ref, c->cprod, p, c->st);
            e = cudaGetLastError();
        }
        if (e != cudaSuccess) rc = fail(WSPR_ERR_CUDA, "subtract_signal2", e);
    }
    if (!rc) rc = wspr_ctx_download(c, nullptr, nullptr, id, qd);
    if (rc) fprintf(stderr, "subtract_signal2 (libwsprd_b200): %s\n", g_err.c_str());
}
