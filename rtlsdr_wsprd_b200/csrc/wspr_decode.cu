// Host side of libwsprd_b200.so: device-buffer ownership, the wave scheduler that drives the kernels in the
// reference's control flow (wsprd/wsprd.c:416-855), and the C ABI declared in include/wspr_b200.h.
//
// Scheduling: the reference walks the candidates of one capture serially because, in pass 0, every successful
// decode is subtracted from the samples before the next candidate is examined (wsprd.c:781-789).  Captures are
// independent, so the batch is processed in *waves*: wave r handles candidate rank r of every capture at once
// (sync search -> soft symbols -> Fano -> unpack/resolve -> subtraction).  Where no subtraction can happen (pass
// >= 1, or subtraction disabled) all ranks go into one wave and only the in-order resolve step stays sequential.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
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

struct wspr_ctx {
    int device = 0, maxcap = 0, np = 0, stride = 0, blocks = 0;
    int ncap = 0;
    int jobcap = 0, failcap = 0;
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr;
    float *I = nullptr, *Q = nullptr, *psT = nullptr, *smspec = nullptr;
    Cand *cands = nullptr;
    CapState *caps = nullptr;
    Spot *spots = nullptr;
    int *nres = nullptr;
    Job *jobs = nullptr;
    int *jobmap = nullptr, *faillist = nullptr, *sublist = nullptr;
    float4 *P0 = nullptr, *P1 = nullptr, *P2 = nullptr;
    Attempt *att0 = nullptr, *att1 = nullptr;
    float *phi0 = nullptr;
    float2 *ref = nullptr, *cprod = nullptr;
    Counters *cnt = nullptr;       // device
    Counters *h_cnt = nullptr;     // pinned host mirror
    int *h_npk = nullptr;          // pinned host copy of per-capture candidate counts
    float last_ms = 0.0f, sync_ms = 0.0f;
    int sync_launches = 0;
    double sync_cells = 0.0;
    bool time_kernels = false;
};

template <class T>
static cudaError_t dalloc(T **p, size_t n) { return cudaMalloc((void **)p, n * sizeof(T)); }

extern "C" void wspr_ctx_destroy(wspr_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    void *ptrs[] = {c->I, c->Q, c->psT, c->smspec, c->cands, c->caps, c->spots, c->nres, c->jobs, c->jobmap, c->faillist,
                    c->sublist, c->P0, c->P1, c->P2, c->att0, c->att1, c->phi0, c->ref, c->cprod, c->cnt};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (c->h_cnt) cudaFreeHost(c->h_cnt);
    if (c->h_npk) cudaFreeHost(c->h_npk);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->evk0) cudaEventDestroy(c->evk0);
    if (c->evk1) cudaEventDestroy(c->evk1);
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
    c->stride = (samples + 127) / 128 * 128;
    c->blocks = 4 * (samples / NFFT) - 1;                    // wsprd.c:516
    c->jobcap = std::max(maxcap, 256);
    c->failcap = std::max(c->jobcap / 4, 64);
    CK(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    CK(cudaEventCreate(&c->evk0));
    CK(cudaEventCreate(&c->evk1));
    size_t B = (size_t)maxcap, J = (size_t)c->jobcap, F = (size_t)c->failcap;
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
    CK(dalloc(&c->jobs, J));
    CK(dalloc(&c->jobmap, B * MAXCAND));
    CK(dalloc(&c->faillist, J));
    CK(dalloc(&c->sublist, B));
    CK(dalloc(&c->P0, J * MAXLAGS * NSYM));
    CK(dalloc(&c->P1, J * NFREQ1 * NSYM));
    CK(dalloc(&c->P2, F * (NJIT - 1) * NSYM));
    CK(dalloc(&c->att0, J));
    CK(dalloc(&c->att1, F * (NJIT - 1)));
    CK(dalloc(&c->phi0, B * NSYM));
    CK(dalloc(&c->ref, B * NSIG));
    CK(dalloc(&c->cprod, B * CPAD));
    CK(dalloc(&c->cnt, 1));
    CK(cudaMallocHost((void **)&c->h_cnt, sizeof(Counters)));
    CK(cudaMallocHost((void **)&c->h_npk, B * sizeof(int)));
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

extern "C" int wspr_ctx_normalise(wspr_ctx *c) {
    if (!c) return fail(WSPR_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    launch_normalise(c->I, c->Q, c->ncap, c->np, c->stride, c->st);
    CK(cudaGetLastError());
    return WSPR_OK;
}

static DecodeParams make_params(const wspr_ctx *c, const decoder_options &o, int ipass) {
    DecodeParams p;
    p.np = c->np;
    p.stride = c->stride;
    p.blocks = c->blocks;
    p.dialfreq = o.freq;
    p.quickmode = o.quickmode;
    p.subtraction = o.subtraction;
    p.ipass = ipass;
    p.maxdrift = (ipass == 2) ? 0 : 4;                       // wsprd.c:524-531
    p.minsync1 = 0.10;
    p.minsync2 = (ipass == 2) ? 0.10 : 0.12;
    p.symfac = 50;
    p.minrms = 52.0 * (p.symfac / 64.0);                     // wsprd.c:429
    p.delta = 60;
    p.maxcycles = 10000;
    p.lagstep = o.quickmode ? 16 : 8;                        // wsprd.c:715-717
    p.nlags = 256 / p.lagstep + 1;
    return p;
}

static int read_counters(wspr_ctx *c) {
    CK(cudaMemcpyAsync(c->h_cnt, c->cnt, sizeof(Counters), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return WSPR_OK;
}

__global__ void k_begin_pass(CapState *caps, int ncap, int ipass) {
    int cap = blockIdx.x * blockDim.x + threadIdx.x;
    if (cap >= ncap) return;
    // wsprd.c:522: no second pass for a capture whose first pass decoded nothing; mark it by an impossible count
    if (ipass >= 1 && caps[cap].uniques == 0) caps[cap].broken = 2;
}
__global__ void k_mask_done(CapState *caps, int ncap) {
    int cap = blockIdx.x * blockDim.x + threadIdx.x;
    if (cap >= ncap) return;
    if (caps[cap].uniques == 0) {
        caps[cap].npk = 0;
    }
}
__global__ void k_gather_npk(const CapState *caps, int *npk, int ncap) {
    int cap = blockIdx.x * blockDim.x + threadIdx.x;
    if (cap < ncap) npk[cap] = caps[cap].npk;
}

// one wave: candidate ranks [r0, r1) of all captures
static int run_wave(wspr_ctx *c, const DecodeParams &p, int r0, int r1) {
    CK(cudaMemsetAsync(c->cnt, 0, sizeof(Counters), c->st));
    launch_make_jobs(c->cands, c->caps, c->jobs, c->jobmap, c->cnt, c->ncap, r0, r1, c->jobcap, c->st);
    if (read_counters(c)) return WSPR_ERR_CUDA;
    int njobs = std::min(c->h_cnt->njobs, c->jobcap);
    if (njobs == 0) return WSPR_OK;
    if (c->time_kernels) CK(cudaEventRecord(c->evk0, c->st));
    launch_sync_lags(c->I, c->Q, c->jobs, njobs, c->P0, p, c->st);
    if (c->time_kernels) {
        CK(cudaEventRecord(c->evk1, c->st));
        CK(cudaEventSynchronize(c->evk1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, c->evk0, c->evk1));
        c->sync_ms += ms;
        c->sync_launches += 1;
        c->sync_cells += (double)njobs * p.nlags * NSYM;
    }
    launch_sync_freqs(c->I, c->Q, c->jobs, njobs, c->P1, c->att0, p, c->st);
    launch_fano(c->att0, njobs, p, c->st);
    launch_collect_failures(c->jobs, njobs, c->att0, c->faillist, c->cnt, c->st);
    if (!p.quickmode) {
        if (read_counters(c)) return WSPR_ERR_CUDA;
        int nfail = c->h_cnt->nfail;
        for (int f0 = 0; f0 < nfail; f0 += c->failcap) {
            int n = std::min(c->failcap, nfail - f0);
            launch_jitter(c->I, c->Q, c->jobs, c->faillist + f0, n, c->P2, c->att1, p, c->st);
            launch_fano(c->att1, n * (NJIT - 1), p, c->st);
            launch_pick_jitter(c->jobs, c->faillist + f0, n, c->att1, c->st);
        }
    }
    launch_resolve(c->jobs, c->jobmap, c->cands, c->caps, c->spots, c->sublist, c->cnt, c->ncap, r0, r1, p, c->st);
    if (p.subtraction && p.ipass == 0) {
        if (read_counters(c)) return WSPR_ERR_CUDA;
        int nsub = c->h_cnt->nsub;
        launch_subtract(c->I, c->Q, c->caps, c->sublist, nsub, c->phi0, c->ref, c->cprod, p, c->st);
    }
    CK(cudaGetLastError());
    return WSPR_OK;
}

extern "C" int wspr_ctx_decode(wspr_ctx *c, decoder_options o) {
    if (!c) return fail(WSPR_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    const int ncap = c->ncap;
    c->sync_ms = 0.0f;
    c->sync_launches = 0;
    c->sync_cells = 0.0;
    CK(cudaEventRecord(c->ev0, c->st));
    launch_reset_caps(c->caps, ncap, c->st);
    for (int ipass = 0; ipass < o.npasses && ncap > 0; ipass++) {
        DecodeParams p = make_params(c, o, ipass);
        launch_spectrogram(c->I, c->Q, c->psT, ncap, p, c->st);
        CK(cudaMemsetAsync(c->cnt, 0, sizeof(Counters), c->st));
        launch_candidates(c->psT, c->cands, c->caps, c->smspec, c->cnt, ncap, p, c->st);
        if (ipass >= 1) {                                    // wsprd.c:522 (per capture)
            k_mask_done<<<(ncap + 127) / 128, 128, 0, c->st>>>(c->caps, ncap);
        }
        k_gather_npk<<<(ncap + 127) / 128, 128, 0, c->st>>>(c->caps, c->nres, ncap);
        CK(cudaMemcpyAsync(c->h_npk, c->nres, (size_t)ncap * sizeof(int), cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        int maxnpk = 0;
        std::vector<int> hist(MAXCAND + 1, 0);               // hist[r] = captures with more than r candidates
        for (int i = 0; i < ncap; i++) {
            maxnpk = std::max(maxnpk, c->h_npk[i]);
            for (int r = 0; r < c->h_npk[i]; r++) hist[r]++;
        }
        if (maxnpk == 0) {
            if (ipass == 0) break;
            continue;
        }
        launch_coarse(c->psT, c->cands, c->caps, ncap, maxnpk, p, c->st);
        const bool serial = (o.subtraction && ipass == 0);
        int r0 = 0;
        while (r0 < maxnpk) {
            int r1 = r0 + 1;
            if (!serial) {                                   // as many ranks as fit the job buffers
                int jobs = hist[r0];
                while (r1 < maxnpk && jobs + hist[r1] <= c->jobcap) jobs += hist[r1++];
            }
            int rc = run_wave(c, p, r0, r1);
            if (rc) return rc;
            r0 = r1;
        }
    }
    launch_finish(c->caps, c->spots, c->nres, ncap, c->st);
    CK(cudaEventRecord(c->ev1, c->st));
    CK(cudaStreamSynchronize(c->st));
    CK(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
    CK(cudaGetLastError());
    return WSPR_OK;
}

extern "C" float wspr_ctx_last_decode_ms(wspr_ctx *c) { return c ? c->last_ms : 0.0f; }
extern "C" void *wspr_ctx_stream(wspr_ctx *c) { return c ? (void *)c->st : nullptr; }
extern "C" float wspr_ctx_last_sync_ms(wspr_ctx *c) { return c ? c->sync_ms : 0.0f; }
extern "C" int wspr_ctx_last_sync_launches(wspr_ctx *c) { return c ? c->sync_launches : 0; }
extern "C" double wspr_ctx_last_sync_cells(wspr_ctx *c) { return c ? c->sync_cells : 0.0; }
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
    CK(cudaStreamSynchronize(c->st));
    return WSPR_OK;
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
    DecodeParams p = make_params(c, o, 0);
    launch_spectrogram(c->I, c->Q, c->psT, c->ncap, p, c->st);
    float *tmp = nullptr;
    size_t n = (size_t)c->ncap * c->blocks * NFFT;
    CK(dalloc(&tmp, n));
    k_untranspose_ps<<<dim3((c->blocks * NFFT + 255) / 256, c->ncap), 256, 0, c->st>>>(c->psT, tmp, c->blocks);
    CK(cudaMemcpyAsync(ps_out, tmp, n * sizeof(float), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    CK(cudaFree(tmp));
    return WSPR_OK;
}

extern "C" int wspr_ctx_candidates(wspr_ctx *c, int maxdrift, cand *cands, int *npk, float *smspec) {
    if (!c || !cands || !npk) return fail(WSPR_ERR_ARG, "wspr_ctx_candidates");
    CK(cudaSetDevice(c->device));
    decoder_options o;
    memset(&o, 0, sizeof o);
    DecodeParams p = make_params(c, o, 0);
    p.maxdrift = maxdrift;
    launch_reset_caps(c->caps, c->ncap, c->st);
    launch_spectrogram(c->I, c->Q, c->psT, c->ncap, p, c->st);
    CK(cudaMemsetAsync(c->cnt, 0, sizeof(Counters), c->st));
    CK(cudaMemsetAsync(c->cands, 0, (size_t)c->ncap * MAXCAND * sizeof(Cand), c->st));
    launch_candidates(c->psT, c->cands, c->caps, c->smspec, c->cnt, c->ncap, p, c->st);
    if (read_counters(c)) return WSPR_ERR_CUDA;
    launch_coarse(c->psT, c->cands, c->caps, c->ncap, c->h_cnt->maxnpk, p, c->st);
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
        DecodeParams p = make_params(c, o, 0);
        cudaError_t e = cudaMemcpyAsync(c->caps, &cs, sizeof cs, cudaMemcpyHostToDevice, c->st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->sublist, &zero, sizeof(int), cudaMemcpyHostToDevice, c->st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->st);
        if (e == cudaSuccess) {
            launch_subtract(c->I, c->Q, c->caps, c->sublist, 1, c->phi0, c->ref, c->cprod, p, c->st);
            e = cudaGetLastError();
        }
        if (e != cudaSuccess) rc = fail(WSPR_ERR_CUDA, "subtract_signal2", e);
    }
    if (!rc) rc = wspr_ctx_download(c, nullptr, nullptr, id, qd);
    if (rc) fprintf(stderr, "subtract_signal2 (libwsprd_b200): %s\n", g_err.c_str());
}
