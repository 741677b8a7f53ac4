// CUDA kernels of the batched WSPR decode path (sm_100a).  Compiled with -fmad=false: every float expression
// below is evaluated with the same roundings, in the same order, as the reference C code it replaces, so the
// results are bit-identical to x86-64 gcc -O3 (no FMA contraction).  File:line citations are relative to the
// reference checkout (wsprd/wsprd.c unless stated).
#include "wspr_kernels.cuh"

#include <math_constants.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "fft512_twiddle.h"
#include "wspr_fano.cuh"
#include "wspr_math.cuh"
#include "wspr_mettab.h"

namespace wspr {

static std::atomic<unsigned long long> g_launches{0};   // contexts may be driven from several host threads
extern unsigned long long g_frontend_launches;   // wspr_frontend.cu
unsigned long long kernel_launch_count() { return g_launches.load() + g_frontend_launches; }
#define LAUNCHED() (g_launches.fetch_add(1, std::memory_order_relaxed))

#ifdef WSPR_EXPERIMENTS
// experiment builds only: how long do the warps of the bulk kernels run on SMs that host Fano worker warps, and does it
// matter whether a worker sits on the warp's own scheduler (SMSP = warp id % 4)?
// g_exp_hist[kernel][class][0: sum of clocks, 1: warps]; class 0: no worker on the SM, 1: worker(s) on the SM but not on this
// warp's scheduler, 2: one worker on this warp's scheduler, 3: two or more
__device__ FanoQueue *g_exp_queue;
__device__ unsigned long long g_exp_hist[2][4][2];
__device__ int g_exp_smsp[256][4];                 // worker warps per (SM, scheduler)
struct ExpTimer {
    long long t0;
    int cls, kernel;
    __device__ ExpTimer(int k) : kernel(k) {
        unsigned smid, wid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        asm("mov.u32 %0, %%warpid;" : "=r"(wid));
        smid &= 255u;
        const volatile int *v = g_exp_smsp[smid];
        const int mine = v[wid & 3u], all = v[0] + v[1] + v[2] + v[3];
        cls = mine >= 2 ? 3 : (mine == 1 ? 2 : (all > 0 ? 1 : 0));
        t0 = clock64();
    }
    __device__ void stop() {
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&g_exp_hist[kernel][cls][0], (unsigned long long)(clock64() - t0));
            atomicAdd(&g_exp_hist[kernel][cls][1], 1ull);
        }
    }
};
__device__ void exp_worker_mark(int delta) {
    unsigned smid, wid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    asm("mov.u32 %0, %%warpid;" : "=r"(wid));
    atomicAdd(&g_exp_smsp[smid & 255u][wid & 3u], delta);
}
void exp_set_queue(FanoQueue *q) { cudaMemcpyToSymbol(g_exp_queue, &q, sizeof q); }
void exp_read_hist(unsigned long long *out16, int reset) {
    cudaMemcpyFromSymbol(out16, g_exp_hist, sizeof(unsigned long long) * 16);
    if (reset) {
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(g_exp_hist, z, sizeof z);
    }
}
#define EXP_TIMER(k) ExpTimer exp_timer(k)
#define EXP_STOP() exp_timer.stop()
#define EXP_WORKER(d) do { if ((threadIdx.x & 31) == 0) exp_worker_mark(d); } while (0)
#else
#define EXP_TIMER(k)
#define EXP_STOP()
#define EXP_WORKER(d)
#endif

// ---- constant tables ----------------------------------------------------------------------------------
__device__ float c_window_g[NFFT];    // (indexed in bit-reversed order by the lanes: global/L1, not the constant bank)
__constant__ float c_lpf_w[NFILT];
__constant__ float c_lpf_psum[NFILT];
__constant__ float c_min_snr;
__constant__ float c_floor_snr;
__constant__ short c_mettab[2][256] = WSPR_METTAB_INIT;
__device__ const double g_tw[256][2] = FFT512_TWIDDLE_INIT;

void upload_tables(const HostTables &t) {
    cudaMemcpyToSymbol(c_window_g, t.window, sizeof t.window);
    cudaMemcpyToSymbol(c_lpf_w, t.lpf_w, sizeof t.lpf_w);
    cudaMemcpyToSymbol(c_lpf_psum, t.lpf_psum, sizeof t.lpf_psum);
    cudaMemcpyToSymbol(c_min_snr, &t.min_snr, sizeof(float));
    cudaMemcpyToSymbol(c_floor_snr, &t.floor_snr, sizeof(float));
}

// rate constants, written as the reference's macro expansions evaluate (wsprd.c:59-69)
__device__ __forceinline__ double twopidt() { return 2.0 * M_PI * 1.0 / 375.0; }
#define W_DF (375.0 / 256.0)
#define W_HALF_DF (375.0 / 256.0 / 2.0)

// ---- packed binary32 pairs (FFMA2) ----------------------------------------------------------------------
// The correlation and low-pass kernels are bound by the FP32 pipe AND by issue slots; sm_100 can carry two
// independent binary32 operations in one instruction (fma.rn.f32x2).  The reference rounds every product and every
// sum separately, so the packed forms below are an exact product (a*b + -0.0: the addend changes neither the value
// nor the sign of any product, zeros included) and an exact sum (a*1.0 + b).  ptxas folds fma(fma(a,b,-0),1,c) into
// fma(a,b,c) when it can see the two constants -- it even contracts mul.rn.f32x2 + add.rn.f32x2 under -fmad=false --
// so they are handed to the kernels as run-time arguments (PK_NEGZERO, PK_ONE), which it cannot fold.
// tests/test_abi_cpu.py counts the FFMA2 instructions of the built kernels to catch a toolchain that fuses anyway.
typedef unsigned long long pk2;                                  // low word = first float
constexpr pk2 PK_NEGZERO = 0x8000000080000000ull, PK_ONE = 0x3f8000003f800000ull;
__device__ __forceinline__ pk2 pk_make(float lo, float hi) {
    pk2 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ float pk_lo(pk2 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float pk_hi(pk2 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ pk2 pk_mul(pk2 a, pk2 b, pk2 negzero) {
    pk2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(negzero));
    return d;
}
__device__ __forceinline__ pk2 pk_add(pk2 a, pk2 b, pk2 one) {
    pk2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(one), "l"(b));
    return d;
}

// =========================================================================================================
// K1  spectrogram: 512-point windowed FFT every 128 samples, |X|^2, fftshift (wsprd.c:536-553).
// The transform is the binary64 radix-2 DIT graph of the oracle's FFTW stand-in (FFTW itself is an
// un-vendored dependency of the reference), so ps is bit-identical to the oracle's.
// Layout: psT[capture][block][bin] (block-major; the reference's ps[bin][block] transposed, which makes the
// stores and the per-bin block sums of K2 coalesced).
// =========================================================================================================
// Work split: 64 threads per transform, 8 points per thread in registers, three register passes of three radix-2
// stages each (exactly the butterflies of the stand-in, in its order per butterfly) with two exchanges through padded
// shared memory; two overlapping blocks per CTA share one staged input window.
constexpr int FFT_PER_CTA = 2;
constexpr int FFT_SPAN = NFFT + (FFT_PER_CTA - 1) * HOP;      // input samples covered by the CTA's blocks
struct Cplx {
    double re, im;
};
__device__ __forceinline__ void fft_bf(Cplx &u, Cplx &v, const double2 w) {   // u' = u + W v, v' = u - W v
    const double tr = w.x * v.re - w.y * v.im;
    const double ti = w.x * v.im + w.y * v.re;
    v.re = u.re - tr;
    v.im = u.im - ti;
    u.re = u.re + tr;
    u.im = u.im + ti;
}
__device__ __forceinline__ void fft_pass(Cplx v[8], const double2 t1, const double2 t2[2], const double2 t3[4]) {
#pragma unroll
    for (int p = 0; p < 8; p += 2) fft_bf(v[p], v[p + 1], t1);
#pragma unroll
    for (int p = 0; p < 8; p++)
        if ((p & 2) == 0) fft_bf(v[p], v[p + 2], t2[p & 1]);
#pragma unroll
    for (int p = 0; p < 4; p++) fft_bf(v[p], v[p + 4], t3[p]);
}
__device__ __forceinline__ int fft_pad(int idx) { return idx + (idx >> 3); }

__global__ void __launch_bounds__(64 * FFT_PER_CTA) k_spectrogram(const float *__restrict__ I, const float *__restrict__ Q,
                                                                  float *__restrict__ psT, const int *__restrict__ list,
                                                                  int stride, int blocks) {
    __shared__ double2 tw[NFFT / 2];
    __shared__ float2 xin[FFT_SPAN];
    __shared__ double s_re[FFT_PER_CTA][NFFT + NFFT / 8], s_im[FFT_PER_CTA][NFFT + NFFT / 8];
    const int cap = list[blockIdx.y], tid = threadIdx.x, f = tid >> 6, t = tid & 63;
    const int b0 = blockIdx.x * FFT_PER_CTA, b = b0 + f;
    const float *ip = I + (size_t)cap * stride + b0 * HOP;
    const float *qp = Q + (size_t)cap * stride + b0 * HOP;
    const int span = min(FFT_SPAN, NFFT + (blocks - 1 - b0) * HOP);            // (the last CTA may hold a single block)
    for (int n = tid; n < span; n += 64 * FFT_PER_CTA) xin[n] = make_float2(ip[n], qp[n]);
    for (int n = tid; n < NFFT / 2; n += 64 * FFT_PER_CTA) tw[n] = make_double2(g_tw[n][0], g_tw[n][1]);
    __syncthreads();
    const bool live = b < blocks;
    Cplx v[8];
    double *re = s_re[f], *im = s_im[f];
    if (live) {
        // pass 1: stages half = 1, 2, 4 on logical indices 8t..8t+7 (bit-reversed input order)
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int n = (int)(__brev((unsigned)(8 * t + k)) >> 23);
            const float w = c_window_g[n];
            const float2 x = xin[f * HOP + n];
            v[k].re = (double)(x.x * w);
            v[k].im = (double)(x.y * w);
        }
        const double2 t2[2] = {tw[0], tw[128]}, t3[4] = {tw[0], tw[64], tw[128], tw[192]};
        fft_pass(v, tw[0], t2, t3);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            re[fft_pad(8 * t + k)] = v[k].re;
            im[fft_pad(8 * t + k)] = v[k].im;
        }
    }
    __syncthreads();
    if (live) {
        // pass 2: stages half = 8, 16, 32 on indices hi*64 + m*8 + lo
        const int hi = t >> 3, lo = t & 7;
#pragma unroll
        for (int m = 0; m < 8; m++) {
            v[m].re = re[fft_pad(hi * 64 + m * 8 + lo)];
            v[m].im = im[fft_pad(hi * 64 + m * 8 + lo)];
        }
        const double2 t2[2] = {tw[lo * 16], tw[(8 + lo) * 16]};
        const double2 t3[4] = {tw[lo * 8], tw[(8 + lo) * 8], tw[(16 + lo) * 8], tw[(24 + lo) * 8]};
        fft_pass(v, tw[lo * 32], t2, t3);
    }
    __syncthreads();
    if (live) {
        const int hi = t >> 3, lo = t & 7;
#pragma unroll
        for (int m = 0; m < 8; m++) {
            re[fft_pad(hi * 64 + m * 8 + lo)] = v[m].re;
            im[fft_pad(hi * 64 + m * 8 + lo)] = v[m].im;
        }
    }
    __syncthreads();
    if (live) {
        // pass 3: stages half = 64, 128, 256 on indices q*64 + t
#pragma unroll
        for (int q = 0; q < 8; q++) {
            v[q].re = re[fft_pad(q * 64 + t)];
            v[q].im = im[fft_pad(q * 64 + t)];
        }
        const double2 t2[2] = {tw[t * 2], tw[(64 + t) * 2]};
        const double2 t3[4] = {tw[t], tw[64 + t], tw[128 + t], tw[192 + t]};
        fft_pass(v, tw[t * 4], t2, t3);
        float *out = psT + ((size_t)cap * blocks + b) * NFFT;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const float fr = (float)v[q].re, fi = (float)v[q].im;
            out[(q * 64 + t + NFFT / 2) & (NFFT - 1)] = fr * fr + fi * fi;
        }
    }
}

void launch_spectrogram(const float *I, const float *Q, float *psT, const int *list, int n, const DecodeParams &p,
                        cudaStream_t st) {
    if (n <= 0 || p.blocks <= 0) return;
    k_spectrogram<<<dim3((p.blocks + FFT_PER_CTA - 1) / FFT_PER_CTA, n), 64 * FFT_PER_CTA, 0, st>>>(I, Q, psT, list, p.stride,
                                                                                                   p.blocks);
    LAUNCHED();
}

// =========================================================================================================
// K2  candidate search: block sums, 7-bin boxcar, 30th-percentile noise floor, SNR normalisation, strict local
// maxima, +-110 Hz filter, stable sort by SNR (wsprd.c:555-631).  One CTA per capture.
// =========================================================================================================
__global__ void __launch_bounds__(512) k_candidates(const float *__restrict__ psT, Cand *__restrict__ cands,
                                                    CapState *__restrict__ caps, float *__restrict__ smspec_dbg,
                                                    const int *__restrict__ list, int blocks) {
    __shared__ float psavg[NFFT];
    __shared__ float sm[NSMOOTH];
    __shared__ float noise;
    const int cap = list[blockIdx.x], t = threadIdx.x;
    const float *ps = psT + (size_t)cap * blocks * NFFT;
    float acc = 0.0f;
    for (int b = 0; b < blocks; b++) acc += ps[(size_t)b * NFFT + t];     // block order, :557-561
    psavg[t] = acc;
    if (t == 0) noise = 0.0f;
    __syncthreads();
    float mine = 0.0f;
    if (t < NSMOOTH) {
        for (int j = -3; j <= 3; j++) mine += psavg[256 - 205 + t + j];   // :567-573
        sm[t] = mine;
    }
    __syncthreads();
    if (t < NSMOOTH) {                                                     // ascending rank 122, :576-583
        int rank = 0;
        for (int k = 0; k < NSMOOTH; k++) {
            float o = sm[k];
            rank += (o < mine) || (o == mine && k < t);
        }
        if (rank == 122) noise = mine;
    }
    __syncthreads();
    if (t < NSMOOTH) {                                                     // :593-596
        float v = (float)((double)(mine / noise) - 1.0);
        if (v < c_min_snr) v = c_floor_snr;
        sm[t] = v;
        if (smspec_dbg) smspec_dbg[(size_t)cap * NSMOOTH + t] = v;
    }
    __syncthreads();
    if (t == 0) {
        Cand *c = cands + (size_t)cap * MAXCAND;
        int npk = 0;
        for (int j = 1; j < NSMOOTH - 1; j++) {                            // :608-629
            if (sm[j] > sm[j - 1] && sm[j] > sm[j + 1] && npk < MAXCAND) {
                float f = (float)((j - 205) * W_HALF_DF);
                if (f >= -110.0f && f <= 110.0f) {
                    Cand x;
                    x.freq = f;
                    x.snr = (float)(10.0 * (double)glibc_log10f(sm[j]) - (double)26.3f);
                    x.shift = 0;
                    x.drift = 0.0f;
                    x.sync = 0.0f;
                    // stable insertion, descending snr (glibc qsort is a stable merge sort), :631
                    int pos = npk;
                    while (pos > 0 && c[pos - 1].snr < x.snr) {
                        c[pos] = c[pos - 1];
                        pos--;
                    }
                    c[pos] = x;
                    npk++;
                }
            }
        }
        // NB: the reference caps the *unfiltered* list at 200; with 409 bins and strict maxima at most 204 exist,
        // so the cap is only reachable on pathological spectra -- count unfiltered maxima to mirror it exactly.
        caps[cap].npk = npk;
        caps[cap].broken = 0;
        caps[cap].rank = 0;
    }
}

void launch_candidates(const float *psT, Cand *cands, CapState *caps, float *smspec_dbg, const int *list, int n,
                       const DecodeParams &p, cudaStream_t st) {
    if (n <= 0) return;
    k_candidates<<<n, 512, 0, st>>>(psT, cands, caps, smspec_dbg, list, p.blocks);
    LAUNCHED();
}

// =========================================================================================================
// K3  coarse sync on the spectrogram (wsprd.c:646-678): 3 frequency bins x 32 half-symbol offsets x drift
// hypotheses, 162 symbols each.  The reference's drift index only changes which of two bins a symbol reads
// (the unparenthesised DF macro makes the drift term ~1e-5 bins), so all negative drifts give one pattern and
// all positive drifts another; with the strict '>' the winners can only be -maxdrift, 0 or +1, which are the
// three hypotheses evaluated here (with the reference's literal index formula).
// =========================================================================================================
constexpr int COARSE_BINS = 12;
constexpr int COARSE_RANKS = 16;            // grid width; a CTA strides over the ranks of its capture
__global__ void __launch_bounds__(288) k_coarse(const float *__restrict__ psT, Cand *__restrict__ cands,
                                                const CapState *__restrict__ caps, const int *__restrict__ list, int blocks) {
    extern __shared__ float sq[];           // [blocks][COARSE_BINS] sqrt(ps)
    __shared__ float s_sync[288 / 32];
    __shared__ int s_arg[288 / 32];
    const int cap = list[blockIdx.y], t = threadIdx.x;
    const int npk = caps[cap].npk, maxdrift = pass_maxdrift(caps[cap].ipass);
  for (int rank = blockIdx.x; rank < npk; rank += COARSE_RANKS) {
    __syncthreads();                        // the previous rank's tables are no longer in use
    Cand *c = cands + (size_t)cap * MAXCAND + rank;
    const int if0 = (int)((double)c->freq / W_HALF_DF + 256);
    const int lo = if0 - 6;
    const float *ps = psT + (size_t)cap * blocks * NFFT;
    for (int i = t; i < blocks * COARSE_BINS; i += 288) {
        int b = i / COARSE_BINS, k = i - b * COARSE_BINS;
        int bin = lo + k;
        sq[i] = (bin >= 0 && bin < NFFT) ? sqrtf(ps[(size_t)b * NFFT + bin]) : 0.0f;
    }
    __syncthreads();
    const int ifr = if0 - 1 + t / 96;
    const int k0 = -10 + (t / 3) % 32;
    const int d = t % 3;
    const int idrift = (d == 0) ? -maxdrift : (d == 1 ? 0 : 1);
    const bool valid = (maxdrift > 0) || (d == 1);
    float sync = CUDART_NAN_F;
    if (valid) {
        float ss = 0.0f, pw = 0.0f;
        for (int k = 0; k < NSYM; k++) {
            // ifd = (int)(ifr + ((k - 81) / 81 * idrift) / 375.0 / 256.0), :655: the drift term is a few 1e-5 of a bin, so the
            // truncation gives ifr - 1 when the term is negative and ifr otherwise (ifr >= 100: the sum stays positive)
            const int ifd = ifr - (((idrift < 0 && k > NBITS) || (idrift > 0 && k < NBITS)) ? 1 : 0);
            int kx = k0 + 2 * k;
            if (kx < blocks) {
                int row = kx, col = ifd - lo;
                if (kx < 0) {               // the reference indexes ps[bin][kx] flat: previous bin's tail
                    row = blocks + kx;
                    col -= 1;
                }
                const float *r = sq + row * COARSE_BINS + col;
                float p0 = r[-3], p1 = r[-1], p2 = r[1], p3 = r[3];
                float m = (p1 + p3) - (p0 + p2);
                ss = sync_bit(k) ? ss + m : ss - m;
                pw = pw + p0 + p1 + p2 + p3;
            }
        }
        sync = ss / pw;
    }
    // arg-max in loop order (:668, strict '>': the first of equal maxima wins, NaN never wins): warp-level (value, index)
    // reduction, then the nine warp winners
    float bv = (sync > -1e30f) ? sync : -CUDART_INF_F;      // (NaN and values that can never win)
    int bi = t;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_down_sync(0xffffffffu, bv, o);
        const int oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) {
            bv = ov;
            bi = oi;
        }
    }
    if ((t & 31) == 0) {
        s_sync[t >> 5] = bv;
        s_arg[t >> 5] = bi;
    }
    __syncthreads();
    if (t == 0) {
        float best = -CUDART_INF_F;
        int arg = -1;
        for (int wq = 0; wq < 288 / 32; wq++)               // (warps hold ascending index ranges)
            if (s_sync[wq] > best) {
                best = s_sync[wq];
                arg = s_arg[wq];
            }
        if (arg >= 0 && best > -1e30f) {
            int bk0 = -10 + (arg / 3) % 32, bd = arg % 3;
            c->shift = 128 * (bk0 + 1);
            c->drift = (float)((bd == 0) ? -maxdrift : (bd == 1 ? 0 : 1));
            c->freq = (float)((if0 - 1 + arg / 96 - 256) * W_HALF_DF);
            c->sync = best;
        }
    }
  }
}

// the capture may start its candidate loop (separate tiny kernel: every k_coarse CTA of the capture must be done)
__global__ void k_setup_done(CapState *__restrict__ caps, const int *__restrict__ list, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) caps[list[i]].phase = PH_READY;
}

void launch_coarse(const float *psT, Cand *cands, const CapState *caps, const int *list, int n, const DecodeParams &p,
                   cudaStream_t st) {
    if (n <= 0) return;
    size_t smem = (size_t)p.blocks * COARSE_BINS * sizeof(float);
    k_coarse<<<dim3(COARSE_RANKS, n), 288, smem, st>>>(psT, cands, caps, list, p.blocks);
    LAUNCHED();
    k_setup_done<<<(n + 127) / 128, 128, 0, st>>>(const_cast<CapState *>(caps), list, n);
    LAUNCHED();
}

// =========================================================================================================
// round planning: one thread per capture advances its state machine and files it into this round's lists
// =========================================================================================================
__global__ void k_plan(CapState *__restrict__ caps, const Cand *__restrict__ cands, Job *__restrict__ jobs,
                       int *__restrict__ setup_list, int *__restrict__ job_list, int *__restrict__ res_list, Counters *cnt,
                       int ncap, int npasses) {
    int cap = blockIdx.x * blockDim.x + threadIdx.x;
    if (cap >= ncap) return;
    CapState &cs = caps[cap];
    int phase = *(volatile int *)&cs.phase;              // PH_RESOLVE is set by a side stream while rounds go on
    if (phase == PH_READY && (cs.rank >= cs.npk || cs.broken)) {
        // the candidate loop of this pass is over (wsprd.c:697, or one of its breaks :787,:793): next pass, if any.
        // wsprd.c:521-523: no second pass when the first found nothing.
        int ip = cs.ipass + 1;
        if (ip >= npasses || (ip == 1 && cs.uniques == 0)) {
            phase = PH_DONE;
        } else {
            phase = PH_SETUP;
            cs.ipass = ip;
        }
        cs.phase = phase;
    }
    if (phase == PH_SETUP) {
        setup_list[atomicAdd(&cnt->nsetup, 1)] = cap;
    } else if (phase == PH_READY) {
        const int slot = atomicAdd(&cnt->njobs, 1);
        job_list[slot] = cap;
        const Cand &c = cands[(size_t)cap * MAXCAND + cs.rank];
        Job j;
        j.cap = cap;
        j.rank = cs.rank;
        j.slot = slot;
        j.ipass = cs.ipass;
        j.freq = c.freq;
        j.drift = c.drift;
        j.shift = c.shift;
        j.sync1 = c.sync;
        j.snr = c.snr;
        j.worth = 0;
        j.fbest = 0;
        j.lbest = -1;
        j.decoded = 0;
        j.idt = 0;
        j.cycles = 0;
        for (int k = 0; k < 12; k++) j.dec[k] = 0;
        jobs[cap] = j;
    } else if (phase == PH_RESOLVE) {
        __threadfence();                                  // the job record was published before the phase flag
        res_list[atomicAdd(&cnt->nres, 1)] = cap;
    } else if (phase == PH_WAIT) {
        atomicAdd(&cnt->nwait, 1);
    } else {
        atomicAdd(&cnt->ndone, 1);
    }
}

void launch_plan(CapState *caps, const Cand *cands, Job *jobs, int *setup_list, int *job_list, int *res_list, Counters *cnt,
                 int ncap, int npasses, cudaStream_t st) {
    if (ncap <= 0) return;
    cudaMemsetAsync(cnt, 0, sizeof(Counters), st);
    k_plan<<<(ncap + 127) / 128, 128, 0, st>>>(caps, cands, jobs, setup_list, job_list, res_list, cnt, ncap, npasses);
    LAUNCHED();
}

// =========================================================================================================
// K4  sync_and_demodulate (wsprd.c:101-259): the 4-tone matched filter.
// For each (frequency hypothesis, lag, symbol) the reference accumulates, strictly in sample order,
//     i_t += I[k]*c_t[j] + Q[k]*s_t[j] ;  q_t += -I[k]*s_t[j] + Q[k]*c_t[j]       (t = 0..3 tones, j = 0..255)
// with phasor tables c_t, s_t generated by a float recurrence seeded by one cosf/sinf pair per tone.  One thread
// owns one (lag, symbol) cell and its eight running sums, so the order of additions is the reference's.
// Tables: without drift (the common case) all symbols share one table set, built once per CTA in shared memory;
// with drift every symbol has its own frequency and the thread advances private phasors in registers.
// =========================================================================================================
__device__ __forceinline__ float symbol_freq(float f0, float drift, int i) {   // :156
    return (float)((double)f0 + ((double)drift / 2.0) * (double)((float)i - (float)NBITS) / (double)(float)NBITS);
}
__device__ __forceinline__ void tone_seeds(float fp, float cd[4], float sd[4]) {   // :158-172
    const double k = twopidt();
    const float d0 = (float)(k * ((double)fp - W_DF * 1.5));
    const float d1 = (float)(k * ((double)fp - W_DF * 0.5));
    const float d2 = (float)(k * ((double)fp + W_DF * 0.5));
    const float d3 = (float)(k * ((double)fp + W_DF * 1.5));
    cd[0] = glibc_cosf(d0); sd[0] = glibc_sinf(d0);
    cd[1] = glibc_cosf(d1); sd[1] = glibc_sinf(d1);
    cd[2] = glibc_cosf(d2); sd[2] = glibc_sinf(d2);
    cd[3] = glibc_cosf(d3); sd[3] = glibc_sinf(d3);
}
struct Acc8 {
    float ai[4], aq[4];
};
__device__ __forceinline__ void acc_step(Acc8 &a, float x, float y, const float c[4], const float s[4]) {   // :200-207
#pragma unroll
    for (int t = 0; t < 4; t++) {
        a.ai[t] = a.ai[t] + x * c[t] + y * s[t];
        a.aq[t] = a.aq[t] - x * s[t] + y * c[t];
    }
}
__device__ __forceinline__ float4 acc_power(const Acc8 &a) {   // :211-214  (double sqrt of a float == correctly rounded)
    float4 p;
    p.x = __fsqrt_rn(a.ai[0] * a.ai[0] + a.aq[0] * a.aq[0]);
    p.y = __fsqrt_rn(a.ai[1] * a.ai[1] + a.aq[1] * a.aq[1]);
    p.z = __fsqrt_rn(a.ai[2] * a.ai[2] + a.aq[2] * a.aq[2]);
    p.w = __fsqrt_rn(a.ai[3] * a.ai[3] + a.aq[3] * a.aq[3]);
    return p;
}

// Drifting candidates (5 % of all): every symbol has its own frequency, so the four phasors advance in registers by the
// reference's recurrence (:181-187: c' = c*cd - s*sd, s' = c*sd + s*cd, every product and sum rounded separately).  Packed
// like the table path: tones (0,1) and (2,3) side by side; a - b is a + (-b) and s*(-sd) is -(s*sd), both exactly.
struct DriftPhasors {
    pk2 c01, s01, c23, s23, cd01, sd01, nsd01, cd23, sd23, nsd23;
};
__device__ __forceinline__ void drift_init(DriftPhasors &p, float fp) {
    float cd[4], sd[4];
    tone_seeds(fp, cd, sd);
    p.cd01 = pk_make(cd[0], cd[1]);
    p.cd23 = pk_make(cd[2], cd[3]);
    p.sd01 = pk_make(sd[0], sd[1]);
    p.sd23 = pk_make(sd[2], sd[3]);
    p.nsd01 = pk_make(-sd[0], -sd[1]);
    p.nsd23 = pk_make(-sd[2], -sd[3]);
    p.c01 = p.c23 = pk_make(1.0f, 1.0f);
    p.s01 = p.s23 = pk_make(0.0f, 0.0f);
}
__device__ __forceinline__ void drift_step(pk2 &ai01, pk2 &aq01, pk2 &ai23, pk2 &aq23, float xf, float yf, DriftPhasors &p, pk2 negzero,
                                           pk2 one) {
    const pk2 x = pk_make(xf, xf), y = pk_make(yf, yf), n = pk_make(-xf, -xf);
    ai01 = pk_add(pk_add(ai01, pk_mul(x, p.c01, negzero), one), pk_mul(y, p.s01, negzero), one);
    aq01 = pk_add(pk_add(aq01, pk_mul(n, p.s01, negzero), one), pk_mul(y, p.c01, negzero), one);
    ai23 = pk_add(pk_add(ai23, pk_mul(x, p.c23, negzero), one), pk_mul(y, p.s23, negzero), one);
    aq23 = pk_add(pk_add(aq23, pk_mul(n, p.s23, negzero), one), pk_mul(y, p.c23, negzero), one);
    const pk2 cn01 = pk_add(pk_mul(p.c01, p.cd01, negzero), pk_mul(p.s01, p.nsd01, negzero), one);
    const pk2 sn01 = pk_add(pk_mul(p.c01, p.sd01, negzero), pk_mul(p.s01, p.cd01, negzero), one);
    const pk2 cn23 = pk_add(pk_mul(p.c23, p.cd23, negzero), pk_mul(p.s23, p.nsd23, negzero), one);
    const pk2 sn23 = pk_add(pk_mul(p.c23, p.sd23, negzero), pk_mul(p.s23, p.cd23, negzero), one);
    p.c01 = cn01;
    p.s01 = sn01;
    p.c23 = cn23;
    p.s23 = sn23;
}

// ---- mode 0: all lags of a group of SYMS_PER_CTA symbols, IQ window staged in shared memory ----------------
// The window is stored transposed, sample m at [m % 8][m / 8], so that lanes holding consecutive lags (8 samples
// apart) read consecutive shared-memory words.  The cells of the group are numbered lag-fastest and dealt to the
// threads in that order (cell q = lag + nlags*symbol): 33 x 18 = 594 cells fill 18.6 of the CTA's 19 warps, and the
// window word a lane reads is q - symbol, i.e. consecutive across a warp.
// Shared phasor tables for the packed form: tabp[j] = {c0,c1,s0,s1}, tabp[256+j] = {c2,c3,s2,s3}; a thread keeps the
// sums of tones (0,1) and (2,3) side by side: AI01 += (x,x)*(c0,c1); AI01 += (y,y)*(s0,s1); AQ01 += (-x,-x)*(s0,s1);
// AQ01 += (y,y)*(c0,c1) -- per tone exactly the reference's (i + x*c) + y*s and (q - x*s) + y*c.
#ifndef WSPR_K4_SYMS
#define WSPR_K4_SYMS 18                                    // (build-time knobs for A/B measurements of the CTA shape)
#define WSPR_K4_MINB 2
#endif
constexpr int SYMS_PER_CTA = WSPR_K4_SYMS;                 // 162 = 9 x 18
constexpr int LAG_WIN = SYMS_PER_CTA * SPS + SPS;          // 4864 samples cover every lag of the group
constexpr int LAG_PITCH = LAG_WIN / 8 + 1;                 // 609
constexpr int LAG_THREADS = (MAXLAGS * SYMS_PER_CTA + 31) / 32 * 32;   // 608

// threads 0..3 run the phasor recurrences (:174-188), one tone each.  Caller synchronises.
__device__ __forceinline__ void build_tables(float fp, float4 *tab, int t) {
    if (t < 4) {
        float cd[4], sd[4];
        tone_seeds(fp, cd, sd);
        const float cdt = cd[t], sdt = sd[t];
        float *base = reinterpret_cast<float *>(tab) + (t >> 1) * (SPS * 4) + (t & 1);
        float c = 1.0f, s = 0.0f;
        for (int j = 0; j < SPS; j++) {
            base[j * 4] = c;
            base[j * 4 + 2] = s;
            float cn = c * cdt - s * sdt;
            float sn = c * sdt + s * cdt;
            c = cn;
            s = sn;
        }
    }
}

// The shared tables of the drift == 0 case, built once per (job, hypothesis) in global memory -- one thread per tone runs
// the 256-step recurrence -- instead of once per CTA behind a barrier (K4 has nine CTAs per job, k_sync_freqs one CTA per
// hypothesis whose main loop is hardly longer than the recurrence).  side = 0: the job's own frequency, slot 2 (K4);
// side = 1: the mode-1 hypotheses freq + (fi - 2) * 0.1 (:151), slots 0..4, after k_pick_lag has settled freq.
__global__ void k_tables(const Job *__restrict__ jobs, const int *__restrict__ job_list, int njobs, float4 *__restrict__ tabs,
                         int side) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int tone = tid & 3, q = tid >> 2, nh = side ? NFREQ1 : 1;
    const int jx = q / nh, fi = side ? q - jx * nh : 2;
    if (jx >= njobs) return;
    const Job &job = jobs[job_list[jx]];
    if (job.drift != 0.0f) return;                             // per-symbol frequencies: phasors advance in registers
    if (side && fi == 2 && job.lbest >= 0) return;             // that row is copied from the lag search, no table needed
    const float fstep = 0.1f;
    const float f0 = side ? job.freq + (float)(fi - 2) * fstep : job.freq;
    float cd[4], sd[4];
    tone_seeds(f0, cd, sd);
    const float cdt = cd[tone], sdt = sd[tone];
    float *base = reinterpret_cast<float *>(tabs + ((size_t)jx * NFREQ1 + fi) * (2 * SPS)) + (tone >> 1) * (SPS * 4) + (tone & 1);
    float c = 1.0f, s = 0.0f;
    for (int j = 0; j < SPS; j++) {
        base[j * 4] = c;
        base[j * 4 + 2] = s;
        float cn = c * cdt - s * sdt;
        float sn = c * sdt + s * cdt;
        c = cn;
        s = sn;
    }
}
__device__ __forceinline__ void load_tables(float4 *tab, const float4 *__restrict__ tabs, int slot, int fi, int t, int nthreads) {
    const float4 *g = tabs + ((size_t)slot * NFREQ1 + fi) * (2 * SPS);
    for (int m = t; m < 2 * SPS; m += nthreads) tab[m] = g[m];
}

__global__ void __launch_bounds__(LAG_THREADS, WSPR_K4_MINB) k_sync_lags(const float *__restrict__ I, const float *__restrict__ Q,
                                                              const Job *__restrict__ jobs, const int *__restrict__ job_list,
                                                              float4 *__restrict__ P0, const float4 *__restrict__ tabs, int np,
                                                              int stride, int lagstep, int nlags, pk2 negzero, pk2 one) {
    __shared__ float4 tab[2 * SPS];
    __shared__ float2 win[8 * LAG_PITCH];
    const Job &job = jobs[job_list[blockIdx.x]];
    const int g = blockIdx.y, t = threadIdx.x;
    const float f0 = job.freq, drift = job.drift;
    const int lagmin = job.shift - 128;
    const bool shared_tab = (drift == 0.0f);
    const float *ip = I + (size_t)job.cap * stride, *qp = Q + (size_t)job.cap * stride;
    const int base = lagmin + g * SYMS_PER_CTA * SPS;       // sample index of window element 0
    // the staging below assumes lagstep 8 or 16 (lags at multiples of 8 samples)
    for (int m = t; m < LAG_WIN; m += LAG_THREADS) {
        int k = base + m;
        float2 v = make_float2(0.0f, 0.0f);
        if (k > 0 && k < np) v = make_float2(ip[k], qp[k]);  // k > 0: the reference never reads sample 0 (:199)
        win[(m & 7) * LAG_PITCH + (m >> 3)] = v;
    }
    if (shared_tab) load_tables(tab, tabs, blockIdx.x, 2, t, LAG_THREADS);
    __syncthreads();
    EXP_TIMER(0);

    const int sym_local = t / nlags, lagidx = t - sym_local * nlags;
    const int sym = g * SYMS_PER_CTA + sym_local;
    if (sym_local >= SYMS_PER_CTA || sym >= NSYM) return;
    const int off = lagidx * lagstep + sym_local * SPS;     // window-relative start of this cell (multiple of 8)
    const float2 *wp = win + (off >> 3);
    float4 power;
    if (shared_tab) {
        const ulonglong2 *tp = reinterpret_cast<const ulonglong2 *>(tab);
        pk2 ai01 = 0, ai23 = 0, aq01 = 0, aq23 = 0;
#pragma unroll 2
        for (int j8 = 0; j8 < SPS / 8; j8++) {
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const float2 v = wp[r * LAG_PITCH + j8];
                const pk2 x = pk_make(v.x, v.x), y = pk_make(v.y, v.y), n = pk_make(-v.x, -v.x);
                const ulonglong2 w01 = tp[j8 * 8 + r], w23 = tp[SPS + j8 * 8 + r];   // .x = (c,c'), .y = (s,s')
                ai01 = pk_add(pk_add(ai01, pk_mul(x, w01.x, negzero), one), pk_mul(y, w01.y, negzero), one);
                aq01 = pk_add(pk_add(aq01, pk_mul(n, w01.y, negzero), one), pk_mul(y, w01.x, negzero), one);
                ai23 = pk_add(pk_add(ai23, pk_mul(x, w23.x, negzero), one), pk_mul(y, w23.y, negzero), one);
                aq23 = pk_add(pk_add(aq23, pk_mul(n, w23.y, negzero), one), pk_mul(y, w23.x, negzero), one);
            }
        }
        Acc8 a;
        a.ai[0] = pk_lo(ai01); a.ai[1] = pk_hi(ai01); a.ai[2] = pk_lo(ai23); a.ai[3] = pk_hi(ai23);
        a.aq[0] = pk_lo(aq01); a.aq[1] = pk_hi(aq01); a.aq[2] = pk_lo(aq23); a.aq[3] = pk_hi(aq23);
        power = acc_power(a);
    } else {
        DriftPhasors ph;
        drift_init(ph, symbol_freq(f0, drift, sym));
        pk2 ai01 = 0, ai23 = 0, aq01 = 0, aq23 = 0;
        for (int j8 = 0; j8 < SPS / 8; j8++) {
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const float2 v = wp[r * LAG_PITCH + j8];
                drift_step(ai01, aq01, ai23, aq23, v.x, v.y, ph, negzero, one);
            }
        }
        Acc8 a;
        a.ai[0] = pk_lo(ai01); a.ai[1] = pk_hi(ai01); a.ai[2] = pk_lo(ai23); a.ai[3] = pk_hi(ai23);
        a.aq[0] = pk_lo(aq01); a.aq[1] = pk_hi(aq01); a.aq[2] = pk_lo(aq23); a.aq[3] = pk_hi(aq23);
        power = acc_power(a);
    }
    P0[((size_t)blockIdx.x * NSYM + sym) * MAXLAGS + lagidx] = power;    // [job][symbol][lag]: consecutive lanes, consecutive words
    EXP_STOP();
}

// per-lag sync metric and arg-max over lags (:216-218,227-232); one warp-sized CTA per job
__global__ void __launch_bounds__(64) k_pick_lag(Job *__restrict__ jobs, const int *__restrict__ job_list,
                                                 const float4 *__restrict__ P0, int lagstep, int nlags) {
    __shared__ float s_ss[64];
    Job &job = jobs[job_list[blockIdx.x]];
    const int t = threadIdx.x;
    float v = CUDART_NAN_F;
    if (t < nlags) {
        const float4 *p = P0 + (size_t)blockIdx.x * NSYM * MAXLAGS + t;   // lane t = lag t: the lanes read consecutive words
        float ss = 0.0f, totp = 0.0f;
        for (int i = 0; i < NSYM; i++) {
            float4 q = p[(size_t)i * MAXLAGS];
            totp = totp + q.x + q.y + q.z + q.w;
            float cmet = (q.y + q.w) - (q.x + q.z);
            ss = sync_bit(i) ? ss + cmet : ss - cmet;
        }
        v = ss / totp;
    }
    s_ss[t] = v;
    __syncthreads();
    if (t == 0) {
        float best = -1e30f, fbest = 0.0f;
        int bl = 0, li = -1;
        const int lagmin = job.shift - 128;
        for (int l = 0; l < nlags; l++)
            if (s_ss[l] > best) {
                best = s_ss[l];
                bl = lagmin + l * lagstep;
                fbest = job.freq;
                li = l;
            }
        job.lbest = li;
        job.shift = bl;          // the reference returns best_shift = 0 / freq = 0 if no lag ever won
        job.freq = fbest;
        job.sync1 = best;
    }
}

void launch_sync_lags(const float *I, const float *Q, Job *jobs, const int *job_list, int njobs, float4 *P0, float4 *tabs,
                      const DecodeParams &p, cudaStream_t st) {
    if (njobs <= 0) return;
    k_tables<<<(njobs * 4 + 127) / 128, 128, 0, st>>>(jobs, job_list, njobs, tabs, 0);
    LAUNCHED();
    k_sync_lags<<<dim3(njobs, NSYM / SYMS_PER_CTA), LAG_THREADS, 0, st>>>(I, Q, jobs, job_list, P0, tabs, p.np, p.stride,
                                                                          p.lagstep, p.nlags, PK_NEGZERO, PK_ONE);
    LAUNCHED();
    k_pick_lag<<<njobs, 64, 0, st>>>(jobs, job_list, P0, p.lagstep, p.nlags);
    LAUNCHED();
}

// ---- one symbol per thread at a fixed lag (modes 1 and 2, and the generic ABI wrapper) ----------------------
__device__ __forceinline__ float4 correlate_symbol(const float *__restrict__ ip, const float *__restrict__ qp, int np,
                                                   int start, bool shared_tab, const float4 *tab, float fp, pk2 negzero,
                                                   pk2 one) {
    Acc8 a;
#pragma unroll
    for (int q = 0; q < 4; q++) a.ai[q] = a.aq[q] = 0.0f;
    const bool inside = (start > 0) && (start + SPS <= np);
    float cd[4], sd[4], c[4] = {1.0f, 1.0f, 1.0f, 1.0f}, s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (shared_tab && inside) {                                // the common case, packed like k_sync_lags
        const ulonglong2 *tp = reinterpret_cast<const ulonglong2 *>(tab);
        pk2 ai01 = 0, ai23 = 0, aq01 = 0, aq23 = 0;
        const bool aligned = (start & 3) == 0;                 // (jittered windows start at shift +- 3k: scalar loads then)
        const float4 *i4 = reinterpret_cast<const float4 *>(ip + (aligned ? start : 0)), *q4 = reinterpret_cast<const float4 *>(qp + (aligned ? start : 0));
#pragma unroll 2
        for (int j4 = 0; j4 < SPS / 4; j4++) {
            float xs[4], ys[4];
            if (aligned) {
                const float4 xi = i4[j4], xq = q4[j4];
                xs[0] = xi.x; xs[1] = xi.y; xs[2] = xi.z; xs[3] = xi.w;
                ys[0] = xq.x; ys[1] = xq.y; ys[2] = xq.z; ys[3] = xq.w;
            } else {
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    xs[r] = ip[start + j4 * 4 + r];
                    ys[r] = qp[start + j4 * 4 + r];
                }
            }
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const pk2 x = pk_make(xs[r], xs[r]), y = pk_make(ys[r], ys[r]), n = pk_make(-xs[r], -xs[r]);
                const ulonglong2 w01 = tp[j4 * 4 + r], w23 = tp[SPS + j4 * 4 + r];
                ai01 = pk_add(pk_add(ai01, pk_mul(x, w01.x, negzero), one), pk_mul(y, w01.y, negzero), one);
                aq01 = pk_add(pk_add(aq01, pk_mul(n, w01.y, negzero), one), pk_mul(y, w01.x, negzero), one);
                ai23 = pk_add(pk_add(ai23, pk_mul(x, w23.x, negzero), one), pk_mul(y, w23.y, negzero), one);
                aq23 = pk_add(pk_add(aq23, pk_mul(n, w23.y, negzero), one), pk_mul(y, w23.x, negzero), one);
            }
        }
        a.ai[0] = pk_lo(ai01); a.ai[1] = pk_hi(ai01); a.ai[2] = pk_lo(ai23); a.ai[3] = pk_hi(ai23);
        a.aq[0] = pk_lo(aq01); a.aq[1] = pk_hi(aq01); a.aq[2] = pk_lo(aq23); a.aq[3] = pk_hi(aq23);
    } else if (!shared_tab && inside) {                        // drifting candidate, whole window inside the capture: packed recurrence
        DriftPhasors ph;
        drift_init(ph, fp);
        pk2 ai01 = 0, ai23 = 0, aq01 = 0, aq23 = 0;
        const bool aligned = (start & 3) == 0;
        const float4 *i4 = reinterpret_cast<const float4 *>(ip + (aligned ? start : 0)), *q4 = reinterpret_cast<const float4 *>(qp + (aligned ? start : 0));
        for (int j4 = 0; j4 < SPS / 4; j4++) {
            float xs[4], ys[4];
            if (aligned) {
                const float4 xi = i4[j4], xq = q4[j4];
                xs[0] = xi.x; xs[1] = xi.y; xs[2] = xi.z; xs[3] = xi.w;
                ys[0] = xq.x; ys[1] = xq.y; ys[2] = xq.z; ys[3] = xq.w;
            } else {
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    xs[r] = ip[start + j4 * 4 + r];
                    ys[r] = qp[start + j4 * 4 + r];
                }
            }
#pragma unroll
            for (int r = 0; r < 4; r++) drift_step(ai01, aq01, ai23, aq23, xs[r], ys[r], ph, negzero, one);
        }
        a.ai[0] = pk_lo(ai01); a.ai[1] = pk_hi(ai01); a.ai[2] = pk_lo(ai23); a.ai[3] = pk_hi(ai23);
        a.aq[0] = pk_lo(aq01); a.aq[1] = pk_hi(aq01); a.aq[2] = pk_lo(aq23); a.aq[3] = pk_hi(aq23);
    } else {
        if (!shared_tab) tone_seeds(fp, cd, sd);
        for (int j = 0; j < SPS; j++) {
            int k = start + j;
            float x = 0.0f, y = 0.0f;
            if (k > 0 && k < np) {
                x = ip[k];
                y = qp[k];
            }
            if (shared_tab) {
                float4 w01 = tab[j], w23 = tab[SPS + j];
                float cc[4] = {w01.x, w01.y, w23.x, w23.y}, ss[4] = {w01.z, w01.w, w23.z, w23.w};
                acc_step(a, x, y, cc, ss);
            } else {
                acc_step(a, x, y, c, s);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    float cn = c[q] * cd[q] - s[q] * sd[q];
                    float sn = c[q] * sd[q] + s[q] * cd[q];
                    c[q] = cn;
                    s[q] = sn;
                }
            }
        }
    }
    return acc_power(a);
}

// mode 1: five frequencies at the best lag (:722-726)
__global__ void __launch_bounds__(192) k_sync_freqs(const float *__restrict__ I, const float *__restrict__ Q,
                                                    const Job *__restrict__ jobs, const int *__restrict__ job_list,
                                                    const float4 *__restrict__ P0, float4 *__restrict__ P1,
                                                    const float4 *__restrict__ tabs, int np, int stride, pk2 negzero, pk2 one) {
    __shared__ float4 tab[2 * SPS];
    const Job &job = jobs[job_list[blockIdx.x]];
    const int fi = blockIdx.y, t = threadIdx.x;
    if (fi == 2 && job.lbest >= 0) {       // the centre hypothesis repeats the winning cell row of the lag search exactly
        if (t < NSYM)
            P1[((size_t)blockIdx.x * NFREQ1 + fi) * NSYM + t] = P0[((size_t)blockIdx.x * NSYM + t) * MAXLAGS + job.lbest];
        return;
    }
    const bool shared_tab = (job.drift == 0.0f);
    if (shared_tab && fi != 2) return;                        // the side hypotheses of drift-free jobs: k_sync_freqs_shared
    const float fstep = 0.1f;
    const float f0 = job.freq + (float)(fi - 2) * fstep;      // :151
    if (shared_tab) load_tables(tab, tabs, blockIdx.x, fi, t, 192);
    __syncthreads();
    if (t >= NSYM) return;
    const float *ip = I + (size_t)job.cap * stride, *qp = Q + (size_t)job.cap * stride;
    float fp = shared_tab ? f0 : symbol_freq(f0, job.drift, t);
    P1[((size_t)blockIdx.x * NFREQ1 + fi) * NSYM + t] = correlate_symbol(ip, qp, np, job.shift + t * SPS, shared_tab, tab, fp, negzero, one);
}

// mode 1, drift == 0 (95 % of the candidates): the four side hypotheses of a job in ONE CTA, so that the 162 symbol windows
// are read from global memory once (in 8-sample chunks, coalesced, double-buffered in shared memory) instead of once per
// hypothesis with a 1 KB stride between lanes -- the per-hypothesis kernel above is bound by L2 traffic, not by arithmetic.
// Thread (g, s) owns hypothesis g and symbol s; the sums are the reference's, sample by sample in order.
constexpr int SF_CHUNK = 8;
constexpr int SF_PITCH = SF_CHUNK + 1;
constexpr int SF_GROUP = 192;                                  // threads per hypothesis (162 used)
constexpr int SF_THREADS = 4 * SF_GROUP;
constexpr int SF_SMEM_BYTES = 4 * 2 * SPS * 16 + 2 * NSYM * SF_PITCH * 8;
__global__ void __launch_bounds__(SF_THREADS) k_sync_freqs_shared(const float *__restrict__ I, const float *__restrict__ Q,
                                                                  const Job *__restrict__ jobs, const int *__restrict__ job_list,
                                                                  float4 *__restrict__ P1, const float4 *__restrict__ tabs, int np,
                                                                  int stride, pk2 negzero, pk2 one) {
    extern __shared__ __align__(16) unsigned char sf_smem[];
    float4 *tab = reinterpret_cast<float4 *>(sf_smem);                              // [4][2 * SPS]
    float2 *buf = reinterpret_cast<float2 *>(sf_smem + 4 * 2 * SPS * 16);           // [2][NSYM][SF_PITCH]
    const Job &job = jobs[job_list[blockIdx.x]];
    if (job.drift != 0.0f) return;                             // per-symbol frequencies: k_sync_freqs
    const int t = threadIdx.x, g = t / SF_GROUP, sym = t - g * SF_GROUP;
    const int fi = g < 2 ? g : g + 1;
    const float *ip = I + (size_t)job.cap * stride, *qp = Q + (size_t)job.cap * stride;
    const int shift = job.shift;
    for (int m = t; m < 4 * 2 * SPS; m += SF_THREADS) {
        const int gg = m / (2 * SPS), f = gg < 2 ? gg : gg + 1;
        tab[m] = tabs[((size_t)blockIdx.x * NFREQ1 + f) * (2 * SPS) + (m - gg * 2 * SPS)];
    }
    // staging: element e of a chunk = (symbol e / 8, sample e % 8)
    constexpr int PER = (NSYM * SF_CHUNK + SF_THREADS - 1) / SF_THREADS;            // 2
    float2 stage[PER];
    auto fetch = [&](int c) {
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const int e = t + u * SF_THREADS;
            float2 v = make_float2(0.0f, 0.0f);
            if (e < NSYM * SF_CHUNK) {
                const int k = shift + (e / SF_CHUNK) * SPS + c * SF_CHUNK + (e % SF_CHUNK);
                if (k > 0 && k < np) v = make_float2(ip[k], qp[k]);                 // (:199)
            }
            stage[u] = v;
        }
    };
    auto park = [&](int b) {
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const int e = t + u * SF_THREADS;
            if (e < NSYM * SF_CHUNK) buf[(b * NSYM + e / SF_CHUNK) * SF_PITCH + (e % SF_CHUNK)] = stage[u];
        }
    };
    fetch(0);
    park(0);
    __syncthreads();
    const ulonglong2 *tp = reinterpret_cast<const ulonglong2 *>(tab + g * 2 * SPS);
    pk2 ai01 = 0, ai23 = 0, aq01 = 0, aq23 = 0;
    for (int c = 0; c < SPS / SF_CHUNK; c++) {
        if (c + 1 < SPS / SF_CHUNK) fetch(c + 1);
        if (sym < NSYM) {
            const float2 *bp = buf + ((c & 1) * NSYM + sym) * SF_PITCH;
#pragma unroll
            for (int r = 0; r < SF_CHUNK; r++) {
                const float2 v = bp[r];
                const pk2 x = pk_make(v.x, v.x), y = pk_make(v.y, v.y), n = pk_make(-v.x, -v.x);
                const ulonglong2 w01 = tp[c * SF_CHUNK + r], w23 = tp[SPS + c * SF_CHUNK + r];
                ai01 = pk_add(pk_add(ai01, pk_mul(x, w01.x, negzero), one), pk_mul(y, w01.y, negzero), one);
                aq01 = pk_add(pk_add(aq01, pk_mul(n, w01.y, negzero), one), pk_mul(y, w01.x, negzero), one);
                ai23 = pk_add(pk_add(ai23, pk_mul(x, w23.x, negzero), one), pk_mul(y, w23.y, negzero), one);
                aq23 = pk_add(pk_add(aq23, pk_mul(n, w23.y, negzero), one), pk_mul(y, w23.x, negzero), one);
            }
        }
        if (c + 1 < SPS / SF_CHUNK) park((c + 1) & 1);
        __syncthreads();
    }
    if (sym < NSYM) {
        Acc8 a;
        a.ai[0] = pk_lo(ai01); a.ai[1] = pk_hi(ai01); a.ai[2] = pk_lo(ai23); a.ai[3] = pk_hi(ai23);
        a.aq[0] = pk_lo(aq01); a.aq[1] = pk_hi(aq01); a.aq[2] = pk_lo(aq23); a.aq[3] = pk_hi(aq23);
        P1[((size_t)blockIdx.x * NFREQ1 + fi) * NSYM + sym] = acc_power(a);
    }
}

// soft symbols from the four tone magnitudes of one (frequency, lag) (:216-225,243-256) followed by the caller's
// rms gate and the deinterleaver (:751-759).  Warp-cooperative: all 32 lanes call it; everything that is a per-symbol
// expression runs across the lanes, every float accumulation of the reference (the two 162-term sync sums, the two moment
// sums, the rms sum) is one lane adding in the reference's order.  `sync` is the metric ss/totp of these sums when the
// caller already has it (the frequency search computes it), else it is computed here.  Returns the mode-2 sync value.
struct SoftScratch {                                           // shared memory, one per calling warp
    float fs[NSYM];
    float y2[NSYM];
    unsigned char u8[NSYM + 2];
    float bc[2];
};
__device__ float sync_metric(const float4 *__restrict__ p) {  // :216-218,227, one lane
    float ss = 0.0f, totp = 0.0f;
    for (int i = 0; i < NSYM; i++) {
        float4 q = p[i];
        totp = totp + q.x + q.y + q.z + q.w;
        float cmet = (q.y + q.w) - (q.x + q.z);
        ss = sync_bit(i) ? ss + cmet : ss - cmet;
    }
    return ss / totp;
}
__device__ float soft_symbols_warp(const float4 *__restrict__ p, unsigned char *sym_out, float *rms_out, int symfac,
                                   bool have_sync, float sync, SoftScratch &sc) {
    const unsigned lane = threadIdx.x & 31u;
    for (int i = lane; i < NSYM; i += 32) {
        float4 q = p[i];
        sc.fs[i] = sync_bit(i) ? q.w - q.y : q.z - q.x;
    }
    __syncwarp();
    if (lane == 0) {
        float ss = have_sync ? sync : sync_metric(p);
        float syncmax = -1e30f;
        if (ss > syncmax) syncmax = ss;
        float fsum = 0.0f, f2sum = 0.0f;
        for (int i = 0; i < NSYM; i++) {
            const float f = sc.fs[i];
            fsum += f / (float)NSYM;
            f2sum += f * f / (float)NSYM;
        }
        sc.bc[0] = __fsqrt_rn(f2sum - fsum * fsum);
        sc.bc[1] = syncmax;
    }
    __syncwarp();
    const float fac = sc.bc[0];
    for (int i = lane; i < NSYM; i += 32) {
        float v = (float)symfac * sc.fs[i] / fac;
        if (v > 127.0f) v = 127.0f;
        if (v < -128.0f) v = -128.0f;
        float w = v + 128.0f;
        unsigned char u = (w >= 0.0f && w < 256.0f) ? (unsigned char)(int)w : (unsigned char)0;   // NaN -> 0 like cvttss2si's low byte
        sc.u8[i] = u;
        float y = (float)((double)(float)u - 128.0);
        sc.y2[i] = y * y;
    }
    __syncwarp();
    if (lane == 0) {
        float sq = 0.0f;
        for (int i = 0; i < NSYM; i++) sq += sc.y2[i];
        *rms_out = sqrtf(sq / (float)NSYM);
    }
    // deinterleave: p-th output is input at the p-th 8-bit-reversed index below 162 (wsprd_utils.c:196-213); of the values
    // v < 256 whose reversal is < 162, the ones below v are counted with a prefix over the warp
    int base = 0;
    for (int v0 = 0; v0 < 256; v0 += 32) {
        const int v = v0 + (int)lane, r = bitrev8(v);
        const unsigned m = __ballot_sync(0xffffffffu, r < NSYM);
        if (r < NSYM) sym_out[base + __popc(m & ((1u << lane) - 1u))] = sc.u8[r];
        base += __popc(m);
    }
    __syncwarp();
    return sc.bc[1];
}

// arg-max over the five frequencies, the minsync1 gate, and the jitter-0 soft symbols, which are exactly the sums
// of the winning hypothesis (mode 2 at the same frequency and lag repeats them) -- one warp per job: five lanes run the
// five 162-term sync sums side by side
constexpr int PICK_WARPS = 4;
__global__ void __launch_bounds__(32 * PICK_WARPS) k_pick_freq(Job *__restrict__ jobs, const int *__restrict__ job_list, int njobs,
                                                               const float4 *__restrict__ P1, Attempt *__restrict__ att,
                                                               float minsync1, float minrms, int symfac) {
    __shared__ SoftScratch scratch[PICK_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int jx = blockIdx.x * PICK_WARPS + warp;
    if (jx >= njobs) return;
    const int cap = job_list[jx];
    Job &job = jobs[cap];
    const float minsync2 = pass_minsync2(job.ipass);
    float mine = 0.0f;
    if (lane < NFREQ1) mine = sync_metric(P1 + ((size_t)jx * NFREQ1 + lane) * NSYM);
    float best = -1e30f, fbest = 0.0f;
    int bf = -1, bshift = 0;
    for (int fi = 0; fi < NFREQ1; fi++) {
        const float ss = __shfl_sync(0xffffffffu, mine, fi);
        if (ss > best) {
            best = ss;
            bf = fi;
            fbest = job.freq + (float)(fi - 2) * 0.1f;
            bshift = job.shift;
        }
    }
    const bool worth = best > minsync1;
    Attempt &a = att[cap];
    float s2 = 0.0f, rms = 0.0f;
    if (worth && bf >= 0) {
        float r = 0.0f;
        s2 = soft_symbols_warp(P1 + ((size_t)jx * NFREQ1 + bf) * NSYM, a.sym, &r, symfac, true, best, scratch[warp]);
        r = __shfl_sync(0xffffffffu, r, 0);
        rms = r;
    }
    if (lane == 0) {
        a.cap = cap;
        a.idt = 0;
        a.ok = 0;
        a.unfinished = 0;
        a.cycles = 0;
        a.sync2 = s2;
        a.gate = (worth && bf >= 0) ? ((s2 > minsync2) && (rms > minrms)) : 0;
        job.freq = fbest;
        job.shift = bshift;
        job.sync1 = best;
        job.fbest = bf;
        job.worth = worth;
    }
}

void launch_sync_freqs(const float *I, const float *Q, Job *jobs, const int *job_list, int njobs, const float4 *P0, float4 *P1,
                       float4 *tabs, Attempt *att0, const DecodeParams &p, cudaStream_t st) {
    if (njobs <= 0) return;
    k_tables<<<(njobs * NFREQ1 * 4 + 127) / 128, 128, 0, st>>>(jobs, job_list, njobs, tabs, 1);
    LAUNCHED();
    k_sync_freqs<<<dim3(njobs, NFREQ1), 192, 0, st>>>(I, Q, jobs, job_list, P0, P1, tabs, p.np, p.stride, PK_NEGZERO, PK_ONE);
    LAUNCHED();
    k_sync_freqs_shared<<<njobs, SF_THREADS, SF_SMEM_BYTES, st>>>(I, Q, jobs, job_list, P1, tabs, p.np, p.stride, PK_NEGZERO, PK_ONE);
    LAUNCHED();
    k_pick_freq<<<(njobs + PICK_WARPS - 1) / PICK_WARPS, 32 * PICK_WARPS, 0, st>>>(jobs, job_list, njobs, P1, att0, p.minsync1, p.minrms,
                                                                                   p.symfac);
    LAUNCHED();
}

// =========================================================================================================
// K5  Fano decoder (fano.c:87-238), metric table in constant memory, lane-dense branch-free form (wspr_fano.cuh).
//   k_fano_round    jitter-0 attempts of a round, 32 attempts per warp, at most fano_budget cycles each: nearly every
//                   decodable candidate finishes within a few hundred cycles; the rest is parked
//   k_fano_workers  the pool of worker warps that serves the device-wide queue of parked attempts (below)
// =========================================================================================================
__global__ void __launch_bounds__(32) k_fano_round(Attempt *__restrict__ att0, const int *__restrict__ job_list, int njobs,
                                                   int delta, unsigned maxcycles, unsigned budget) {
    extern __shared__ __align__(16) unsigned char fano_smem[];
    const int i = blockIdx.x * 32 + threadIdx.x;
    Attempt *a = (i < njobs) ? &att0[job_list[i]] : nullptr;
    const bool want = a != nullptr && a->gate;
    FanoOneShot feed{want ? a->sym : nullptr, budget, {}};
    fano_run<false>(feed, FanoSmem::at(fano_smem), &c_mettab[0][0], delta, maxcycles);
    if (want) {
        const FanoResult &r = feed.res;
        a->ok = (r.rc == 0);
        a->unfinished = (r.rc == FANO_STOPPED);
        a->cycles = r.cycles;
        for (int k = 0; k < 12; k++) a->dec[k] = r.data[k];
    }
}

void launch_fano_round(Attempt *att0, const int *job_list, int njobs, const DecodeParams &p, cudaStream_t st) {
    if (njobs <= 0) return;
    k_fano_round<<<(njobs + 31) / 32, 32, FANO_WARP_SMEM_BYTES, st>>>(att0, job_list, njobs, p.delta, p.maxcycles, p.fano_budget);
    LAUNCHED();
}

// triage after the jitter-0 attempt: finished candidates go to this round's resolve list; candidates that are worth
// a try but still undecided (unfinished Fano run, or not decoded and the jitter search is still to come, :741-766)
// are parked and handed to the Fano workers.
__global__ void k_collect(Job *__restrict__ jobs, const Attempt *__restrict__ att0, CapState *__restrict__ caps,
                          const int *__restrict__ job_list, int njobs, int *__restrict__ res_list,
                          int *__restrict__ defer_list, Counters *cnt, int quickmode) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= njobs) return;
    const int cap = job_list[i];
    Job &job = jobs[cap];
    const Attempt &a = att0[cap];
    bool defer = false;
    if (a.gate && !a.unfinished) job.cycles = a.cycles;      // `cycles` keeps the last completed decoder call's count
    if (a.gate && a.ok) {
        job.decoded = 1;
        job.idt = 0;
        for (int k = 0; k < 12; k++) job.dec[k] = a.dec[k];
    } else if (job.worth && (a.unfinished || !quickmode)) {
        defer = true;
    }
    if (defer) {
        caps[cap].phase = PH_WAIT;
        defer_list[atomicAdd(&cnt->ndefer, 1)] = cap;
    } else {
        res_list[atomicAdd(&cnt->nres, 1)] = cap;
    }
}

void launch_collect(Job *jobs, const Attempt *att0, CapState *caps, const int *job_list, int njobs, int *res_list,
                    int *defer_list, Counters *cnt, const DecodeParams &p, cudaStream_t st) {
    if (njobs <= 0) return;
    k_collect<<<(njobs + 127) / 128, 128, 0, st>>>(jobs, att0, caps, job_list, njobs, res_list, defer_list, cnt, p.quickmode);
    LAUNCHED();
}

// ---- parked candidates -------------------------------------------------------------------------------------------
// What is left of a parked candidate's jitter loop (wsprd.c:741-766): attempt 0 is the unfinished jitter-0 Fano run (now
// with the reference's full cycle budget), attempts 1..42 are the jittered ones, shift + 3*(+-1..21).
//   k_jitter_soft   grid (candidate, jitter): the jittered soft-symbol vectors and their gates (mode 2 of
//                   sync_and_demodulate + the rms test), written to the capture's scratch record;
//   k_fano_enqueue  the round's candidates go into the device-wide queue (wspr_kernels.cuh, FanoQueue);
//   k_fano_workers  a pool of one-warp CTAs, every LANE of which takes an attempt, decodes it, publishes the result and
//                   takes the next one (the lanes of a warp work on unrelated attempts; wspr_fano.cuh).  Attempt 0 of every
//                   queued candidate is handed out before any jittered attempt -- what the reference would run first is
//                   served first.  An attempt is skipped or abandoned as soon as a lower-numbered attempt of its candidate
//                   has decoded (the reference's sequential loop would never have reached it), and the lane that decodes
//                   retires the candidate's unclaimed attempts itself.  Winner = lowest successful attempt, exactly what
//                   the sequential loop returns; the lane that accounts for a candidate's last attempt writes the outcome
//                   into the job and hands the capture back to the rounds.
// Round 1 gave every parked candidate its own 1.5 warps for as long as its slowest attempt ran (~1 200 one-warp CTAs of
// 84 KB per 4096-capture batch, a third of the lanes idle, jittered attempts started whether or not attempt 0 was about to
// succeed); the pool keeps its lanes full, runs attempt 0 first and is a fixed, small number of warps.
__global__ void __launch_bounds__(192) k_jitter_soft(const float *__restrict__ I, const float *__restrict__ Q,
                                                     Job *__restrict__ jobs, const Attempt *__restrict__ att0,
                                                     CapState *__restrict__ caps, const int *__restrict__ defer_list,
                                                     const Counters *__restrict__ cnt, ChainScratch *__restrict__ scratch,
                                                     const float4 *__restrict__ tabs,
                                                     int *__restrict__ stats, int *__restrict__ host_done, int nattempts,
                                                     int np, int stride, float minrms, int symfac, pk2 negzero, pk2 one) {
    __shared__ float4 tab[2 * SPS];
    __shared__ float4 P[NSYM];
    __shared__ SoftScratch soft;
    const int idt = blockIdx.y, t = threadIdx.x;
    const int n = cnt->ndefer;                                 // (the grid is sized from an upper bound; see launch_deferred)
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    __syncthreads();                                           // the previous candidate's tables are no longer in use
    const int cap = defer_list[e];
    const Job &job = jobs[cap];
    ChainScratch &cs = scratch[cap];
    if (idt == 0) {                                            // attempt 0: the parked jitter-0 soft symbols
        const Attempt &a = att0[cap];
        for (int i = t; i < NSYM; i += 192) cs.sym[0][i] = a.sym[i];
        if (t == 0) {
            cs.gate[0] = a.gate && a.unfinished;             // (else it already ran to its end inside the round, or is gated off)
            cs.ok[0] = cs.unfinished[0] = 0;
            cs.best = NJIT;                                    // no attempt has decoded yet
            cs.done = 0;
            cs.nattempts = nattempts;
            cs.job = &jobs[cap];
            cs.phase = &caps[cap].phase;
            cs.stats = stats;
            cs.host_done = host_done;
        }
        continue;
    }
    int ii = (idt + 1) / 2;
    if (idt % 2 == 1) ii = -ii;
    ii = 3 * ii;
    // the phasor table of the job's final frequency is the one the frequency search of this round built for the winning
    // hypothesis (same expression for the frequency, same recurrence): still in the round's table buffer
    const bool shared_tab = (job.drift == 0.0f);
    if (shared_tab) load_tables(tab, tabs, job.slot, job.fbest, t, 192);
    __syncthreads();
    if (t < NSYM) {
        const float *ip = I + (size_t)cap * stride, *qp = Q + (size_t)cap * stride;
        const float fp = shared_tab ? job.freq : symbol_freq(job.freq, job.drift, t);
        P[t] = correlate_symbol(ip, qp, np, job.shift + ii + t * SPS, shared_tab, tab, fp, negzero, one);
    }
    __syncthreads();
    if (t < 32) {
        float rms = 0.0f;
        const float s2 = soft_symbols_warp(P, cs.sym[idt], &rms, symfac, false, 0.0f, soft);
        if (t == 0) {
            cs.gate[idt] = (s2 > pass_minsync2(job.ipass)) && (rms > minrms);
            cs.ok[idt] = cs.unfinished[idt] = 0;
        }
    }
  }
}

__global__ void __launch_bounds__(256) k_fano_enqueue(FanoQueue *__restrict__ q, ChainScratch *__restrict__ scratch,
                                                      const int *__restrict__ defer_list, const Counters *__restrict__ cnt,
                                                      int *__restrict__ stats) {
    __shared__ unsigned s_base;
    const int n = cnt->ndefer;
    if (n <= 0) return;
    if (threadIdx.x == 0) {
        atomicAdd(stats + 4, n);                               // candidates parked during this decode
        const unsigned base = atomicAdd(&q->tail, (unsigned)n);
        const unsigned h0 = *(volatile unsigned *)&q->head0, h1 = *(volatile unsigned *)&q->head1;
        const unsigned oldest = (int)(h0 - h1) < 0 ? h0 : h1;
        if (base + (unsigned)n - oldest > q->mask + 1u) q->overflow = 1;   // (then entries are overwritten: the host reports it)
        s_base = base;
    }
    __syncthreads();
    const unsigned base = s_base;
    for (unsigned i = threadIdx.x; i < (unsigned)n; i += blockDim.x) {
        ChainScratch *cs = scratch + defer_list[i];
        *(volatile int *)&cs->next = 1;                        // (everything k_jitter_soft wrote is complete: previous kernel)
        FanoQueueEntry &x = q->ring[(base + i) & q->mask];
        *(ChainScratch *volatile *)&x.cs = cs;
        __threadfence();
        *(volatile unsigned *)&x.seq = base + i + 1u;
    }
}

// `count` attempts of the capture's parked candidate are accounted for; the lane that accounts for the last one leaves the
// outcome in the job and hands the capture back
__device__ void fano_settle(ChainScratch *cs, int count) {
    __threadfence();
    const int arrived = atomicAdd(&cs->done, count);
    const int na = *(volatile int *)&cs->nattempts;
    if (arrived + count != na) return;
    __threadfence();                                           // the other attempts' results are complete and visible
    Job &job = *cs->job;
    const volatile ChainScratch &v = *cs;
    for (int y = 0; y < na; y++) {
        if (v.gate[y] && !v.unfinished[y]) job.cycles = v.cycles[y];
        if (v.gate[y] && v.ok[y]) {
            job.decoded = 1;
            job.idt = y;
            for (int k = 0; k < 12; k++) job.dec[k] = v.dec[y][k];
            break;
        }
    }
    atomicAdd(cs->stats + (job.decoded ? (job.idt == 0 ? 0 : 1) : 2), 1);
    __threadfence();
    *(volatile int *)cs->phase = PH_RESOLVE;
    __threadfence_system();
    atomicAdd_system(cs->host_done, 1);                        // (mapped host memory: wakes a host thread with nothing else to do)
}

struct FanoQueueFeed {
    FanoQueue *q;
    ChainScratch *cs;
    int idt;
    bool overflow;                                             // an overflow worker: takes work only while a backlog exists
    unsigned periods, busy_periods, attempts, dropped;         // statistics of this lane
    __device__ bool backlog() const {
        return (int)(*(volatile unsigned *)&q->tail - *(volatile unsigned *)&q->head1) >= *(volatile int *)&q->ovf_backlog;
    }
    __device__ void period(bool active) {
        periods++;
        busy_periods += active ? 1u : 0u;
    }
    __device__ const unsigned char *next(unsigned &stop_after) {
        stop_after = 0;
        for (;;) {
            if (overflow && !backlog()) return nullptr;
            const unsigned tail = *(volatile unsigned *)&q->tail;
            unsigned h = *(volatile unsigned *)&q->head0;
            // (tail was read BEFORE head0: producers and other lanes may have moved both on in between, so head0 can be
            // past the tail read here -- `!=` would then pop an entry that does not exist yet and the lane, with its whole
            // warp, would wait for a candidate that may never come; tools/fano_queue_host_check.cpp found exactly that)
            if ((int)(tail - h) > 0) {                         // attempt 0 of the next candidate nobody has started
                if (atomicCAS(&q->head0, h, h + 1u) != h) continue;
                FanoQueueEntry &x = q->ring[h & q->mask];
                while (*(volatile unsigned *)&x.seq != h + 1u) __nanosleep(100);   // reserved, being written by a running kernel
                __threadfence();
                cs = *(ChainScratch *volatile *)&x.cs;
                idt = 0;
            } else {                                           // a jittered attempt of the oldest candidate that has any left
                h = *(volatile unsigned *)&q->head1;
                if (h == tail) return nullptr;
                FanoQueueEntry &x = q->ring[h & q->mask];
                if (*(volatile unsigned *)&x.seq != h + 1u) continue;       // not written yet, or the cursor has moved on
                __threadfence();
                ChainScratch *c = *(ChainScratch *volatile *)&x.cs;
                __threadfence();
                if (*(volatile unsigned *)&x.seq != h + 1u) continue;
                const int na = *(volatile int *)&c->nattempts;
                const int k = *(volatile int *)&c->next >= na ? na : atomicAdd(&c->next, 1);
                if (k >= na) {                                 // nothing left to claim there: move the cursor on
                    atomicCAS(&q->head1, h, h + 1u);
                    continue;
                }
                cs = c;
                idt = k;
            }
            if (cs->gate[idt] && *(volatile int *)&cs->best > idt) return cs->sym[idt];
            cs->unfinished[idt] = 1;                           // never run: gated off, or a lower-numbered attempt has decoded
            dropped++;
            fano_settle(cs, 1);
        }
    }
    __device__ bool abandon() const { return *(volatile int *)&cs->best < idt; }
    __device__ void finish(const FanoResult &r) {
        cs->ok[idt] = (r.rc == 0);
        cs->unfinished[idt] = (r.rc == FANO_STOPPED);
        cs->cycles[idt] = r.cycles;
        for (int k = 0; k < 12; k++) cs->dec[idt][k] = r.data[k];
        if (r.rc == FANO_STOPPED) dropped++;
        else attempts++;
        int count = 1;
        if (r.rc == 0) {
            atomicMin(&cs->best, idt);
            // the attempts nobody has claimed yet all come after this one: the reference's loop never reaches them
            const int na = cs->nattempts, k = atomicAdd(&cs->next, NJIT + 1);
            for (int y = k; y < na; y++) cs->unfinished[y] = 1;
            if (k < na) count += na - k;
        }
        fano_settle(cs, count);
    }
};

// The pool: at most q->pool worker warps are alive at any time.  A warp that finds the pool complete leaves at once; a warp
// whose lanes are all idle with the queue empty gives its place up FIRST and looks at the queue once more afterwards (a
// producer may have added work, and the warps launched for that work may have found the pool complete and left), taking
// the place back if there is something to do -- so work is never stranded, and nobody ever waits for work.
// (The place on the SM -- sm_workers, the per-SM cap -- is given up only at the very end, a few hundred clocks after the last
// look at the queue: a warp launched for new work that lands on this SM inside that window finds it full and leaves.  The
// launch of such a warp follows the enqueue by microseconds and a launch spreads its CTAs over all SMs, so not ALL of them
// can be turned away that way in practice; tools/fano_queue_host_check.cpp strands candidates only when the exit is
// artificially slowed by 300 us on a one-SM model.  Giving the SM place up before the look, like the pool place, closes the
// window for good; it was not done in round 2 because it re-allocates the registers of the whole kernel and the change
// could no longer be run on a GPU.)
// A CTA carries blockDim.x / 32 worker warps, completely independent of each other (no CTA barrier): with more than one the
// warps of a CTA sit on different schedulers of their SM, which single-warp CTAs leave to chance.
__global__ void __launch_bounds__(128) k_fano_workers(FanoQueue *__restrict__ q, int delta, unsigned maxcycles, int overflow) {
    extern __shared__ __align__(16) unsigned char fano_smem_all[];
    unsigned char *fano_smem = fano_smem_all + (size_t)(threadIdx.x >> 5) * FANO_WARP_SMEM_BYTES;
    const unsigned lane = threadIdx.x & 31u;
    unsigned smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    smid &= 255u;
    int *active = overflow ? &q->active2 : &q->active;
    const volatile int *pool = overflow ? &q->pool2 : &q->pool;
    FanoQueueFeed feed{q, nullptr, 0, overflow != 0, 0u, 0u, 0u, 0u};
    const int per_sm = *(volatile int *)&q->per_sm;
    int mine = 0;
    if (lane == 0) {
        mine = !overflow || feed.backlog();
        if (mine && per_sm > 0) {                              // (this SM has its share of workers: leave)
            mine = atomicAdd(&q->sm_workers[smid], 1) < per_sm;
            if (!mine) atomicSub(&q->sm_workers[smid], 1);
        }
        if (mine) {
            mine = atomicAdd(active, 1) < *pool;
            if (!mine) {
                atomicSub(active, 1);
                if (per_sm > 0) atomicSub(&q->sm_workers[smid], 1);
            }
        }
    }
    mine = __shfl_sync(0xffffffffu, mine, 0);
    if (!mine) return;
    EXP_WORKER(1);
    while (mine) {
        fano_run<false>(feed, FanoSmem::at(fano_smem), &c_mettab[0][0], delta, maxcycles);
        mine = 0;
        if (lane == 0) {
            atomicSub(active, 1);
            __threadfence();
            const unsigned tail = *(volatile unsigned *)&q->tail;
            const bool pending = overflow ? feed.backlog() : (*(volatile unsigned *)&q->head0 != tail || *(volatile unsigned *)&q->head1 != tail);
            if (pending) {
                mine = atomicAdd(active, 1) < *pool;
                if (!mine) atomicSub(active, 1);
            }
        }
        mine = __shfl_sync(0xffffffffu, mine, 0);
    }
    const unsigned busy = __reduce_add_sync(0xffffffffu, feed.busy_periods), att = __reduce_add_sync(0xffffffffu, feed.attempts),
                   drop = __reduce_add_sync(0xffffffffu, feed.dropped);
    EXP_WORKER(-1);
    if (lane == 0) {
        if (per_sm > 0) atomicSub(&q->sm_workers[smid], 1);
        atomicAdd(&q->st_warp_periods, (unsigned long long)feed.periods);
        if (overflow) atomicAdd(&q->st_ovf_warp_periods, (unsigned long long)feed.periods);
        atomicAdd(&q->st_lane_periods, (unsigned long long)busy);
        atomicAdd(&q->st_attempts, (unsigned long long)att);
        atomicAdd(&q->st_dropped, (unsigned long long)drop);
        atomicAdd(&q->st_warps, 1ull);
    }
}

// n_max: upper bound of the candidates parked by this round (the exact number is cnt->ndefer, on the device)
void launch_deferred(const float *I, const float *Q, Job *jobs, const Attempt *att0, CapState *caps, const int *defer_list, int n_max,
                     const Counters *cnt, ChainScratch *scratch, const float4 *tabs, int *stats, int *host_done, FanoQueue *queue,
                     const DecodeParams &p, cudaStream_t st) {
    if (n_max <= 0) return;
    const int nattempts = p.quickmode ? 1 : NJIT;
    k_jitter_soft<<<dim3(std::min(n_max, 256), nattempts), 192, 0, st>>>(I, Q, jobs, att0, caps, defer_list, cnt, scratch, tabs, stats,
                                                                         host_done, nattempts, p.np, p.stride, p.minrms, p.symfac,
                                                                         PK_NEGZERO, PK_ONE);
    LAUNCHED();
    k_fano_enqueue<<<1, 256, 0, st>>>(queue, scratch, defer_list, cnt, stats);
    LAUNCHED();
}

int fano_warp_smem_bytes() { return FANO_WARP_SMEM_BYTES; }
void launch_fano_workers(FanoQueue *queue, int nwarps, int cta_warps, bool overflow, const DecodeParams &p, cudaStream_t st) {
    if (nwarps <= 0) return;
    cta_warps = cta_warps == 2 ? 2 : (cta_warps == 4 ? 4 : 1);
    unsigned maxcycles = p.maxcycles;
#ifdef WSPR_EXPERIMENTS
    // experiment builds only (make exp -> libwsprd_b200_exp.so): cut the long Fano runs short -- WRONG results -- to measure
    // what the bulk of the decode costs without them
    static const unsigned dbg = [] { const char *e = getenv("WSPR_DEBUG_CHAIN_MAXCYCLES"); return e ? (unsigned)atoi(e) : 0u; }();
    if (dbg) maxcycles = dbg;
#endif
    k_fano_workers<<<(nwarps + cta_warps - 1) / cta_warps, 32 * cta_warps, (size_t)cta_warps * FANO_WARP_SMEM_BYTES, st>>>(
        queue, p.delta, maxcycles, overflow ? 1 : 0);
    LAUNCHED();
}

// the Fano kernel on caller-supplied soft symbols (wspr_fano_batch): solo & 1: one attempt per warp (latency of a lone
// attempt), solo & 4: the decode instantiation instead of the exact one
__global__ void __launch_bounds__(32) k_fano_test(const unsigned char *__restrict__ symbols, int n, int delta,
                                                  unsigned maxcycles, unsigned stop_after, int solo, int *__restrict__ rc,
                                                  unsigned *__restrict__ metric, unsigned *__restrict__ cycles,
                                                  unsigned *__restrict__ maxnp, unsigned char *__restrict__ data,
                                                  unsigned long long *__restrict__ clocks) {
    extern __shared__ __align__(16) unsigned char fano_smem[];
    const bool fast = (solo & 4) != 0;
    solo &= 1;
    const int i = solo ? (int)blockIdx.x : (int)(blockIdx.x * 32 + threadIdx.x);
    const bool want = i < n && (!solo || threadIdx.x == 0);
    const long long t0 = clock64();
    FanoOneShot feed{want ? symbols + (size_t)i * NSYM : nullptr, stop_after, {}};
    if (fast) fano_run<false>(feed, FanoSmem::at(fano_smem), &c_mettab[0][0], delta, maxcycles);   // what the decode kernels run
    else fano_run<true>(feed, FanoSmem::at(fano_smem), &c_mettab[0][0], delta, maxcycles);
    if (!want) return;
    const FanoResult &r = feed.res;
    if (clocks) clocks[i] = (unsigned long long)(clock64() - t0);
    rc[i] = r.rc;
    metric[i] = r.metric;
    cycles[i] = r.cycles;
    maxnp[i] = r.maxnp;
    for (int k = 0; k < 12; k++) data[(size_t)i * 12 + k] = r.data[k];
}
void launch_fano_test(const unsigned char *symbols, int n, int delta, unsigned maxcycles, unsigned stop_after, int solo, int *rc,
                      unsigned *metric, unsigned *cycles, unsigned *maxnp, unsigned char *data, unsigned long long *clocks,
                      cudaStream_t st) {
    if (n <= 0) return;
    const int blocks = (solo & 1) ? n : (n + 31) / 32;
    k_fano_test<<<blocks, 32, FANO_WARP_SMEM_BYTES, st>>>(symbols, n, delta, maxcycles, stop_after, solo & 5, rc, metric, cycles,
                                                         maxnp, data, clocks);
    LAUNCHED();
}

// =========================================================================================================
// resolve: the per-capture, in-order tail of the candidate loop (wsprd.c:768-822): unpack the message, decide on
// subtraction, apply the two `break`s, drop duplicates, append the spot.  One thread per capture of the round's
// resolve list handles that capture's current candidate, then releases the capture for the next round.
// =========================================================================================================
__global__ void k_resolve(Job *__restrict__ jobs, CapState *__restrict__ caps, Spot *__restrict__ spots,
                          const int *__restrict__ res_list, int *__restrict__ sublist, Counters *cnt, DecodeParams p) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= cnt->nres) return;
    const int cap = res_list[e];
    CapState &cs = caps[cap];
    cs.sub_pending = 0;
    ListHashStore hs{cs.hash, &cs.nhash, HASH_CAP, p.preload, &cs.hash_overflow};
    const Job &job = jobs[cap];
    for (int once = 0; once < 1; once++) {                    // (the reference's `continue` / `break` targets)
        if (!(job.worth && job.decoded)) continue;
        signed char message[12];
        for (int i = 0; i < 11; i++) message[i] = (signed char)job.dec[i];
        message[11] = 0;
        char callsign[CALL_LEN], call_loc_pow[23], call[CALL_LEN], loc[7], pwr[3];
        for (int i = 0; i < CALL_LEN; i++) callsign[i] = call[i] = 0;
        for (int i = 0; i < 23; i++) call_loc_pow[i] = 0;
        for (int i = 0; i < 7; i++) loc[i] = 0;
        for (int i = 0; i < 3; i++) pwr[i] = 0;
        int noprint = unpack_message(message, hs, call_loc_pow, call, loc, pwr, callsign);
        if (p.subtraction && job.ipass == 0 && !noprint) {
            if (channel_symbols(call_loc_pow, hs, cs.chan)) {
                cs.sub_pending = 1;
                cs.sub_f0 = job.freq;
                cs.sub_shift = job.shift;
                cs.sub_drift = job.drift;
                sublist[atomicAdd(&cnt->nsub, 1)] = cap;
            } else {
                cs.broken = 1;
                break;
            }
        }
        if (s_eq(loc, "A000AA")) {
            cs.broken = 1;
            break;
        }
        bool dupe = false;
        for (int i = 0; i < cs.uniques; i++)
            if (s_eq(callsign, cs.allcalls[i]) && fabs((double)(job.freq - cs.allfreqs[i])) < 3.0) dupe = true;
        if (dupe || cs.uniques >= MAXUNIQ) continue;
        s_copy(cs.allcalls[cs.uniques], CALL_LEN, callsign);
        cs.allfreqs[cs.uniques] = job.freq;
        Spot &r = spots[(size_t)cap * MAXUNIQ + cs.uniques];
        cs.uniques++;
        int idt = job.idt, ii = (idt + 1) / 2;
        if (idt % 2 == 1) ii = -ii;
        ii = 3 * ii;
        double dialfreq = (double)p.dialfreq / 1e6;
        r.freq = dialfreq + (1500.0 + (double)job.freq) / 1e6;
        r.sync = job.sync1;
        r.snr = job.snr;
        r.dt = (float)((double)job.shift * 1.0 / 375.0 - 2.0);
        r.drift = job.drift;
        r.jitter = ii;
        r.cycles = (int)job.cycles;
        for (int i = 0; i < 23; i++) r.message[i] = 0;
        for (int i = 0; i < 13; i++) r.call[i] = 0;
        for (int i = 0; i < 7; i++) r.loc[i] = 0;
        for (int i = 0; i < 3; i++) r.pwr[i] = 0;
        s_copy(r.message, 23, call_loc_pow);
        s_copy(r.call, 13, call);
        s_copy(r.loc, 7, loc);
        s_copy(r.pwr, 3, pwr);
    }
    cs.rank = job.rank + 1;
    cs.phase = PH_READY;
}

void launch_resolve(Job *jobs, CapState *caps, Spot *spots, const int *res_list, int nres_max, int *sub_list, Counters *cnt,
                    const DecodeParams &p, cudaStream_t st) {
    if (nres_max <= 0) return;
    k_resolve<<<(nres_max + 31) / 32, 32, 0, st>>>(jobs, caps, spots, res_list, sub_list, cnt, p);
    LAUNCHED();
}

// =========================================================================================================
// K6  subtract_signal2 (wsprd.c:316-413)
//   (a) the reference phase is a float running sum over all 41 472 samples; one thread per job replays the
//       additions and records the phase every PHI_SEG samples;
//   (b) per sample: at most PHI_SEG - 1 more additions from the recorded phase of its segment, cos/sin (glibc-faithful),
//       s(t)*conj(r(t)) into a zero-padded buffer;
//   (c) 360-tap low-pass (each output a sequential 360-term sum, four consecutive outputs per thread with a
//       sliding register window), edge renormalisation, subtraction in place.
// =========================================================================================================
__device__ __forceinline__ float sub_dphi(float f0, float drift, int i, unsigned char sym) {   // :341-343
    float cs = (float)sym;
    return (float)(twopidt() * ((double)f0 +
                                ((double)drift / 2.0) * ((double)(float)i - (double)(float)NSYM / 2.0) / ((double)(float)NSYM / 2.0) +
                                ((double)cs - 1.5) * 375.0 / 256.0));
}

__global__ void k_sub_phase(const CapState *__restrict__ caps, const int *__restrict__ sublist, const Counters *cnt,
                            float *__restrict__ phi_seg) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= cnt->nsub) return;
    const CapState &cs = caps[sublist[s]];
    float *out = phi_seg + (size_t)s * (NSIG / PHI_SEG);
    float phi = 0.0f;
    for (int i = 0; i < NSYM; i++) {
        const float dphi = sub_dphi(cs.sub_f0, cs.sub_drift, i, cs.chan[i]);
        for (int j8 = 0; j8 < SPS / PHI_SEG; j8++) {
            out[i * (SPS / PHI_SEG) + j8] = phi;
#pragma unroll
            for (int j = 0; j < PHI_SEG; j++) phi = phi + dphi;
        }
    }
}

// Four consecutive samples per thread, four symbols per CTA: the chain of dependent loads that every thread starts with
// (list entry -> capture state -> recorded phase) and the rate set-up are paid once per four samples, and the four
// sincos evaluations of a thread overlap.
constexpr int SUBREF_SYMS = 4;                                 // symbols per CTA (64 threads each)
constexpr int SUBREF_CTAS = (NSYM + SUBREF_SYMS - 1) / SUBREF_SYMS + 1;   // + one CTA that zeroes the pads of the product buffer
__global__ void __launch_bounds__(SPS) k_sub_ref(const float *__restrict__ I, const float *__restrict__ Q,
                                                 const CapState *__restrict__ caps, const int *__restrict__ sublist,
                                                 const Counters *cnt, const float *__restrict__ phi_seg,
                                                 float2 *__restrict__ ref, float2 *__restrict__ cprod, int np, int stride) {
    const int s = blockIdx.x, t = threadIdx.x;
    if (s >= cnt->nsub) return;
    if (blockIdx.y == SUBREF_CTAS - 1) {                       // pads: [0, 360) and [360 + NSIG, CPAD)
        float2 *c = cprod + (size_t)s * CPAD;
        constexpr int tailn = CPAD - (NFILT + NSIG);
        for (int z = t; z < NFILT + tailn; z += SPS) c[z < NFILT ? z : NSIG + z] = make_float2(0.0f, 0.0f);
        return;
    }
    const int i = blockIdx.y * SUBREF_SYMS + (t >> 6), j0 = (t & 63) * 4;
    if (i >= NSYM) return;
    const int cap = sublist[s];
    const CapState &cs = caps[cap];
    const int ii = i * SPS + j0, k0 = cs.sub_shift + ii;
    // the phase of sample ii: the recorded phase of its segment, then the reference's additions up to ii
    float phi = phi_seg[(size_t)s * (NSIG / PHI_SEG) + ii / PHI_SEG];
    const float dphi = sub_dphi(cs.sub_f0, cs.sub_drift, i, cs.chan[i]);
    static_assert(PHI_SEG == 8, "a thread's four samples start at offset 0 or 4 of a recorded segment");
    const bool second_half = (j0 & 4) != 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {                              // (uniform instruction stream: lanes differ only in the select)
        const float nxt = phi + dphi;
        phi = second_half ? nxt : phi;
    }
    const float *ip = I + (size_t)cap * stride, *qp = Q + (size_t)cap * stride;
    float xs[4], ys[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int k = k0 + u;
        const bool in = k > 0 && k < np;                       // :375-381
        xs[u] = in ? ip[k] : 0.0f;
        ys[u] = in ? qp[k] : 0.0f;
    }
    float2 r4[4], c4[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        float rc, rs;
        glibc_sincosf(phi, &rs, &rc);
        phi = phi + dphi;
        const int k = k0 + u;
        float2 c = make_float2(0.0f, 0.0f);
        if (k > 0 && k < np) {
            c.x = xs[u] * rc + ys[u] * rs;
            c.y = ys[u] * rc - xs[u] * rs;
        }
        r4[u] = make_float2(rc, rs);
        c4[u] = c;
    }
    float4 *rp = reinterpret_cast<float4 *>(ref + (size_t)s * NSIG + ii);
    float4 *cp = reinterpret_cast<float4 *>(cprod + (size_t)s * CPAD + NFILT + ii);
    rp[0] = make_float4(r4[0].x, r4[0].y, r4[1].x, r4[1].y);
    rp[1] = make_float4(r4[2].x, r4[2].y, r4[3].x, r4[3].y);
    cp[0] = make_float4(c4[0].x, c4[0].y, c4[1].x, c4[1].y);
    cp[1] = make_float4(c4[2].x, c4[2].y, c4[3].x, c4[3].y);
}

constexpr int LPF_R = 4;                                     // consecutive outputs per thread
#ifndef WSPR_LPF_THREADS
#define WSPR_LPF_THREADS 256                                 // (build-time knob for A/B measurements of the CTA shape)
#endif

// LPF_THREADS = 256: 1024 outputs per CTA.  (One-warp CTAs of 128 outputs were tried, to let a scheduler that also hosts a
// long-running Fano warp simply take fewer of them: 3 % slower end to end -- 3.8x the staging traffic and 8x the CTAs.)
template <int LPF_THREADS>
__global__ void __launch_bounds__(LPF_THREADS) k_sub_lpf(float *__restrict__ I, float *__restrict__ Q,
                                                         const CapState *__restrict__ caps, const int *__restrict__ sublist,
                                                         const Counters *cnt, const float2 *__restrict__ ref,
                                                         const float2 *__restrict__ cprod, int np, int stride, pk2 negzero,
                                                         pk2 one) {
    constexpr int LPF_TILE = LPF_THREADS * LPF_R;            // outputs per CTA
    constexpr int LPF_SPAN = LPF_TILE + NFILT;               // inputs per tile (multiple of 4)
    constexpr int LPF_PITCH = LPF_SPAN / LPF_R + 1;
    __shared__ __align__(8) float2 sc[LPF_R * LPF_PITCH];
    const int s = blockIdx.x, tile = blockIdx.y, t = threadIdx.x;
    if (s >= cnt->nsub) return;
    const int cap = sublist[s];
    const CapState &cs = caps[cap];
    const int i0 = tile * LPF_TILE;                          // first output (signal sample index) of the tile
    // output i needs cprod[i + NFILT/2 + tap], tap = 0..359 (cf index i+360, window starts 180 earlier)
    const float2 *src = cprod + (size_t)s * CPAD + i0 + NFILT / 2;
    for (int m = t; m < LPF_SPAN; m += LPF_THREADS) {
        int g = i0 + NFILT / 2 + m;
        sc[(m % LPF_R) * LPF_PITCH + m / LPF_R] = (g < CPAD) ? src[m] : make_float2(0.0f, 0.0f);
    }
    __syncthreads();
    EXP_TIMER(1);
    pk2 acc[LPF_R];                                          // (i, q) sums side by side (packed pairs, see pk_mul/pk_add)
    pk2 w[LPF_R];                                            // sliding window: inputs 4t+tap .. 4t+tap+3
    const pk2 *scp = reinterpret_cast<const pk2 *>(sc);
#pragma unroll
    for (int r = 0; r < LPF_R; r++) {
        acc[r] = 0;
        w[r] = scp[r * LPF_PITCH + t];
    }
    for (int tap = 0; tap < NFILT; tap += LPF_R) {
#pragma unroll
        for (int u = 0; u < LPF_R; u++) {
            const float wt = c_lpf_w[tap + u];
            const pk2 wt2 = pk_make(wt, wt);
            // at tap+u the window holds inputs (4t + tap+u + r); rotate by u
#pragma unroll
            for (int r = 0; r < LPF_R; r++)                   // :388-389  i += w*c.x ; q += w*c.y
                acc[r] = pk_add(acc[r], pk_mul(wt2, w[(u + r) % LPF_R], negzero), one);
            // slot u now leaves the window; refill it with input 4t + tap+u + 4
            int m = LPF_R * t + tap + u + LPF_R;
            w[u] = scp[(m % LPF_R) * LPF_PITCH + m / LPF_R];
        }
    }
    float ai[LPF_R], aq[LPF_R];
#pragma unroll
    for (int r = 0; r < LPF_R; r++) {
        ai[r] = pk_lo(acc[r]);
        aq[r] = pk_hi(acc[r]);
    }
    EXP_STOP();
#pragma unroll
    for (int r = 0; r < LPF_R; r++) {                         // :397-410
        int i = i0 + LPF_R * t + r;
        if (i >= NSIG) continue;
        float norm;
        if (i < NFILT / 2) norm = c_lpf_psum[NFILT / 2 + i];
        else if (i > NSIG - 1 - NFILT / 2) norm = c_lpf_psum[NFILT / 2 + NSIG - 1 - i];
        else norm = 1.0f;
        int k = cs.sub_shift + i;
        if (k > 0 && k < np) {
            float2 rr = ref[(size_t)s * NSIG + i];
            size_t g = (size_t)cap * stride + k;
            I[g] = I[g] - (ai[r] * rr.x - aq[r] * rr.y) / norm;
            Q[g] = Q[g] - (ai[r] * rr.y + aq[r] * rr.x) / norm;
        }
    }
}

// grids are sized for nsub_max entries; the actual count is read from cnt->nsub on the device
void launch_subtract(float *I, float *Q, const CapState *caps, const int *sublist, int nsub_max, const Counters *cnt,
                     float *phi0, float2 *ref, float2 *cprod, const DecodeParams &p, cudaStream_t st) {
    if (nsub_max <= 0) return;
    k_sub_phase<<<(nsub_max + 31) / 32, 32, 0, st>>>(caps, sublist, cnt, phi0);
    LAUNCHED();
    k_sub_ref<<<dim3(nsub_max, SUBREF_CTAS), SPS, 0, st>>>(I, Q, caps, sublist, cnt, phi0, ref, cprod, p.np, p.stride);
    LAUNCHED();
    constexpr int T = WSPR_LPF_THREADS;
    k_sub_lpf<T><<<dim3(nsub_max, (NSIG + 4 * T - 1) / (4 * T)), T, 0, st>>>(I, Q, caps, sublist, cnt, ref, cprod, p.np, p.stride,
                                                                             PK_NEGZERO, PK_ONE);
    LAUNCHED();
}

// =========================================================================================================
// bookkeeping kernels
// =========================================================================================================
__global__ void k_reset_caps(CapState *caps, int ncap, int npasses) {
    int cap = blockIdx.x * blockDim.x + threadIdx.x;
    if (cap >= ncap) return;
    CapState &cs = caps[cap];
    cs.phase = npasses > 0 ? PH_SETUP : PH_DONE;
    cs.ipass = 0;
    cs.rank = 0;
    cs.npk = 0;
    cs.uniques = 0;
    cs.broken = 0;
    cs.nhash = 0;
    cs.sub_pending = 0;
    cs.hash_overflow = 0;
}
void launch_reset_caps(CapState *caps, int ncap, int npasses, cudaStream_t st) {
    if (ncap <= 0) return;
    k_reset_caps<<<(ncap + 127) / 128, 128, 0, st>>>(caps, ncap, npasses);
    LAUNCHED();
}

// final stable sort by SNR, descending (wsprd.c:827) and the spot count
__global__ void k_finish(CapState *__restrict__ caps, Spot *__restrict__ spots, int *__restrict__ nres, int *__restrict__ stats,
                         int ncap) {
    int cap = blockIdx.x * blockDim.x + threadIdx.x;
    if (cap >= ncap) return;
    if (caps[cap].hash_overflow) atomicAdd(stats + 3, 1);
    Spot *r = spots + (size_t)cap * MAXUNIQ;
    int n = caps[cap].uniques;
    for (int i = 1; i < n; i++) {
        Spot x = r[i];
        int j = i;
        while (j > 0 && r[j - 1].snr < x.snr) {
            r[j] = r[j - 1];
            j--;
        }
        r[j] = x;
    }
    nres[cap] = n;
}
void launch_finish(CapState *caps, Spot *spots, int *nres, int *stats, int ncap, cudaStream_t st) {
    if (ncap <= 0) return;
    k_finish<<<(ncap + 31) / 32, 32, 0, st>>>(caps, spots, nres, stats, ncap);
    LAUNCHED();
}

// peak normalisation of the hand-off (rtlsdr_wsprd.c:291-305): scale = (float)(0.5 / max(|I|,|Q|, 1e-24))
__global__ void __launch_bounds__(256) k_normalise(float *__restrict__ I, float *__restrict__ Q, int n, int stride) {
    __shared__ float red[256];
    const int cap = blockIdx.x, t = threadIdx.x;
    float *ip = I + (size_t)cap * stride, *qp = Q + (size_t)cap * stride;
    float m = 1e-24f;
    for (int i = t; i < n; i += 256) {
        float a = fabsf(ip[i]), b = fabsf(qp[i]);
        if (a > m) m = a;
        if (b > m) m = b;
    }
    red[t] = m;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (t < s && red[t + s] > red[t]) red[t] = red[t + s];
        __syncthreads();
    }
    const float scale = (float)(0.5 / (double)red[0]);
    for (int i = t; i < n; i += 256) {
        ip[i] *= scale;
        qp[i] *= scale;
    }
}
void launch_normalise(float *I, float *Q, int ncap, int n, int stride, cudaStream_t st) {
    if (ncap <= 0 || n <= 0) return;
    k_normalise<<<ncap, 256, 0, st>>>(I, Q, n, stride);
    LAUNCHED();
}

// =========================================================================================================
// generic sync_and_demodulate correlation grid for the reference-ABI wrapper: P[(f*nlags + l)*162 + sym]
// =========================================================================================================
__global__ void __launch_bounds__(192) k_sync_generic(const float *__restrict__ I, const float *__restrict__ Q, int np,
                                                      float freq, int ifmin, float fstep, int lagmin, int lagstep,
                                                      float drift, float4 *__restrict__ P, pk2 negzero, pk2 one) {
    __shared__ float4 tab[2 * SPS];
    const int l = blockIdx.x, fi = blockIdx.y, t = threadIdx.x;
    const float f0 = freq + (float)(ifmin + fi) * fstep;
    const bool shared_tab = (drift == 0.0f);
    if (shared_tab) build_tables(f0, tab, t);
    __syncthreads();
    if (t >= NSYM) return;
    float fp = shared_tab ? f0 : symbol_freq(f0, drift, t);
    P[((size_t)fi * gridDim.x + l) * NSYM + t] = correlate_symbol(I, Q, np, lagmin + l * lagstep + t * SPS, shared_tab, tab, fp, negzero, one);
}
void launch_sync_generic(const float *I, const float *Q, int np, float freq, int ifmin, int ifmax, float fstep, int lagmin,
                         int lagmax, int lagstep, float drift, float4 *P, cudaStream_t st) {
    int nf = ifmax - ifmin + 1, nl = (lagmax - lagmin) / lagstep + 1;
    if (nf <= 0 || nl <= 0) return;
    k_sync_generic<<<dim3(nl, nf), 192, 0, st>>>(I, Q, np, freq, ifmin, fstep, lagmin, lagstep, drift, P, PK_NEGZERO, PK_ONE);
    LAUNCHED();
}

// =========================================================================================================
// subtract_signal (wsprd.c:263-312): the per-symbol variant the reference exports but never calls.  One thread per symbol
// (the symbols touch disjoint samples): phasor recurrence, the two 256-term sums in sample order, subtraction.
// =========================================================================================================
__global__ void __launch_bounds__(192) k_sub_symbolwise(float *__restrict__ I, float *__restrict__ Q, int np, float f0, int shift,
                                                        float drift, const unsigned char *__restrict__ chan) {
    const int i = threadIdx.x;
    if (i >= NSYM) return;
    const float fp = symbol_freq(f0, drift, i);                                             // :274
    const float dphi = (float)(twopidt() * ((double)fp + ((double)(float)chan[i] - 1.5) * 375.0 / 256.0));   // :276
    const float cd = glibc_cosf(dphi), sd = glibc_sinf(dphi);
    float i0 = 0.0f, q0 = 0.0f, c = 1.0f, s = 0.0f;
    for (int j = 0; j < SPS; j++) {
        const int k = shift + i * SPS + j;
        if (k > 0 && k < np) {
            i0 = i0 + I[k] * c + Q[k] * s;
            q0 = q0 - I[k] * s + Q[k] * c;
        }
        const float cn = c * cd - s * sd, sn = c * sd + s * cd;
        c = cn;
        s = sn;
    }
    i0 = i0 / (float)SPS;
    q0 = q0 / (float)SPS;
    c = 1.0f;
    s = 0.0f;
    for (int j = 0; j < SPS; j++) {
        const int k = shift + i * SPS + j;
        if (k > 0 && k < np) {
            I[k] = I[k] - (i0 * c - q0 * s);
            Q[k] = Q[k] - (q0 * c + i0 * s);
        }
        const float cn = c * cd - s * sd, sn = c * sd + s * cd;
        c = cn;
        s = sn;
    }
}
void launch_subtract_symbolwise(float *I, float *Q, int np, float f0, int shift, float drift, const unsigned char *chan, cudaStream_t st) {
    k_sub_symbolwise<<<1, 192, 0, st>>>(I, Q, np, f0, shift, drift, chan);
    LAUNCHED();
}

// Opt in to > 48 KB of dynamic shared memory (k_sync_freqs_shared).  Function attributes belong to the device that is
// current, so this runs once per device (wspr_decode.cu calls it when it sets a device up).
// carveout_kb > 0: every decode kernel asks for the SAME shared-memory carve-out.  An SM cannot change its L1/shared split
// while a CTA is resident, and the Fano worker warps are resident practically all the time: with the driver's per-kernel
// choice an SM stays frozen at whatever split it had when a worker landed on it, and kernels that need more shared memory
// than that split leaves (K4: 2 x 47 KB) can only use the SM partly or not at all.  With one common split nothing ever has
// to wait for an SM to drain.
void init_kernel_attributes(int carveout_kb) {
    cudaFuncSetAttribute(k_sync_freqs_shared, cudaFuncAttributeMaxDynamicSharedMemorySize, SF_SMEM_BYTES);
    cudaFuncSetAttribute(k_fano_workers, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * FANO_WARP_SMEM_BYTES);
    if (carveout_kb <= 0) return;
    const int pct = std::min(100, (carveout_kb * 100 + 227) / 228);
#define WSPR_CARVE(k) cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, pct)
    WSPR_CARVE(k_spectrogram); WSPR_CARVE(k_candidates); WSPR_CARVE(k_coarse); WSPR_CARVE(k_setup_done); WSPR_CARVE(k_plan);
    WSPR_CARVE(k_tables); WSPR_CARVE(k_sync_lags); WSPR_CARVE(k_pick_lag); WSPR_CARVE(k_sync_freqs); WSPR_CARVE(k_sync_freqs_shared);
    WSPR_CARVE(k_pick_freq); WSPR_CARVE(k_fano_round); WSPR_CARVE(k_collect); WSPR_CARVE(k_jitter_soft); WSPR_CARVE(k_fano_enqueue);
    WSPR_CARVE(k_fano_workers); WSPR_CARVE(k_resolve); WSPR_CARVE(k_sub_phase); WSPR_CARVE(k_sub_ref);
    WSPR_CARVE(k_sub_lpf<WSPR_LPF_THREADS>); WSPR_CARVE(k_reset_caps); WSPR_CARVE(k_finish); WSPR_CARVE(k_normalise);
#undef WSPR_CARVE
}

}  // namespace wspr
