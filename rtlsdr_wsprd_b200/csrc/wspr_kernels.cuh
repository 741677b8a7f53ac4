// Device data structures and kernel declarations of the batched WSPR decode path.
// (The kernels live in wspr_kernels.cu, the host-side wave scheduler in wspr_decode.cu.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "wspr_codec.cuh"

namespace wspr {

constexpr int NFFT = 512;
constexpr int HOP = 128;
constexpr int MAXCAND = 200;     // MAX_CANDIDATES, wsprd/wsprd.h:40
constexpr int MAXUNIQ = 100;     // MAX_UNIQUES,    wsprd/wsprd.h:41
constexpr int NSMOOTH = 411;     // smoothed-spectrum bins, wsprd.c:566
constexpr int MAXLAGS = 33;      // (128+128)/8 + 1 lags of the mode-0 search, wsprd.c:713-715
constexpr int NFREQ1 = 5;        // mode-1 frequency hypotheses, wsprd.c:722-724
constexpr int NJIT = 43;         // jitter attempts 0..42, wsprd.c:741
constexpr int NFILT = 360;       // subtract_signal2 low-pass length, wsprd.c:325
constexpr int NSIG = NSYM * SPS; // 41472 samples of one transmission
constexpr int CPAD = 42240;      // padded length of the s*conj(r) product (NSIG + 2*NFILT, rounded up)
constexpr int PHI_SEG = 8;      // subtract_signal2: the running phase is recorded every PHI_SEG samples (divides SPS)
constexpr int HASH_CAP = 224;    // per-capture callsign-hash entries kept on the device

// struct cand, wsprd/wsprd.h:54-60
struct Cand {
    float freq, snr;
    int shift;
    float drift, sync;
};

// struct decoder_results, wsprd/wsprd.h:62-74 (80 bytes; layout checked with static_assert in wspr_decode.cu)
struct Spot {
    double freq;
    float sync, snr, dt, drift;
    int jitter;
    char message[23];
    char call[13];
    char loc[7];
    char pwr[3];
    int cycles;
};

// decode thresholds and options (wsprd.c:424-433, wsprd.h:44-52); per-pass values come from pass_maxdrift()/pass_minsync2()
struct DecodeParams {
    int np;             // samples per capture (45000)
    int stride;         // floats between captures in the I/Q planes
    int blocks;         // spectrogram columns (347)
    int dialfreq;       // options.freq
    int quickmode, subtraction, npasses;
    float minsync1, minrms;
    int symfac, delta;
    unsigned maxcycles;
    int lagstep;        // 8, or 16 in quick mode
    int nlags;          // 33 / 17
    unsigned fano_budget;   // cycles a jitter-0 Fano attempt may spend inside a round before it is deferred
    const char *preload;    // device [32768][13]: callsign hash table loaded from hashtable.txt (options.usehashtable), or null
};
// wsprd.c:524-531
// (only passes 0/1 and 2 assign them, so later passes keep the values of pass 2)
__host__ __device__ inline int pass_maxdrift(int ipass) { return ipass >= 2 ? 0 : 4; }
__host__ __device__ inline float pass_minsync2(int ipass) { return ipass >= 2 ? 0.10f : 0.12f; }

// Scheduling.  The reference walks the candidates of one capture strictly in order (a decode is subtracted from the
// samples before the next candidate is looked at), but captures are independent.  Every capture therefore runs its
// own little state machine and a *round* advances every capture that is ready by one candidate; a capture whose
// Fano attempt needs more than fano_budget cycles (or that must go through the 42-attempt jitter search) is parked
// (PH_WAIT) while the pool of Fano worker warps finishes it, and rejoins a later round (PH_RESOLVE) -- it never holds the
// others up.
enum Phase { PH_SETUP = 0, PH_READY = 1, PH_WAIT = 2, PH_RESOLVE = 3, PH_DONE = 4 };

// the candidate a capture is currently working on (at most one per capture; indexed by capture)
struct Job {
    int cap, rank, slot; // slot = position in the round's job list (indexes the P0/P1 scratch)
    int ipass;
    float freq, drift;
    int shift;
    float sync1;        // sync after the frequency search (what the spot reports)
    float snr;
    int worth;          // sync1 > minsync1
    int fbest;          // winning mode-1 hypothesis (its sums are the jitter-0 soft symbols)
    int lbest;          // index of the winning lag of the mode-0 search (-1: none)
    int decoded;        // a Fano attempt succeeded
    int idt;            // winning jitter attempt
    unsigned cycles;
    unsigned char dec[12];
};

// one soft-decision attempt (candidate x jitter)
struct Attempt {
    int cap;
    int idt;
    int gate;           // sync and rms gates passed -> run the Fano decoder
    int ok;             // decoder result: 1 = decoded
    int unfinished;     // the attempt ran out of its in-round budget and must be re-run to completion
    unsigned cycles;
    float sync2;
    unsigned char dec[12];
    unsigned char sym[NSYM + 2];   // deinterleaved soft symbols
};

// per-capture bookkeeping across candidates and passes (everything wspr_decode keeps on its stack)
struct CapState {
    int phase;          // enum Phase
    int ipass;          // current pass
    int npk;            // candidates of the current pass
    int rank;           // next candidate to examine
    int uniques;
    int broken;         // the reference hit one of its `break`s in this pass
    int nhash;
    int sub_pending;    // a subtraction has been scheduled by the resolve step
    int hash_overflow;  // more than HASH_CAP distinct callsign hashes in one capture: an entry was dropped
    float sub_f0, sub_drift;
    int sub_shift;
    float allfreqs[MAXUNIQ];
    char allcalls[MAXUNIQ][CALL_LEN];
    HashEntry hash[HASH_CAP];
    unsigned char chan[NSYM + 2];
};

// soft symbols, gates and results of the 43 attempts of a capture's parked candidate (attempt 0 = jitter 0); one record per
// capture (a capture has at most one candidate parked at a time), filled by k_jitter_soft, worked off by the Fano workers.
// The records live in memory owned by the per-device Fano service and are never freed while the process runs: a worker may
// still look at a record (and find nothing left to claim in it) after its candidate has been settled.
struct ChainScratch {
    int best;           // lowest attempt number that has decoded so far
    int done;           // attempts accounted for (decoded, timed out, abandoned or skipped)
    int next;           // next jittered attempt to hand out (>= nattempts: nothing left to claim)
    int nattempts;      // 43, or 1 in quick mode
    Job *job;           // where the lane that accounts for the last attempt leaves the outcome ...
    int *phase;         // ... before it hands the capture back to the rounds (CapState::phase = PH_RESOLVE)
    int *stats;         // [0] settled by the full-budget jitter-0 run, [1] by a jittered attempt, [2] never decoded
    int *host_done;     // mapped host counter of the context: captures handed back so far
    int gate[NJIT], ok[NJIT], unfinished[NJIT];
    unsigned cycles[NJIT];
    unsigned char dec[NJIT][12];
    unsigned char sym[NJIT][NSYM + 2];
};

// Device-wide queue of parked candidates (one per GPU, shared by every context of the process).  Producers reserve slots
// with one atomicAdd on `tail` and stamp each entry with its sequence number once it is written.  Two cursors walk the
// same ring: head0 hands out attempt 0 of each candidate (one pop per candidate), head1 points at the oldest candidate
// that may still have jittered attempts to claim (ChainScratch::next) and moves on when there are none.  Attempt 0 of
// every queued candidate is handed out before any jittered attempt.  `active`/`pool`: worker warps alive / allowed.
struct FanoQueueEntry {
    ChainScratch *cs;
    unsigned seq;       // slot index + 1 once the entry is valid
    unsigned pad;
};
struct FanoQueue {
    unsigned head0, head1, tail;
    int active, pool;
    unsigned mask;      // capacity - 1 (power of two)
    int overflow;       // a producer found the ring full (the decode reports an error)
    int per_sm;         // worker warps allowed on one SM (0: no limit)
    // overflow workers (only with an SM partition for the pool): warps that share the SMs of the other kernels and take work
    // only while at least `ovf_backlog` candidates are waiting for lanes -- they absorb the bursts a round produces, the
    // partition carries the base load
    int active2, pool2, ovf_backlog;
    FanoQueueEntry *ring;
    // statistics (wspr_fano_stats): housekeeping periods (256 loop trips) worker warps were alive for / their lanes had an
    // attempt in, attempts decoded to the end, attempts skipped or abandoned, worker warps started
    unsigned long long st_warp_periods, st_lane_periods, st_attempts, st_dropped, st_warps, st_ovf_warp_periods;
    int sm_workers[256];
};

struct Counters {
    int nsetup, njobs, nres, ndefer, nsub, nwait, ndone, maxnpk;
};

// constant tables uploaded once per process (host libm values where the reference computes them with libm)
struct HostTables {
    float window[NFFT];       // sinf(0.006147931*i), wsprd.c:510-513
    float lpf_w[NFILT];       // normalised half-sine taps, wsprd.c:359-365
    float lpf_psum[NFILT];    // running sums, wsprd.c:366-368
    float min_snr;            // powf(10, -0.8), wsprd.c:590
    float floor_snr;          // (float)(0.1*min_snr), wsprd.c:595
};
void upload_tables(const HostTables &t);

// ---- launchers (all asynchronous on `st`); lists are device arrays of capture indices ----
void launch_reset_caps(CapState *caps, int ncap, int npasses, cudaStream_t st);
void launch_plan(CapState *caps, const Cand *cands, Job *jobs, int *setup_list, int *job_list, int *res_list, Counters *cnt,
                 int ncap, int npasses, cudaStream_t st);
void launch_spectrogram(const float *I, const float *Q, float *psT, const int *list, int n, const DecodeParams &p,
                        cudaStream_t st);
void launch_candidates(const float *psT, Cand *cands, CapState *caps, float *smspec_dbg, const int *list, int n,
                       const DecodeParams &p, cudaStream_t st);
void launch_coarse(const float *psT, Cand *cands, const CapState *caps, const int *list, int n, const DecodeParams &p,
                   cudaStream_t st);
// tabs: [njobs][NFREQ1][2 * SPS] float4 scratch for the shared phasor tables (slot 2 = the job's own frequency)
void launch_sync_lags(const float *I, const float *Q, Job *jobs, const int *job_list, int njobs, float4 *P0, float4 *tabs,
                      const DecodeParams &p, cudaStream_t st);
void launch_sync_freqs(const float *I, const float *Q, Job *jobs, const int *job_list, int njobs, const float4 *P0, float4 *P1,
                       float4 *tabs, Attempt *att0, const DecodeParams &p, cudaStream_t st);
// jitter-0 Fano attempts of the round (budgeted) and their triage into the resolve list / the deferred list
void launch_fano_round(Attempt *att0, const int *job_list, int njobs, const DecodeParams &p, cudaStream_t st);
void launch_collect(Job *jobs, const Attempt *att0, CapState *caps, const int *job_list, int njobs, int *res_list,
                    int *defer_list, Counters *cnt, const DecodeParams &p, cudaStream_t st);
// parked candidates (wsprd.c:741-766): soft symbols of the jittered attempts into scratch[capture], then all attempts of the
// n candidates into the device queue; launch_fano_workers starts up to `nwarps` worker warps on `st` (they leave at once
// when the pool is already complete, and when the queue runs dry)
void launch_deferred(const float *I, const float *Q, Job *jobs, const Attempt *att0, CapState *caps, const int *defer_list, int n_max,
                     const Counters *cnt, ChainScratch *scratch, const float4 *tabs, int *stats, int *host_done, FanoQueue *queue,
                     const DecodeParams &p, cudaStream_t st);
void launch_fano_workers(FanoQueue *queue, int nwarps, int cta_warps, bool overflow, const DecodeParams &p, cudaStream_t st);
void init_kernel_attributes(int carveout_kb);   // per device: opt-in to > 48 KB of dynamic shared memory, common carve-out
int fano_warp_smem_bytes();                  // shared memory one worker warp holds
void launch_resolve(Job *jobs, CapState *caps, Spot *spots, const int *res_list, int nres_max, int *sub_list, Counters *cnt,
                    const DecodeParams &p, cudaStream_t st);
void launch_subtract(float *I, float *Q, const CapState *caps, const int *sub_list, int nsub_max, const Counters *cnt,
                     float *phi0, float2 *ref, float2 *cprod, const DecodeParams &p, cudaStream_t st);
void launch_finish(CapState *caps, Spot *spots, int *nres, int *stats, int ncap, cudaStream_t st);
// stand-alone per-symbol subtraction (subtract_signal, wsprd.c:263-312) on one capture resident on the device
void launch_subtract_symbolwise(float *I, float *Q, int np, float f0, int shift, float drift, const unsigned char *chan, cudaStream_t st);
void launch_normalise(float *I, float *Q, int ncap, int n, int stride, cudaStream_t st);

// stand-alone single-call forms used by the reference-ABI wrappers (sync_and_demodulate / subtract_signal2)
void launch_sync_generic(const float *I, const float *Q, int np, float freq, int ifmin, int ifmax, float fstep, int lagmin,
                         int lagmax, int lagstep, float drift, float4 *P, cudaStream_t st);

void launch_fano_test(const unsigned char *symbols, int n, int delta, unsigned maxcycles, unsigned stop_after, int solo, int *rc,
                      unsigned *metric, unsigned *cycles, unsigned *maxnp, unsigned char *data, unsigned long long *clocks,
                      cudaStream_t st);

// front end (rtlsdr_wsprd.c:126-244)
void launch_decimate(const uint8_t *raw, size_t n_iq, int nstreams, size_t stream_stride_bytes, uint4 *moments, float *I,
                     float *Q, int out_stride, int max_out, cudaStream_t st);
int decimate_outputs(size_t n_iq);

unsigned long long kernel_launch_count();
#ifdef WSPR_EXPERIMENTS
void exp_set_queue(FanoQueue *q);
void exp_read_hist(unsigned long long *out16, int reset);
#endif

}  // namespace wspr
