/* C ABI of libwsprd_b200.so -- the B200 drop-in for the decode hot path of Guenael/rtlsdr-wsprd.
 *
 * Plain C: pointers and sizes only.  Section 1 re-exports the reference's own entry points with identical
 * signatures, so the reference application (rtlsdr_wsprd.c) and its unit tests (tests/test_wsprd.c) link against
 * this library instead of wsprd/*.o without source changes.  Section 2 adds the batch / device-resident entry
 * points a throughput caller uses.  Section 3 is the front end.  Every compute entry point needs a CUDA device and
 * returns WSPR_ERR_CUDA (or, for the void reference signatures, prints to stderr and leaves outputs untouched)
 * when there is none -- there is no CPU fallback.
 *
 * file:line citations are relative to the reference checkout.
 */
#ifndef WSPR_B200_H
#define WSPR_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WSPR_OK 0
#define WSPR_ERR_CUDA (-2)
#define WSPR_ERR_ARG (-3)
#define WSPR_ERR_LIMIT (-4) /* an internal capacity was exceeded (results would differ from the reference's): nothing is returned */
#define WSPR_MAX_UNIQUES 100 /* MAX_UNIQUES, wsprd/wsprd.h:41 */
#define WSPR_CAPTURE_SAMPLES 45000

/* ---- wsprd/wsprd.h:44-74 ---- */
struct decoder_options {
    int freq;          /* dial frequency, Hz */
    char rcall[13];
    char rloc[7];
    int quickmode;
    int usehashtable;  /* persistent hashtable.txt in the CWD (reference -H, wsprd.c:481-494,842-852) */
    int npasses;
    int subtraction;
};
struct cand {
    float freq;
    float snr;
    int shift;
    float drift;
    float sync;
};
struct decoder_results {
    double freq;
    float sync;
    float snr;
    float dt;
    float drift;
    int jitter;
    char message[23];
    char call[13];
    char loc[7];
    char pwr[3];
    int cycles;
};

/* ================= 1. reference entry points ================= */
/* wsprd/wsprd.h:106-111.  idat/qdat (host, `samples` floats each) are mutated by the signal subtraction exactly like
 * the reference does; decodes must hold WSPR_MAX_UNIQUES entries.  Always returns 0 like the reference; on a CUDA
 * failure *n_results is set to 0 and the error is printed to stderr.
 * Only `samples` floats are read: where the reference's spectrogram loop reaches past the end of a short capture (up to
 * 255 samples when samples % 512 < 256, wsprd.c:516,536-541) this library reads zeros, which is what the reference itself
 * finds there in its own program (full-size buffers with a zeroed tail, rtlsdr_wsprd.c:285-288,575-589). */
int wspr_decode(float *idat, float *qdat, int samples, struct decoder_options options,
                struct decoder_results *decodes, int *n_results);
/* wsprd/wsprd.h:76-91 */
void sync_and_demodulate(float *id, float *qd, long np, unsigned char *symbols, float *freq, int ifmin, int ifmax,
                         float fstep, int *shift, int lagmin, int lagmax, int lagstep, float *drift, int symfac,
                         float *sync, int mode);
/* wsprd/wsprd.h:92-98 (exported by the reference, never called by it) */
void subtract_signal(float *id, float *qd, long np, float f0, int shift, float drift, const unsigned char *channel_symbols);
/* wsprd/wsprd.h:99-105 */
void subtract_signal2(float *id, float *qd, long np, float f0, int shift, float drift,
                      const unsigned char *channel_symbols);
/* wsprd/fano.h:14-28 (host evaluation of the same inline code the Fano kernel runs) */
int fano(unsigned int *metric, unsigned int *cycles, unsigned int *maxnp, unsigned char *data,
         unsigned char *symbols, unsigned int nbits, int mettab[2][256], int delta, unsigned int maxcycles);
int encode(unsigned char *symbols, unsigned char *data, unsigned int nbytes);
extern unsigned char Partab[256];
/* wsprd/wsprd_utils.h:32-42 */
void unpack50(signed char *dat, int32_t *n1, int32_t *n2);
int unpackcall(int32_t ncall, char *call);
int unpackgrid(int32_t ngrid, char *grid);
int unpackpfx(int32_t nprefix, char *call);
void deinterleave(unsigned char *sym);
int doublecomp(const void *elem1, const void *elem2);
int floatcomp(const void *elem1, const void *elem2);
int unpk_(signed char *message, char *hashtab, char *loctab, char *call_loc_pow, char *call, char *loc, char *pwr,
          char *callsign);
/* wsprd/wsprsim_utils.h:3-9 */
char get_locator_character_code(char ch);
char get_callsign_character_code(char ch);
long unsigned int pack_grid4_power(char const *grid4, int power);
long unsigned int pack_call(char const *callsign);
void pack_prefix(char *callsign, int32_t *n, int32_t *m, int32_t *nadd);
void interleave(unsigned char *sym);
int get_wspr_channel_symbols(char *rawmessage, char *hashtab, char *loctab, unsigned char *symbols);
/* wsprd/nhash.h:3 */
uint32_t nhash(const void *key, size_t length, uint32_t initval);

/* ================= 2. batch / device-resident decode ================= */
/* One-shot batch: I/Q are [ncaptures][samples] row-major host arrays (not modified); out is
 * [ncaptures][WSPR_MAX_UNIQUES]; n_results is [ncaptures].  wspr_decode() is this with ncaptures = 1 plus the
 * write-back of the subtracted samples.  device < 0 selects the current CUDA device. */
int wspr_decode_batch(const float *I, const float *Q, int ncaptures, int samples, struct decoder_options options,
                      struct decoder_results *out, int *n_results, int device);

typedef struct wspr_ctx wspr_ctx;
/* A context owns all device buffers for up to max_captures captures of `samples` samples on one GPU. */
wspr_ctx *wspr_ctx_create(int device, int max_captures, int samples);
void wspr_ctx_destroy(wspr_ctx *ctx);
const char *wspr_last_error(void);
/* host -> device copy of ncaptures captures ([ncaptures][samples] each); pinned host memory makes it asynchronous */
int wspr_ctx_upload(wspr_ctx *ctx, const float *I, const float *Q, int ncaptures);
/* use captures already resident in device memory (device pointers, row stride in floats); copied device-to-device */
int wspr_ctx_upload_device(wspr_ctx *ctx, const float *dI, const float *dQ, int ncaptures, int row_stride);
/* raw 2.4 Msps u8 IQ streams resident on the device -> the context's sample planes: rtlsdr_callback (rtlsdr_wsprd.c:126-244)
 * for nstreams whole streams of n_iq samples (16-byte aligned, stream_stride_bytes apart), tail zeroed like the hand-off
 * (rtlsdr_wsprd.c:285-288).  Asynchronous on the context's stream; follow with wspr_ctx_normalise and wspr_ctx_decode.
 * Returns the samples each stream produced (n_iq / 6401, at most the context's capture length). */
int wspr_ctx_decimate(wspr_ctx *ctx, const uint8_t *d_raw, int nstreams, size_t n_iq, size_t stream_stride_bytes);
/* peak-normalise every resident capture to 0.5 (rtlsdr_wsprd.c:291-305) */
int wspr_ctx_normalise(wspr_ctx *ctx);
/* decode the resident captures (both passes, subtraction, ...); results stay on the device until downloaded */
int wspr_ctx_decode(wspr_ctx *ctx, struct decoder_options options);
/* device -> host: results [ncaptures][100], n_results [ncaptures]; I_out/Q_out (may be NULL) receive the
 * post-subtraction samples */
int wspr_ctx_download(wspr_ctx *ctx, struct decoder_results *out, int *n_results, float *I_out, float *Q_out);
/* device time of the last wspr_ctx_decode in ms (CUDA events on the context's stream) and kernels launched so far */
float wspr_ctx_last_decode_ms(wspr_ctx *ctx);
/* scheduling statistics of the last decode: rounds driven, candidates finished on side streams */
int wspr_ctx_last_rounds(wspr_ctx *ctx);
int wspr_ctx_last_deferred(wspr_ctx *ctx);
/* how the deferred candidates were settled: out8[0] by the full-budget jitter-0 Fano run, [1] by a jittered attempt,
 * [2] never decoded */
int wspr_ctx_last_stats(wspr_ctx *ctx, int *out8);
/* the cudaStream_t all of the context's copies and kernels are issued on (for callers that record their own events) */
void *wspr_ctx_stream(wspr_ctx *ctx);
unsigned long long wspr_kernel_launches(void);
/* kernel-level timing of the last decode (enable with wspr_ctx_time_kernels(ctx, 1); adds a stream synchronisation per
 * wave, so leave it off for throughput runs): accumulated ms of the mode-0 sync correlation kernel, its launch count
 * and the (lag, symbol) cells it evaluated */
int wspr_ctx_time_kernels(wspr_ctx *ctx, int on);
float wspr_ctx_last_sync_ms(wspr_ctx *ctx);
int wspr_ctx_last_sync_launches(wspr_ctx *ctx);
double wspr_ctx_last_sync_cells(wspr_ctx *ctx);

/* Counters of the per-device pool of Fano worker warps that finishes parked candidates (wsprd.c:741-766 when fano() does
 * not converge at once): out8[0] worker warps allowed, [1] SMs set aside for them (0: they share the SMs with the other
 * kernels), [2] housekeeping periods (256 decoder-loop trips) worker warps were alive for, [3] lane-periods with an attempt
 * in the lane (utilisation = [3] / (32 x [2])), [4] attempts decoded to their end, [5] attempts skipped or abandoned,
 * [6] worker warps started, [7] the part of [2] spent by overflow workers (section 6b of DESIGN.md).  device -1: the current
 * device; reset != 0 clears the counters. */
int wspr_fano_stats(int device, unsigned long long *out8, int reset);

/* the Fano decoder kernel (K5) on caller-supplied soft symbols: n vectors of 162 deinterleaved bytes, the batch / device
 * counterpart of fano() (wsprd/fano.h:14-28; metric table = the one wspr_decode builds, wsprd.c:467-473).  stop_after != 0
 * cuts a run short after that many cycles (rc 2); solo bit 0: one attempt per warp instead of 32; bit 2: the instantiation
 * the decode kernels use (time-out test every 256 trips, maxnp not tracked: rc, cycles and data as fano.c's, metric too for
 * successful decodes).
 * rc/metric/cycles/maxnp: n entries each, data: n x 12 bytes (host memory); clocks (may be NULL): SM clock ticks each
 * attempt took. */
int wspr_fano_batch(const unsigned char *symbols, int n, int delta, unsigned maxcycles, unsigned stop_after, int solo, int *rc,
                    unsigned *metric, unsigned *cycles, unsigned *maxnp, unsigned char *data, unsigned long long *clocks);

/* stage-level access for parity tests (run the first stages of pass 0 on the resident captures) */
int wspr_ctx_spectrogram(wspr_ctx *ctx, float *ps_out /* [ncaptures][512][blocks], reference layout */);
int wspr_ctx_candidates(wspr_ctx *ctx, int maxdrift, struct cand *cands /* [ncaptures][200] */, int *npk,
                        float *smspec /* [ncaptures][411] or NULL */);

/* ================= 3. front end ================= */
/* rtlsdr_callback (rtlsdr_wsprd.c:126-244) for whole streams: raw is nstreams x n_iq interleaved u8 (I,Q) pairs
 * (host memory), zero initial filter state; I/Q receive [nstreams][max_out] floats (zero padded); returns the
 * number of outputs per stream (n_iq / 6401) or a negative error. */
int wspr_decimate_batch(const uint8_t *raw, int nstreams, size_t n_iq, float *I, float *Q, int max_out, int device);
/* same with device-resident input/output (device pointers) */
int wspr_decimate_device(const uint8_t *d_raw, int nstreams, size_t n_iq, size_t stream_stride_bytes, float *dI,
                         float *dQ, int out_stride, int max_out, int device);
/* device time (ms, CUDA events) of the kernels of the last wspr_decimate_device call; text of the last front-end error */
float wspr_decimate_last_ms(void);
const char *wspr_frontend_last_error(void);


/* ---- streaming form: what the live daemon's receive thread and main loop do ----
 * A front-end context owns, for `nstreams` receivers advancing in lockstep, the filter state rtlsdr_callback keeps in
 * function statics (integrators, decimation phase, comb delay lines, FIR history: rtlsdr_wsprd.c:130-136,155-156; zero at
 * creation, never reset afterwards) and the reference's double buffer of `slot_samples` outputs (rx_state.iSamples /
 * qSamples / iqIndex / bufferIndex, rtlsdr_wsprd.c:80-87; 0 selects 45000). */
typedef struct wspr_frontend wspr_frontend;
wspr_frontend *wspr_frontend_create(int device, int nstreams, int slot_samples);
void wspr_frontend_destroy(wspr_frontend *fe);
/* rtlsdr_callback(samples, samples_count, ctx) (rtlsdr_wsprd.c:126) for every stream: raw + s * stream_stride_bytes holds
 * nbytes interleaved u8 (I,Q) bytes of stream s (host memory, not modified; nbytes must be a multiple of 8 as the
 * reference's mixer loop assumes, :171).  Outputs beyond slot_samples are dropped but still advance the filters (:238).
 * Returns the number of outputs produced per stream, or a negative error. */
int wspr_frontend_push(wspr_frontend *fe, const uint8_t *raw, size_t stream_stride_bytes, uint32_t nbytes);
/* outputs held by the slot being filled (rx_state.iqIndex[bufferIndex]) */
int wspr_frontend_samples(wspr_frontend *fe);
/* the main loop's slot switch (rtlsdr_wsprd.c:1181-1183): the other buffer becomes current and starts empty.  Returns the
 * number of samples in the slot that just ended; its tail is zeroed like decoder() does (rtlsdr_wsprd.c:285-288). */
int wspr_frontend_swap(wspr_frontend *fe);
/* the slot that ended at the last swap: host copy ([nstreams][slot_samples] floats each) / device pointers (row stride in
 * floats) to hand to wspr_ctx_upload_device + wspr_ctx_normalise.  Both return the samples the slot holds. */
int wspr_frontend_read(wspr_frontend *fe, float *I, float *Q);
int wspr_frontend_slot_device(wspr_frontend *fe, const float **dI, const float **dQ, int *stride);

#ifdef __cplusplus
}
#endif
#endif
