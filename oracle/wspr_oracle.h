/* TEST INFRASTRUCTURE -- CPU restatement of the reference's WSPR decode hot path (plain C, glibc libm).
 *
 * This is the checker, never the product: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  Parity status: PINNED -- validated in
 * tests/test_oracle_vs_ref.py against the reference's own compiled sources (oracle/_ref, built from
 * /root/reference by oracle/Makefile), the reference's golden spot lines
 * (documentation/bug-fix/REPORT.md:197-203), its Fano/unpack known-answer tests (tests/test_wsprd.c:168-220,
 * :345-384) and the committed fixtures under tests/golden/.
 *
 * The exported names and signatures equal the reference ABI (wsprd/wsprd.h:76-111, fano.h:14-28,
 * wsprd_utils.h:32-42, wsprsim_utils.h:3-9, nhash.h:3) so the same ctypes bindings drive the compiled
 * reference, this restatement and (for the decode entry points) the CUDA library.
 */
#ifndef WSPR_ORACLE_H
#define WSPR_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

struct decoder_options {          /* wsprd/wsprd.h:44-52 */
    int freq;
    char rcall[13];
    char rloc[7];
    int quickmode;
    int usehashtable;
    int npasses;
    int subtraction;
};

struct cand {                     /* wsprd/wsprd.h:54-60 */
    float freq;
    float snr;
    int shift;
    float drift;
    float sync;
};

struct decoder_results {          /* wsprd/wsprd.h:62-74 */
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

/* ---- reference ABI ---- */
int wspr_decode(float *idat, float *qdat, int samples, struct decoder_options options,
                struct decoder_results *decodes, int *n_results);
void sync_and_demodulate(float *id, float *qd, long np, unsigned char *symbols, float *freq, int ifmin,
                         int ifmax, float fstep, int *shift, int lagmin, int lagmax, int lagstep,
                         float *drift, int symfac, float *sync, int mode);
void subtract_signal2(float *id, float *qd, long np, float f0, int shift, float drift,
                      const unsigned char *channel_symbols);
int fano(unsigned int *metric, unsigned int *cycles, unsigned int *maxnp, unsigned char *data,
         unsigned char *symbols, unsigned int nbits, int mettab[2][256], int delta, unsigned int maxcycles);
int encode(unsigned char *symbols, unsigned char *data, unsigned int nbytes);
void deinterleave(unsigned char *sym);
void interleave(unsigned char *sym);
void unpack50(signed char *dat, int32_t *n1, int32_t *n2);
int unpackcall(int32_t ncall, char *call);
int unpackgrid(int32_t ngrid, char *grid);
int unpackpfx(int32_t nprefix, char *call);
int unpk_(signed char *message, char *hashtab, char *loctab, char *call_loc_pow, char *call, char *loc,
          char *pwr, char *callsign);
char get_locator_character_code(char ch);
char get_callsign_character_code(char ch);
long unsigned int pack_grid4_power(char const *grid4, int power);
long unsigned int pack_call(char const *callsign);
void pack_prefix(char *callsign, int32_t *n, int32_t *m, int32_t *nadd);
int get_wspr_channel_symbols(char *rawmessage, char *hashtab, char *loctab, unsigned char *symbols);
uint32_t nhash(const void *key, size_t length, uint32_t initval);

/* ---- intermediates the reference keeps private inside wspr_decode (exposed for stage-level parity) ---- */
void oracle_mettab(int mettab[2][256]);                                   /* wsprd.c:467-473 (derived ints) */
void oracle_window(float *win /*[512]*/);                                /* wsprd.c:510-513 */
int oracle_blocks(int samples);                                           /* wsprd.c:516 */
void oracle_spectrogram(const float *idat, const float *qdat, int samples, float *ps /*[512][blocks]*/);
/* candidate finder + coarse sync of one pass (wsprd.c:555-678); returns npk, fills cands[<=200] */
int oracle_candidates(const float *ps, int blocks, int maxdrift, struct cand *cands,
                      float *smspec_out /*[411] after normalisation, may be NULL*/);

/* ---- front end (rtlsdr_wsprd.c:126-244) and hand-off normalisation (:285-305) ---- */
/* raw: interleaved u8 (I,Q) pairs, n_iq pairs, zero initial filter state; returns the number of outputs
 * written (at most max_out).  Streaming state is internal to one call (= first slot after start). */
int oracle_decimate(const uint8_t *raw, size_t n_iq, float *i_out, float *q_out, int max_out);
void oracle_normalise(float *idat, float *qdat, int n);

#ifdef __cplusplus
}
#endif
#endif
