/* TEST INFRASTRUCTURE -- implementation of the FFTW stand-in declared in fftw3.h (see there). */
#include <stdlib.h>
#include "fftw3.h"
#include "../fft512_twiddle.h"

static const double TW[256][2] = FFT512_TWIDDLE_INIT;

struct fftwf_standin_plan_s {
    fftwf_complex *in, *out;
};

/* 512-point forward DFT: 9-bit bit reversal, then 9 radix-2 DIT stages in binary64.
 * Butterfly (u, v, W): t = W*v with t.re = W.re*v.re - W.im*v.im, t.im = W.re*v.im + W.im*v.re
 * (four roundings for the products, two for the sums); u' = u + t, v' = u - t. */
#ifdef FFT_STANDIN_FLOAT
/* Sensitivity variants (tools/fft_sensitivity.py; never the checker the parity tests use): the same transform evaluated
 * in binary32 like FFTW3f does -- FFTW's codelets are not reproduced (the library is an un-vendored dependency of the
 * reference), the point is a perturbation of the spectrogram of the size of FFTW3f's own rounding error (~1e-7 relative).
 * FFT_STANDIN_FLOAT = 1: radix-2 decimation in time (the graph above in float); 2: radix-2 decimation in frequency (a
 * different order of roundings). */
void oracle_dft512(const float *in, float *out) {
    float re[512], im[512];
#if FFT_STANDIN_FLOAT == 1
    for (int n = 0; n < 512; n++) {
        unsigned r = 0;
        for (int b = 0; b < 9; b++) r |= ((n >> b) & 1u) << (8 - b);
        re[r] = in[2 * n];
        im[r] = in[2 * n + 1];
    }
    for (int half = 1; half < 512; half <<= 1) {
        int stride = 256 / half;
        for (int base = 0; base < 512; base += 2 * half)
            for (int j = 0; j < half; j++) {
                float wr = (float)TW[j * stride][0], wi = (float)TW[j * stride][1];
                int a = base + j, b = a + half;
                float tr = wr * re[b] - wi * im[b];
                float ti = wr * im[b] + wi * re[b];
                re[b] = re[a] - tr;
                im[b] = im[a] - ti;
                re[a] = re[a] + tr;
                im[a] = im[a] + ti;
            }
    }
    for (int k = 0; k < 512; k++) {
        out[2 * k] = re[k];
        out[2 * k + 1] = im[k];
    }
#else
    for (int n = 0; n < 512; n++) {
        re[n] = in[2 * n];
        im[n] = in[2 * n + 1];
    }
    for (int half = 256; half >= 1; half >>= 1) {
        int stride = 256 / half;
        for (int base = 0; base < 512; base += 2 * half)
            for (int j = 0; j < half; j++) {
                float wr = (float)TW[j * stride][0], wi = (float)TW[j * stride][1];
                int a = base + j, b = a + half;
                float dr = re[a] - re[b], di = im[a] - im[b];
                re[a] = re[a] + re[b];
                im[a] = im[a] + im[b];
                re[b] = wr * dr - wi * di;
                im[b] = wr * di + wi * dr;
            }
    }
    for (int k = 0; k < 512; k++) {
        unsigned r = 0;
        for (int b = 0; b < 9; b++) r |= ((k >> b) & 1u) << (8 - b);
        out[2 * k] = re[r];
        out[2 * k + 1] = im[r];
    }
#endif
}
#else
void oracle_dft512(const float *in, float *out) {
    double re[512], im[512];
    for (int n = 0; n < 512; n++) {
        unsigned r = 0;
        for (int b = 0; b < 9; b++) r |= ((n >> b) & 1u) << (8 - b);
        re[r] = (double)in[2 * n];
        im[r] = (double)in[2 * n + 1];
    }
    for (int half = 1; half < 512; half <<= 1) {
        int stride = 256 / half;
        for (int base = 0; base < 512; base += 2 * half) {
            for (int j = 0; j < half; j++) {
                double wr = TW[j * stride][0], wi = TW[j * stride][1];
                int a = base + j, b = a + half;
                double tr = wr * re[b] - wi * im[b];
                double ti = wr * im[b] + wi * re[b];
                re[b] = re[a] - tr;
                im[b] = im[a] - ti;
                re[a] = re[a] + tr;
                im[a] = im[a] + ti;
            }
        }
    }
    for (int k = 0; k < 512; k++) {
        out[2 * k] = (float)re[k];
        out[2 * k + 1] = (float)im[k];
    }
}
#endif

void *fftwf_malloc(size_t n) { return malloc(n); }
void fftwf_free(void *p) { free(p); }

fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags) {
    (void)flags;
    if (n != 512 || sign != FFTW_FORWARD) {
        fprintf(stderr, "fftw stand-in: only the forward 512-point transform is provided\n");
        abort();
    }
    fftwf_plan p = (fftwf_plan)malloc(sizeof(*p));
    p->in = in;
    p->out = out;
    return p;
}

void fftwf_execute(const fftwf_plan p) { oracle_dft512(&p->in[0][0], &p->out[0][0]); }
void fftwf_destroy_plan(fftwf_plan p) { free(p); }
int fftwf_import_wisdom_from_file(FILE *f) { (void)f; return 1; }
void fftwf_export_wisdom_to_file(FILE *f) { (void)f; }
