/* TEST INFRASTRUCTURE -- stand-in for the 8 FFTW3 single-precision symbols the reference decoder
 * uses (wsprd/wsprd.c:38,497-507,544,832-840).  FFTW3 (libfftw3f, version unpinned by the reference:
 * README.md:17, Dockerfile:17) is not vendored under /root/reference and not installed in this image,
 * so the reference's own wsprd.c is compiled against this header instead.
 *
 * Semantics kept: unnormalised forward DFT, X[k] = sum_n x[n] exp(-2*pi*i*n*k/N), N = 512 only.
 * Arithmetic: binary64 radix-2 decimation-in-time with the correctly rounded twiddle table in
 * ../fft512_twiddle.h, result rounded once to binary32 (error <= ~1 float ulp of the exact DFT, i.e.
 * tighter than FFTW's own float codelets).  The CUDA spectrogram kernel evaluates the identical DAG.
 */
#ifndef ORACLE_FFTW3_STANDIN_H
#define ORACLE_FFTW3_STANDIN_H
#include <stddef.h>
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef float fftwf_complex[2];
typedef struct fftwf_standin_plan_s *fftwf_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_EXHAUSTIVE (1U << 3)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)
void *fftwf_malloc(size_t n);
void fftwf_free(void *p);
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
int fftwf_import_wisdom_from_file(FILE *f);
void fftwf_export_wisdom_to_file(FILE *f);
/* not an FFTW symbol: the bare transform, used by the oracle restatement too */
void oracle_dft512(const float *in_interleaved, float *out_interleaved);
#ifdef __cplusplus
}
#endif
#endif
