// Host exports of the integer/byte codec stages under the reference's own names and signatures
// (wsprd/fano.h:14-28, wsprd_utils.h:32-42, wsprsim_utils.h:3-9, nhash.h:3), so that the reference's unit tests
// (tests/test_wsprd.c) and any caller of those helpers link against libwsprd_b200.so unchanged.  Each function is a
// thin wrapper over the SAME inline code (wspr_codec.cuh) that the resolve and Fano kernels execute on the device;
// nothing here is a separate implementation.
#include <stdint.h>
#include <string.h>

#include "../../include/wspr_b200.h"
#include "wspr_codec.cuh"
#include "wspr_math.cuh"

using namespace wspr;

extern "C" {

// wsprd/tab.c:7-40 -- 8-bit parity lookup (the device code uses popc instead); filled when the library is loaded
unsigned char Partab[256];
__attribute__((constructor)) static void fill_partab() {
    for (int i = 0; i < 256; i++) Partab[i] = (unsigned char)parity_u32((uint32_t)i);
}

// wsprd/fano.c:63-82
int encode(unsigned char *symbols, unsigned char *data, unsigned int nbytes) {
    conv_encode(symbols, data, nbytes);
    return 0;
}

// wsprd/fano.c:87-238
int fano(unsigned int *metric, unsigned int *cycles, unsigned int *maxnp, unsigned char *data, unsigned char *symbols,
         unsigned int nbits, int mettab[2][256], int delta, unsigned int maxcycles) {
    return fano_decode<int>(metric, cycles, maxnp, data, symbols, nbits, &mettab[0][0], delta, maxcycles);
}

// wsprd/wsprd_utils.c:40-194
void unpack50(signed char *dat, int32_t *n1, int32_t *n2) { unpack_50(dat, n1, n2); }
int unpackcall(int32_t ncall, char *call) { return unpack_call(ncall, call); }
int unpackgrid(int32_t ngrid, char *grid) { return unpack_grid(ngrid, grid); }
int unpackpfx(int32_t nprefix, char *call) { return unpack_pfx(nprefix, call); }
// wsprd/wsprd_utils.c:196-213
void deinterleave(unsigned char *sym) { deinterleave162(sym); }
// wsprd/wsprd_utils.c:216-226
int doublecomp(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return x < y ? -1 : (x > y);
}
int floatcomp(const void *a, const void *b) {
    float x = *(const float *)a, y = *(const float *)b;
    return x < y ? -1 : (x > y);
}
// wsprd/wsprd_utils.c:228-313 (dense 32768-slot tables owned by the caller, like the reference)
int unpk_(signed char *message, char *hashtab, char *loctab, char *call_loc_pow, char *call, char *loc, char *pwr,
          char *callsign) {
    DenseHashStore hs{hashtab, loctab};
    return unpack_message(message, hs, call_loc_pow, call, loc, pwr, callsign);
}

// wsprd/wsprsim_utils.c:15-39
char get_locator_character_code(char ch) { return (char)loc_code(ch); }
char get_callsign_character_code(char ch) { return (char)call_code(ch); }
// wsprd/wsprsim_utils.c:41-47 -- grid4 holds character CODES (signed chars), not letters
long unsigned int pack_grid4_power(char const *grid4, int power) {
    return (long unsigned int)pack_grid_power((const signed char *)grid4, power);
}
// wsprd/wsprsim_utils.c:49-80
long unsigned int pack_call(char const *callsign) { return (long unsigned int)pack_callsign(callsign); }
// wsprd/wsprsim_utils.c:82-142
void pack_prefix(char *callsign, int32_t *n, int32_t *m, int32_t *nadd) { pack_compound(callsign, n, m, nadd); }
// wsprd/wsprsim_utils.c:144-161
void interleave(unsigned char *sym) { interleave162(sym); }
// wsprd/wsprsim_utils.c:163-316
int get_wspr_channel_symbols(char *rawmessage, char *hashtab, char *loctab, unsigned char *symbols) {
    DenseHashStore hs{hashtab, loctab};
    return channel_symbols(rawmessage, hs, symbols);
}

// wsprd/nhash.c:205-451
uint32_t nhash(const void *key, size_t length, uint32_t initval) { return nhash15(key, length, initval); }

// ---- test hooks: the glibc float-function replicas the kernels use (wspr_math.cuh), evaluated on the host ----
float wspr_test_sinf(float x) { return glibc_sinf(x); }
float wspr_test_cosf(float x) { return glibc_cosf(x); }
float wspr_test_log10f(float x) { return glibc_log10f(x); }

}  // extern "C"
