/* TEST INFRASTRUCTURE -- thin export shim around the reference's own front-end callback.
 *
 * `rtlsdr_callback` (rtlsdr_wsprd.c:126-244) is `static` and keeps all CIC/FIR state in function-static
 * variables, so the only way to call the unmodified code is to compile it in this translation unit.
 * oracle/Makefile pipes the reference's line ranges 35-43 (rate macros), 75-91 (struct receiver_state)
 * and 125-244 (the callback) into two frontend_extract_*.inc files in a temporary directory at build time; nothing
 * of the reference is stored in this repository.
 *
 * The filter state cannot be reset (function statics): load a fresh copy of the .so per raw stream.
 */
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <time.h>

#include "frontend_extract_types.inc"

struct receiver_state rx_state;   /* the global the callback writes into (rtlsdr_wsprd.c:114) */

#include "frontend_extract_callback.inc"

/* feed one librtlsdr-sized buffer (the callback mutates it in place, as in the reference) */
void ref_frontend_push(unsigned char *buf, uint32_t nbytes) { rtlsdr_callback(buf, nbytes, NULL); }

/* samples produced so far in buffer 0 */
uint32_t ref_frontend_count(void) { return rx_state.iqIndex[0]; }

void ref_frontend_read(float *i_out, float *q_out, uint32_t n) {
    memcpy(i_out, rx_state.iSamples[0], n * sizeof(float));
    memcpy(q_out, rx_state.qSamples[0], n * sizeof(float));
}
