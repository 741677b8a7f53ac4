/* TEST INFRASTRUCTURE -- CPU restatement of the reference's WSPR decode hot path (plain C, glibc libm).
 *
 * This is the checker, never the product (see wspr_oracle.h for who may load it).  Every function cites
 * the reference lines whose behaviour it restates; all paths are relative to /root/reference.
 * Parity status: PINNED (tests/test_oracle_*.py compare every entry point with the reference's own
 * sources compiled into oracle/_ref, with the reference's unit-test known answers and its two golden
 * spot lines).
 *
 * Arithmetic notes (what has to be kept for bit parity with x86-64 gcc -O3, no -march, no fast-math):
 *   - float expressions are evaluated in binary32, anything touching a double literal in binary64,
 *     each assignment to a float rounds once; no fused multiply-add anywhere;
 *   - the unparenthesised rate macros of wsprd/wsprd.c:59-69 are expanded textually below (OR_* macros
 *     keep the same token sequence so precedence quirks survive);
 *   - sums are accumulated strictly left to right in the reference's loop order.
 */
#include "wspr_oracle.h"

#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fftw_standin/fftw3.h"
#include "wspr_mettab.h"

#define OR_NSYM 162
#define OR_NBITS 81
#define OR_SPS 256       /* samples per symbol */
#define OR_NFFT 512
#define OR_MAXCAND 200
#define OR_MAXUNIQ 100
#define OR_HASHN 32768
#define OR_HLEN 13
#define OR_LLEN 5
#define OR_CAPTURE 45000 /* SIGNAL_LENGHT * SIGNAL_SAMPLE_RATE, wsprd.c:59-61 */
/* token-for-token expansions of wsprd.c:65-69 */
#define OR_DF 375.0 / 256.0
#define OR_DT 1.0 / 375.0
#define OR_TWOPIDT 2.0 * M_PI * OR_DT

/* sync vector, wsprd.c:84-93 (also wsprsim_utils.c:167-176); stored packed, one bit per symbol */
static const unsigned char SYNC_BITS[OR_NSYM] = {
    1,1,0,0,0,0,0,0,1,0,0,0,1,1,1,0,0,0,1,0,0,1,0,1,1,1,1,0,0,0,0,0,0,0,1,0,0,1,0,1,0,0,0,0,0,0,1,0,
    1,1,0,0,1,1,0,1,0,0,0,1,1,0,1,0,0,0,0,1,1,0,1,0,1,0,1,0,1,0,0,1,0,0,1,0,1,1,0,0,0,1,1,0,1,0,1,0,
    0,0,1,0,0,0,0,0,1,0,0,1,0,0,1,1,1,0,1,1,0,0,1,1,0,1,0,0,0,1,1,1,0,0,0,0,0,1,0,1,0,0,1,1,0,0,0,0,
    0,0,0,1,1,0,1,0,1,1,0,0,0,1,1,0,0,0};

/* ------------------------------------------------------------------------------------------------
 * nhash: Bob Jenkins' lookup3 hashlittle(), byte-at-a-time form, masked to 15 bits
 * (wsprd/nhash.c:205-451; only the alignment-independent result matters).
 * ---------------------------------------------------------------------------------------------- */
static inline uint32_t rol32(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }

uint32_t nhash(const void *key, size_t length, uint32_t initval) {
    const uint8_t *p = (const uint8_t *)key;
    uint32_t a, b, c;
    a = b = c = 0xdeadbeefu + (uint32_t)length + initval;
    while (length > 12) {
        a += p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
        b += p[4] | (uint32_t)p[5] << 8 | (uint32_t)p[6] << 16 | (uint32_t)p[7] << 24;
        c += p[8] | (uint32_t)p[9] << 8 | (uint32_t)p[10] << 16 | (uint32_t)p[11] << 24;
        a -= c; a ^= rol32(c, 4);  c += b;
        b -= a; b ^= rol32(a, 6);  a += c;
        c -= b; c ^= rol32(b, 8);  b += a;
        a -= c; a ^= rol32(c, 16); c += b;
        b -= a; b ^= rol32(a, 19); a += c;
        c -= b; c ^= rol32(b, 4);  b += a;
        length -= 12;
        p += 12;
    }
    if (length == 0) return c;                      /* nhash.c:442-443: unmasked early return */
    uint32_t w[3] = {0, 0, 0};
    for (size_t i = 0; i < length; i++) w[i >> 2] += (uint32_t)p[i] << (8 * (i & 3));
    a += w[0]; b += w[1]; c += w[2];
    c ^= b; c -= rol32(b, 14);
    a ^= c; a -= rol32(c, 11);
    b ^= a; b -= rol32(a, 25);
    c ^= b; c -= rol32(b, 16);
    a ^= c; a -= rol32(c, 4);
    b ^= a; b -= rol32(a, 14);
    c ^= b; c -= rol32(b, 24);
    return c & 32767u;                              /* nhash.c:448 */
}

/* ------------------------------------------------------------------------------------------------
 * Convolutional code: K=32, r=1/2 Layland-Lushbaugh polynomials (fano.c:47-53), parity by table in the
 * reference (tab.c:7-40), by folding here.
 * ---------------------------------------------------------------------------------------------- */
#define OR_POLY1 0xf2d05351u
#define OR_POLY2 0xe4613c47u

static inline unsigned parity32(uint32_t v) {
    v ^= v >> 16;
    v ^= v >> 8;
    v ^= v >> 4;
    v ^= v >> 2;
    v ^= v >> 1;
    return v & 1u;
}
/* fano.h:35-44 ENCODE: bit1 from POLY1, bit0 from POLY2 */
static inline unsigned conv_pair(uint32_t state) {
    return (parity32(state & OR_POLY1) << 1) | parity32(state & OR_POLY2);
}

/* fano.c:63-82 */
int encode(unsigned char *symbols, unsigned char *data, unsigned int nbytes) {
    uint32_t st = 0;
    for (unsigned int b = 0; b < nbytes; b++)
        for (int bit = 7; bit >= 0; bit--) {
            st = (st << 1) | ((data[b] >> bit) & 1u);
            unsigned pr = conv_pair(st);
            *symbols++ = (unsigned char)(pr >> 1);
            *symbols++ = (unsigned char)(pr & 1u);
        }
    return 0;
}

/* fano.c:87-238.  Node arrays instead of a struct list; same moves, same cycle accounting:
 * returns 0 on success, -1 when the loop counter reached maxcycles*nbits (note: also when the
 * final forward move happened exactly on the last allowed cycle, fano.c:234). */
/* instrumentation for scheduling studies (tools/oracle_stats.py); not part of any result */
long oracle_stat[16];
int fano(unsigned int *metric, unsigned int *cycles, unsigned int *maxnp, unsigned char *data,
         unsigned char *symbols, unsigned int nbits, int mettab[2][256], int delta, unsigned int maxcycles) {
    enum { MAXN = 256 };
    if (nbits + 1 > MAXN || nbits < 32) return -1;
    uint32_t enc[MAXN];
    int gam[MAXN], bm[MAXN][4], tm[MAXN][2];
    unsigned char sel[MAXN];
    const int last = (int)nbits - 1, tail = (int)nbits - 31;

    for (unsigned int n = 0; n < nbits; n++) {       /* fano.c:118-124 */
        int a0 = mettab[0][symbols[2 * n]], a1 = mettab[1][symbols[2 * n]];
        int b0 = mettab[0][symbols[2 * n + 1]], b1 = mettab[1][symbols[2 * n + 1]];
        bm[n][0] = a0 + b0;
        bm[n][1] = a0 + b1;
        bm[n][2] = a1 + b0;
        bm[n][3] = a1 + b1;
    }
    *maxnp = 0;
    int pos = 0, thr = 0;
    enc[0] = 0;
    {
        unsigned ls = conv_pair(enc[0]);
        int m0 = bm[0][ls], m1 = bm[0][3 ^ ls];
        if (m0 > m1) { tm[0][0] = m0; tm[0][1] = m1; }
        else { tm[0][0] = m1; tm[0][1] = m0; enc[0]++; }
    }
    sel[0] = 0;
    gam[0] = 0;
    const unsigned int limit = maxcycles * nbits;
    unsigned int it;
    for (it = 1; it <= limit; it++) {
        if (pos > (int)*maxnp) *maxnp = (unsigned)pos;
        int ng = gam[pos] + tm[pos][sel[pos]];
        oracle_stat[14]++;
        if (ng >= thr) {                              /* forward, fano.c:158-197 */
            oracle_stat[15]++;
            if (gam[pos] < thr + delta)
                while (ng >= thr + delta) thr += delta;
            gam[pos + 1] = ng;
            enc[pos + 1] = enc[pos] << 1;
            pos++;
            if (pos == last + 1) break;
            unsigned ls = conv_pair(enc[pos]);
            if (pos >= tail) {
                tm[pos][0] = bm[pos][ls];
            } else {
                int m0 = bm[pos][ls], m1 = bm[pos][3 ^ ls];
                if (m0 > m1) { tm[pos][0] = m0; tm[pos][1] = m1; }
                else { tm[pos][0] = m1; tm[pos][1] = m0; enc[pos]++; }
            }
            sel[pos] = 0;
            continue;
        }
        for (;;) {                                    /* backward, fano.c:199-219 */
            if (pos == 0 || gam[pos - 1] < thr) {
                thr -= delta;
                if (sel[pos] != 0) { sel[pos] = 0; enc[pos] ^= 1u; }
                break;
            }
            pos--;
            oracle_stat[13]++;
            if (pos < tail && sel[pos] != 1) {
                sel[pos]++;
                enc[pos] ^= 1u;
                break;
            }
        }
    }
    *metric = (unsigned)gam[pos];
    for (unsigned int b = 0; b < (nbits >> 3); b++) data[b] = (unsigned char)enc[7 + 8 * b];
    *cycles = it + 1;
    return (it >= limit) ? -1 : 0;
}

/* ------------------------------------------------------------------------------------------------
 * Interleaver: 8-bit bit reversal of a running counter, keep values < 162
 * (wsprd_utils.c:196-213, wsprsim_utils.c:144-161)
 * ---------------------------------------------------------------------------------------------- */
static void bitrev_order(unsigned char order[OR_NSYM]) {
    int filled = 0;
    for (unsigned v = 0; filled < OR_NSYM; v++) {
        unsigned r = 0;
        for (int b = 0; b < 8; b++) r |= ((v >> b) & 1u) << (7 - b);
        if (r < OR_NSYM) order[filled++] = (unsigned char)r;
    }
}
void deinterleave(unsigned char *sym) {
    unsigned char ord[OR_NSYM], t[OR_NSYM];
    bitrev_order(ord);
    for (int p = 0; p < OR_NSYM; p++) t[p] = sym[ord[p]];
    memcpy(sym, t, OR_NSYM);
}
void interleave(unsigned char *sym) {
    unsigned char ord[OR_NSYM], t[OR_NSYM];
    bitrev_order(ord);
    for (int p = 0; p < OR_NSYM; p++) t[ord[p]] = sym[p];
    memcpy(sym, t, OR_NSYM);
}

/* ------------------------------------------------------------------------------------------------
 * Message unpacking (wsprd_utils.c:40-194, :228-313)
 * ---------------------------------------------------------------------------------------------- */
static const char ALNUM37[] = "0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ ";

void unpack50(signed char *dat, int32_t *n1, int32_t *n2) {       /* wsprd_utils.c:40-70 */
    uint32_t b[7];
    for (int i = 0; i < 7; i++) b[i] = (uint32_t)(dat[i] & 255);
    *n1 = (int32_t)((b[0] << 20) + (b[1] << 12) + (b[2] << 4) + ((b[3] >> 4) & 15));
    *n2 = (int32_t)(((b[3] & 15) << 18) + (b[4] << 10) + (b[5] << 2) + ((b[6] >> 6) & 3));
}

int unpackcall(int32_t ncall, char *call) {                        /* wsprd_utils.c:72-116 */
    snprintf(call, 13, "......");
    if (ncall >= 262177560) return 0;
    char t[7];
    int32_t n = ncall;
    t[5] = ALNUM37[n % 27 + 10]; n /= 27;
    t[4] = ALNUM37[n % 27 + 10]; n /= 27;
    t[3] = ALNUM37[n % 27 + 10]; n /= 27;
    t[2] = ALNUM37[n % 10];      n /= 10;
    t[1] = ALNUM37[n % 36];      n /= 36;
    t[0] = ALNUM37[n];
    t[6] = 0;
    int lead = 0;
    while (lead < 5 && t[lead] == ' ') lead++;
    snprintf(call, 13, "%-6s", t + lead);
    for (int i = 0; i < 6; i++)
        if (call[i] == ' ') call[i] = 0;                           /* every blank, not only trailing */
    return 1;
}

int unpackgrid(int32_t ngrid, char *grid) {                        /* wsprd_utils.c:118-147 */
    ngrid >>= 7;
    if (ngrid >= 32400) {
        snprintf(grid, 5, "XXXX");
        return 0;
    }
    int dlat = ngrid % 180 - 90;
    int dlong = (ngrid / 180) * 2 - 180 + 2;
    if (dlong < -180) dlong += 360;
    if (dlong > 180) dlong += 360;
    int nlong = 60.0 * (180.0 - dlong) / 5.0;
    int nlat = 60.0 * (dlat + 90) / 2.5;
    grid[0] = ALNUM37[10 + nlong / 240];
    grid[2] = ALNUM37[(nlong - 240 * (nlong / 240)) / 24];
    grid[1] = ALNUM37[10 + nlat / 240];
    grid[3] = ALNUM37[(nlat - 240 * (nlat / 240)) / 24];
    return 1;
}

int unpackpfx(int32_t nprefix, char *call) {                       /* wsprd_utils.c:149-194 */
    char base[13];
    snprintf(base, sizeof base, "%s", call);
    if (nprefix < 60000) {
        char pfx[4] = {0, 0, 0, 0};
        int32_t n = nprefix;
        for (int i = 2; i >= 0; i--) {
            char nc = (char)(n % 37);
            pfx[i] = (nc >= 0 && nc <= 9) ? (char)(nc + 48) : (nc >= 10 && nc <= 35) ? (char)(nc + 55) : ' ';
            n /= 37;
        }
        const char *sp = strrchr(pfx, ' ');
        snprintf(call, 13, "%s/%s", sp ? sp + 1 : pfx, base);
        return 1;
    }
    char nc = (char)(nprefix - 60000);                             /* narrowing to char as in the reference */
    if (nc >= 0 && nc <= 9) snprintf(call, 13, "%s/%c", base, nc + 48);
    else if (nc >= 10 && nc <= 35) snprintf(call, 13, "%s/%c", base, nc + 55);
    else if (nc >= 36 && nc <= 125) snprintf(call, 13, "%s/%c%c", base, (nc - 26) / 10 + 48, (nc - 26) % 10 + 48);
    else return 0;
    return 1;
}

static int is_power_digit(int nu) { return nu == 0 || nu == 3 || nu == 7; }

int unpk_(signed char *message, char *hashtab, char *loctab, char *call_loc_pow, char *call, char *loc,
          char *pwr, char *callsign) {                             /* wsprd_utils.c:228-313 */
    int32_t n1, n2;
    char grid[5], cdbm[4];
    int noprint = 0;
    unpack50(message, &n1, &n2);
    if (!unpackcall(n1, callsign)) return 1;
    if (!unpackgrid(n2, grid)) return 1;
    int ntype = (n2 & 127) - 64;
    callsign[12] = 0;
    grid[4] = 0;
    if (ntype >= 0 && ntype <= 62) {
        int nu = ntype % 10;
        if (is_power_digit(nu)) {                                  /* type 1 */
            snprintf(cdbm, sizeof cdbm, "%02d", ntype);
            snprintf(call_loc_pow, 23, "%s %s %s", callsign, grid, cdbm);
            uint32_t h = nhash(callsign, strlen(callsign), 146u);
            snprintf(hashtab + h * OR_HLEN, OR_HLEN, "%s", callsign);
            snprintf(loctab + h * OR_LLEN, OR_LLEN, "%s", grid);
            snprintf(call, OR_HLEN, "%s", callsign);
            snprintf(loc, 7, "%s", grid);
            snprintf(pwr, 3, "%s", cdbm);
        } else {                                                   /* type 2 */
            int nadd = nu;
            if (nu > 3) nadd = nu - 3;
            if (nu > 7) nadd = nu - 7;
            int n3 = n2 / 128 + OR_HASHN * (nadd - 1);
            if (!unpackpfx(n3, callsign)) return 1;
            int ndbm = ntype - nadd;
            snprintf(cdbm, sizeof cdbm, "%2d", ndbm);
            snprintf(call_loc_pow, 23, "%s %s", callsign, cdbm);
            if (is_power_digit(ndbm % 10)) {
                uint32_t h = nhash(callsign, strlen(callsign), 146u);
                snprintf(hashtab + h * OR_HLEN, OR_HLEN, "%s", callsign);
            } else {
                noprint = 1;
            }
        }
    } else if (ntype < 0) {                                        /* type 3 */
        int ndbm = -(ntype + 1);
        char grid6[7];
        memset(grid6, 0, sizeof grid6);
        snprintf(grid6, sizeof grid6, "%c%.*s", callsign[5], 5, callsign);
        if (!is_power_digit(ndbm % 10) || !isalpha((unsigned char)grid6[0]) || !isalpha((unsigned char)grid6[1]) ||
            !isdigit((unsigned char)grid6[2]) || !isdigit((unsigned char)grid6[3]))
            noprint = 1;
        int h = (n2 - ntype - 64) / 128;
        if (hashtab[h * OR_HLEN] != 0) snprintf(callsign, OR_HLEN, "<%s>", hashtab + h * OR_HLEN);
        else snprintf(callsign, OR_HLEN, "<...>");
        snprintf(cdbm, sizeof cdbm, "%2d", ndbm);
        snprintf(call_loc_pow, 23, "%s %s %s", callsign, grid6, cdbm);
        snprintf(call, OR_HLEN, "%s", callsign);
        snprintf(loc, 7, "%s", grid6);
        snprintf(pwr, 3, "%s", cdbm);
        if (ntype == -64) noprint = 1;
    }
    return noprint;
}

/* ------------------------------------------------------------------------------------------------
 * Message packing / channel symbols (wsprsim_utils.c:15-316)
 * ---------------------------------------------------------------------------------------------- */
char get_locator_character_code(char ch) {                         /* wsprsim_utils.c:15-26 */
    if (ch >= '0' && ch <= '9') return ch - '0';
    if (ch == ' ') return 36;
    if (ch >= 'A' && ch <= 'R') return ch - 'A';
    return -1;
}
char get_callsign_character_code(char ch) {                        /* wsprsim_utils.c:28-39 */
    if (ch >= '0' && ch <= '9') return ch - '0';
    if (ch == ' ') return 36;
    if (ch >= 'A' && ch <= 'Z') return ch - 'A' + 10;
    return -1;
}
long unsigned int pack_grid4_power(char const *g, int power) {     /* wsprsim_utils.c:41-47 */
    long unsigned int m = (179 - 10 * g[0] - g[2]) * 180 + 10 * g[1] + g[3];
    return m * 128 + power + 64;
}
long unsigned int pack_call(char const *callsign) {                /* wsprsim_utils.c:49-78 */
    size_t len = strlen(callsign);
    if (len > 6) return 0;
    char c6[8];
    memset(c6, ' ', sizeof c6);
    if (isdigit((unsigned char)callsign[2])) {
        for (size_t i = 0; i < len; i++) c6[i] = callsign[i];
    } else if (isdigit((unsigned char)callsign[1])) {
        for (size_t i = 1; i < len + 1 && i < 7; i++) c6[i] = callsign[i - 1];   /* i==6 would overflow there */
    }
    long unsigned int n = 0;
    static const int radix[6] = {1, 36, 10, 27, 27, 27};
    static const int bias[6] = {0, 0, 0, 10, 10, 10};
    for (int i = 0; i < 6; i++) n = n * radix[i] + get_callsign_character_code(c6[i]) - bias[i];
    return n;
}

static int alnum_value(int ch, int other) {
    if (ch >= '0' && ch <= '9') return ch - '0';
    if (ch >= 'A' && ch <= 'Z') return ch - 'A' + 10;
    return other;
}

void pack_prefix(char *callsign, int32_t *n, int32_t *m, int32_t *nadd) {   /* wsprsim_utils.c:80-142 */
    char c6[16];
    memset(c6, 0, sizeof c6);
    size_t slash = strcspn(callsign, "/");
    if (callsign[slash + 2] == 0) {                                /* CALL/x */
        memcpy(c6, callsign, slash < 12 ? slash : 12);
        *n = (int32_t)pack_call(c6);
        *nadd = 1;
        *m = 60000 - 32768 + alnum_value(callsign[slash + 1], 38);
    } else if (callsign[slash + 3] == 0) {                         /* CALL/nn */
        memcpy(c6, callsign, slash < 12 ? slash : 12);
        *n = (int32_t)pack_call(c6);
        *nadd = 1;
        *m = 60000 + 26 + 10 * (callsign[slash + 1] - 48) + (callsign[slash + 2] - 48);
    } else {                                                       /* PFX/CALL (strtok cuts callsign at '/') */
        const char *pfx = strtok(callsign, "/");
        const char *rest = strtok(NULL, " ");
        *n = rest ? (int32_t)pack_call(rest) : 0;   /* NULL would crash the reference */
        size_t plen = strlen(pfx);
        *m = (plen == 1) ? 37 * 36 + 36 : (plen == 2) ? 36 : 0;
        for (size_t i = 0; i < plen; i++) *m = 37 * (*m) + alnum_value(callsign[i], 36);
        *nadd = 0;
        if (*m > 32768) {
            *m -= 32768;
            *nadd = 1;
        }
    }
}

int get_wspr_channel_symbols(char *rawmessage, char *hashtab, char *loctab, unsigned char *symbols) {
    /* wsprsim_utils.c:163-316 */
    static const int round_pwr[10] = {0, -1, 1, 0, -1, 2, 1, 0, -1, 1};
    char msg[24];
    memset(msg, 0, sizeof msg);
    for (int i = 0; i < 23 && rawmessage[i]; i++) msg[i] = rawmessage[i];
    size_t sp = strcspn(msg, " "), sl = strcspn(msg, "/"), lt = strcspn(msg, "<"), gt = strcspn(msg, ">");
    size_t mlen = strlen(msg);
    long unsigned int n = 0;
    int m = 0;

    if (sp > 3 && sp < 7 && sl == mlen && lt == mlen) {            /* type 1 */
        char *cs = strtok(msg, " "), *grid = strtok(NULL, " "), *ps = strtok(NULL, " ");
        if (!cs || !grid || !ps) return 0;                         /* the reference would crash here */
        int power = atoi(ps);
        n = pack_call(cs);
        char g4[4];
        for (int i = 0; i < 4; i++) g4[i] = get_locator_character_code(grid[i]);
        m = (int)pack_grid4_power(g4, power);
    } else if (lt == 0 && gt < mlen) {                             /* type 3 */
        char *cs = strtok(msg, "<> "), *grid = strtok(NULL, " "), *ps = strtok(NULL, " ");
        if (!cs || !grid || !ps) return 0;
        int power = atoi(ps);
        if (power < 0) power = 0;
        if (power > 60) power = 60;
        power += round_pwr[power % 10];
        int ntype = -(power + 1);
        int h = (int)nhash(cs, strlen(cs), 146u);
        m = 128 * h + ntype + 64;
        char g6[8];
        memset(g6, 0, sizeof g6);
        int gl = (int)strlen(grid);
        for (int i = 0; i < gl - 1 && i < 7; i++) g6[i] = grid[i + 1];
        g6[5] = grid[0];
        n = pack_call(g6);
    } else if (sl < mlen) {                                        /* type 2 */
        char *cs = strtok(msg, " ");
        if (!cs || sl == 0 || sl > strlen(cs)) return 0;
        char *ps = strtok(NULL, " ");
        if (!ps) return 0;
        int power = atoi(ps);
        if (power < 0) power = 0;
        if (power > 60) power = 60;
        power += round_pwr[power % 10];
        int32_t n1, ng, nadd;
        pack_prefix(cs, &n1, &ng, &nadd);
        int ntype = power + 1 + nadd;
        m = 128 * ng + ntype + 64;
        n = (long unsigned int)(long)n1;                           /* int -> unsigned long sign-extends */
    } else {
        return 0;
    }

    unsigned char data[11];
    memset(data, 0, sizeof data);
    data[0] = 0xFF & (n >> 20);
    data[1] = 0xFF & (n >> 12);
    data[2] = 0xFF & (n >> 4);
    data[3] = (unsigned char)(((n & 0x0F) << 4) + ((m >> 18) & 0x0F));
    data[4] = 0xFF & (m >> 10);
    data[5] = 0xFF & (m >> 2);
    data[6] = (unsigned char)((m & 0x03) << 6);

    /* the reference re-unpacks the packed bytes (wsprsim_utils.c:277-295); only the hash-table side
       effect of that call survives */
    char t_clp[23] = {0}, t_cs[13] = {0}, t_call[13] = {0}, t_loc[7] = {0}, t_pwr[3] = {0};
    signed char chk[11];
    memcpy(chk, data, 11);
    unpk_(chk, hashtab, loctab, t_clp, t_call, t_loc, t_pwr, t_cs);

    unsigned char bits[176];
    memset(bits, 0, sizeof bits);
    encode(bits, data, 11);
    interleave(bits);
    for (int i = 0; i < OR_NSYM; i++) symbols[i] = (unsigned char)(2 * bits[i] + SYNC_BITS[i]);
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * sync_and_demodulate (wsprd.c:101-259)
 * ---------------------------------------------------------------------------------------------- */
/* four tone phasor tables for one symbol frequency fp, wsprd.c:158-188 */
static void tone_tables(float fp, float c[4][OR_SPS], float s[4][OR_SPS]) {
    float dphi[4];
    dphi[0] = OR_TWOPIDT * (fp - OR_DF * 1.5);
    dphi[1] = OR_TWOPIDT * (fp - OR_DF * 0.5);
    dphi[2] = OR_TWOPIDT * (fp + OR_DF * 0.5);
    dphi[3] = OR_TWOPIDT * (fp + OR_DF * 1.5);
    for (int t = 0; t < 4; t++) {
        float cd = cosf(dphi[t]), sd = sinf(dphi[t]);
        c[t][0] = 1;
        s[t][0] = 0;
        for (int j = 1; j < OR_SPS; j++) {
            c[t][j] = c[t][j - 1] * cd - s[t][j - 1] * sd;
            s[t][j] = c[t][j - 1] * sd + s[t][j - 1] * cd;
        }
    }
}

void sync_and_demodulate(float *id, float *qd, long np, unsigned char *symbols, float *freq, int ifmin,
                         int ifmax, float fstep, int *shift, int lagmin, int lagmax, int lagstep,
                         float *drift, int symfac, float *sync, int mode) {
    float c[4][OR_SPS], s[4][OR_SPS];
    float soft[OR_NSYM];
    float best_f = 0.0, best_sync = -1e30;
    int best_lag = 0;

    if (mode == 0) { ifmin = ifmax = 0; fstep = 0.0; }
    else { lagmin = lagmax = *shift; if (mode == 2) ifmin = ifmax = 0; }

    for (int ifq = ifmin; ifq <= ifmax; ifq++) {
        float f0 = *freq + ifq * fstep;
        for (int lag = lagmin; lag <= lagmax; lag += lagstep) {
            float ss = 0.0, totp = 0.0, fplast = 0.0;
            for (int i = 0; i < OR_NSYM; i++) {
                float fp = f0 + (*drift / 2.0) * ((float)i - (float)OR_NBITS) / (float)OR_NBITS;
                if (i == 0 || fp != fplast) {       /* wsprd.c:157 (the static fplast only matters via i==0) */
                    tone_tables(fp, c, s);
                    fplast = fp;
                }
                float p[4];
                for (int t = 0; t < 4; t++) {
                    float ai = 0.0, aq = 0.0;
                    for (int j = 0; j < OR_SPS; j++) {
                        int k = lag + i * OR_SPS + j;
                        if (k > 0 && k < np) {
                            ai = ai + id[k] * c[t][j] + qd[k] * s[t][j];
                            aq = aq - id[k] * s[t][j] + qd[k] * c[t][j];
                        }
                    }
                    p[t] = sqrt(ai * ai + aq * aq);
                }
                totp = totp + p[0] + p[1] + p[2] + p[3];
                float cmet = (p[1] + p[3]) - (p[0] + p[2]);
                ss = SYNC_BITS[i] ? ss + cmet : ss - cmet;
                if (mode == 2) soft[i] = SYNC_BITS[i] ? p[3] - p[1] : p[2] - p[0];
            }
            ss = ss / totp;
            if (ss > best_sync) {
                best_sync = ss;
                best_lag = lag;
                best_f = f0;
            }
        }
    }
    *sync = best_sync;
    if (mode <= 1) {
        *shift = best_lag;
        *freq = best_f;
        return;
    }
    if (mode == 2) {                                /* wsprd.c:243-256 */
        float fsum = 0.0, f2sum = 0.0;
        for (int i = 0; i < OR_NSYM; i++) {
            fsum += soft[i] / OR_NSYM;
            f2sum += soft[i] * soft[i] / OR_NSYM;
        }
        float fac = sqrt(f2sum - fsum * fsum);
        for (int i = 0; i < OR_NSYM; i++) {
            float v = symfac * soft[i] / fac;
            if (v > 127) v = 127.0;
            if (v < -128) v = -128.0;
            symbols[i] = v + 128;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * subtract_signal2 (wsprd.c:316-413)
 * ---------------------------------------------------------------------------------------------- */
#define OR_NFILT 360
void subtract_signal2(float *id, float *qd, long np, float f0, int shift, float drift,
                      const unsigned char *channel_symbols) {
    enum { NS = OR_NSYM * OR_SPS, NTOT = OR_CAPTURE };
    float *buf = (float *)calloc(6 * (size_t)NTOT, sizeof(float));
    float *refi = buf, *refq = buf + NTOT, *ci = buf + 2 * NTOT, *cq = buf + 3 * NTOT,
          *cfi = buf + 4 * NTOT, *cfq = buf + 5 * NTOT;
    float phi = 0.0;
    for (int i = 0; i < OR_NSYM; i++) {             /* wsprd.c:340-351 */
        float cs = (float)channel_symbols[i];
        float dphi = OR_TWOPIDT * (f0 + (drift / 2.0) * ((float)i - (float)OR_NSYM / 2.0) / ((float)OR_NSYM / 2.0) +
                                   (cs - 1.5) * OR_DF);
        for (int j = 0; j < OR_SPS; j++) {
            refi[OR_SPS * i + j] = cosf(phi);
            refq[OR_SPS * i + j] = sinf(phi);
            phi = phi + dphi;
        }
    }
    float w[OR_NFILT], psum[OR_NFILT], norm = 0;    /* wsprd.c:353-368 */
    for (int i = 0; i < OR_NFILT; i++) {
        w[i] = sinf(M_PI * (float)i / (float)(OR_NFILT - 1));
        norm = norm + w[i];
    }
    for (int i = 0; i < OR_NFILT; i++) w[i] = w[i] / norm;
    psum[0] = 0.0;
    for (int i = 1; i < OR_NFILT; i++) psum[i] = psum[i - 1] + w[i];

    for (int i = 0; i < NS; i++) {                  /* wsprd.c:375-381 */
        int k = shift + i;
        if (k > 0 && k < np) {
            ci[i + OR_NFILT] = id[k] * refi[i] + qd[k] * refq[i];
            cq[i + OR_NFILT] = qd[k] * refi[i] - id[k] * refq[i];
        }
    }
    for (int i = OR_NFILT / 2; i < NTOT - OR_NFILT / 2; i++) {     /* wsprd.c:384-391 */
        float ai = 0.0, aq = 0.0;
        for (int j = 0; j < OR_NFILT; j++) {
            ai = ai + w[j] * ci[i - OR_NFILT / 2 + j];
            aq = aq + w[j] * cq[i - OR_NFILT / 2 + j];
        }
        cfi[i] = ai;
        cfq[i] = aq;
    }
    for (int i = 0; i < NS; i++) {                  /* wsprd.c:397-411 */
        if (i < OR_NFILT / 2) norm = psum[OR_NFILT / 2 + i];
        else if (i > NS - 1 - OR_NFILT / 2) norm = psum[OR_NFILT / 2 + NS - 1 - i];
        else norm = 1.0;
        int k = shift + i, j = i + OR_NFILT;
        if (k > 0 && k < np) {
            id[k] = id[k] - (cfi[j] * refi[i] - cfq[j] * refq[i]) / norm;
            qd[k] = qd[k] - (cfi[j] * refq[i] + cfq[j] * refi[i]) / norm;
        }
    }
    free(buf);
}

/* ------------------------------------------------------------------------------------------------
 * Spectrogram, candidate search, coarse sync (wsprd.c:509-678)
 * ---------------------------------------------------------------------------------------------- */
void oracle_mettab(int mettab[2][256]) {
    static const int T[2][256] = WSPR_METTAB_INIT;
    memcpy(mettab, T, sizeof T);
}

int oracle_blocks(int samples) { return 4 * floor(samples / OR_NFFT) - 1; }   /* wsprd.c:516 */

void oracle_window(float *win) {                                   /* wsprd.c:510-513 */
    for (int i = 0; i < OR_NFFT; i++) win[i] = sinf(0.006147931 * i);
}

void oracle_spectrogram(const float *idat, const float *qdat, int samples, float *ps) {   /* wsprd.c:536-553 */
    const int blocks = oracle_blocks(samples);
    float win[OR_NFFT], in[2 * OR_NFFT], out[2 * OR_NFFT];
    oracle_window(win);
    for (int b = 0; b < blocks; b++) {
        for (int j = 0; j < OR_NFFT; j++) {
            in[2 * j] = idat[b * 128 + j] * win[j];
            in[2 * j + 1] = qdat[b * 128 + j] * win[j];
        }
        oracle_dft512(in, out);
        for (int j = 0; j < OR_NFFT; j++) {
            int k = (j + OR_NFFT / 2) % OR_NFFT;
            ps[(size_t)j * blocks + b] = out[2 * k] * out[2 * k] + out[2 * k + 1] * out[2 * k + 1];
        }
    }
}

/* stable insertion sort, descending snr (the reference calls glibc qsort, a stable merge sort, with
 * cand_snr_desc, wsprd.c:47-51,631) */
static void sort_cands(struct cand *c, int n) {
    for (int i = 1; i < n; i++) {
        struct cand x = c[i];
        int j = i;
        while (j > 0 && c[j - 1].snr < x.snr) {
            c[j] = c[j - 1];
            j--;
        }
        c[j] = x;
    }
}
static int cmp_float_asc(const void *a, const void *b) {
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

int oracle_candidates(const float *ps, int blocks, int maxdrift, struct cand *cands, float *smspec_out) {
    float psavg[OR_NFFT];
    memset(psavg, 0, sizeof psavg);
    for (int b = 0; b < blocks; b++)                               /* wsprd.c:556-561 */
        for (int j = 0; j < OR_NFFT; j++) psavg[j] += ps[(size_t)j * blocks + b];

    float sm[411], sorted[411];                                    /* wsprd.c:565-583 */
    for (int i = 0; i < 411; i++) {
        sm[i] = 0.0;
        for (int j = -3; j <= 3; j++) sm[i] += 1 * psavg[256 - 205 + i + j];
        sorted[i] = sm[i];
    }
    qsort(sorted, 411, sizeof(float), cmp_float_asc);
    float noise_level = sorted[122];
    float min_snr = powf(10.0, -8.0 / 10.0);                       /* wsprd.c:590-597 */
    float snr_scaling_factor = 26.3;
    for (int j = 0; j < 411; j++) {
        sm[j] = sm[j] / noise_level - 1.0;
        if (sm[j] < min_snr) sm[j] = 0.1 * min_snr;
    }
    if (smspec_out) memcpy(smspec_out, sm, sizeof sm);

    memset(cands, 0, OR_MAXCAND * sizeof(struct cand));            /* wsprd.c:600-631 */
    int npk = 0;
    for (int j = 1; j < 410; j++) {
        if (sm[j] > sm[j - 1] && sm[j] > sm[j + 1] && npk < OR_MAXCAND) {
            cands[npk].freq = (j - 205) * (OR_DF / 2.0);
            cands[npk].snr = 10.0 * log10f(sm[j]) - snr_scaling_factor;
            npk++;
        }
    }
    int kept = 0;
    for (int j = 0; j < npk; j++)
        if (cands[j].freq >= -110.0f && cands[j].freq <= 110.0f) cands[kept++] = cands[j];
    npk = kept;
    sort_cands(cands, npk);

    for (int j = 0; j < npk; j++) {                                /* wsprd.c:646-678 */
        float sync = 0.0, sync_max = -1e30;
        int if0 = cands[j].freq / (OR_DF / 2.0) + OR_SPS;
        for (int ifr = if0 - 1; ifr <= if0 + 1; ifr++)
            for (int k0 = -10; k0 < 22; k0++)
                for (int idrift = -maxdrift; idrift <= maxdrift; idrift++) {
                    float ss = 0.0, pw = 0.0;
                    for (int k = 0; k < OR_NSYM; k++) {
                        int ifd = ifr + ((float)k - (float)OR_NBITS) / (float)OR_NBITS * ((float)idrift) / OR_DF;
                        int kx = k0 + 2 * k;
                        if (kx < blocks) {                          /* kx may be negative: flat indexing */
                            float p0 = sqrtf(ps[(long)(ifd - 3) * blocks + kx]);
                            float p1 = sqrtf(ps[(long)(ifd - 1) * blocks + kx]);
                            float p2 = sqrtf(ps[(long)(ifd + 1) * blocks + kx]);
                            float p3 = sqrtf(ps[(long)(ifd + 3) * blocks + kx]);
                            ss = ss + (2 * SYNC_BITS[k] - 1) * ((p1 + p3) - (p0 + p2));
                            pw = pw + p0 + p1 + p2 + p3;
                            sync = ss / pw;
                        }
                    }
                    if (sync > sync_max) {
                        sync_max = sync;
                        cands[j].shift = 128 * (k0 + 1);
                        cands[j].drift = idrift;
                        cands[j].freq = (ifr - OR_SPS) * (OR_DF / 2.0);
                        cands[j].sync = sync;
                    }
                }
    }
    return npk;
}

/* ------------------------------------------------------------------------------------------------
 * wspr_decode (wsprd.c:416-855).  File side effects (fftw_wisdom.dat; hashtable.txt) are kept only for
 * hashtable.txt, which changes results when options.usehashtable is set.
 * ---------------------------------------------------------------------------------------------- */
static void stable_sort_results(struct decoder_results *r, int n) {   /* wsprd.c:53-57,827 */
    for (int i = 1; i < n; i++) {
        struct decoder_results x = r[i];
        int j = i;
        while (j > 0 && r[j - 1].snr < x.snr) {
            r[j] = r[j - 1];
            j--;
        }
        r[j] = x;
    }
}

int wspr_decode(float *idat, float *qdat, int samples, struct decoder_options options,
                struct decoder_results *decodes, int *n_results) {
    const float minsync1 = 0.10;
    float minsync2 = 0.12;
    const int iifac = 3, symfac = 50, delta = 60, maxcycles = 10000;
    int maxdrift = 4;
    const float minrms = 52.0 * (symfac / 64.0);
    int uniques = 0;
    unsigned int metric, cycles = 0, maxnp;
    unsigned char symbols[OR_NSYM] = {0};
    unsigned char decdata[11] = {0};
    signed char message[12] = {0};
    char callsign[OR_HLEN] = {0}, call_loc_pow[23] = {0}, call[OR_HLEN] = {0}, loc[7] = {0}, pwr[3] = {0};
    float allfreqs[OR_MAXUNIQ] = {0};
    char allcalls[OR_MAXUNIQ][OR_HLEN];
    memset(allcalls, 0, sizeof allcalls);
    int mettab[2][256];
    oracle_mettab(mettab);

    char *hashtab = (char *)calloc(OR_HASHN * OR_HLEN, 1);
    char *loctab = (char *)calloc(OR_HASHN * OR_LLEN, 1);
    if (options.usehashtable) {                                    /* wsprd.c:481-494 */
        FILE *fh = fopen("hashtable.txt", "r+");
        if (fh) {
            char line[80], hcall[80], hgrid[80];
            int nh;
            while (fgets(line, sizeof line, fh)) {
                hgrid[0] = 0;
                hcall[0] = 0;
                nh = -1;
                sscanf(line, "%d %12s %4s", &nh, hcall, hgrid);
                if (nh >= 0 && nh < OR_HASHN) {
                    snprintf(hashtab + nh * OR_HLEN, OR_HLEN, "%s", hcall);
                    if (strlen(hgrid) > 0) snprintf(loctab + nh * OR_LLEN, OR_LLEN, "%s", hgrid);
                }
            }
            fclose(fh);
        }
    }

    const int blocks = oracle_blocks(samples);
    float *ps = (float *)calloc((size_t)OR_NFFT * (blocks > 0 ? blocks : 1), sizeof(float));
    struct cand cands[OR_MAXCAND];

    for (int ipass = 0; ipass < options.npasses; ipass++) {
        if (ipass == 1 && uniques == 0) break;
        if (ipass < 2) { maxdrift = 4; minsync2 = 0.12; }
        if (ipass == 2) { maxdrift = 0; minsync2 = 0.10; }

        oracle_spectrogram(idat, qdat, samples, ps);
        int npk = oracle_candidates(ps, blocks, maxdrift, cands, NULL);

        for (int j = 0; j < npk; j++) {                            /* wsprd.c:697-823 */
            oracle_stat[0 + (ipass > 0)]++;                        /* candidates examined, per pass */
            memset(callsign, 0, 13);
            memset(call_loc_pow, 0, 23);
            memset(call, 0, 13);
            memset(loc, 0, 7);
            memset(pwr, 0, 3);
            float freq = cands[j].freq, drift = cands[j].drift, sync = cands[j].sync;
            int shift = cands[j].shift;
            int lagstep = options.quickmode ? 16 : 8;
            sync_and_demodulate(idat, qdat, samples, symbols, &freq, 0, 0, 0.0, &shift, shift - 128, shift + 128,
                                lagstep, &drift, symfac, &sync, 0);
            float fstep = 0.1;
            sync_and_demodulate(idat, qdat, samples, symbols, &freq, -2, 2, fstep, &shift, 0, 0, lagstep, &drift,
                                symfac, &sync, 1);
            cands[j].freq = freq;
            cands[j].shift = shift;
            cands[j].drift = drift;
            cands[j].sync = sync;
            int worth = sync > minsync1;
            oracle_stat[2 + (ipass > 0)] += worth;                 /* passed the minsync1 gate */
            int idt = 0, ii = 0, not_decoded = 1;
            while (worth && not_decoded && idt <= (128 / iifac)) {
                ii = (idt + 1) / 2;
                if (idt % 2 == 1) ii = -ii;
                ii = iifac * ii;
                int jig = shift + ii;
                sync_and_demodulate(idat, qdat, samples, symbols, &freq, -2, 2, fstep, &jig, 0, 0, lagstep, &drift,
                                    symfac, &sync, 2);
                float sq = 0.0;
                for (int i = 0; i < OR_NSYM; i++) {
                    float y = (float)symbols[i] - 128.0;
                    sq += y * y;
                }
                float rms = sqrtf(sq / (float)OR_NSYM);
                if (sync > minsync2 && rms > minrms) {
                    deinterleave(symbols);
                    not_decoded = fano(&metric, &cycles, &maxnp, decdata, symbols, OR_NBITS, mettab, delta, maxcycles);
                    oracle_stat[4]++;                              /* Fano calls */
                    oracle_stat[5] += not_decoded != 0;            /* ... that timed out */
                    if (!not_decoded) {
                        oracle_stat[6 + (idt > 0)]++;              /* decoded at jitter 0 / at a later jitter */
                        if (cycles > 4096) oracle_stat[8]++;       /* successes needing more than 4096 / 32768 cycles */
                        if (cycles > 32768) oracle_stat[9]++;
                        if (idt > 0) oracle_stat[12] += idt;       /* sum of winning idt */
                    }
                }
                idt++;
                if (options.quickmode) break;
            }
            if (worth && not_decoded) oracle_stat[10 + (ipass > 0)]++;   /* worth a try, never decoded (per pass) */
            if (!(worth && !not_decoded)) continue;

            for (int i = 0; i < 11; i++) message[i] = (signed char)(decdata[i] > 127 ? decdata[i] - 256 : decdata[i]);
            int noprint = unpk_(message, hashtab, loctab, call_loc_pow, call, loc, pwr, callsign);
            if (options.subtraction && ipass == 0 && !noprint) {
                unsigned char chan[OR_NSYM];
                if (get_wspr_channel_symbols(call_loc_pow, hashtab, loctab, chan))
                    subtract_signal2(idat, qdat, samples, freq, shift, drift, chan);
                else
                    break;
            }
            if (!strcmp(loc, "A000AA")) break;
            int dupe = 0;
            for (int i = 0; i < uniques; i++)
                if (!strcmp(callsign, allcalls[i]) && fabs(freq - allfreqs[i]) < 3.0) dupe = 1;
            if (dupe) continue;
            if (uniques >= OR_MAXUNIQ) continue;                   /* the reference would overflow its arrays */
            snprintf(allcalls[uniques], OR_HLEN, "%s", callsign);
            allfreqs[uniques] = freq;
            struct decoder_results *r = &decodes[uniques++];
            double dialfreq = (double)options.freq / 1e6;
            r->sync = cands[j].sync;
            r->snr = cands[j].snr;
            r->dt = shift * OR_DT - 2.0;
            r->freq = dialfreq + (1500.0 + freq) / 1e6;
            r->drift = drift;
            r->cycles = cycles;
            r->jitter = ii;
            snprintf(r->message, sizeof r->message, "%s", call_loc_pow);
            snprintf(r->call, sizeof r->call, "%s", call);
            snprintf(r->loc, sizeof r->loc, "%s", loc);
            snprintf(r->pwr, sizeof r->pwr, "%s", pwr);
        }
    }
    stable_sort_results(decodes, uniques);
    *n_results = uniques;

    if (options.usehashtable) {                                    /* wsprd.c:842-852 */
        FILE *fh = fopen("hashtable.txt", "w");
        if (fh) {
            for (int i = 0; i < OR_HASHN; i++)
                if (hashtab[i * OR_HLEN]) fprintf(fh, "%5d %s %s\n", i, hashtab + i * OR_HLEN, loctab + i * OR_LLEN);
            fclose(fh);
        }
    }
    free(ps);
    free(hashtab);
    free(loctab);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Front end (rtlsdr_wsprd.c:126-244) and hand-off normalisation (:285-305, same as :575-589)
 * ---------------------------------------------------------------------------------------------- */
/* FIR coefficients are data of the reference (rtlsdr_wsprd.c:142-152): a symmetric 33-tap compensation
 * filter; stored here as the 17 distinct values. */
static const float FIR_HALF[17] = {
    -0.0027772683, -0.0005058826, 0.0049745750, -0.0034059318, -0.0077557814, 0.0139375423,
    0.0039896935,  -0.0299394142, 0.0162250643, 0.0405130860,  -0.0580746013, -0.0272104968,
    0.1183705475,  -0.0306029022, -0.2011241667, 0.1615898423, 0.5000000000};

int oracle_decimate(const uint8_t *raw, size_t n_iq, float *i_out, float *q_out, int max_out) {
    float z[33], hi[32], hq[32];
    for (int j = 0; j < 33; j++) z[j] = FIR_HALF[j <= 16 ? j : 32 - j];
    memset(hi, 0, sizeof hi);
    memset(hq, 0, sizeof hq);
    uint32_t acc1[2] = {0, 0}, acc2[2] = {0, 0};      /* integrators (wrap like the reference's int32) */
    uint32_t c1[2][2] = {{0, 0}, {0, 0}}, c2[2][2] = {{0, 0}, {0, 0}};   /* comb delay lines [ch][y,z] */
    uint32_t dec = 0;
    int nout = 0;
    for (size_t n = 0; n < n_iq; n++) {
        int8_t a = (int8_t)(raw[2 * n] ^ 0x80), b = (int8_t)(raw[2 * n + 1] ^ 0x80);
        int8_t x[2];
        switch (n & 3) {                              /* fs/4 rotation, rtlsdr_wsprd.c:171-182 */
            case 0: x[0] = a; x[1] = b; break;
            case 1: x[0] = (int8_t)(-b); x[1] = a; break;
            case 2: x[0] = (int8_t)(-a); x[1] = (int8_t)(-b); break;
            default: x[0] = b; x[1] = (int8_t)(-a); break;
        }
        for (int ch = 0; ch < 2; ch++) {
            acc1[ch] += (uint32_t)(int32_t)x[ch];
            acc2[ch] += acc1[ch];
        }
        dec++;
        if (dec <= 2400000 / 375) continue;           /* one output per 6401 inputs, :198-202 */
        dec = 0;
        float fresh[2];
        for (int ch = 0; ch < 2; ch++) {              /* two combs with delay 2, :204-218 */
            uint32_t y1 = acc2[ch] - c1[ch][1];
            c1[ch][1] = c1[ch][0];
            c1[ch][0] = acc2[ch];
            uint32_t y2 = y1 - c2[ch][1];
            c2[ch][1] = c2[ch][0];
            c2[ch][0] = y1;
            fresh[ch] = (float)(int32_t)y2;
        }
        float si = 0.0, sq = 0.0;                     /* FIR, :221-234 */
        for (int j = 0; j < 32; j++) {
            si += hi[j] * z[j];
            sq += hq[j] * z[j];
        }
        memmove(hi, hi + 1, 31 * sizeof(float));
        memmove(hq, hq + 1, 31 * sizeof(float));
        hi[31] = fresh[0];
        hq[31] = fresh[1];
        si += hi[31] * z[32];
        sq += hq[31] * z[32];
        if (nout < max_out) {
            i_out[nout] = si;
            q_out[nout] = sq;
            nout++;
        }
    }
    return nout;
}

void oracle_normalise(float *idat, float *qdat, int n) {
    float m = 1e-24f;
    for (int i = 0; i < n; i++) {
        float ai = fabs(idat[i]), aq = fabs(qdat[i]);
        if (ai > m) m = ai;
        if (aq > m) m = aq;
    }
    m = 0.5 / m;
    for (int i = 0; i < n; i++) {
        idat[i] *= m;
        qdat[i] *= m;
    }
}
