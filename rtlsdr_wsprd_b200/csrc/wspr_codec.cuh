// WSPR message codec, convolutional code and Fano sequential decoder as host+device inline functions.
//
// These are the integer/byte stages of the decode path.  On the GPU they run
// inside the per-capture resolve kernel and the one-thread-per-attempt Fano kernel; the same functions back the
// C-ABI exports that mirror the reference's codec entry points (include/wspr_b200.h), so the reference's unit
// tests can link against this library.  No libc string functions are used (device code); the small helpers
// below reproduce the truncation semantics of the snprintf() calls in the reference.
//
// Behaviour follows (file:line relative to the reference checkout):
//   nhash            wsprd/nhash.c:205-451          unpack50/unpackcall/unpackgrid/unpackpfx/unpk_
//   ENCODE / encode  wsprd/fano.h:35-44, fano.c:63-82                    wsprd/wsprd_utils.c:40-194,228-313
//   fano             wsprd/fano.c:87-238            deinterleave      wsprd/wsprd_utils.c:196-213
//   pack_*, interleave, get_wspr_channel_symbols    wsprd/wsprsim_utils.c:15-316
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __CUDACC__
#define WHD __host__ __device__ __forceinline__
#define WHD_NOINLINE inline __host__ __device__
#else
#define WHD inline
#define WHD_NOINLINE inline
#endif

namespace wspr {

constexpr int NSYM = 162;
constexpr int NBITS = 81;
constexpr int SPS = 256;
constexpr int HASH_SLOTS = 32768;
constexpr int CALL_LEN = 13;
constexpr int LOC_LEN = 5;

// 162-bit sync vector (wsprd.c:84-93), packed LSB-first into six words.
WHD_NOINLINE unsigned sync_bit(int i) {
    const uint32_t w[6] = {0x07a47103u, 0x58b340a4u, 0x56349558u, 0xe2cdc904u, 0x63580ca0u, 0x00000000u};
    return (w[i >> 5] >> (i & 31)) & 1u;
}

// ---------------------------------------------------------------------------------------------------------
// tiny string helpers (NUL-terminated char arrays)
// ---------------------------------------------------------------------------------------------------------
WHD int s_len(const char *s) {
    int n = 0;
    while (s[n]) n++;
    return n;
}
WHD bool s_eq(const char *a, const char *b) {
    int i = 0;
    while (a[i] && a[i] == b[i]) i++;
    return a[i] == b[i];
}
// first index whose char is in `set` (or the length): strcspn
WHD int s_cspn(const char *s, const char *set) {
    int i = 0;
    for (; s[i]; i++)
        for (int k = 0; set[k]; k++)
            if (s[i] == set[k]) return i;
    return i;
}
WHD bool c_alpha(char c) { return (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'); }
WHD bool c_digit(char c) { return c >= '0' && c <= '9'; }

// bounded string builder: the result is what snprintf(dst, cap, ...) would leave (at most cap-1 chars + NUL)
struct StrOut {
    char *dst;
    int cap, n;
    WHD StrOut(char *d, int c) : dst(d), cap(c), n(0) {
        if (cap > 0) dst[0] = 0;
    }
    WHD void ch(char c) {
        if (n < cap - 1) {
            dst[n++] = c;
            dst[n] = 0;
        }
    }
    WHD void str(const char *s) {
        for (int i = 0; s[i]; i++) ch(s[i]);
    }
    WHD void strn(const char *s, int maxn) {
        for (int i = 0; i < maxn && s[i]; i++) ch(s[i]);
    }
};
WHD void s_copy(char *dst, int cap, const char *src) {
    StrOut o(dst, cap);
    o.str(src);
}
// "%02d" / "%2d" for 0 <= v <= 99 (wider values print all their digits like printf)
WHD void fmt_int2(char *dst, int cap, int v, bool zero_pad) {
    char tmp[12];
    int n = 0;
    bool neg = v < 0;
    unsigned u = neg ? (unsigned)(-v) : (unsigned)v;
    do {
        tmp[n++] = (char)('0' + u % 10);
        u /= 10;
    } while (u);
    StrOut o(dst, cap);
    int width = n + (neg ? 1 : 0);
    if (zero_pad) {
        if (neg) o.ch('-');
        for (int i = width; i < 2; i++) o.ch('0');
    } else {
        for (int i = width; i < 2; i++) o.ch(' ');
        if (neg) o.ch('-');
    }
    while (n) o.ch(tmp[--n]);
}
// strtok(): skip leading delimiters, cut at the next one.  *save carries the scan position between calls.
WHD char *s_tok(char **save, const char *delims) {
    char *p = *save;
    if (!p) return nullptr;
    auto is_delim = [&](char c) {
        for (int k = 0; delims[k]; k++)
            if (c == delims[k]) return true;
        return false;
    };
    while (*p && is_delim(*p)) p++;
    if (!*p) {
        *save = nullptr;
        return nullptr;
    }
    char *tok = p;
    while (*p && !is_delim(*p)) p++;
    if (*p) {
        *p = 0;
        *save = p + 1;
    } else {
        *save = nullptr;
    }
    return tok;
}
WHD int s_atoi(const char *s) {
    int i = 0, sign = 1, v = 0;
    while (s[i] == ' ' || (s[i] >= 9 && s[i] <= 13)) i++;
    if (s[i] == '-') { sign = -1; i++; }
    else if (s[i] == '+') i++;
    while (c_digit(s[i])) v = v * 10 + (s[i++] - '0');
    return sign * v;
}

// ---------------------------------------------------------------------------------------------------------
// nhash: lookup3 hashlittle, 15-bit result
// ---------------------------------------------------------------------------------------------------------
WHD uint32_t rotl32(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }

WHD_NOINLINE uint32_t nhash15(const void *key, size_t length, uint32_t initval) {
    const uint8_t *p = (const uint8_t *)key;
    uint32_t a, b, c;
    a = b = c = 0xdeadbeefu + (uint32_t)length + initval;
    while (length > 12) {
        a += p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
        b += p[4] | (uint32_t)p[5] << 8 | (uint32_t)p[6] << 16 | (uint32_t)p[7] << 24;
        c += p[8] | (uint32_t)p[9] << 8 | (uint32_t)p[10] << 16 | (uint32_t)p[11] << 24;
        a -= c; a ^= rotl32(c, 4);  c += b;
        b -= a; b ^= rotl32(a, 6);  a += c;
        c -= b; c ^= rotl32(b, 8);  b += a;
        a -= c; a ^= rotl32(c, 16); c += b;
        b -= a; b ^= rotl32(a, 19); a += c;
        c -= b; c ^= rotl32(b, 4);  b += a;
        length -= 12;
        p += 12;
    }
    if (length == 0) return c;   // reference quirk: the empty tail skips the final mix and the mask
    uint32_t w0 = 0, w1 = 0, w2 = 0;
    for (size_t i = 0; i < length; i++) {
        uint32_t v = (uint32_t)p[i] << (8 * (i & 3));
        if (i < 4) w0 += v;
        else if (i < 8) w1 += v;
        else w2 += v;
    }
    a += w0; b += w1; c += w2;
    c ^= b; c -= rotl32(b, 14);
    a ^= c; a -= rotl32(c, 11);
    b ^= a; b -= rotl32(a, 25);
    c ^= b; c -= rotl32(b, 16);
    a ^= c; a -= rotl32(c, 4);
    b ^= a; b -= rotl32(a, 14);
    c ^= b; c -= rotl32(b, 24);
    return c & 32767u;
}

// ---------------------------------------------------------------------------------------------------------
// K=32 r=1/2 convolutional code (Layland-Lushbaugh polynomials)
// ---------------------------------------------------------------------------------------------------------
constexpr uint32_t POLY_A = 0xf2d05351u;
constexpr uint32_t POLY_B = 0xe4613c47u;

WHD unsigned parity_u32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __popc(v) & 1u;
#else
    return (unsigned)__builtin_parity(v);
#endif
}
// 2-bit branch symbol for an encoder state: bit1 from POLY_A, bit0 from POLY_B
WHD unsigned branch_sym(uint32_t state) { return (parity_u32(state & POLY_A) << 1) | parity_u32(state & POLY_B); }

WHD void conv_encode(unsigned char *symbols, const unsigned char *data, unsigned nbytes) {
    uint32_t st = 0;
    for (unsigned b = 0; b < nbytes; b++)
        for (int bit = 7; bit >= 0; bit--) {
            st = (st << 1) | ((data[b] >> bit) & 1u);
            unsigned s = branch_sym(st);
            *symbols++ = (unsigned char)(s >> 1);
            *symbols++ = (unsigned char)(s & 1u);
        }
}

// interleaver order: p-th kept value of the bit-reversed 8-bit counter
WHD int bitrev8(int v) { return ((v * 0x0802u & 0x22110u) | (v * 0x8020u & 0x88440u)) * 0x10101u >> 16 & 0xff; }

WHD void deinterleave162(unsigned char *sym) {
    unsigned char t[NSYM];
    int p = 0;
    for (int v = 0; p < NSYM; v++) {
        int r = bitrev8(v);
        if (r < NSYM) t[p++] = sym[r];
    }
    for (int i = 0; i < NSYM; i++) sym[i] = t[i];
}
WHD void interleave162(unsigned char *sym) {
    unsigned char t[NSYM];
    int p = 0;
    for (int v = 0; p < NSYM; v++) {
        int r = bitrev8(v);
        if (r < NSYM) t[r] = sym[p++];
    }
    for (int i = 0; i < NSYM; i++) sym[i] = t[i];
}

// ---------------------------------------------------------------------------------------------------------
// Fano sequential decoder.  Same moves and cycle accounting as the reference; state kept in small arrays.
// Returns 0 on success, -1 on timeout (cycle counter reached maxcycles*nbits).
// ---------------------------------------------------------------------------------------------------------
constexpr int FANO_MAXBITS = 128;
constexpr int FANO_STOPPED = 2;      // return value: the run was cut short (budget reached or poll() asked to stop)

// poll hook: called every 1024 cycles; returning true abandons the run (result FANO_STOPPED)
struct FanoNoPoll {
    WHD bool operator()(unsigned) const { return false; }
};

// stop_after: 0 = run to the reference's limit (maxcycles*nbits); otherwise give up with FANO_STOPPED once that many
// cycles have been spent without finishing (a later full run from scratch reproduces the identical result).
template <typename MetT, typename Poll = FanoNoPoll>
WHD_NOINLINE int fano_decode(unsigned *metric_out, unsigned *cycles_out, unsigned *maxnp_out, unsigned char *data,
                             const unsigned char *symbols, unsigned nbits, const MetT *mettab /*[2][256]*/, int delta,
                             unsigned maxcycles, unsigned stop_after = 0, Poll poll = Poll()) {
    if (nbits > (unsigned)FANO_MAXBITS - 1 || nbits < 32) return -1;
    uint32_t enc[FANO_MAXBITS];
    int gam[FANO_MAXBITS];
    short bm[FANO_MAXBITS][4];
    short tm[FANO_MAXBITS][2];
    unsigned char sel[FANO_MAXBITS];
    const int last = (int)nbits - 1, tail = (int)nbits - 31;
    for (unsigned n = 0; n <= nbits; n++) enc[n] = 0;   // nodes never reached read as zero in the output bytes
    for (unsigned n = 0; n < nbits; n++) {
        int a0 = mettab[symbols[2 * n]], a1 = mettab[256 + symbols[2 * n]];
        int b0 = mettab[symbols[2 * n + 1]], b1 = mettab[256 + symbols[2 * n + 1]];
        bm[n][0] = (short)(a0 + b0);
        bm[n][1] = (short)(a0 + b1);
        bm[n][2] = (short)(a1 + b0);
        bm[n][3] = (short)(a1 + b1);
    }
    int pos = 0, thr = 0, maxnp = 0;
    enc[0] = 0;
    {
        unsigned ls = branch_sym(0);
        int m0 = bm[0][ls], m1 = bm[0][3 ^ ls];
        if (m0 > m1) { tm[0][0] = (short)m0; tm[0][1] = (short)m1; }
        else { tm[0][0] = (short)m1; tm[0][1] = (short)m0; enc[0] = 1; }
    }
    sel[0] = 0;
    gam[0] = 0;
    const unsigned limit = maxcycles * nbits;
    const unsigned stop = (stop_after != 0 && stop_after < limit) ? stop_after : 0;
    unsigned it;
    bool stopped = false;
    for (it = 1; it <= limit; it++) {
        if ((it & 1023u) == 0 && ((stop != 0 && it >= stop) || poll(it))) {
            stopped = true;
            break;
        }
        if (pos > maxnp) maxnp = pos;
        int ng = gam[pos] + tm[pos][sel[pos]];
        if (ng >= thr) {
            if (gam[pos] < thr + delta)
                while (ng >= thr + delta) thr += delta;
            gam[pos + 1] = ng;
            uint32_t e = enc[pos] << 1;
            pos++;
            if (pos == last + 1) {
                enc[pos] = e;
                break;
            }
            unsigned ls = branch_sym(e);
            if (pos >= tail) {
                tm[pos][0] = bm[pos][ls];
            } else {
                int m0 = bm[pos][ls], m1 = bm[pos][3 ^ ls];
                if (m0 > m1) { tm[pos][0] = (short)m0; tm[pos][1] = (short)m1; }
                else { tm[pos][0] = (short)m1; tm[pos][1] = (short)m0; e |= 1u; }
            }
            enc[pos] = e;
            sel[pos] = 0;
            continue;
        }
        for (;;) {
            if (pos == 0 || gam[pos - 1] < thr) {
                thr -= delta;
                if (sel[pos] != 0) {
                    sel[pos] = 0;
                    enc[pos] ^= 1u;
                }
                break;
            }
            pos--;
            if (pos < tail && sel[pos] != 1) {
                sel[pos]++;
                enc[pos] ^= 1u;
                break;
            }
        }
    }
    *metric_out = (unsigned)gam[pos];
    for (unsigned b = 0; b < (nbits >> 3); b++) data[b] = (unsigned char)enc[7 + 8 * b];
    *cycles_out = it + 1;
    *maxnp_out = (unsigned)maxnp;
    if (stopped) return FANO_STOPPED;
    return (it >= limit) ? -1 : 0;
}

// ---------------------------------------------------------------------------------------------------------
// callsign hash stores.  The reference keeps a dense 32768 x 13 char table per wspr_decode() call
// (wsprd.c:478-479); on the device a capture only ever inserts a few dozen entries, so a short list with
// last-writer-wins lookup is equivalent and 400x smaller.
// ---------------------------------------------------------------------------------------------------------
struct DenseHashStore {
    char *calls;   // [32768][13]
    char *locs;    // [32768][5]
    WHD const char *get(int h) const { return calls + (size_t)h * CALL_LEN; }
    WHD void put_call(int h, const char *s) { s_copy(calls + (size_t)h * CALL_LEN, CALL_LEN, s); }
    WHD void put_loc(int h, const char *s) { s_copy(locs + (size_t)h * LOC_LEN, LOC_LEN, s); }
};
struct HashEntry {
    int h;
    char call[CALL_LEN];
    char loc[LOC_LEN];             // locator stored with the call by type-1 messages ("" = none); only read by the -H write-back
    char pad[2];
};
struct ListHashStore {
    HashEntry *e;
    int *count;
    int cap;
    const char *preload;           // [32768][13] calls read from hashtable.txt (reference -H, wsprd.c:481-494), or null
    int *overflow;                 // set when an entry had to be dropped because the list is full (the decode reports it)
    WHD const char *get(int h) const {
        for (int i = *count - 1; i >= 0; i--)
            if (e[i].h == h) return e[i].call;
        return preload ? preload + (size_t)h * CALL_LEN : "";
    }
    WHD HashEntry *slot(int h) {
        for (int i = 0; i < *count; i++)
            if (e[i].h == h) return &e[i];
        if (*count >= cap) {
            if (overflow) *overflow = 1;
            return nullptr;
        }
        HashEntry *n = &e[(*count)++];
        n->h = h;
        n->call[0] = 0;
        n->loc[0] = 0;
        return n;
    }
    WHD void put_call(int h, const char *s) {
        if (HashEntry *n = slot(h)) s_copy(n->call, CALL_LEN, s);
    }
    WHD void put_loc(int h, const char *s) {   // (always follows put_call of the same hash, wsprd_utils.c:258-259)
        if (HashEntry *n = slot(h)) s_copy(n->loc, LOC_LEN, s);
    }
};

// ---------------------------------------------------------------------------------------------------------
// unpacking
// ---------------------------------------------------------------------------------------------------------
WHD char alnum37(int i) {
    return (i < 10) ? (char)('0' + i) : (i < 36) ? (char)('A' + i - 10) : ' ';
}

WHD void unpack_50(const signed char *dat, int32_t *n1, int32_t *n2) {
    uint32_t b[7];
    for (int i = 0; i < 7; i++) b[i] = (uint32_t)(dat[i] & 255);
    *n1 = (int32_t)((b[0] << 20) + (b[1] << 12) + (b[2] << 4) + ((b[3] >> 4) & 15));
    *n2 = (int32_t)(((b[3] & 15) << 18) + (b[4] << 10) + (b[5] << 2) + ((b[6] >> 6) & 3));
}

WHD int unpack_call(int32_t ncall, char *call /*[13]*/) {
    s_copy(call, 13, "......");
    if (ncall >= 262177560) return 0;
    char t[7];
    int32_t n = ncall;
    t[5] = alnum37(n % 27 + 10); n /= 27;
    t[4] = alnum37(n % 27 + 10); n /= 27;
    t[3] = alnum37(n % 27 + 10); n /= 27;
    t[2] = alnum37(n % 10);      n /= 10;
    t[1] = alnum37(n % 36);      n /= 36;
    t[0] = alnum37(n);
    t[6] = 0;
    int lead = 0;
    while (lead < 5 && t[lead] == ' ') lead++;
    // "%-6s": left-justified in a field of six
    StrOut o(call, 13);
    o.str(t + lead);
    while (o.n < 6) o.ch(' ');
    for (int i = 0; i < 6; i++)
        if (call[i] == ' ') call[i] = 0;
    return 1;
}

WHD int unpack_grid(int32_t ngrid, char *grid /*[5], only [0..3] written on success*/) {
    ngrid >>= 7;
    if (ngrid >= 32400) {
        s_copy(grid, 5, "XXXX");
        return 0;
    }
    int dlat = ngrid % 180 - 90;
    int dlong = (ngrid / 180) * 2 - 180 + 2;
    if (dlong < -180) dlong += 360;
    if (dlong > 180) dlong += 360;
    int nlong = (int)(60.0 * (180.0 - dlong) / 5.0);
    int nlat = (int)(60.0 * (dlat + 90) / 2.5);
    int a = nlong / 240, b = nlat / 240;
    grid[0] = alnum37(10 + a);
    grid[2] = alnum37((nlong - 240 * a) / 24);
    grid[1] = alnum37(10 + b);
    grid[3] = alnum37((nlat - 240 * b) / 24);
    return 1;
}

WHD int unpack_pfx(int32_t nprefix, char *call /*[13]*/) {
    char base[13];
    s_copy(base, 13, call);
    if (nprefix < 60000) {
        char pfx[4] = {0, 0, 0, 0};
        int32_t n = nprefix;
        for (int i = 2; i >= 0; i--) {
            int nc = n % 37;
            pfx[i] = (nc <= 9) ? (char)(nc + 48) : (nc <= 35) ? (char)(nc + 55) : ' ';
            n /= 37;
        }
        int start = 0;   // text after the last blank
        for (int i = 0; i < 3; i++)
            if (pfx[i] == ' ') start = i + 1;
        StrOut o(call, 13);
        o.str(pfx + start);
        o.ch('/');
        o.str(base);
        return 1;
    }
    int nc = (int)(signed char)(nprefix - 60000);   // the reference narrows to char here
    StrOut o(call, 13);
    if (nc >= 0 && nc <= 9) {
        o.str(base); o.ch('/'); o.ch((char)(nc + 48));
    } else if (nc >= 10 && nc <= 35) {
        o.str(base); o.ch('/'); o.ch((char)(nc + 55));
    } else if (nc >= 36 && nc <= 125) {
        o.str(base); o.ch('/'); o.ch((char)((nc - 26) / 10 + 48)); o.ch((char)((nc - 26) % 10 + 48));
    } else {
        s_copy(call, 13, base);
        return 0;
    }
    return 1;
}

WHD bool pwr_digit_ok(int nu) { return nu == 0 || nu == 3 || nu == 7; }

// Returns the reference's `noprint`.  Outputs are only written on the paths where the reference writes them.
template <class Store>
WHD_NOINLINE int unpack_message(const signed char *message, Store &hs, char *call_loc_pow /*[23]*/, char *call /*[13]*/,
                                char *loc /*[7]*/, char *pwr /*[3]*/, char *callsign /*[13]*/) {
    int32_t n1, n2;
    char grid[5], cdbm[4];
    int noprint = 0;
    unpack_50(message, &n1, &n2);
    if (!unpack_call(n1, callsign)) return 1;
    if (!unpack_grid(n2, grid)) return 1;
    int ntype = (n2 & 127) - 64;
    callsign[12] = 0;
    grid[4] = 0;
    if (ntype >= 0 && ntype <= 62) {
        int nu = ntype % 10;
        if (pwr_digit_ok(nu)) {   // type 1: CALL GRID4 PWR
            fmt_int2(cdbm, 4, ntype, true);
            StrOut o(call_loc_pow, 23);
            o.str(callsign); o.ch(' '); o.str(grid); o.ch(' '); o.str(cdbm);
            int h = (int)nhash15(callsign, (size_t)s_len(callsign), 146u);
            hs.put_call(h, callsign);
            hs.put_loc(h, grid);
            s_copy(call, CALL_LEN, callsign);
            s_copy(loc, 7, grid);
            s_copy(pwr, 3, cdbm);
        } else {                  // type 2: compound callsign + PWR
            int nadd = nu;
            if (nu > 3) nadd = nu - 3;
            if (nu > 7) nadd = nu - 7;
            int n3 = n2 / 128 + HASH_SLOTS * (nadd - 1);
            if (!unpack_pfx(n3, callsign)) return 1;
            int ndbm = ntype - nadd;
            fmt_int2(cdbm, 4, ndbm, false);
            StrOut o(call_loc_pow, 23);
            o.str(callsign); o.ch(' '); o.str(cdbm);
            if (pwr_digit_ok(ndbm % 10)) {
                int h = (int)nhash15(callsign, (size_t)s_len(callsign), 146u);
                hs.put_call(h, callsign);
            } else {
                noprint = 1;
            }
        }
    } else if (ntype < 0) {       // type 3: <hashed call> GRID6 PWR
        int ndbm = -(ntype + 1);
        char grid6[7] = {0, 0, 0, 0, 0, 0, 0};
        {   // "%c%.5s": callsign[5] (possibly NUL, which still occupies a byte) then up to five chars
            int n = 0;
            grid6[n++] = callsign[5];
            for (int i = 0; i < 5 && callsign[i]; i++) grid6[n++] = callsign[i];
            grid6[n] = 0;
        }
        if (!pwr_digit_ok(ndbm % 10) || !c_alpha(grid6[0]) || !c_alpha(grid6[1]) || !c_digit(grid6[2]) ||
            !c_digit(grid6[3]))
            noprint = 1;
        int h = (n2 - ntype - 64) / 128;
        const char *known = hs.get(h);
        {
            char tmp[CALL_LEN];
            StrOut o(tmp, CALL_LEN);
            if (known[0]) { o.ch('<'); o.str(known); o.ch('>'); }
            else o.str("<...>");
            s_copy(callsign, CALL_LEN, tmp);
        }
        fmt_int2(cdbm, 4, ndbm, false);
        StrOut o(call_loc_pow, 23);
        o.str(callsign); o.ch(' '); o.str(grid6); o.ch(' '); o.str(cdbm);
        s_copy(call, CALL_LEN, callsign);
        s_copy(loc, 7, grid6);
        s_copy(pwr, 3, cdbm);
        if (ntype == -64) noprint = 1;
    }
    return noprint;
}

// ---------------------------------------------------------------------------------------------------------
// packing
// ---------------------------------------------------------------------------------------------------------
WHD int loc_code(char ch) {
    if (ch >= '0' && ch <= '9') return ch - '0';
    if (ch == ' ') return 36;
    if (ch >= 'A' && ch <= 'R') return ch - 'A';
    return -1;
}
WHD int call_code(char ch) {
    if (ch >= '0' && ch <= '9') return ch - '0';
    if (ch == ' ') return 36;
    if (ch >= 'A' && ch <= 'Z') return ch - 'A' + 10;
    return -1;
}
// codes are the (signed char) values loc_code() returns; the arithmetic is int, widened to 64 bits like the
// reference's `long unsigned`
WHD uint64_t pack_grid_power(const signed char *g, int power) {
    uint64_t m = (uint64_t)(int64_t)((179 - 10 * g[0] - g[2]) * 180 + 10 * g[1] + g[3]);
    return m * 128 + (uint64_t)(int64_t)power + 64;
}
WHD uint64_t pack_callsign(const char *cs) {
    int len = s_len(cs);
    if (len > 6) return 0;
    char c6[8] = {' ', ' ', ' ', ' ', ' ', ' ', ' ', ' '};
    // cs[2] / cs[1] are read even past a short string's terminator in the reference; mimic with 0 there
    char c1 = len >= 1 ? cs[1] : 0, c2 = len >= 2 ? cs[2] : 0;
    if (c_digit(c2)) {
        for (int i = 0; i < len; i++) c6[i] = cs[i];
    } else if (c_digit(c1)) {
        for (int i = 1; i < len + 1 && i < 7; i++) c6[i] = cs[i - 1];
    }
    uint64_t n = (uint64_t)(int64_t)call_code(c6[0]);
    n = n * 36 + (uint64_t)(int64_t)call_code(c6[1]);
    n = n * 10 + (uint64_t)(int64_t)call_code(c6[2]);
    n = n * 27 + (uint64_t)(int64_t)call_code(c6[3]) - 10;
    n = n * 27 + (uint64_t)(int64_t)call_code(c6[4]) - 10;
    n = n * 27 + (uint64_t)(int64_t)call_code(c6[5]) - 10;
    return n;
}
WHD int alnum_or(int ch, int other) {
    if (ch >= '0' && ch <= '9') return ch - '0';
    if (ch >= 'A' && ch <= 'Z') return ch - 'A' + 10;
    return other;
}
// callsign is modified (cut at '/') on the prefix path, like strtok does in the reference
WHD void pack_compound(char *callsign, int32_t *n, int32_t *m, int32_t *nadd) {
    char c6[16];
    for (int i = 0; i < 16; i++) c6[i] = 0;
    int slash = s_cspn(callsign, "/");
    int len = s_len(callsign);
    char after1 = slash + 1 <= len ? callsign[slash + 1] : 0;
    char after2 = slash + 2 <= len ? callsign[slash + 2] : 0;
    char after3 = slash + 3 <= len ? callsign[slash + 3] : 0;
    if (after2 == 0) {            // CALL/x
        for (int i = 0; i < slash && i < 12; i++) c6[i] = callsign[i];
        *n = (int32_t)pack_callsign(c6);
        *nadd = 1;
        *m = 60000 - 32768 + alnum_or(after1, 38);
    } else if (after3 == 0) {     // CALL/nn
        for (int i = 0; i < slash && i < 12; i++) c6[i] = callsign[i];
        *n = (int32_t)pack_callsign(c6);
        *nadd = 1;
        *m = 60000 + 26 + 10 * (after1 - 48) + (after2 - 48);
    } else {                      // PFX/CALL
        char *save = callsign;
        char *pfx = s_tok(&save, "/");
        char *rest = s_tok(&save, " ");
        *n = rest ? (int32_t)pack_callsign(rest) : 0;
        int plen = pfx ? s_len(pfx) : 0;
        *m = (plen == 1) ? 37 * 36 + 36 : (plen == 2) ? 36 : 0;
        for (int i = 0; i < plen; i++) *m = 37 * (*m) + alnum_or(callsign[i], 36);
        *nadd = 0;
        if (*m > 32768) {
            *m -= 32768;
            *nadd = 1;
        }
    }
}

// message text -> 162 channel symbols (0..3).  Returns 0 when the text is not one of the three message shapes.
template <class Store>
WHD_NOINLINE int channel_symbols(const char *rawmessage, Store &hs, unsigned char *symbols /*[162]*/) {
    const int pwr_round[10] = {0, -1, 1, 0, -1, 2, 1, 0, -1, 1};
    char msg[24];
    for (int i = 0; i < 24; i++) msg[i] = 0;
    for (int i = 0; i < 23 && rawmessage[i]; i++) msg[i] = rawmessage[i];
    int sp = s_cspn(msg, " "), sl = s_cspn(msg, "/"), lt = s_cspn(msg, "<"), gt = s_cspn(msg, ">");
    int mlen = s_len(msg);
    uint64_t n = 0;
    int m = 0;
    char *save = msg;
    if (sp > 3 && sp < 7 && sl == mlen && lt == mlen) {
        char *cs = s_tok(&save, " "), *grid = s_tok(&save, " "), *ps = s_tok(&save, " ");
        if (!cs || !grid || !ps) return 0;
        int power = s_atoi(ps);
        n = pack_callsign(cs);
        signed char g4[4];
        int glen = s_len(grid);
        for (int i = 0; i < 4; i++) g4[i] = (signed char)loc_code(i <= glen ? grid[i] : 0);
        m = (int)pack_grid_power(g4, power);
    } else if (lt == 0 && gt < mlen) {
        char *cs = s_tok(&save, "<> "), *grid = s_tok(&save, " "), *ps = s_tok(&save, " ");
        if (!cs || !grid || !ps) return 0;
        int power = s_atoi(ps);
        if (power < 0) power = 0;
        if (power > 60) power = 60;
        power += pwr_round[power % 10];
        int ntype = -(power + 1);
        int h = (int)nhash15(cs, (size_t)s_len(cs), 146u);
        m = 128 * h + ntype + 64;
        char g6[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int gl = s_len(grid);
        for (int i = 0; i < gl - 1 && i < 7; i++) g6[i] = grid[i + 1];
        g6[5] = grid[0];
        n = pack_callsign(g6);
    } else if (sl < mlen) {
        char *cs = s_tok(&save, " ");
        if (!cs || sl == 0 || sl > s_len(cs)) return 0;
        char *ps = s_tok(&save, " ");
        if (!ps) return 0;
        int power = s_atoi(ps);
        if (power < 0) power = 0;
        if (power > 60) power = 60;
        power += pwr_round[power % 10];
        int32_t n1, ng, nadd;
        pack_compound(cs, &n1, &ng, &nadd);
        int ntype = power + 1 + nadd;
        m = 128 * ng + ntype + 64;
        n = (uint64_t)(int64_t)n1;
    } else {
        return 0;
    }
    unsigned char data[11];
    for (int i = 0; i < 11; i++) data[i] = 0;
    data[0] = (unsigned char)(0xFF & (n >> 20));
    data[1] = (unsigned char)(0xFF & (n >> 12));
    data[2] = (unsigned char)(0xFF & (n >> 4));
    data[3] = (unsigned char)(((n & 0x0F) << 4) + ((m >> 18) & 0x0F));
    data[4] = (unsigned char)(0xFF & (m >> 10));
    data[5] = (unsigned char)(0xFF & (m >> 2));
    data[6] = (unsigned char)((m & 0x03) << 6);

    // the reference unpacks its own packing once more; the only surviving effect is on the hash store
    {
        char t_clp[23], t_cs[13], t_call[13], t_loc[7], t_pwr[3];
        t_clp[0] = t_cs[0] = t_call[0] = t_loc[0] = t_pwr[0] = 0;
        signed char chk[11];
        for (int i = 0; i < 11; i++) chk[i] = (signed char)data[i];
        unpack_message(chk, hs, t_clp, t_call, t_loc, t_pwr, t_cs);
    }
    unsigned char bits[176];
    for (int i = 0; i < 176; i++) bits[i] = 0;
    conv_encode(bits, data, 11);
    interleave162(bits);
    for (int i = 0; i < NSYM; i++) symbols[i] = (unsigned char)(2 * bits[i] + sync_bit(i));
    return 1;
}

}  // namespace wspr
