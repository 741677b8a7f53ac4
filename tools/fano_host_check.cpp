// Host build of the device Fano loop (rtlsdr_wsprd_b200/csrc/wspr_fano.cuh) for CPU-side checks of its logic: the same
// template source, one lane, the scratch in ordinary memory.  tests/test_fano_host.py compiles this file with g++ and
// compares every instantiation (exact / decode, plain / pipelined) with the oracle's fano() on random symbol vectors,
// time-outs included -- so a change to the loop is checked before any GPU time is spent on it.
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector>

struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static struct { unsigned x; } threadIdx = {0};
#define __device__
#define __forceinline__ inline
static inline bool __any_sync(unsigned, bool p) { return p; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __float2int_rz(float f) { return (int)f; }
using std::max;
static inline unsigned long long __cvta_generic_to_shared(const void *p) { return (unsigned long long)(uintptr_t)p; }

#include "../rtlsdr_wsprd_b200/csrc/wspr_fano.cuh"
#include "../rtlsdr_wsprd_b200/csrc/wspr_mettab.h"

namespace {
const short g_mettab[2][256] = WSPR_METTAB_INIT;
struct HostMem {                                   // one lane: record e at e * 16
    unsigned char *base;
    unsigned row;
    uint4 ld(unsigned off) const { uint4 v; memcpy(&v, base + off, 16); return v; }
    void st(unsigned off, unsigned x, unsigned y, unsigned z, unsigned w) const { uint4 v{x, y, z, w}; memcpy(base + off, &v, 16); }
};
}  // namespace

// variant bit 0: decode instantiation (time-out test every 256 trips, maxnp not tracked); bit 1: pipelined loop
extern "C" int fano_host(int variant, const unsigned char *sym, int delta, unsigned maxcycles, unsigned stop_after,
                         unsigned *metric, unsigned *cycles, unsigned *maxnp, unsigned char *data /*[12]*/) {
    using namespace wspr;
    // guard records on both sides: the loop may address one record before the level table and a few after the node stack
    std::vector<unsigned char> scratch((size_t)(FANO_LVL_RECORDS + FANO_NODE_RECORDS + 8) * 16, 0);
    HostMem mem{scratch.data() + 4 * 16, 16u};
    FanoResult r;
    switch (variant & 3) {
        case 0: fano_dense<true, false>(r, true, sym, &g_mettab[0][0], delta, maxcycles, stop_after, FanoNoStop(), mem); break;
        case 1: fano_dense<false, false>(r, true, sym, &g_mettab[0][0], delta, maxcycles, stop_after, FanoNoStop(), mem); break;
        case 2: fano_dense<true, true>(r, true, sym, &g_mettab[0][0], delta, maxcycles, stop_after, FanoNoStop(), mem); break;
        default: fano_dense<false, true>(r, true, sym, &g_mettab[0][0], delta, maxcycles, stop_after, FanoNoStop(), mem); break;
    }
    *metric = r.metric;
    *cycles = r.cycles;
    *maxnp = r.maxnp;
    memcpy(data, r.data, 12);
    return r.rc;
}
