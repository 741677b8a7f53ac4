// Host build of the device Fano loop (rtlsdr_wsprd_b200/csrc/wspr_fano.cuh) for CPU-side checks of its logic: the same
// template source, one lane, the scratch in ordinary memory.  tests/test_fano_host.py compiles this file with g++ and
// compares both instantiations (exact / decode) with the oracle's fano() on random symbol vectors, time-outs included,
// and runs several attempts through ONE lane back to back (the re-arming the queue-fed worker warps rely on) -- so a
// change to the loop is checked before any GPU time is spent on it.
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector>

struct uint2 { unsigned x, y; };
static struct { unsigned x; } threadIdx = {0};
#define __device__
#define __forceinline__ inline
static inline bool __any_sync(unsigned, bool p) { return p; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __float2int_rz(float f) { return (int)f; }
using std::max;
using std::min;

#include "../rtlsdr_wsprd_b200/csrc/wspr_fano.cuh"
#include "../rtlsdr_wsprd_b200/csrc/wspr_mettab.h"

namespace {
const short g_mettab[2][256] = WSPR_METTAB_INIT;
struct HostMem {                                   // one lane: level record n at lvl[n], node record pos at node[pos + 1]
    uint2 *lvl;
    unsigned *node;
    uint2 ldl(int n) const { return lvl[n]; }
    void stl(int n, unsigned a, unsigned b) const { lvl[n] = uint2{a, b}; }
    unsigned ldn(int pos) const { return node[pos + 1]; }
    void stn(int pos, unsigned v) const { node[pos + 1] = v; }
};
// `count` attempts handed to the lane one after the other
struct ListFeed {
    const unsigned char *sym;
    int count, next_i, done_i;
    unsigned stop_after;
    wspr::FanoResult *out;
    const unsigned char *next(unsigned &stop) {
        stop = stop_after;
        return next_i < count ? sym + 162 * (size_t)next_i++ : nullptr;
    }
    bool abandon() const { return false; }
    void finish(const wspr::FanoResult &r) { out[done_i++] = r; }
    void period(bool) {}
};
}  // namespace

// variant bit 0: decode instantiation (time-out test every 256 trips, maxnp not tracked).  `count` vectors of 162 symbols
// are decoded back to back by the one lane; outputs are arrays of `count` entries (data: count x 12 bytes).
extern "C" int fano_host(int variant, const unsigned char *sym, int count, int delta, unsigned maxcycles, unsigned stop_after,
                         int *rc, unsigned *metric, unsigned *cycles, unsigned *maxnp, unsigned char *data) {
    using namespace wspr;
    std::vector<uint2> lvl(FANO_LEVELS, uint2{0xdeadbeefu, 0xdeadbeefu});   // (poisoned: an idle lane reads, never uses, them)
    std::vector<unsigned> node(FANO_NODES + 1, 0xdeadbeefu);
    std::vector<FanoResult> res(count);
    HostMem mem{lvl.data(), node.data()};
    ListFeed feed{sym, count, 0, 0, stop_after, res.data()};
    if (variant & 1) fano_run<false>(feed, mem, &g_mettab[0][0], delta, maxcycles);
    else fano_run<true>(feed, mem, &g_mettab[0][0], delta, maxcycles);
    if (feed.done_i != count) return -100;
    for (int i = 0; i < count; i++) {
        rc[i] = res[i].rc;
        metric[i] = res[i].metric;
        cycles[i] = res[i].cycles;
        maxnp[i] = res[i].maxnp;
        memcpy(data + 12 * (size_t)i, res[i].data, 12);
    }
    return 0;
}
