// Device-side Fano sequential decoder (K=32, r=1/2; wsprd/fano.c:87-238): lane-dense, branch-free, compact state.
//
// A decode that fails costs maxcycles*nbits = 810 000 strictly sequential tree moves, and a hopeless candidate brings 43 such
// attempts (wsprd.c:741-766).  Three things decide what that costs the rest of the GPU: the instructions one move takes,
// how many attempts share one instruction stream, and how much on-chip memory a resident decoder warp holds.
//   * Every lane of a warp decodes its own attempt and all lanes execute the SAME instruction stream: one loop trip performs
//     whichever of {forward move, tighten threshold, step back} the lane needs, chosen by predicates/selects, so 32 attempts
//     advance per issued instruction.  A walk back over several nodes costs one trip per node but only one Fano cycle.
//   * The node a lane stands on (encoder state, path metric, the two branch metrics, branch index) and its parent's path
//     metric live in registers; a forward move touches no memory on its critical path.
//   * Per tree level the branch metrics are precomputed as TWO words (one per complementary symbol pair {0,3} / {1,2}):
//     arriving at a node is one select on the parity of the encoder state plus a shift that picks the bit of the better
//     branch.  8 bytes per level and lane.
//   * A node left forwards is remembered in ONE word: its two branch metrics, which branch is being tried, the metric of the
//     branch that led to it (so the grandparent's path metric is a subtraction away when the decoder steps back) and the
//     one encoder-state bit a step back cannot recover by shifting.  4 bytes per node and lane.
//   * Records are interleaved by lane in shared memory ([record][lane]): bank-conflict free whatever depth each lane is at.
//     31.6 KB per warp (round 1 kept 16-byte records: 84 KB, which is what kept co-resident decoder warps to one or two
//     per SM and evicted the bulk kernels' CTAs).
//   * Lanes re-arm: a lane whose attempt has ended publishes its result and asks its Feed for the next attempt at the next
//     housekeeping point (every 256 trips), so a warp stays full as long as there is work -- the long runs of parked
//     candidates are served by a pool of such warps from a device-side queue (wspr_kernels.cu, k_fano_workers).
// The move sequence, cycle count, final metric and decoded bytes are identical to fano.c: tests/test_fano_host.py compiles
// this loop for the host and compares it with the oracle, tests/test_gpu_parity.py does the same on the GPU.
#pragma once
#include "wspr_codec.cuh"

namespace wspr {

struct FanoResult {
    int rc;              // 0 decoded, -1 timeout, FANO_STOPPED cut short
    unsigned metric, cycles, maxnp;
    unsigned char data[12];
};

// stored branch metrics: u = metric + FANO_BIAS, 9 bits (a branch metric is the sum of two table entries in [-137, 5])
constexpr int FANO_BIAS = 300;
constexpr int FANO_LEVELS = NBITS + 1;           // level records 0..81 per lane (81: zeros, read by a move past the last node)
constexpr int FANO_NODES = NBITS + 2;            // node records -1..81 per lane (-1: read, never used, by a lane at the root)
constexpr int FANO_WARP_SMEM_BYTES = FANO_LEVELS * 32 * 8 + FANO_NODES * 32 * 4;   // 31 616
// level word: [0..8] x  [9..17] y  [18..26] x  [28] swap when the 0-branch symbol is the pair's lower one  [29] ... upper one
//             [30] bit of the first-choice branch when the 0-branch symbol is the lower one  [31] ... upper one
//   view (W & 0x3ffff) = (x, y), view ((W >> 9) & 0x3ffff) = (y, x): levels that carry data store (better, worse) and never
//   swap; tail levels (a single branch, fano.c:176-180) store (metric of the lower symbol, of the upper one) and swap when the
//   encoder asks for the upper one.
// node word:  [0..8] first-choice metric  [9..17] second-choice metric  [18..26] metric of the branch that led here
//             [27] bit 31 of the node's encoder state  [31] second choice being tried
constexpr unsigned FANO_W_KEEP = 0x8003ffffu;

#ifdef __CUDACC__
// this warp's scratch in shared memory: addresses pinned in registers (otherwise the window base of dynamic shared memory
// is re-derived from special registers on every loop trip)
struct FanoSmem {
    unsigned lvl, node;                          // shared-space byte address of this lane's level record 0 / node record 0
    static __device__ __forceinline__ FanoSmem at(const void *warp_smem) {
        const unsigned lane = threadIdx.x & 31u;
        unsigned b = (unsigned)__cvta_generic_to_shared(warp_smem);
        unsigned l = b + lane * 8u, n = b + (unsigned)FANO_LEVELS * 256u + 128u + lane * 4u;
        asm volatile("" : "+r"(l), "+r"(n));
        return FanoSmem{l, n};
    }
    __device__ __forceinline__ uint2 ldl(int n) const {
        uint2 v;
        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(lvl + (unsigned)n * 256u) : "memory");
        return v;
    }
    __device__ __forceinline__ void stl(int n, unsigned a, unsigned b) const {
        asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(lvl + (unsigned)n * 256u), "r"(a), "r"(b) : "memory");
    }
    __device__ __forceinline__ unsigned ldn(int pos) const {
        unsigned v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(node + (unsigned)(pos * 128)) : "memory");
        return v;
    }
    __device__ __forceinline__ void stn(int pos, unsigned v) const {
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(node + (unsigned)(pos * 128)), "r"(v) : "memory");
    }
};
#endif

// A Feed hands attempts to the lanes of a warp and takes their results back.  All three are called at housekeeping points
// only, by the lanes concerned (the warp is diverged there):
//   const unsigned char *next(unsigned &stop_after)  an idle lane asks for work: the 162 deinterleaved soft symbols of the
//                                                    next attempt (stop_after: 0 = run to the reference's limit, else give
//                                                    up with FANO_STOPPED after that many cycles), or nullptr: nothing left
//   bool abandon()                                   an active lane asks whether its attempt is still wanted
//   void finish(const FanoResult &)                  the lane's attempt has ended
//   void period(bool active)                         once per housekeeping point and lane (statistics)
// One attempt per lane, results kept in the object: the form the in-round kernel and the tests use.
struct FanoOneShot {
    const unsigned char *sym;                    // nullptr: this lane has nothing to decode
    unsigned stop_after;
    FanoResult res;
    __device__ __forceinline__ const unsigned char *next(unsigned &stop) {
        const unsigned char *s = sym;
        sym = nullptr;
        stop = stop_after;
        return s;
    }
    __device__ __forceinline__ bool abandon() const { return false; }
    __device__ __forceinline__ void finish(const FanoResult &r) { res = r; }
    __device__ __forceinline__ void period(bool) {}
};

// Every lane of the warp must call this together.  mem: this warp's FANO_WARP_SMEM_BYTES of scratch.
// A lane without an attempt, or whose attempt has ended, keeps executing the (uniform) loop as a harmless zombie -- its
// threshold is parked so high that it only ever tightens it in place -- so the hot loop carries no per-lane predicate.
//   * `w` carries the node's two branch metrics AND, in bit 31, which branch is being tried;
//   * the metric of the branch being tried (`cur`) is extracted at the end of the previous trip, off the critical path;
//   * EXACT = false (the decode kernels): the time-out test is made every 256 trips instead of every trip -- a lane past
//     the limit can no longer succeed (the decode test checks the count) and walks on harmlessly until it is noticed --
//     and maxnp, which wspr_decode never looks at, is not tracked.  `cycles` of a time-out is the reference's constant
//     either way; only `metric` of a time-out (unused by wspr_decode, fano.c:236) is then not the reference's.
//     EXACT = true (fano() parity tests): every output field as fano.c produces it.
//   * SMALLSTEP: delta > 10; a branch metric is at most +10, so one threshold step per move suffices.
template <bool EXACT, bool SMALLSTEP, typename Feed, typename Mem>
__device__ __forceinline__ void fano_run_impl(Feed &feed, Mem mem, const short *__restrict__ mettab, int delta, unsigned maxcycles) {
    constexpr int nbits = NBITS;
    constexpr int tail = nbits - 31;
    constexpr int PARKED = 0x3fffffff;
    const unsigned limit = maxcycles * (unsigned)nbits;
    const float inv_delta = 1.0f / (float)delta;

    bool act = false, busy = false;
    unsigned inback = 0;                           // 1: this trip continues a walk back (no new Fano cycle)
    int pos = 0, thr = PARKED, gam = 0, pgam = 0, maxnp = 0;
    unsigned it = 0, stop = 0xffffffffu;           // Fano cycles started so far; cycle budget of this attempt
    unsigned w = 0, enc = 0, cur = 0;              // cur: biased metric of the branch being tried
    int r_rc = -1;                                 // result registers, filled when the attempt ends
    unsigned r_metric = 0, r_cycles = 0, r_maxnp = 0, r_enc = 0;
#pragma unroll 1
    for (unsigned trip = 0;; trip++) {
        if ((trip & 255u) == 0u) {                 // housekeeping
            if (!EXACT && act && it >= limit) {    // the reference's loop ended somewhere in the last 256 trips: time-out
                act = false;
                r_rc = -1;
                r_metric = (unsigned)gam;
                r_cycles = limit + 2u;
                r_maxnp = 0u;
            }
            if (act && inback == 0 && (it >= stop || feed.abandon())) {
                act = false;
                r_rc = FANO_STOPPED;
                r_metric = (unsigned)gam;
                r_cycles = it + 1u;
                r_maxnp = (unsigned)maxnp;
            }
            if (!act) {
                if (busy) {                        // publish the attempt that has ended
                    FanoResult r;
                    r.rc = r_rc;
                    r.metric = r_metric;
                    r.cycles = r_cycles;
                    r.maxnp = r_maxnp;
#pragma unroll
                    for (int b = 0; b < 12; b++) r.data[b] = 0;
                    if (r_rc == 0) {
                        // decoded bits: bit j of the path is bit 31 of the encoder state of node j + 31 (kept in that node's
                        // record) for j <= 48, and bit 80 - j of the last node's state beyond (fano.c:224-230 reads the low
                        // byte of the states of nodes 7, 15, ...)
#pragma unroll 1
                        for (int b = 0; b < (nbits >> 3); b++) {
                            unsigned v = 0;
                            for (int k = 0; k < 8; k++) {
                                const int j = 8 * b + k;
                                const unsigned bit = (j <= nbits - 33) ? (mem.ldn(j + 31) >> 27) & 1u : (r_enc >> (nbits - 1 - j)) & 1u;
                                v = (v << 1) | bit;
                            }
                            r.data[b] = (unsigned char)v;
                        }
                    }
                    feed.finish(r);
                    busy = false;
                }
                unsigned stop_after = 0;
                const unsigned char *symbols = feed.next(stop_after);
                if (symbols != nullptr) {          // arm the lane: level records of the attempt, decoder at the root
#pragma unroll 1
                    for (int n = 0; n <= nbits; n++) {
                        unsigned wa = 0, wb = 0;
                        if (n < nbits) {
                            const int a0 = mettab[symbols[2 * n]], a1 = mettab[256 + symbols[2 * n]];
                            const int b0 = mettab[symbols[2 * n + 1]], b1 = mettab[256 + symbols[2 * n + 1]];
                            const int lo[2] = {a0 + b0, a0 + b1}, hi[2] = {a1 + b1, a1 + b0};   // pairs {0,3} and {1,2}
                            unsigned word[2];
#pragma unroll
                            for (int p = 0; p < 2; p++) {
                                const unsigned ulo = (unsigned)(lo[p] + FANO_BIAS), uhi = (unsigned)(hi[p] + FANO_BIAS);
                                unsigned x, y, flags;
                                if (n >= tail) {   // one branch only: its metric is the arriving symbol's
                                    x = ulo;
                                    y = uhi;
                                    flags = 1u << 29;
                                } else {           // (fano.c:127-135: the 1-branch is the better one on a tie)
                                    x = max(ulo, uhi);
                                    y = min(ulo, uhi);
                                    flags = (lo[p] > hi[p] ? 0u : 1u << 30) | (hi[p] > lo[p] ? 0u : 1u << 31);
                                }
                                word[p] = x | (y << 9) | (x << 18) | flags;
                            }
                            wa = word[0];
                            wb = word[1];
                        }
                        mem.stl(n, wa, wb);
                    }
                    const unsigned w0 = mem.ldl(0).x;  // root: encoder state 0, branch symbol 0 -> pair {0,3}, lower symbol
                    act = busy = true;
                    inback = 0;
                    pos = 0;
                    thr = 0;
                    gam = pgam = 0;
                    maxnp = 0;
                    it = 0;
                    stop = (stop_after != 0 && stop_after < limit) ? stop_after : 0xffffffffu;
                    w = w0 & 0x3ffffu;
                    enc = (w0 >> 30) & 1u;
                    cur = w & 0x1ffu;
                } else {                           // park
                    thr = PARKED;
                    inback = 0;
                }
            }
            feed.period(act);
            if (!__any_sync(0xffffffffu, act)) break;
        }
        // the level we would move down to, the node we would step back to (never used at the root)
        const uint2 nl = mem.ldl(pos + 1);
        const unsigned nd = mem.ldn(pos - 1);
        const int ng = gam + (int)cur - FANO_BIAS;
        const bool newc = inback == 0;             // this trip opens a new Fano cycle
        const bool fwd = newc && (ng >= thr);
        const bool tig = newc && !fwd && (pos == 0 || pgam < thr);
        const bool bck = !fwd && !tig;
        int posN = pos + (fwd ? 1 : (bck ? -1 : 0));
        const bool arrived = posN == nbits;        // a move past the last node: decoded, the lane parks where it stands
        posN = arrived ? nbits - 1 : posN;
        if (EXACT && newc && it >= limit && act) { // the reference's loop ends here: time-out (rare, once per attempt)
            act = false;
            r_rc = -1;
            r_metric = (unsigned)gam;
            r_cycles = limit + 2u;
            r_maxnp = (unsigned)maxnp;
        }
        it += 1u - inback;
        if (EXACT) maxnp = newc ? max(maxnp, pos) : maxnp;
        // ---- forward: raise the threshold on a first visit, remember the node, descend along the better branch
        int thrF = thr;
        if (SMALLSTEP) {
            const int t1 = thr + delta;
            thrF = (gam < t1 && ng >= t1) ? t1 : thr;
        } else if (gam < thr + delta) {            // while (ng >= thr + delta) thr += delta
            const int d = ng - thr;
            int k = __float2int_rz((float)d * inv_delta);
            k += ((k + 1) * delta <= d) ? 1 : 0;
            k -= (k * delta > d) ? 1 : 0;
            thrF = thr + k * delta;
        }
        if (fwd) mem.stn(pos, w | ((unsigned)(gam - pgam + FANO_BIAS) << 18) | ((enc >> 4) & 0x08000000u));
        const unsigned e = enc << 1;
        const unsigned pa = (unsigned)__popc(e & POLY_A) & 1u;                 // branch symbol of the 0-branch = 2*pa + pb
        const bool pair = (__popc(e & (POLY_A ^ POLY_B)) & 1) != 0;           // pa != pb: symbols {1,2}
        const unsigned lw = pair ? nl.y : nl.x;
        const unsigned lt = lw >> pa;              // bit 28: swap, bit 30: bit of the first-choice branch
        const unsigned wF = (lt & 0x10000000u) ? ((lw >> 9) & 0x3ffffu) : (lw & 0x3ffffu);
        const unsigned encF = e | ((lt >> 30) & 1u);
        // ---- step back onto the parent
        const unsigned wB0 = nd & FANO_W_KEEP;
        const bool selB0 = (int)nd < 0;
        const int pgamB = pgam - (int)((nd >> 18) & 0x1ffu) + FANO_BIAS;      // the grandparent's path metric
        const bool b1 = (pos <= tail) && !selB0;                              // take the parent's other branch
        const bool b2 = !b1 && (pos == 1 || pgamB < thr);                     // cannot go higher: tighten there
        const unsigned wB = b1 ? (wB0 | 0x80000000u) : (b2 ? (wB0 & 0x7fffffffu) : wB0);
        const unsigned encB = ((enc >> 1) | ((nd << 4) & 0x80000000u)) ^ ((b1 || (b2 && selB0)) ? 1u : 0u);
        // ---- merge
        const int dthr = (tig || (bck && b2)) ? delta : 0;
        thr = fwd ? thrF : thr - dthr;
        const int gamN = fwd ? ng : (bck ? pgam : gam);                       // (the parent's path metric is what pgam holds)
        pgam = fwd ? gam : (bck ? pgamB : pgam);
        gam = gamN;
        const unsigned encT = enc;
        enc = fwd ? encF : (bck ? encB : (enc ^ (w >> 31)));
        w = fwd ? wF : (bck ? wB : (w & 0x7fffffffu));
        cur = (((int)w < 0) ? (w >> 9) : w) & 0x1ffu;
        inback = (bck && !b1 && !b2) ? 1u : 0u;
        pos = posN;
        if (arrived) {                             // reached the last node: decoded (rare, once per attempt)
            if (act) {
                act = false;
                r_rc = (it >= limit) ? -1 : 0;     // (a decode in the very last cycle counts as a timeout, fano.c:234)
                r_metric = (unsigned)gam;
                r_cycles = (it > limit) ? limit + 2u : it + 1u;   // (it > limit: a lane past the limit, not yet noticed)
                r_maxnp = (unsigned)maxnp;
                r_enc = encT;                      // encoder state of the last node
            }
            thr = PARKED;                          // park: stay put (pos = nbits - 1), tightening an unreachable threshold
            inback = 0;
        }
    }
}

template <bool EXACT, typename Feed, typename Mem>
__device__ __forceinline__ void fano_run(Feed &feed, Mem mem, const short *__restrict__ mettab, int delta, unsigned maxcycles) {
    if (delta > 10) fano_run_impl<EXACT, true>(feed, mem, mettab, delta, maxcycles);
    else fano_run_impl<EXACT, false>(feed, mem, mettab, delta, maxcycles);
}

}  // namespace wspr
