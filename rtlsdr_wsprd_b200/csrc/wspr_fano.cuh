// Device-side Fano sequential decoder (K=32, r=1/2; wsprd/fano.c:87-238): lane-dense, branch-free.
//
// A decode that fails costs maxcycles*nbits = 810 000 strictly sequential tree moves, so two things matter on a GPU:
// the dependent latency of ONE move, and how many issue slots a move costs.  A thread-per-attempt port of fano.c
// (the portable version in wspr_codec.cuh) is poor at both: the tree lives in five indexed local arrays (every move a
// chain of dependent memory round trips) and the 32 lanes of a warp sit on different paths of a branchy loop, so the
// warp serialises them.  Here every lane of a warp decodes its own attempt and all lanes execute the SAME instruction
// stream: one loop trip performs whichever of {forward move, tighten threshold, step back} the lane needs, chosen by
// predicates/selects, so 32 attempts advance per issued instruction.
//   * The node a lane stands on (encoder state, path metric, packed branch metrics, branch index) and its parent's
//     metric live in registers; a forward move touches no memory on its critical path.
//   * Per tree level the four possible (better metric, worse metric, bit) triples are precomputed, so arriving at a
//     node is a 4-way register select.
//   * A node record is 16 bytes {enc, gam, metrics | branch index, gam of the parent}: one vector store when a node is
//     left forwards, one vector load per step back.  Records are interleaved by lane in shared memory
//     ([element][lane]), which makes the 128-bit accesses of a warp bank-conflict free whatever depth each lane is at.
//   * A walk back over several nodes costs one loop trip per node but only one Fano cycle, as in fano.c.
// The move sequence, cycle count, final metric and decoded bytes are identical to fano.c (tests/test_gpu_parity.py
// checks them against the oracle on random symbol vectors, time-outs included).
#pragma once
#include "wspr_codec.cuh"

namespace wspr {

struct FanoResult {
    int rc;              // 0 decoded, -1 timeout, FANO_STOPPED cut short
    unsigned metric, cycles, maxnp;
    unsigned char data[12];
};

constexpr int FANO_LVL_RECORDS = NBITS + 1;      // per lane
constexpr int FANO_NODE_RECORDS = NBITS + 2;
// shared memory of one warp: [FANO_LVL_RECORDS + FANO_NODE_RECORDS][32 lanes] uint4
constexpr int FANO_WARP_SMEM_BYTES = (FANO_LVL_RECORDS + FANO_NODE_RECORDS) * 32 * 16;

__device__ __forceinline__ unsigned fano_pack(int tm0, int tm1, unsigned bit) {   // tm0: bits 0..15, tm1: 16..29, bit: 30
    return ((unsigned)tm0 & 0xffffu) | (((unsigned)tm1 & 0x3fffu) << 16) | (bit << 30);
}
__device__ __forceinline__ int fano_tm0(unsigned w) { return (int)(short)(w & 0xffffu); }
__device__ __forceinline__ int fano_tm1(unsigned w) { return ((int)(w << 2)) >> 18; }
__device__ __forceinline__ uint4 lds128(unsigned addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(unsigned addr, unsigned x, unsigned y, unsigned z, unsigned w) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// Where a warp keeps its level tables and node records: shared memory (lowest latency), or global memory read through
// L1 (ld.global.ca).  The long-running chains of parked candidates use the latter: a CTA that sits on 170 KB of shared
// memory for 100+ ms starves the bulk kernels sharing its SM, while L1 lines are a soft claim.
// `row` = bytes between consecutive records of one lane = 16 * (lanes the scratch is laid out for): 512 for a full
// warp, 256 when only the first 16 lanes decode (the scratch is then half the size).
struct FanoSmem {
    unsigned base, row;
    // base: shared-state-space address of the scratch, pinned in a register (otherwise the window base of dynamic shared
    // memory is re-derived from special registers on every loop trip)
    static __device__ __forceinline__ FanoSmem at(const void *smem, unsigned row) {
        unsigned b = (unsigned)__cvta_generic_to_shared(smem);
        asm volatile("" : "+r"(b));
        return FanoSmem{b, row};
    }
    __device__ __forceinline__ uint4 ld(unsigned off) const { return lds128(base + off); }
    __device__ __forceinline__ void st(unsigned off, unsigned x, unsigned y, unsigned z, unsigned w) const { sts128(base + off, x, y, z, w); }
};
struct FanoGmem {
    unsigned char *base;
    unsigned row;
    __device__ __forceinline__ uint4 ld(unsigned off) const {
        uint4 v;
        asm volatile("ld.global.ca.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(base + off) : "memory");
        return v;
    }
    __device__ __forceinline__ void st(unsigned off, unsigned x, unsigned y, unsigned z, unsigned w) const {
        asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(base + off), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
    }
};

// hook: stop() polled every 256 trips by active lanes; success() called by a lane the moment it decodes, with its cycle
// count and the means to read the decoded bytes (byte b = the low byte of the encoder state stored with node 7 + 8 b)
struct FanoNoStop {
    __device__ bool stop() const { return false; }
    template <typename Mem>
    __device__ void success(unsigned, const Mem &, unsigned) const {}
};

// Every lane of the warp must call this together.  `want`: this lane has an attempt to decode (symbols valid).
// mem: this warp's FANO_WARP_SMEM_BYTES of scratch, FanoSmem{cvta'd shared address} or FanoGmem{pointer}.
// stop_after: 0 = run to the reference's limit; else give up (FANO_STOPPED) once that many cycles were spent.
// hook.stop(): evaluated every 256 trips by active lanes, true abandons the attempt (FANO_STOPPED);
// hook.success(cycles, mem, node_base): called by a lane the moment it decodes.
// A lane that has finished keeps executing the (uniform) loop as a harmless zombie -- its threshold is parked so high
// that it only ever tightens it in place -- while its result waits in separate registers; the hot loop therefore
// carries no per-lane "active" predicate.
//
// The throughput of the whole decode is very sensitive to the length of one trip (a hopeless candidate keeps 43 lanes
// busy for 810 000 cycles, and the SMs they sit on are shared with the bulk kernels), so the loop body is kept minimal:
//   * `w` carries the node's packed branch metrics AND, in bit 31, which branch is being tried (sel); the record a node
//     is left with is stored and reloaded in that form, no packing or unpacking;
//   * the metric of the branch being tried (`cur`) is extracted at the end of the previous trip, off the critical path;
//   * EXACT = false (the decode kernels): the time-out test is made every 256 trips instead of every trip -- a lane past
//     the limit can no longer succeed (the decode test checks the count) and walks on harmlessly until it is noticed --
//     and maxnp, which wspr_decode never looks at, is not tracked.  `cycles` of a time-out is the reference's constant
//     either way; only `metric` of a time-out (unused by wspr_decode, fano.c:236) is then not the reference's.
//     EXACT = true (fano() parity tests): every output field as fano.c produces it.
//   * SMALLSTEP: delta > 10; a branch metric is at most +10, so one threshold step per move suffices.
// Measured alone on a B200 (tools/fano_microbench.py): 249 SM clocks per Fano cycle (277 with EXACT), i.e. ~190 clocks
// for a trip of 81 instructions -- the trip is bound by the DEPTH of its predicate/select dataflow, not by issue slots
// and not by the shared-memory latency: holding the two records in registers and fetching the successors' records one
// trip ahead (three loads in flight across the loop edge, 111 instructions) was slower, 263 clocks per cycle.
template <bool EXACT, bool SMALLSTEP, bool PIPE, typename Hook, typename Mem>
__device__ __forceinline__ void fano_dense_impl(FanoResult &out, bool want, const unsigned char *__restrict__ symbols,
                                                const short *__restrict__ mettab, int delta, unsigned maxcycles,
                                                unsigned stop_after, Hook hook, Mem mem) {
    constexpr int nbits = NBITS;
    constexpr int tail = nbits - 31;
    constexpr int PARKED = 0x3fffffff;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned row = mem.row;                                                  // record e of this lane: e*row + lane*16
    unsigned lvl_base = lane * 16u;
    unsigned node_base = (unsigned)FANO_LVL_RECORDS * row + lane * 16u;
    asm volatile("" : "+r"(lvl_base), "+r"(node_base));                            // keep both in registers
#pragma unroll 1
    for (int n = 0; n <= nbits; n++) {                                             // (lanes without an attempt get zeros)
        unsigned w[4] = {0, 0, 0, 0};
        if (want && n < nbits) {
            const int a0 = mettab[symbols[2 * n]], a1 = mettab[256 + symbols[2 * n]];
            const int b0 = mettab[symbols[2 * n + 1]], b1 = mettab[256 + symbols[2 * n + 1]];
            const int m[4] = {a0 + b0, a0 + b1, a1 + b0, a1 + b1};
#pragma unroll
            for (int ls = 0; ls < 4; ls++) {
                const int m0 = m[ls], m1 = m[3 ^ ls];
                if (n >= tail) w[ls] = fano_pack(m0, 0, 0);
                else w[ls] = (m0 > m1) ? fano_pack(m0, m1, 0) : fano_pack(m1, m0, 1);
            }
        }
        mem.st(lvl_base + row * (unsigned)n, w[0], w[1], w[2], w[3]);
    }
#pragma unroll 1
    for (int n = 0; n < FANO_NODE_RECORDS; n++) mem.st(node_base + row * (unsigned)n, 0u, 0u, 0u, 0u);
    const unsigned limit = maxcycles * (unsigned)nbits;
    const unsigned stop = (stop_after != 0 && stop_after < limit) ? stop_after : 0xffffffffu;
    const float inv_delta = 1.0f / (float)delta;

    bool act = want, inback = false;
    int pos = 0, thr = want ? 0 : PARKED, gam = 0, pgam = 0, maxnp = 0;
    unsigned it = 0;                               // Fano cycles started so far
    unsigned w = mem.ld(lvl_base).x;               // root: branch_sym(0) == 0; bit 31 (sel) = 0: the better branch first
    unsigned enc = w >> 30;
    int cur = fano_tm0(w);                         // metric of the branch being tried
    int r_rc = -1;                                 // result registers, filled when the lane finishes
    unsigned r_metric = 0, r_cycles = 0, r_maxnp = 0;
    // PIPE: the two records of the node a trip stands on are fetched during the PREVIOUS trip, as soon as that trip knows
    // where it moves to (see below); they are carried in registers.
    uint4 nl_c = make_uint4(0u, 0u, 0u, 0u), nd_c = make_uint4(0u, 0u, 0u, 0u);
    if (PIPE) {
        nl_c = mem.ld(lvl_base + row);
        nd_c = mem.ld(node_base - row);             // (the root has no parent: never used)
    }
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
            if (act && !inback && (it >= stop || hook.stop())) {
                act = false;
                r_rc = FANO_STOPPED;
                r_metric = (unsigned)gam;
                r_cycles = it + 1u;
                r_maxnp = (unsigned)maxnp;
                thr = PARKED;
            }
            if (!__any_sync(0xffffffffu, act)) break;
        }
        // the level we would move down to, the node we would step back to (never used at the root, where that address is
        // the last level record)
        uint4 nl, nd;
        if (PIPE) {
            nl = nl_c;
            nd = nd_c;
        } else {
            nl = mem.ld(lvl_base + row * (unsigned)(pos + 1));
            nd = mem.ld(node_base + row * (unsigned)pos - row);
        }
        const int ng = gam + cur;
        const bool newc = !inback;                 // this trip opens a new Fano cycle
        const bool fwd = newc && (ng >= thr);
        const bool tig = newc && !fwd && (pos == 0 || pgam < thr);
        const bool bck = !fwd && !tig;
        // PIPE: which way the decoder moves depends on registers only, so the move is decided first, the node is pushed, and
        // the records of the node it lands on are requested at once; everything below works on the records fetched a trip
        // ago and the shared-memory latency overlaps with it instead of heading the loop-carried dependence chain.  (The
        // push precedes the fetch in program order, so a forward move reads back the record it has just written.)
        int posN = pos + (fwd ? 1 : (bck ? -1 : 0));
        const bool arrived = posN == nbits;          // a move past the last node: decoded, the lane parks where it stands
        posN = arrived ? nbits - 1 : posN;
        if (PIPE) {
            if (fwd) mem.st(node_base + row * (unsigned)pos, enc, (unsigned)gam, w, (unsigned)pgam);
            nl_c = mem.ld(lvl_base + row * (unsigned)(posN + 1));
            nd_c = mem.ld(node_base + row * (unsigned)posN - row);
        }
        if (EXACT && newc && it >= limit && act) {  // the reference's loop ends here: time-out (rare, once per lane)
            act = false;
            r_rc = -1;
            r_metric = (unsigned)gam;
            r_cycles = limit + 2u;
            r_maxnp = (unsigned)maxnp;
        }
        it += newc ? 1u : 0u;
        if (EXACT) maxnp = newc ? max(maxnp, pos) : maxnp;
        // ---- forward: raise the threshold on a first visit, push the node, descend along the better branch
        int thrF = thr;
        if (SMALLSTEP) {
            const int t1 = thr + delta;
            thrF = (gam < t1 && ng >= t1) ? t1 : thr;
        } else if (gam < thr + delta) {             // while (ng >= thr + delta) thr += delta
            const int d = ng - thr;
            int k = __float2int_rz((float)d * inv_delta);
            k += ((k + 1) * delta <= d) ? 1 : 0;
            k -= (k * delta > d) ? 1 : 0;
            thrF = thr + k * delta;
        }
        if (!PIPE && fwd) mem.st(node_base + row * (unsigned)pos, enc, (unsigned)gam, w, (unsigned)pgam);
        const unsigned e = enc << 1;
        const bool pa = (__popc(e & POLY_A) & 1) != 0, pb = (__popc(e & POLY_B) & 1) != 0;   // branch symbol = 2*pa + pb
        const unsigned wlo = pb ? nl.y : nl.x, whi = pb ? nl.w : nl.z;
        const unsigned wF = pa ? whi : wlo;          // (bit 31 clear: the better branch first)
        const unsigned encF = e | (wF >> 30);
        // ---- step back onto the parent
        const unsigned wB0 = nd.z;
        const bool selB0 = (int)wB0 < 0;
        const int pgamB = (int)nd.w;
        const bool b1 = (pos <= tail) && !selB0;                     // take the parent's other branch
        const bool b2 = !b1 && (pos == 1 || pgamB < thr);            // cannot go higher: tighten there
        const unsigned wB = b1 ? (wB0 | 0x80000000u) : (b2 ? (wB0 & 0x7fffffffu) : wB0);
        const unsigned encB = nd.x ^ ((b1 || (b2 && selB0)) ? 1u : 0u);
        // ---- merge
        const int dthr = (tig || (bck && b2)) ? delta : 0;
        thr = fwd ? thrF : thr - dthr;
        const int gamN = fwd ? ng : (bck ? pgam : gam);              // (the parent's path metric is what pgam holds)
        pgam = fwd ? gam : (bck ? pgamB : pgam);
        gam = gamN;
        enc = fwd ? encF : (bck ? encB : (enc ^ (w >> 31)));
        w = fwd ? wF : (bck ? wB : (w & 0x7fffffffu));
        cur = ((int)w < 0) ? fano_tm1(w) : fano_tm0(w);
        inback = bck && !b1 && !b2;
        pos = posN;
        if (arrived) {                             // reached the last node: decoded (rare, once per lane)
            if (act) {
                act = false;
                r_rc = (it >= limit) ? -1 : 0;     // (a decode in the very last cycle counts as a timeout, fano.c:234)
                r_metric = (unsigned)gam;
                r_cycles = (it > limit) ? limit + 2u : it + 1u;   // (it > limit: a lane past the limit, not yet noticed)
                r_maxnp = (unsigned)maxnp;
                if (r_rc == 0) hook.success(r_cycles, mem, node_base);
            }
            thr = PARKED;                          // park: stay put (pos = nbits - 1), tightening an unreachable threshold
            inback = false;
        }
    }
    out.rc = r_rc;
    out.metric = r_metric;
    out.cycles = r_cycles;
    out.maxnp = r_maxnp;
#pragma unroll
    for (int b = 0; b < 12; b++) out.data[b] = 0;
    if (want && r_rc == 0) {
#pragma unroll
        for (int b = 0; b < (nbits >> 3); b++) out.data[b] = (unsigned char)mem.ld(node_base + row * (unsigned)(7 + 8 * b)).x;
    }
}

template <bool EXACT, bool PIPE = false, typename Hook, typename Mem>
__device__ __forceinline__ void fano_dense(FanoResult &out, bool want, const unsigned char *__restrict__ symbols,
                                           const short *__restrict__ mettab, int delta, unsigned maxcycles, unsigned stop_after,
                                           Hook hook, Mem mem) {
    if (delta > 10) fano_dense_impl<EXACT, true, PIPE>(out, want, symbols, mettab, delta, maxcycles, stop_after, hook, mem);
    else fano_dense_impl<EXACT, false, PIPE>(out, want, symbols, mettab, delta, maxcycles, stop_after, hook, mem);
}

}  // namespace wspr
