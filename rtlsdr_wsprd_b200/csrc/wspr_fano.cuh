// Device-side Fano sequential decoder (K=32, r=1/2; wsprd/fano.c:87-238), latency-tuned.
//
// A decode that fails costs maxcycles*nbits = 810 000 strictly sequential tree moves, so what matters on a GPU thread
// is the dependent latency of ONE move.  The portable version in wspr_codec.cuh keeps the tree in five indexed
// arrays (every move is a chain of dependent local-memory round trips, ~30 cycles each).  Here
//   * the node the decoder stands on (encoder state, path metric, the two sorted branch metrics, branch index) and
//     the path metric of its parent live in registers: a forward move touches no memory on its critical path,
//   * a node is one 16-byte record, so stepping back is a single vector load,
//   * the four branch metrics of a tree level are one 8-byte record fetched at the top of the move.
// The sequence of moves, the cycle count, the final metric and the decoded bytes are identical to fano.c (checked
// against the oracle on random symbol vectors, including timeouts, in tests/test_gpu_parity.py).
#pragma once
#include "wspr_codec.cuh"

namespace wspr {

struct FanoResult {
    int rc;              // 0 decoded, -1 timeout, FANO_STOPPED cut short
    unsigned metric, cycles, maxnp;
    unsigned char data[12];
};

__device__ __forceinline__ int fano_pick(uint2 m, unsigned ls) {   // m = four int16 metrics, index ls
    const unsigned long long v = ((unsigned long long)m.y << 32) | m.x;
    return (int)(short)(unsigned short)(v >> (16 * ls));
}

constexpr int FANO_NODE_WORDS = NBITS + 2;   // uint4 records
constexpr int FANO_BM_WORDS = NBITS + 1;     // uint2 records
constexpr int FANO_STATE_BYTES = FANO_NODE_WORDS * 16 + FANO_BM_WORDS * 8 + 8;   // multiple of 16

// node / bm: working storage supplied by the caller -- per-thread local arrays when every lane of a warp decodes
// (the local-memory interleave keeps that cache friendly), shared memory when a single lane per warp decodes (a lone
// lane would use 1/32 of every local-memory line and fall out of L1).
template <typename Poll>
__device__ __forceinline__ void fano_fast(FanoResult &out, const unsigned char *__restrict__ symbols, const short *__restrict__ mettab,
                                          int delta, unsigned maxcycles, unsigned stop_after, Poll poll, uint4 *node, uint2 *bm) {
    constexpr int nbits = NBITS;
    constexpr int last = nbits - 1, tail = nbits - 31;
    // node[n] = {enc, gam, tm0 | tm1 << 16, sel} of the nodes on the current path below `pos`
    // bm[n]   = branch metrics of level n: (symbol pair 00, 01), (10, 11)
#pragma unroll 1
    for (int n = 0; n < nbits; n++) {
        const int a0 = mettab[symbols[2 * n]], a1 = mettab[256 + symbols[2 * n]];
        const int b0 = mettab[symbols[2 * n + 1]], b1 = mettab[256 + symbols[2 * n + 1]];
        bm[n] = make_uint2(((unsigned)(a0 + b0) & 0xffffu) | ((unsigned)(a0 + b1) << 16),
                           ((unsigned)(a1 + b0) & 0xffffu) | ((unsigned)(a1 + b1) << 16));
        node[n] = make_uint4(0, 0, 0, 0);
    }
    bm[nbits] = make_uint2(0, 0);
    node[nbits] = node[nbits + 1] = make_uint4(0, 0, 0, 0);

    int pos = 0, thr = 0, maxnp = 0;
    unsigned enc = 0;
    int gam = 0, pgam = 0, tm0, tm1, sel = 0;     // pgam = path metric of the parent node (valid when pos > 0)
    {
        const int m0 = fano_pick(bm[0], 0), m1 = fano_pick(bm[0], 3);   // branch_sym(0) == 0
        if (m0 > m1) { tm0 = m0; tm1 = m1; }
        else { tm0 = m1; tm1 = m0; enc = 1; }
    }
    const unsigned limit = maxcycles * (unsigned)nbits;
    const unsigned stop = (stop_after != 0 && stop_after < limit) ? stop_after : 0;
    unsigned it;
    bool stopped = false;
#pragma unroll 1
    for (it = 1; it <= limit; it++) {
        if ((it & 1023u) == 0 && ((stop != 0 && it >= stop) || poll(it))) {
            stopped = true;
            break;
        }
        const uint2 nbm = bm[pos + 1];            // issued early; only the forward move consumes it
        if (pos > maxnp) maxnp = pos;
        const int ng = gam + (sel ? tm1 : tm0);
        if (ng >= thr) {                          // ---- forward
            if (gam < thr + delta)
                while (ng >= thr + delta) thr += delta;
            node[pos] = make_uint4(enc, (unsigned)gam, ((unsigned)tm0 & 0xffffu) | ((unsigned)tm1 << 16), (unsigned)sel);
            pgam = gam;
            gam = ng;
            unsigned e = enc << 1;
            pos++;
            if (pos == last + 1) {
                enc = e;
                break;
            }
            const unsigned ls = branch_sym(e);
            const int m0 = fano_pick(nbm, ls);
            if (pos >= tail) {
                tm0 = m0;
            } else {
                const int m1 = fano_pick(nbm, 3u ^ ls);
                if (m0 > m1) { tm0 = m0; tm1 = m1; }
                else { tm0 = m1; tm1 = m0; e |= 1u; }
            }
            enc = e;
            sel = 0;
            continue;
        }
        for (;;) {                                // ---- backward
            if (pos == 0 || pgam < thr) {
                thr -= delta;
                if (sel != 0) {
                    sel = 0;
                    enc ^= 1u;
                }
                break;
            }
            pos--;
            const uint4 nd = node[pos];
            const int gp = (pos > 0) ? (int)node[pos - 1].y : 0;
            enc = nd.x;
            gam = (int)nd.y;
            tm0 = (int)(short)(nd.z & 0xffffu);
            tm1 = (int)(short)(nd.z >> 16);
            sel = (int)nd.w;
            pgam = gp;
            if (pos < tail && sel != 1) {
                sel++;
                enc ^= 1u;
                break;
            }
        }
    }
    node[pos] = make_uint4(enc, (unsigned)gam, 0, 0);
    out.metric = (unsigned)gam;
    for (int b = 0; b < 12; b++) out.data[b] = 0;
    for (int b = 0; b < (nbits >> 3); b++) out.data[b] = (unsigned char)node[7 + 8 * b].x;
    out.cycles = it + 1;
    out.maxnp = (unsigned)maxnp;
    out.rc = stopped ? FANO_STOPPED : ((it >= limit) ? -1 : 0);   // (a decode in the very last cycle counts as a timeout, fano.c:234)
}

// ---------------------------------------------------------------------------------------------------------
// Single-lane form for the long runners, tree state in shared memory (a lone lane would use 1/32 of every
// local-memory line and fall out of L1).  Tuned for the dependent latency of one move:
//   * per tree level the four possible (better metric, worse metric, bit) triples are precomputed, so arriving at a
//     node is a 4-way register select instead of two metric fetches, a compare and a swap;
//   * a node record is 16 bytes {enc, gam, packed metrics | branch index, gam of the parent}: one vector store when a
//     node is left forwards, ONE vector load per step back (the parent's metric needed for the next test rides along);
//   * shared memory is addressed through a 32-bit window address held in a register (ld/st.shared), and the
//     threshold loop is kept a loop -- left alone the compiler re-derives the window base and inserts an integer
//     division inside the hot loop.
// ---------------------------------------------------------------------------------------------------------
struct FanoSharedState {               // one per attempt, 16-byte aligned
    uint4 lvl[NBITS + 1];              // per level, indexed by the 2-bit branch symbol: packed (tm0, tm1, bit)
    uint4 node[NBITS + 2];
};

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

template <typename Poll>
__device__ __forceinline__ void fano_shared(FanoResult &out, const unsigned char *__restrict__ symbols,
                                            const short *__restrict__ mettab, int delta, unsigned maxcycles, unsigned stop_after,
                                            Poll poll, FanoSharedState &st) {
    constexpr int nbits = NBITS;
    constexpr int last = nbits - 1, tail = nbits - 31;
#pragma unroll 1
    for (int n = 0; n <= nbits; n++) {
        uint4 e = make_uint4(0, 0, 0, 0);
        if (n < nbits) {
            const int a0 = mettab[symbols[2 * n]], a1 = mettab[256 + symbols[2 * n]];
            const int b0 = mettab[symbols[2 * n + 1]], b1 = mettab[256 + symbols[2 * n + 1]];
            const int m[4] = {a0 + b0, a0 + b1, a1 + b0, a1 + b1};
            unsigned w[4];
#pragma unroll
            for (int ls = 0; ls < 4; ls++) {
                const int m0 = m[ls], m1 = m[3 ^ ls];
                if (n >= tail) w[ls] = fano_pack(m0, 0, 0);
                else w[ls] = (m0 > m1) ? fano_pack(m0, m1, 0) : fano_pack(m1, m0, 1);
            }
            e = make_uint4(w[0], w[1], w[2], w[3]);
        }
        st.lvl[n] = e;
        st.node[n] = make_uint4(0, 0, 0, 0);
    }
    st.node[nbits + 1] = make_uint4(0, 0, 0, 0);
    unsigned lvl_base = (unsigned)__cvta_generic_to_shared(&st.lvl[0]);
    unsigned node_base = (unsigned)__cvta_generic_to_shared(&st.node[0]);
    asm volatile("" : "+r"(lvl_base), "+r"(node_base));   // opaque: keep the window addresses in registers

    int pos = 0, thr = 0, maxnp = 0;
    unsigned w = st.lvl[0].x;                     // branch_sym(0) == 0
    unsigned enc = w >> 30;
    int gam = 0, pgam = 0, sel = 0;
    const unsigned limit = maxcycles * (unsigned)nbits;
    const unsigned stop = (stop_after != 0 && stop_after < limit) ? stop_after : 0;
    unsigned it;
    bool stopped = false;
#pragma unroll 1
    for (it = 1; it <= limit; it++) {
        if ((it & 1023u) == 0 && ((stop != 0 && it >= stop) || poll(it))) {
            stopped = true;
            break;
        }
        const uint4 nl = lds128(lvl_base + 16u * (unsigned)(pos + 1));   // issued early; only the forward move uses it
        maxnp = max(maxnp, pos);
        const int ng = gam + (sel ? fano_tm1(w) : fano_tm0(w));
        if (ng >= thr) {                          // ---- forward
            if (gam < thr + delta) {
#pragma unroll 1
                while (ng >= thr + delta) {
                    thr += delta;
                    asm volatile("" : "+r"(thr));  // (keeps this a loop: the closed form needs an integer division)
                }
            }
            sts128(node_base + 16u * (unsigned)pos, enc, (unsigned)gam, w | ((unsigned)sel << 31), (unsigned)pgam);
            pgam = gam;
            gam = ng;
            const unsigned e = enc << 1;
            pos++;
            if (pos == last + 1) {
                enc = e;
                break;
            }
            const unsigned pa = __popc(e & POLY_A), pb = __popc(e & POLY_B);
            const unsigned wlo = (pb & 1u) ? nl.y : nl.x, whi = (pb & 1u) ? nl.w : nl.z;   // ls = 2*(pa&1) + (pb&1)
            w = (pa & 1u) ? whi : wlo;
            enc = e | (w >> 30);
            sel = 0;
            continue;
        }
        if (pos == 0 || pgam < thr) {             // ---- tighten the threshold, stay on this node
            thr -= delta;
            enc ^= (unsigned)sel;                 // (sel is 0 or 1: undo the branch flip)
            sel = 0;
            continue;
        }
#pragma unroll 1
        for (;;) {                                // ---- step back
            pos--;
            const uint4 nd = lds128(node_base + 16u * (unsigned)pos);
            enc = nd.x;
            gam = (int)nd.y;
            w = nd.z & 0x7fffffffu;
            sel = (int)(nd.z >> 31);
            pgam = (int)nd.w;
            if (pos < tail && sel != 1) {         // take the other branch of this node
                sel = 1;
                enc ^= 1u;
                break;
            }
            if (pos == 0 || pgam < thr) {         // cannot go further up: tighten, stay here
                thr -= delta;
                enc ^= (unsigned)sel;
                sel = 0;
                break;
            }
        }
    }
    sts128(node_base + 16u * (unsigned)pos, enc, (unsigned)gam, 0u, 0u);
    out.metric = (unsigned)gam;
    for (int b = 0; b < 12; b++) out.data[b] = 0;
    for (int b = 0; b < (nbits >> 3); b++) out.data[b] = (unsigned char)st.node[7 + 8 * b].x;
    out.cycles = it + 1;
    out.maxnp = (unsigned)maxnp;
    out.rc = stopped ? FANO_STOPPED : ((it >= limit) ? -1 : 0);
}

}  // namespace wspr
