// K0  front end: raw RTL-SDR stream (interleaved u8 I/Q at 2.4 Msps) -> 375 sps float I/Q, bit-identical to
// rtlsdr_callback (rtlsdr_wsprd.c:126-244): fs/4 mixer (:171-182), two int32 integrators (:190-195), decimation by
// 6401 (:198-202), two delay-2 combs (:204-218), 33-tap float FIR evaluated in tap order (:221-234).
//
// The reference runs the integrators sample by sample.  They are linear over Z/2^32, so the value of the second
// integrator at the decimation instant of block m (a block = 6401 consecutive samples) is
//       v[m] = (m+1)*6401 * sum_{b<=m} S0[b]  -  sum_{b<=m} (b*6401*S0[b] + S1[b])        (mod 2^32)
// with per-block moments S0[b] = sum x[t], S1[b] = sum t*x[t] (t = index inside the block, x = mixer output).
//   k_block_moments : the HBM-bound pass -- one warp per block, 16-byte loads, byte sums with dp4a
//   k_comb_fir      : a wrapping prefix scan of the moments, the two combs and the FIR, 1024 outputs per CTA
// Compiled with -fmad=false (the FIR multiplies and adds round separately, like the reference's x86-64 build).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <string>

#include "../../include/wspr_b200.h"
#include "wspr_kernels.cuh"

namespace wspr {

extern unsigned long long g_frontend_launches;
unsigned long long g_frontend_launches = 0;

constexpr int DECIM = 6401;                 // DOWNSAMPLING + 1, see the header comment
constexpr int BLOCK_BYTES = 2 * DECIM;      // 12802
constexpr int FIR_TAPS = 32;

// The 33 FIR coefficients are data of the reference (rtlsdr_wsprd.c:142-152), symmetric; the 17 distinct values:
__constant__ float c_fir[33];
static const float h_fir_half[17] = {-0.0027772683, -0.0005058826, 0.0049745750,  -0.0034059318, -0.0077557814, 0.0139375423,
                                     0.0039896935,  -0.0299394142, 0.0162250643,  0.0405130860,  -0.0580746013, -0.0272104968,
                                     0.1183705475,  -0.0306029022, -0.2011241667, 0.1615898423,  0.5000000000};

int decimate_outputs(size_t n_iq) { return (int)std::min<size_t>(n_iq / DECIM, 1u << 30); }

__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c) {   // sum of (unsigned byte of a) * (signed byte of b) + c
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {        // read-once data: bypass L1 allocation
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

// Mixer.  A 16-byte vector holds samples n = 8V .. 8V+7 as bytes (I0 Q0 I1 Q1 | I2 Q2 I3 Q3 | ...), and 8V = 0 mod 4, so
// word 0/2 holds mixer phases 0,1 and word 1/3 phases 2,3:
//   phase 0: (I,Q) ; 1: (-Q, I) ; 2: (-I,-Q) ; 3: (Q,-I)                       (rtlsdr_wsprd.c:171-182)
// The reference negates in int8, so -(-128) stays -128.  In the offset-binary form the bytes arrive in (u = s + 128) that
// wrap-around is exactly the negation of the BYTE modulo 256: u' = (-u) & 255 (u = 0 -> 0, i.e. s = -128 -> -128).  So the
// bytes at the negated positions -- byte 3 of words 0/2 (phase 1: -Q), bytes 0,1,2 of words 1/3 (phase 2: -I, -Q; phase 3:
// -I) -- are negated byte-wise (carry-free SWAR), after which EVERY word holds the mixer outputs in the same places:
// I_out of its two samples in bytes 0 and 3, Q_out in bytes 1 and 2, all as unsigned bytes with offset 128.  The moments
// are then plain dp4a sums with non-negative weights; no value-dependent path is left (round 1 detected zero bytes and
// corrected them in a divergent slow path, which uniform random test data took for two thirds of all vectors).
#define PACK4(a, b, c, d) ((int)(((unsigned)(a)&255u) | (((unsigned)(b)&255u) << 8) | (((unsigned)(c)&255u) << 16) | (((unsigned)(d)&255u) << 24)))

struct Moments {
    unsigned s0i, s0q, s1i, s1q;                   // sums over Z/2^32 (the reference's int32 integrators wrap, rtlsdr_wsprd.c:130-135,192-195)
};

__device__ __forceinline__ unsigned neg_byte3(unsigned u) { return (u ^ 0xff000000u) + 0x01000000u; }
__device__ __forceinline__ unsigned neg_bytes012(unsigned u) {
    const unsigned lo = (u ^ 0x00ffffffu) & 0x7f7f7f7fu, hi = (u ^ 0x00ffffffu) & 0x80808080u;
    return (lo + 0x00010101u) ^ hi;
}

// accumulate one 16-byte vector whose first sample has in-block index t0 (may be negative at the leading edge;
// bytes outside the block are already forced to 128, which stays 128 under the byte negation)
__device__ __forceinline__ void accumulate_vector(Moments &m, const uint4 v, int t0) {
    const unsigned x = neg_byte3(v.x), y = neg_bytes012(v.y), z = neg_byte3(v.z), w = neg_bytes012(v.w);
    // zeroth moment: the 8 mixer outputs of each channel; -128 per byte = -1024
    int ai = dp4a_us(x, PACK4(1, 0, 0, 1), -1024);
    ai = dp4a_us(y, PACK4(1, 0, 0, 1), ai);
    ai = dp4a_us(z, PACK4(1, 0, 0, 1), ai);
    ai = dp4a_us(w, PACK4(1, 0, 0, 1), ai);
    int aq = dp4a_us(x, PACK4(0, 1, 1, 0), -1024);
    aq = dp4a_us(y, PACK4(0, 1, 1, 0), aq);
    aq = dp4a_us(z, PACK4(0, 1, 1, 0), aq);
    aq = dp4a_us(w, PACK4(0, 1, 1, 0), aq);
    // first moment about the vector's first sample (offsets 0..7; -128 * 28 = -3584), accumulated straight into S1
    int bi = dp4a_us(x, PACK4(0, 0, 0, 1), (int)(m.s1i - 3584u));
    bi = dp4a_us(y, PACK4(2, 0, 0, 3), bi);
    bi = dp4a_us(z, PACK4(4, 0, 0, 5), bi);
    bi = dp4a_us(w, PACK4(6, 0, 0, 7), bi);
    int bq = dp4a_us(x, PACK4(0, 0, 1, 0), (int)(m.s1q - 3584u));
    bq = dp4a_us(y, PACK4(0, 2, 3, 0), bq);
    bq = dp4a_us(z, PACK4(0, 4, 5, 0), bq);
    bq = dp4a_us(w, PACK4(0, 6, 7, 0), bq);
    m.s0i += (unsigned)ai;                         // (unsigned: wrap-around is the arithmetic here, not an accident)
    m.s0q += (unsigned)aq;
    m.s1i = (unsigned)t0 * (unsigned)ai + (unsigned)bi;
    m.s1q = (unsigned)t0 * (unsigned)aq + (unsigned)bq;
}

// force the bytes of a vector that lie outside [lo, hi) (byte offsets relative to the vector start) to 128
__device__ __forceinline__ uint4 mask_vector(uint4 v, int lo, int hi) {
    unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        unsigned keep = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            int pos = 4 * k + b;
            if (pos >= lo && pos < hi) keep |= 0xffu << (8 * b);
        }
        w[k] = (w[k] & keep) | (0x80808080u & ~keep);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

constexpr int MOM_WARPS = 8;
constexpr int MOM_UNROLL = 5;

// one warp per 6401-sample block.  raw must be 16-byte aligned per stream.
// first_byte: offset of block 0 inside each stream's buffer; the buffer's byte 0 must be sample 0 (mod 8) of the stream,
// because the mixer phase is taken from the position inside the 16-byte vector.
__global__ void __launch_bounds__(MOM_WARPS * 32) k_block_moments(const uint8_t *__restrict__ raw, size_t stream_stride,
                                                                  int nblk, uint4 *__restrict__ moments, int first_byte) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * MOM_WARPS + warp, stream = blockIdx.y;
    if (b >= nblk) return;
    const uint8_t *base = raw + (size_t)stream * stream_stride;
    const long long byte0 = (long long)first_byte + (long long)b * BLOCK_BYTES, byte1 = byte0 + BLOCK_BYTES;
    const long long v0 = byte0 >> 4, v1 = (byte1 + 15) >> 4;        // vectors [v0, v1) cover the block
    const int nvec = (int)(v1 - v0);                                 // 801 or 802
    const uint4 *vp = reinterpret_cast<const uint4 *>(base) + v0;
    const int lead = (int)(byte0 - (v0 << 4));                       // bytes of vector 0 before the block
    const int tail = (int)((v1 << 4) - byte1);                       // bytes of the last vector after the block
    const int tfirst = -(lead >> 1);                                 // in-block index of vector 0's first sample
    Moments m = {0, 0, 0, 0};
    // interior vectors 1 .. nvec-2 need no masking
    int i = 1 + lane;
    for (; i + 32 * (MOM_UNROLL - 1) < nvec - 1; i += 32 * MOM_UNROLL) {
        uint4 v[MOM_UNROLL];
#pragma unroll
        for (int u = 0; u < MOM_UNROLL; u++) v[u] = ld_stream(vp + i + 32 * u);
#pragma unroll
        for (int u = 0; u < MOM_UNROLL; u++) accumulate_vector(m, v[u], tfirst + 8 * (i + 32 * u));
    }
    for (; i < nvec - 1; i += 32) accumulate_vector(m, ld_stream(vp + i), tfirst + 8 * i);
    if (lane == 0) accumulate_vector(m, mask_vector(ld_stream(vp), lead, 16), tfirst);
    if (lane == 1) accumulate_vector(m, mask_vector(ld_stream(vp + nvec - 1), 0, 16 - tail), tfirst + 8 * (nvec - 1));
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        m.s0i += __shfl_xor_sync(0xffffffffu, m.s0i, s);
        m.s0q += __shfl_xor_sync(0xffffffffu, m.s0q, s);
        m.s1i += __shfl_xor_sync(0xffffffffu, m.s1i, s);
        m.s1q += __shfl_xor_sync(0xffffffffu, m.s1q, s);
    }
    if (lane == 0) moments[(size_t)stream * nblk + b] = make_uint4((unsigned)m.s0i, (unsigned)m.s0q, (unsigned)m.s1i, (unsigned)m.s1q);
}

// The same pass with the block staged through shared memory by the bulk-copy engine (cp.async.bulk + mbarrier, the 1-D form
// of TMA): one warp per block as above, the block's covering vectors fetched as MOMB_PIECES bulk copies of 16-byte-aligned
// ranges into the warp's slab, each completing on its own mbarrier, so the first piece is summed while the others are in
// flight.  Selected with WSPR_K0_BULK=1 (an A/B of the load path: profiles/r2_k0_bulk_ab.txt); the arithmetic is identical.
constexpr int MOMB_WARPS = 4;
constexpr int MOMB_PIECES = 4;
constexpr int MOMB_SLAB = 803 * 16;                           // a block spans at most 802 vectors
constexpr int MOMB_SMEM = MOMB_WARPS * MOMB_SLAB;
__device__ __forceinline__ uint4 lds_vec(unsigned addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__global__ void __launch_bounds__(MOMB_WARPS * 32) k_block_moments_bulk(const uint8_t *__restrict__ raw, size_t stream_stride,
                                                                       int nblk, uint4 *__restrict__ moments, int first_byte) {
    extern __shared__ __align__(128) unsigned char momb_smem[];
    __shared__ __align__(8) unsigned long long mbar[MOMB_WARPS][MOMB_PIECES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * MOMB_WARPS + warp, stream = blockIdx.y;
    if (b >= nblk) return;                                     // (warps are independent: no CTA barrier below)
    const uint8_t *base = raw + (size_t)stream * stream_stride;
    const long long byte0 = (long long)first_byte + (long long)b * BLOCK_BYTES, byte1 = byte0 + BLOCK_BYTES;
    const long long v0 = byte0 >> 4, v1 = (byte1 + 15) >> 4;
    const int nvec = (int)(v1 - v0);
    const uint4 *vp = reinterpret_cast<const uint4 *>(base) + v0;
    const int lead = (int)(byte0 - (v0 << 4)), tail = (int)((v1 << 4) - byte1);
    const int tfirst = -(lead >> 1);
    const unsigned slab = (unsigned)__cvta_generic_to_shared(momb_smem + (size_t)warp * MOMB_SLAB);
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&mbar[warp][0]);
    const int per = (nvec + MOMB_PIECES - 1) / MOMB_PIECES;
    if (lane == 0) {
#pragma unroll
        for (int p = 0; p < MOMB_PIECES; p++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * p), "r"(1));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
        for (int p = 0; p < MOMB_PIECES; p++) {
            const int p0 = p * per, p1 = min(p0 + per, nvec);
            const unsigned bytes = (unsigned)(p1 - p0) * 16u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * p), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(slab + 16u * p0),
                         "l"(vp + p0), "r"(bytes), "r"(bar0 + 8u * p)
                         : "memory");
        }
    }
    __syncwarp();
    Moments m = {0, 0, 0, 0};
#pragma unroll 1
    for (int p = 0; p < MOMB_PIECES; p++) {
        unsigned ok;
        do {
            asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }"
                         : "=r"(ok)
                         : "r"(bar0 + 8u * p), "r"(0)
                         : "memory");
        } while (!ok);
        const int p0 = max(p * per, 1), p1 = min(min(p * per + per, nvec), nvec - 1);   // interior vectors of the piece
        int i = p0 + lane;
        for (; i + 32 * (MOM_UNROLL - 1) < p1; i += 32 * MOM_UNROLL) {
            uint4 v[MOM_UNROLL];
#pragma unroll
            for (int u = 0; u < MOM_UNROLL; u++) v[u] = lds_vec(slab + 16u * (i + 32 * u));
#pragma unroll
            for (int u = 0; u < MOM_UNROLL; u++) accumulate_vector(m, v[u], tfirst + 8 * (i + 32 * u));
        }
        for (; i < p1; i += 32) accumulate_vector(m, lds_vec(slab + 16u * i), tfirst + 8 * i);
    }
    if (lane == 0) accumulate_vector(m, mask_vector(lds_vec(slab), lead, 16), tfirst);
    if (lane == 1) accumulate_vector(m, mask_vector(lds_vec(slab + 16u * (nvec - 1)), 0, 16 - tail), tfirst + 8 * (nvec - 1));
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        m.s0i += __shfl_xor_sync(0xffffffffu, m.s0i, s);
        m.s0q += __shfl_xor_sync(0xffffffffu, m.s0q, s);
        m.s1i += __shfl_xor_sync(0xffffffffu, m.s1i, s);
        m.s1q += __shfl_xor_sync(0xffffffffu, m.s1q, s);
    }
    if (lane == 0) moments[(size_t)stream * nblk + b] = make_uint4((unsigned)m.s0i, (unsigned)m.s0q, (unsigned)m.s1i, (unsigned)m.s1q);
}

// Second pass: v = (b+1)*6401*P - W with the wrapping prefix sums P = sum S0, W = sum (b*6401*S0 + S1), the two combs and
// the FIR.  One CTA per CF_CHUNK outputs of a stream: it first sums the moments of every earlier block of its stream (a
// redundant prefix, 16 bytes per block out of L2 -- 3 % of the traffic of the first pass), then scans its own blocks
// (plus the 36 earlier ones the combs and the FIR reach back to) in shared memory.
constexpr int CF_THREADS = 256;
constexpr int CF_CHUNK = 1024;                                // outputs per CTA
constexpr int CF_HALO = FIR_TAPS + 4;                         // 32 FIR taps back + two delay-2 combs
constexpr int CF_SPAN = CF_CHUNK + CF_HALO;
__global__ void __launch_bounds__(CF_THREADS) k_comb_fir(const uint4 *__restrict__ moments, int nblk, float *__restrict__ I,
                                                         float *__restrict__ Q, int out_stride, int max_out) {
    __shared__ uint4 red[CF_THREADS];
    __shared__ uint4 loc[CF_SPAN];                             // inclusive (P_i, P_q, W_i, W_q) of the chunk's blocks
    __shared__ uint2 val[CF_SPAN];
    const int stream = blockIdx.y, t = threadIdx.x;
    const int m0 = blockIdx.x * CF_CHUNK;                      // first output of this CTA
    const uint4 *mom = moments + (size_t)stream * nblk;
    const int nout = min(nblk, max_out);
    if (m0 < nout) {
        const int h0 = max(m0 - CF_HALO, 0);                   // first block whose v we need
        const int h1 = min(m0 + CF_CHUNK, nout);
        // (1) prefix over blocks [0, h0)
        uint4 acc = make_uint4(0, 0, 0, 0);
        for (int b = t; b < h0; b += CF_THREADS) {
            const uint4 s = mom[b];
            const unsigned off = (unsigned)b * (unsigned)DECIM;
            acc.x += s.x;
            acc.y += s.y;
            acc.z += off * s.x + s.z;
            acc.w += off * s.y + s.w;
        }
        red[t] = acc;
        __syncthreads();
        for (int d = CF_THREADS / 2; d > 0; d >>= 1) {
            if (t < d) {
                const uint4 o = red[t + d], p = red[t];
                red[t] = make_uint4(p.x + o.x, p.y + o.y, p.z + o.z, p.w + o.w);
            }
            __syncthreads();
        }
        const uint4 base = red[0];
        // (2) inclusive scan of blocks [h0, h1): per-thread runs, then a scan of the run totals
        const int n = h1 - h0, per = (n + CF_THREADS - 1) / CF_THREADS;
        const int r0 = min(t * per, n), r1 = min(r0 + per, n);
        uint4 run = make_uint4(0, 0, 0, 0);
        for (int k = r0; k < r1; k++) {
            const uint4 s = mom[h0 + k];
            const unsigned off = (unsigned)(h0 + k) * (unsigned)DECIM;
            run.x += s.x;
            run.y += s.y;
            run.z += off * s.x + s.z;
            run.w += off * s.y + s.w;
            loc[k] = run;
        }
        __syncthreads();
        red[t] = run;
        __syncthreads();
        for (int d = 1; d < CF_THREADS; d <<= 1) {             // Hillis-Steele (wrapping adds are associative)
            uint4 o = make_uint4(0, 0, 0, 0);
            if (t >= d) o = red[t - d];
            __syncthreads();
            if (t >= d) {
                const uint4 p = red[t];
                red[t] = make_uint4(p.x + o.x, p.y + o.y, p.z + o.z, p.w + o.w);
            }
            __syncthreads();
        }
        const uint4 before = (t == 0) ? make_uint4(0, 0, 0, 0) : red[t - 1];
        for (int k = r0; k < r1; k++) {
            const uint4 l = loc[k];
            const unsigned px = base.x + before.x + l.x, py = base.y + before.y + l.y;
            const unsigned wz = base.z + before.z + l.z, ww = base.w + before.w + l.w;
            const unsigned n1 = (unsigned)(h0 + k + 1) * (unsigned)DECIM;
            val[k] = make_uint2(n1 * px - wz, n1 * py - ww);
        }
        __syncthreads();
        // (3) combs f[k] = v[k] - 2 v[k-2] + v[k-4] (zero history) and the FIR in tap order
        for (int m = m0 + t; m < h1; m += CF_THREADS) {
            float si = 0.0f, sq = 0.0f;
#pragma unroll 1
            for (int j = 0; j <= FIR_TAPS; j++) {
                const int k = m - FIR_TAPS + j;
                float fi = 0.0f, fq = 0.0f;
                if (k >= 0) {
                    const uint2 a = val[k - h0];
                    const uint2 c = (k >= 2) ? val[k - 2 - h0] : make_uint2(0, 0);
                    const uint2 e = (k >= 4) ? val[k - 4 - h0] : make_uint2(0, 0);
                    fi = (float)(int)(a.x - 2u * c.x + e.x);
                    fq = (float)(int)(a.y - 2u * c.y + e.y);
                }
                const float z = c_fir[j];
                si = si + fi * z;
                sq = sq + fq * z;
            }
            I[(size_t)stream * out_stride + m] = si;
            Q[(size_t)stream * out_stride + m] = sq;
        }
    }
    for (int m = max(m0, nout) + t; m < min(m0 + CF_CHUNK, max_out); m += CF_THREADS) {   // zero tail
        I[(size_t)stream * out_stride + m] = 0.0f;
        Q[(size_t)stream * out_stride + m] = 0.0f;
    }
}

static bool g_fir_uploaded[64] = {false};
static cudaError_t upload_fir() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && g_fir_uploaded[dev]) return cudaSuccess;
    float z[33];
    for (int j = 0; j < 33; j++) z[j] = h_fir_half[j <= 16 ? j : 32 - j];
    cudaError_t e = cudaMemcpyToSymbol(c_fir, z, sizeof z);
    if (e == cudaSuccess && dev >= 0 && dev < 64) g_fir_uploaded[dev] = true;
    return e;
}

// moments: [nstreams][nblk] uint4 (scratch, device)
void launch_decimate(const uint8_t *raw, size_t n_iq, int nstreams, size_t stream_stride_bytes, uint4 *moments, float *I,
                     float *Q, int out_stride, int max_out, cudaStream_t st) {
    const int nblk = decimate_outputs(n_iq);
    if (nstreams <= 0 || max_out <= 0) return;
    upload_fir();
    if (nblk > 0) {
#ifdef WSPR_EXPERIMENTS
        static const bool bulk = [] { const char *e = getenv("WSPR_K0_BULK"); return e && e[0] == '1'; }();   // load-path A/B
#else
        constexpr bool bulk = false;                   // (1 % slower than plain vector loads, profiles/r2_k0_bulk_ab.txt)
#endif
        if (bulk) {
            static bool attr[64] = {false};
            int dev = 0;
            cudaGetDevice(&dev);
            if (dev >= 0 && dev < 64 && !attr[dev]) {
                cudaFuncSetAttribute(k_block_moments_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, MOMB_SMEM);
                attr[dev] = true;
            }
            k_block_moments_bulk<<<dim3((nblk + MOMB_WARPS - 1) / MOMB_WARPS, nstreams), MOMB_WARPS * 32, MOMB_SMEM, st>>>(
                raw, stream_stride_bytes, nblk, (uint4 *)moments, 0);
        } else {
            k_block_moments<<<dim3((nblk + MOM_WARPS - 1) / MOM_WARPS, nstreams), MOM_WARPS * 32, 0, st>>>(raw, stream_stride_bytes,
                                                                                                          nblk, (uint4 *)moments, 0);
        }
        g_frontend_launches++;
    }
    k_comb_fir<<<dim3((max_out + CF_CHUNK - 1) / CF_CHUNK, nstreams), CF_THREADS, 0, st>>>(moments, nblk, I, Q, out_stride, max_out);
    g_frontend_launches++;
}

// ---- streaming form: the filter state of rtlsdr_callback carried from push to push (rtlsdr_wsprd.c:130-136,155-156) ----
struct FeChan {
    uint32_t x1, x2;                 // integrators Ix1, Ix2 (wrapping)
    uint32_t t1y, t1z, t2y, t2z;     // comb delay lines
    float fir[FIR_TAPS];             // firI / firQ
};
struct FeState {
    FeChan ch[2];
};
// One thread per (stream, channel) replays the per-output part of the callback for the nb blocks whose byte moments
// k_block_moments has just produced: over one block of n = 6401 samples x2 += n*x1 + n*S0 - S1 and x1 += S0 (what the
// sample-by-sample integrators add up to), then the two combs and the FIR exactly as the reference writes them.
__global__ void k_stream_advance(FeState *__restrict__ state, const uint4 *__restrict__ moments, int nb, int nstreams,
                                 float *__restrict__ I, float *__restrict__ Q, int out_stride, int out_pos, int cap) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int stream = tid >> 1, ch = tid & 1;
    if (stream >= nstreams) return;
    FeChan c = state[stream].ch[ch];
    float *out = (ch ? Q : I) + (size_t)stream * out_stride;
    for (int b = 0; b < nb; b++) {
        const uint4 m = moments[(size_t)stream * nb + b];
        const uint32_t s0 = ch ? m.y : m.x, s1 = ch ? m.w : m.z;
        c.x2 += (uint32_t)DECIM * c.x1 + (uint32_t)DECIM * s0 - s1;
        c.x1 += s0;
        const uint32_t y1 = c.x2 - c.t1z;                      // rtlsdr_wsprd.c:204-218
        c.t1z = c.t1y;
        c.t1y = c.x2;
        const uint32_t y2 = y1 - c.t2z;
        c.t2z = c.t2y;
        c.t2y = y1;
        float sum = 0.0f;                                      // rtlsdr_wsprd.c:221-234
        for (int j = 0; j < FIR_TAPS; j++) {
            sum = sum + c.fir[j] * c_fir[j];
            if (j < FIR_TAPS - 1) c.fir[j] = c.fir[j + 1];
        }
        c.fir[FIR_TAPS - 1] = (float)(int32_t)y2;
        sum = sum + c.fir[FIR_TAPS - 1] * c_fir[FIR_TAPS];
        if (out_pos + b < cap) out[out_pos + b] = sum;         // rtlsdr_wsprd.c:237-242
    }
    state[stream].ch[ch] = c;
}

}  // namespace wspr

using namespace wspr;

static thread_local std::string g_fe_err;
extern "C" const char *wspr_frontend_last_error(void) { return g_fe_err.c_str(); }
static int fe_fail(int code, const char *what, cudaError_t e = cudaSuccess) {
    g_fe_err = what;
    if (e != cudaSuccess) {
        g_fe_err += ": ";
        g_fe_err += cudaGetErrorString(e);
    }
    return code;
}
#define FCK(call)                                                        \
    do {                                                                 \
        cudaError_t e_ = (call);                                         \
        if (e_ != cudaSuccess) { rc = fe_fail(WSPR_ERR_CUDA, #call, e_); goto done; } \
    } while (0)

static float g_fe_last_ms = 0.0f;
extern "C" float wspr_decimate_last_ms(void) { return g_fe_last_ms; }

extern "C" int wspr_decimate_device(const uint8_t *d_raw, int nstreams, size_t n_iq, size_t stream_stride_bytes, float *dI,
                                    float *dQ, int out_stride, int max_out, int device) {
    if (nstreams < 0 || !d_raw || !dI || !dQ || max_out < 0 || out_stride < max_out) return fe_fail(WSPR_ERR_ARG, "wspr_decimate_device: bad arguments");
    if (((uintptr_t)d_raw & 15) || (stream_stride_bytes & 15) || stream_stride_bytes < 2 * n_iq)
        return fe_fail(WSPR_ERR_ARG, "wspr_decimate_device: raw streams must be 16-byte aligned and strided");
    if (nstreams == 0 || max_out == 0) return 0;
    int rc = WSPR_OK;
    const int nblk = decimate_outputs(n_iq);
    uint4 *mom = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (device >= 0) FCK(cudaSetDevice(device));
    FCK(cudaMalloc((void **)&mom, (size_t)nstreams * std::max(nblk, 1) * sizeof(uint4)));
    FCK(cudaEventCreate(&e0));
    FCK(cudaEventCreate(&e1));
    FCK(cudaEventRecord(e0, 0));
    launch_decimate(d_raw, n_iq, nstreams, stream_stride_bytes, mom, dI, dQ, out_stride, max_out, 0);
    FCK(cudaEventRecord(e1, 0));
    FCK(cudaGetLastError());
    FCK(cudaEventSynchronize(e1));
    FCK(cudaEventElapsedTime(&g_fe_last_ms, e0, e1));
    rc = std::min(nblk, max_out);
done:
    if (mom) cudaFree(mom);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return rc;
}

extern "C" int wspr_decimate_batch(const uint8_t *raw, int nstreams, size_t n_iq, float *I, float *Q, int max_out, int device) {
    if (nstreams < 0 || !raw || !I || !Q || max_out < 0) return fe_fail(WSPR_ERR_ARG, "wspr_decimate_batch: bad arguments");
    if (nstreams == 0 || max_out == 0) return 0;
    int rc = WSPR_OK, nout = 0;
    const size_t bytes = 2 * n_iq, stride = (bytes + 15) / 16 * 16 + 16;
    // bound device memory: at most ~8 GiB of raw samples resident at a time
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)nstreams, ((size_t)8 << 30) / std::max<size_t>(stride, 1)));
    uint8_t *d_raw = nullptr;
    float *dI = nullptr, *dQ = nullptr;
    if (device >= 0) FCK(cudaSetDevice(device));
    FCK(cudaMalloc((void **)&d_raw, (size_t)chunk * stride));
    FCK(cudaMalloc((void **)&dI, (size_t)chunk * max_out * sizeof(float)));
    FCK(cudaMalloc((void **)&dQ, (size_t)chunk * max_out * sizeof(float)));
    for (int s0 = 0; s0 < nstreams; s0 += chunk) {
        const int n = std::min(chunk, nstreams - s0);
        FCK(cudaMemset(d_raw, 0x80, (size_t)n * stride));
        if (bytes) FCK(cudaMemcpy2D(d_raw, stride, raw + (size_t)s0 * bytes, bytes, bytes, n, cudaMemcpyHostToDevice));
        nout = wspr_decimate_device(d_raw, n, n_iq, stride, dI, dQ, max_out, max_out, -1);
        if (nout < 0) { rc = nout; goto done; }
        FCK(cudaMemcpy(I + (size_t)s0 * max_out, dI, (size_t)n * max_out * sizeof(float), cudaMemcpyDeviceToHost));
        FCK(cudaMemcpy(Q + (size_t)s0 * max_out, dQ, (size_t)n * max_out * sizeof(float), cudaMemcpyDeviceToHost));
    }
    rc = nout;
done:
    if (d_raw) cudaFree(d_raw);
    if (dI) cudaFree(dI);
    if (dQ) cudaFree(dQ);
    return rc;
}

// ================= streaming front end (the live daemon's callback + double buffer) =================
struct wspr_frontend {
    int device = 0, nstreams = 0, cap = 0;
    cudaStream_t st = nullptr;
    FeState *state = nullptr;
    uint8_t *stage[2] = {nullptr, nullptr};     // [nstreams][stage_stride]: the unfinished block + the chunk being pushed
    size_t stage_stride = 0;
    int cur_stage = 0;
    uint4 *moments = nullptr;
    size_t moments_cap = 0;
    float *I[2] = {nullptr, nullptr}, *Q[2] = {nullptr, nullptr};   // the reference's double buffer (rtlsdr_wsprd.c:80-87)
    int iq_index[2] = {0, 0};
    int buffer_index = 0;
    unsigned long long blocks_done = 0;         // decimation blocks consumed since the stream started
    size_t carry = 0;                           // bytes of the unfinished block waiting in the staging buffer
};

extern "C" void wspr_frontend_destroy(wspr_frontend *fe) {
    if (!fe) return;
    cudaSetDevice(fe->device);
    if (fe->st) cudaStreamSynchronize(fe->st);
    void *ptrs[] = {fe->state, fe->stage[0], fe->stage[1], fe->moments, fe->I[0], fe->I[1], fe->Q[0], fe->Q[1]};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (fe->st) cudaStreamDestroy(fe->st);
    delete fe;
}

static int fe_init(wspr_frontend *fe, int device, int nstreams, int slot_samples) {
    int rc = WSPR_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fe_fail(WSPR_ERR_CUDA, "no CUDA device");
    if (device < 0) FCK(cudaGetDevice(&device));
    FCK(cudaSetDevice(device));
    fe->device = device;
    fe->nstreams = nstreams;
    fe->cap = slot_samples > 0 ? slot_samples : WSPR_CAPTURE_SAMPLES;
    FCK(cudaStreamCreateWithFlags(&fe->st, cudaStreamNonBlocking));
    FCK(cudaMalloc((void **)&fe->state, (size_t)nstreams * sizeof(FeState)));
    FCK(cudaMemset(fe->state, 0, (size_t)nstreams * sizeof(FeState)));       // the reference's statics start at zero
    for (int b = 0; b < 2; b++) {
        FCK(cudaMalloc((void **)&fe->I[b], (size_t)nstreams * fe->cap * sizeof(float)));
        FCK(cudaMalloc((void **)&fe->Q[b], (size_t)nstreams * fe->cap * sizeof(float)));
        FCK(cudaMemset(fe->I[b], 0, (size_t)nstreams * fe->cap * sizeof(float)));
        FCK(cudaMemset(fe->Q[b], 0, (size_t)nstreams * fe->cap * sizeof(float)));
    }
    FCK(upload_fir());
done:
    return rc;
}

extern "C" wspr_frontend *wspr_frontend_create(int device, int nstreams, int slot_samples) {
    if (nstreams <= 0) {
        fe_fail(WSPR_ERR_ARG, "wspr_frontend_create: bad arguments");
        return nullptr;
    }
    wspr_frontend *fe = new wspr_frontend();
    if (fe_init(fe, device, nstreams, slot_samples) != WSPR_OK) {
        std::string keep = g_fe_err;
        wspr_frontend_destroy(fe);
        g_fe_err = keep;
        return nullptr;
    }
    return fe;
}

// room for `need` bytes per stream in both staging buffers (contents of the current one are kept)
static int fe_reserve(wspr_frontend *fe, size_t need) {
    int rc = WSPR_OK;
    const size_t stride = (need + 15) / 16 * 16 + 32;         // whole 16-byte vectors are read around every block
    if (stride <= fe->stage_stride) return WSPR_OK;
    uint8_t *nb[2] = {nullptr, nullptr};
    for (int b = 0; b < 2; b++) {
        FCK(cudaMalloc((void **)&nb[b], (size_t)fe->nstreams * stride));
        FCK(cudaMemsetAsync(nb[b], 0x80, (size_t)fe->nstreams * stride, fe->st));
    }
    if (fe->stage[fe->cur_stage] && fe->stage_stride)
        FCK(cudaMemcpy2DAsync(nb[fe->cur_stage], stride, fe->stage[fe->cur_stage], fe->stage_stride, fe->stage_stride,
                              fe->nstreams, cudaMemcpyDeviceToDevice, fe->st));
    FCK(cudaStreamSynchronize(fe->st));
    for (int b = 0; b < 2; b++) {
        if (fe->stage[b]) cudaFree(fe->stage[b]);
        fe->stage[b] = nb[b];
        nb[b] = nullptr;
    }
    fe->stage_stride = stride;
done:
    for (int b = 0; b < 2; b++)
        if (nb[b]) cudaFree(nb[b]);
    return rc;
}

extern "C" int wspr_frontend_push(wspr_frontend *fe, const uint8_t *raw, size_t stream_stride_bytes, uint32_t nbytes) {
    if (!fe || (nbytes && !raw) || (nbytes & 7u) || (fe->nstreams > 1 && stream_stride_bytes < nbytes))
        return fe_fail(WSPR_ERR_ARG, "wspr_frontend_push: bad arguments (the byte count must be a multiple of 8, rtlsdr_wsprd.c:171)");
    if (nbytes == 0) return 0;
    int rc = WSPR_OK;
    FCK(cudaSetDevice(fe->device));
    {
        // byte offset of the unfinished block inside the stream, modulo the 16-byte vector grid (keeps the mixer phase)
        const unsigned long long g0 = fe->blocks_done * (unsigned long long)BLOCK_BYTES;
        const int pad = (int)(g0 & 15ull);
        const size_t total = fe->carry + nbytes;
        if (fe_reserve(fe, (size_t)pad + total) != WSPR_OK) return WSPR_ERR_CUDA;
        uint8_t *cur = fe->stage[fe->cur_stage];
        FCK(cudaMemcpy2DAsync(cur + pad + fe->carry, fe->stage_stride, raw, stream_stride_bytes, nbytes, fe->nstreams,
                              cudaMemcpyHostToDevice, fe->st));
        const int nb = (int)(total / BLOCK_BYTES);
        if (nb > 0) {
            if ((size_t)nb * fe->nstreams > fe->moments_cap) {
                FCK(cudaStreamSynchronize(fe->st));
                if (fe->moments) cudaFree(fe->moments);
                fe->moments = nullptr;
                fe->moments_cap = (size_t)nb * fe->nstreams;
                FCK(cudaMalloc((void **)&fe->moments, fe->moments_cap * sizeof(uint4)));
            }
            k_block_moments<<<dim3((nb + MOM_WARPS - 1) / MOM_WARPS, fe->nstreams), MOM_WARPS * 32, 0, fe->st>>>(
                cur, fe->stage_stride, nb, fe->moments, pad);
            g_frontend_launches++;
            const int buf = fe->buffer_index;
            k_stream_advance<<<(2 * fe->nstreams + 63) / 64, 64, 0, fe->st>>>(fe->state, fe->moments, nb, fe->nstreams, fe->I[buf],
                                                                            fe->Q[buf], fe->cap, fe->iq_index[buf], fe->cap);
            g_frontend_launches++;
            FCK(cudaGetLastError());
            fe->iq_index[buf] = std::min(fe->cap, fe->iq_index[buf] + nb);
            // the unfinished tail moves to the other staging buffer, re-aligned to the vector grid
            const size_t left = total - (size_t)nb * BLOCK_BYTES;
            fe->blocks_done += (unsigned long long)nb;
            const int pad2 = (int)((fe->blocks_done * (unsigned long long)BLOCK_BYTES) & 15ull);
            uint8_t *nxt = fe->stage[fe->cur_stage ^ 1];
            if (left)
                FCK(cudaMemcpy2DAsync(nxt + pad2, fe->stage_stride, cur + pad + (size_t)nb * BLOCK_BYTES, fe->stage_stride, left,
                                      fe->nstreams, cudaMemcpyDeviceToDevice, fe->st));
            fe->cur_stage ^= 1;
            fe->carry = left;
        } else {
            fe->carry = total;
        }
        FCK(cudaStreamSynchronize(fe->st));       // `raw` belongs to the caller again (librtlsdr reuses its buffers)
        rc = nb;
    }
done:
    return rc;
}

// The main loop's slot switch (rtlsdr_wsprd.c:1181-1183): the other buffer becomes current and starts empty; returns the
// number of samples in the slot that just ended, whose tail is zeroed like decoder() does (rtlsdr_wsprd.c:285-288).
extern "C" int wspr_frontend_swap(wspr_frontend *fe) {
    if (!fe) return fe_fail(WSPR_ERR_ARG, "null front end");
    int rc = WSPR_OK;
    FCK(cudaSetDevice(fe->device));
    {
        const int prev = fe->buffer_index, n = fe->iq_index[prev];
        fe->buffer_index ^= 1;
        fe->iq_index[fe->buffer_index] = 0;
        if (n < fe->cap) {
            FCK(cudaMemset2DAsync(fe->I[prev] + n, (size_t)fe->cap * sizeof(float), 0, (size_t)(fe->cap - n) * sizeof(float),
                                  fe->nstreams, fe->st));
            FCK(cudaMemset2DAsync(fe->Q[prev] + n, (size_t)fe->cap * sizeof(float), 0, (size_t)(fe->cap - n) * sizeof(float),
                                  fe->nstreams, fe->st));
        }
        FCK(cudaStreamSynchronize(fe->st));
        rc = n;
    }
done:
    return rc;
}

// device pointers of the slot that ended at the last swap: [nstreams][stride] floats, zero padded
extern "C" int wspr_frontend_slot_device(wspr_frontend *fe, const float **dI, const float **dQ, int *stride) {
    if (!fe || !dI || !dQ || !stride) return fe_fail(WSPR_ERR_ARG, "wspr_frontend_slot_device");
    const int prev = fe->buffer_index ^ 1;
    *dI = fe->I[prev];
    *dQ = fe->Q[prev];
    *stride = fe->cap;
    return fe->iq_index[prev];
}

// copy of that slot in host memory ([nstreams][slot_samples] each); returns the samples it holds
extern "C" int wspr_frontend_read(wspr_frontend *fe, float *I, float *Q) {
    if (!fe || !I || !Q) return fe_fail(WSPR_ERR_ARG, "wspr_frontend_read");
    int rc = WSPR_OK;
    const int prev = fe->buffer_index ^ 1;
    FCK(cudaSetDevice(fe->device));
    FCK(cudaMemcpyAsync(I, fe->I[prev], (size_t)fe->nstreams * fe->cap * sizeof(float), cudaMemcpyDeviceToHost, fe->st));
    FCK(cudaMemcpyAsync(Q, fe->Q[prev], (size_t)fe->nstreams * fe->cap * sizeof(float), cudaMemcpyDeviceToHost, fe->st));
    FCK(cudaStreamSynchronize(fe->st));
    rc = fe->iq_index[prev];
done:
    return rc;
}

extern "C" int wspr_frontend_samples(wspr_frontend *fe) { return fe ? fe->iq_index[fe->buffer_index] : WSPR_ERR_ARG; }
