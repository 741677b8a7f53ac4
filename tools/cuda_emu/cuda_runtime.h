// A stand-in for <cuda_runtime.h> that lets the library's .cu sources be compiled by g++ and RUN ON THE HOST, one CUDA thread
// = one cooperative fiber (tools/cuda_emu/emu_runtime.cpp): test infrastructure, see tools/cuda_emu/README.md.  It implements
// what the decode path and the front end use of the execution model (grids, CTAs, __syncthreads, warp collectives over the 32
// fibers of a warp, static and dynamic shared memory, atomics) and of the runtime API (memory, copies, streams and events --
// everything executes synchronously at the call).  Nothing here is fast and nothing is meant to be complete.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <algorithm>
#include <functional>

#define WSPR_CUDA_EMULATION 1
#ifndef __CUDACC__
#define __CUDACC__ 1                      // the sources show their device side (but never __CUDA_ARCH__: host fall-backs apply)
#endif
#define __host__
#define __device__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__
#define __align__(n) __attribute__((aligned(n)))

// ---- vector types -------------------------------------------------------------------------------------------------
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }

// ---- execution model ----------------------------------------------------------------------------------------------
namespace emu {
struct ThreadState {
    uint3 tid, bid;
    dim3 bdim, gdim;
    int lane, warp;
};
extern ThreadState *g_cur;                                   // the fiber that is running
void launch(const char *kernel, dim3 grid, dim3 block, size_t dynamic_smem, const std::function<void()> &body);
void run_deferred(bool from_query);                          // EMU_DEFER_WORKERS: the worker-pool launches that were put off
void syncthreads();
void syncwarp();
void yield();
void *dynamic_smem();
unsigned long long warp_gather(unsigned long long mine, int src_lane);            // every lane gives, every lane takes one
unsigned warp_ballot(bool pred);
unsigned long long kernel_launches();
void unsupported_asm(const char *what);
static inline dim3 d3(dim3 v) { return v; }
static inline dim3 d3(long long v) { return dim3((unsigned)v); }
// shared-window addresses: offsets from a symbol inside this library (all "shared" memory is static storage of it)
extern char g_anchor;
static inline unsigned to_shared(const void *p) { return (unsigned)(uintptr_t)((const char *)p - &g_anchor); }
static inline char *from_shared(unsigned a) { return &g_anchor + (intptr_t)(int32_t)a; }
static inline unsigned long long pack2(float lo, float hi) {
    unsigned a, b;
    memcpy(&a, &lo, 4);
    memcpy(&b, &hi, 4);
    return (unsigned long long)a | ((unsigned long long)b << 32);
}
static inline unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {   // fma.rn.f32x2
    float al, ah, bl, bh, cl, ch;
    unsigned u;
    u = (unsigned)a; memcpy(&al, &u, 4); u = (unsigned)(a >> 32); memcpy(&ah, &u, 4);
    u = (unsigned)b; memcpy(&bl, &u, 4); u = (unsigned)(b >> 32); memcpy(&bh, &u, 4);
    u = (unsigned)c; memcpy(&cl, &u, 4); u = (unsigned)(c >> 32); memcpy(&ch, &u, 4);
    return pack2(fmaf(al, bl, cl), fmaf(ah, bh, ch));                              // one rounding each, like the instruction
}
static inline int dp4a_u32_s32(unsigned a, int b, int c) {
    for (int k = 0; k < 4; k++) c += (int)((a >> (8 * k)) & 255u) * (int)(int8_t)((unsigned)b >> (8 * k));
    return c;
}
}  // namespace emu
#define threadIdx (emu::g_cur->tid)
#define blockIdx (emu::g_cur->bid)
#define blockDim (emu::g_cur->bdim)
#define gridDim (emu::g_cur->gdim)
static inline void __syncthreads() { emu::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::syncwarp(); }
static inline void __threadfence() {}
static inline void __threadfence_system() {}
static inline void __threadfence_block() {}
static inline void __nanosleep(unsigned) { emu::yield(); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) {
    unsigned long long b = 0;
    memcpy(&b, &v, sizeof(T));
    b = emu::warp_gather(b, src & 31);
    T r;
    memcpy(&r, &b, sizeof(T));
    return r;
}
template <class T> static inline T __shfl_down_sync(unsigned m, T v, unsigned delta, int = 32) {
    const int lane = emu::g_cur->lane;
    return __shfl_sync(m, v, lane + (int)delta < 32 ? lane + (int)delta : lane);
}
template <class T> static inline T __shfl_xor_sync(unsigned m, T v, int x, int = 32) { return __shfl_sync(m, v, emu::g_cur->lane ^ x); }
static inline unsigned __ballot_sync(unsigned, int pred) { return emu::warp_ballot(pred != 0); }
static inline int __any_sync(unsigned, int pred) { return emu::warp_ballot(pred != 0) != 0u; }
static inline int __all_sync(unsigned m, int pred) { return emu::warp_ballot(pred == 0) == 0u; }
static inline unsigned __reduce_add_sync(unsigned m, unsigned v) {
    unsigned s = 0;
    for (int l = 0; l < 32; l++) {                           // (lanes that have left give nothing: warp_gather returns 0 for them)
        s += (unsigned)emu::warp_gather(v, l);
    }
    return s;
}
static inline int __reduce_add_sync(unsigned m, int v) { return (int)__reduce_add_sync(m, (unsigned)v); }
static inline long long clock64() { return 0; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned __brev(unsigned v) {
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
    return __builtin_bswap32(v);
}
static inline float __fsqrt_rn(float v) { return sqrtf(v); }
static inline int __float2int_rz(float f) { return (int)f; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline size_t __cvta_generic_to_shared(const void *p) { return emu::to_shared(p); }
using std::max;
using std::min;
static inline unsigned max(unsigned a, int b) { return a > (unsigned)b ? a : (unsigned)b; }
static inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }

// atomics: one fiber runs at a time (launches are serialised), so these are plain read-modify-writes
template <class T, class U> static inline T atomicAdd(T *p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U> static inline T atomicSub(T *p, U v) { T o = *p; *p = (T)(o - (T)v); return o; }
template <class T, class U> static inline T atomicAdd_system(T *p, U v) { return atomicAdd(p, v); }
template <class T, class U> static inline T atomicMin(T *p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicMax(T *p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U, class V> static inline T atomicCAS(T *p, U cmp, V val) { T o = *p; if (o == (T)cmp) *p = (T)val; return o; }
template <class T, class U> static inline T atomicExch(T *p, U v) { T o = *p; *p = (T)v; return o; }
template <class T, class U> static inline T atomicOr(T *p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }

// ---- runtime API (everything is synchronous) ----------------------------------------------------------------------
enum cudaError_t { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600, cudaErrorUnknown = 999 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1, cudaDriverEntryPointVersionNotSufficent = 2 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaEventBlockingSync = 1, cudaEnableDefault = 0, cudaHostAllocMapped = 2, cudaHostAllocDefault = 0 };
struct emuStream_ { int id; };
struct emuEvent_ { int id; };
typedef emuStream_ *cudaStream_t;
typedef emuEvent_ *cudaEvent_t;
struct cudaDeviceProp {
    char name[256];
    int multiProcessorCount, major, minor;
    size_t totalGlobalMem, sharedMemPerMultiprocessor, sharedMemPerBlockOptin;
};
static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { emu::run_deferred(false); return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    memset(p, 0, sizeof *p);
    strcpy(p->name, "host emulation");
    p->multiProcessorCount = 4;                                // (a small device: the Fano pool is 2 worker warps per SM)
    p->major = 10;
    p->totalGlobalMem = (size_t)16 << 30;
    p->sharedMemPerMultiprocessor = 228 * 1024;
    p->sharedMemPerBlockOptin = 227 * 1024;
    return cudaSuccess;
}
namespace emu { void *device_alloc(size_t n); }
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)emu::device_alloc(n); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t n) { *p = (T *)calloc(1, n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> static inline cudaError_t cudaHostAlloc(T **p, size_t n, unsigned) { return cudaMallocHost(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
template <class T> static inline cudaError_t cudaHostGetDevicePointer(T **d, void *h, unsigned) { *d = (T *)h; return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = nullptr) {
    for (size_t r = 0; r < h; r++) memmove((char *)d + r * dp, (const char *)s + r * sp, w);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy2D(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind k) { return cudaMemcpy2DAsync(d, dp, s, sp, w, h, k); }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset2DAsync(void *d, size_t pitch, int v, size_t w, size_t h, cudaStream_t = nullptr) {
    for (size_t r = 0; r < h; r++) memset((char *)d + r * pitch, v, w);
    return cudaSuccess;
}
#define cudaMemcpyToSymbol(sym, src, ...) emu_copy_to_symbol((void *)&(sym), sizeof(sym), src, __VA_ARGS__)
#define cudaMemcpyFromSymbol(dst, sym, ...) emu_copy_from_symbol(dst, (const void *)&(sym), sizeof(sym), __VA_ARGS__)
static inline cudaError_t emu_copy_to_symbol(void *sym, size_t cap, const void *src, size_t n, size_t off = 0, cudaMemcpyKind = cudaMemcpyHostToDevice) {
    if (off + n > cap) return cudaErrorInvalidValue;
    memcpy((char *)sym + off, src, n);
    return cudaSuccess;
}
static inline cudaError_t emu_copy_from_symbol(void *dst, const void *sym, size_t cap, size_t n, size_t off = 0, cudaMemcpyKind = cudaMemcpyDeviceToHost) {
    if (off + n > cap) return cudaErrorInvalidValue;
    memcpy(dst, (const char *)sym + off, n);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = new emuStream_{0}; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { return cudaStreamCreate(s); }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { return cudaStreamCreate(s); }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamQuery(cudaStream_t) { emu::run_deferred(true); return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emuEvent_{0}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaGetDriverEntryPoint(const char *, void **fn, unsigned long long, cudaDriverEntryPointQueryResult *q = nullptr) {
    *fn = nullptr;                                             // no driver: no green contexts, the pool shares the "SMs"
    if (q) *q = cudaDriverEntryPointSymbolNotFound;
    return cudaSuccess;
}
