// Execution model of the host emulation (see cuda_runtime.h in this directory): a kernel launch runs its CTAs one after the
// other; the threads of a CTA are cooperative fibers (ucontext) scheduled round-robin on the calling host thread, which give
// the processor up only where CUDA threads can wait for each other -- __syncthreads, the warp collectives, __nanosleep.
// Launches from several host threads are serialised by one lock, so "device" code never runs concurrently with itself and
// its atomics can be plain read-modify-writes.
#include "cuda_runtime.h"

#include <stdio.h>
#include <sys/mman.h>
#include <ucontext.h>

#include <atomic>
#include <mutex>
#include <vector>

namespace emu {
char g_anchor;
ThreadState *g_cur = nullptr;

namespace {
constexpr size_t STACK_BYTES = 256 * 1024;
constexpr size_t DYN_SMEM_BYTES = 256 * 1024;
alignas(128) char g_dyn_smem[DYN_SMEM_BYTES];

struct Warp {
    int alive = 0, arrived = 0;
    unsigned gen = 0;
    unsigned long long slot[32];
    bool active[32];
};
struct Fiber {
    ucontext_t ctx;
    ThreadState ts;
    bool done = false;
};
struct Block {
    int nthreads = 0, alive = 0, bar_arrived = 0;
    unsigned bar_gen = 0;
    std::vector<Warp> warps;
};

std::mutex g_launch_mu;
std::atomic<unsigned long long> g_launches{0};
std::vector<char *> g_stacks;
std::vector<Fiber> g_fibers;
Block g_block;
ucontext_t g_main;
const std::function<void()> *g_body = nullptr;
Fiber *g_fiber = nullptr;
const int g_order = [] {
    const char *e = getenv("EMU_ORDER");
    return !e ? 0 : (e[0] == 'r' && e[1] == 'e') ? 1 : (e[0] == 'r' && e[1] == 'a') ? 2 : 0;
}();
unsigned long long g_rng = 0x853c49e6748fea9bull;

char *stack_for(size_t i) {
    while (g_stacks.size() <= i) {
        void *p = mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) {
            fprintf(stderr, "cuda_emu: cannot map a fiber stack\n");
            abort();
        }
        g_stacks.push_back((char *)p);
    }
    return g_stacks[i];
}

void release_barriers_after_exit(Fiber *f) {
    Block &b = g_block;
    Warp &w = b.warps[f->ts.warp];
    w.active[f->ts.lane] = false;
    w.alive--;
    b.alive--;
    if (w.alive > 0 && w.arrived == w.alive) {              // the lanes that are waiting were only waiting for this one
        w.arrived = 0;
        w.gen++;
    }
    if (b.alive > 0 && b.bar_arrived == b.alive) {
        b.bar_arrived = 0;
        b.bar_gen++;
    }
}

void trampoline() {
    Fiber *f = g_fiber;
    (*g_body)();
    f->done = true;
    release_barriers_after_exit(f);
    // (returning resumes uc_link = the scheduler)
}

void warp_barrier() {
    Warp &w = g_block.warps[g_cur->warp];
    const unsigned gen = w.gen;
    if (++w.arrived == w.alive) {
        w.arrived = 0;
        w.gen++;
    } else {
        while (w.gen == gen) yield();
    }
}
}  // namespace

void yield() {
    Fiber *f = g_fiber;
    swapcontext(&f->ctx, &g_main);
}

void syncthreads() {
    Block &b = g_block;
    const unsigned gen = b.bar_gen;
    if (++b.bar_arrived == b.alive) {
        b.bar_arrived = 0;
        b.bar_gen++;
    } else {
        while (b.bar_gen == gen) yield();
    }
}

void syncwarp() { warp_barrier(); }

unsigned long long warp_gather(unsigned long long mine, int src_lane) {
    Warp &w = g_block.warps[g_cur->warp];
    w.slot[g_cur->lane] = mine;
    warp_barrier();                                           // everybody has given
    const unsigned long long r = w.active[src_lane & 31] ? w.slot[src_lane & 31] : 0ull;
    warp_barrier();                                           // everybody has taken: the slots may be written again
    return r;
}

unsigned warp_ballot(bool pred) {
    Warp &w = g_block.warps[g_cur->warp];
    w.slot[g_cur->lane] = pred ? 1ull : 0ull;
    warp_barrier();
    unsigned m = 0;
    for (int l = 0; l < 32; l++)
        if (w.active[l] && w.slot[l]) m |= 1u << l;
    warp_barrier();
    return m;
}

void *dynamic_smem() { return g_dyn_smem; }

void *device_alloc(size_t n) {
    // uninitialised device memory is NOT zero: poison it so that code which relies on a fresh allocation being zero shows
    void *p = malloc(n ? n : 1);
    if (p) memset(p, 0xA5, n);
    return p;
}

unsigned long long kernel_launches() { return g_launches.load(); }

void unsupported_asm(const char *what) {
    fprintf(stderr, "cuda_emu: inline PTX without a host form was executed: %s\n", what);
    abort();
}

// EMU_DEFER_WORKERS=n: launches of k_fano_workers do not run at once but when the host has polled cudaStreamQuery n times since
// the first of them (wait_parked does so every 1024 spins) or synchronises the device -- the worker pool then lags behind the
// rounds as it does on a GPU: captures stay parked (PH_WAIT) across rounds, the host lingers and waits for hand-backs, and
// with n large enough it reaches its quarter-second relaunch.  Results must not depend on any of it.
struct Deferred {
    dim3 grid, block;
    size_t smem;
    std::function<void()> body;
};
std::vector<Deferred> g_deferred;
std::mutex g_deferred_mu;
int g_queries_since_deferral = 0;
const int g_defer_workers = [] {
    const char *e = getenv("EMU_DEFER_WORKERS");
    return e ? atoi(e) : 0;
}();

static void run_now(dim3 grid, dim3 block, size_t dynamic_smem_bytes, const std::function<void()> &body);

void run_deferred(bool from_query) {
    std::vector<Deferred> todo;
    {
        std::lock_guard<std::mutex> lock(g_deferred_mu);
        if (g_deferred.empty()) return;
        if (from_query && ++g_queries_since_deferral < g_defer_workers) return;
        todo.swap(g_deferred);
        g_queries_since_deferral = 0;
    }
    for (Deferred &d : todo) run_now(d.grid, d.block, d.smem, d.body);
}

void launch(const char *kernel, dim3 grid, dim3 block, size_t dynamic_smem_bytes, const std::function<void()> &body) {
    if (g_defer_workers > 0 && strcmp(kernel, "k_fano_workers") == 0) {
        std::lock_guard<std::mutex> lock(g_deferred_mu);
        g_deferred.push_back(Deferred{grid, block, dynamic_smem_bytes, body});
        g_launches++;
        return;
    }
    g_launches++;
    run_now(grid, block, dynamic_smem_bytes, body);
}

static void run_now(dim3 grid, dim3 block, size_t dynamic_smem_bytes, const std::function<void()> &body) {
    std::lock_guard<std::mutex> lock(g_launch_mu);
    const int nthreads = (int)(block.x * block.y * block.z);
    if (nthreads <= 0 || nthreads > 1024 || dynamic_smem_bytes > DYN_SMEM_BYTES) {
        fprintf(stderr, "cuda_emu: launch configuration out of range (%d threads, %zu bytes of dynamic shared memory)\n", nthreads,
                dynamic_smem_bytes);
        abort();
    }
    if ((int)g_fibers.size() < nthreads) g_fibers.resize(nthreads);
    g_body = &body;
    // (EMU_ORDER=reverse | random also applies to the CTAs of a grid: lists that kernels build with atomics come out in
    // another order, which no result may depend on)
    const unsigned long long nblocks = (unsigned long long)grid.x * grid.y * grid.z;
    std::vector<unsigned long long> order(nblocks);
    for (unsigned long long i = 0; i < nblocks; i++) order[i] = g_order == 1 ? nblocks - 1 - i : i;
    if (g_order == 2)
        for (unsigned long long i = nblocks; i > 1; i--) {
            g_rng = g_rng * 6364136223846793005ull + 1442695040888963407ull;
            std::swap(order[i - 1], order[(g_rng >> 33) % i]);
        }
    for (unsigned long long idx : order) {
                const unsigned bx = (unsigned)(idx % grid.x), by = (unsigned)((idx / grid.x) % grid.y), bz = (unsigned)(idx / ((unsigned long long)grid.x * grid.y));
                Block &b = g_block;
                b.nthreads = b.alive = nthreads;
                b.bar_arrived = 0;
                b.bar_gen = 0;
                b.warps.assign((nthreads + 31) / 32, Warp());
                memset(g_dyn_smem, 0xA5, dynamic_smem_bytes);
                for (int t = 0; t < nthreads; t++) {
                    Fiber &f = g_fibers[t];
                    f.done = false;
                    f.ts.tid = uint3{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
                    f.ts.bid = uint3{bx, by, bz};
                    f.ts.bdim = block;
                    f.ts.gdim = grid;
                    f.ts.lane = t & 31;
                    f.ts.warp = t >> 5;
                    Warp &w = b.warps[t >> 5];
                    w.alive++;
                    w.active[t & 31] = true;
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = stack_for(t);
                    f.ctx.uc_stack.ss_size = STACK_BYTES;
                    f.ctx.uc_link = &g_main;
                    makecontext(&f.ctx, trampoline, 0);
                }
                for (Warp &w : b.warps)
                    for (int l = w.alive; l < 32; l++) w.active[l] = false;
                // EMU_ORDER=reverse | random: the fibers of a CTA are resumed in another order than 0, 1, 2, ... -- results that
                // change with it point at a missing barrier (threads reading what other threads wrote in the same phase)
                while (b.alive > 0) {
                    for (int k = 0; k < nthreads; k++) {
                        int t = k;
                        if (g_order == 1) t = nthreads - 1 - k;
                        else if (g_order == 2) {
                            g_rng = g_rng * 6364136223846793005ull + 1442695040888963407ull;
                            t = (int)((g_rng >> 33) % (unsigned)nthreads);
                        }
                        Fiber &f = g_fibers[t];
                        if (f.done) continue;
                        g_fiber = &f;
                        g_cur = &f.ts;
                        swapcontext(&g_main, &f.ctx);
                    }
                }
            }
    g_cur = nullptr;
    g_fiber = nullptr;
    g_body = nullptr;
}
}  // namespace emu
