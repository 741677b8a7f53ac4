// Host model of the per-device Fano service (wspr_kernels.cu: fano_settle, FanoQueueFeed, the enqueue step and the pool
// protocol of k_fano_workers) for CPU-side checks of its logic under real concurrency: tests/test_fano_queue_host.py cuts
// fano_settle + FanoQueueFeed out of wspr_kernels.cu VERBATIM (fano_queue_extracted.inc), and this file runs them with host
// threads in place of warps -- every worker thread is a one-lane warp running the real decoder loop (wspr_fano.cuh) fed by
// the real FanoQueueFeed -- against several producer threads ("contexts") that park candidates on their captures the way
// k_jitter_soft + k_fano_enqueue do and start workers the way the host scheduler does.  The ring is tiny (it wraps every
// few candidates), the scratch records are reused as soon as a capture has been handed back, workers come and go.
// Checked for every parked candidate: it is handed back exactly once; the outcome in the job is the LOWEST gated attempt
// that decodes (wsprd.c:741-766), with its bytes and cycle count; every gated attempt below the winner ran to its end;
// nothing is left in the queue and no worker is left alive when the producers are done; no wait lasts longer than the
// watchdog allows (a stranded candidate -- work in the queue and nobody to do it -- would).
// x86 orders memory more strongly than a GPU does, so this checks the protocol's logic, not its fences.
#include <cuda_runtime.h>      // vector types only: nothing of the CUDA runtime is called
#include <sched.h>
#include <time.h>
#include <cstdlib>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>

// ---- device intrinsics the extracted code uses ----
// g_chaos > 0: every atomic and every fence gives the core away with probability 1 / g_chaos first, which shuffles the
// interleavings far beyond what the scheduler does by itself (the test also builds a copy of the extracted code with a
// yield between the reads of `tail` and `head0`, the one window no intrinsic sits in)
static int g_chaos = 0;
static inline void chaos() {
    if (g_chaos <= 0) return;
    static thread_local unsigned long long x = 0x9e3779b97f4a7c15ull ^ (unsigned long long)(uintptr_t)&x;
    x ^= x << 13;
    x ^= x >> 7;
    x ^= x << 17;
    if ((unsigned)(x >> 33) % (unsigned)g_chaos == 0u) sched_yield();
}
template <class T> static inline T atomicAdd(T *p, T v) { chaos(); return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicSub(T *p, T v) { chaos(); return __atomic_fetch_sub(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicAdd_system(T *p, T v) { chaos(); return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicCAS(unsigned *p, unsigned cmp, unsigned val) {
    chaos();
    __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;                                        // (the value found, like the device function)
}
static inline int atomicMin(int *p, int v) {
    chaos();
    int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); chaos(); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); chaos(); }
static inline void __nanosleep(unsigned) { sched_yield(); }
static inline bool __any_sync(unsigned, bool p) { return p; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __float2int_rz(float f) { return (int)f; }
static struct { unsigned x; } threadIdx = {0};
using std::max;
using std::min;

#include "wspr_kernels.cuh"
#include "wspr_fano.cuh"
#include "wspr_mettab.h"

namespace wspr {
#include "fano_queue_extracted.inc"
}

namespace {
using namespace wspr;
const short g_mettab[2][256] = WSPR_METTAB_INIT;

struct HostMem {                                   // one lane: level record n at lvl[n], node record pos at node[pos + 1]
    uint2 *lvl;
    unsigned *node;
    uint2 ldl(int n) const { return lvl[n]; }
    void stl(int n, unsigned a, unsigned b) const { lvl[n] = uint2{a, b}; }
    unsigned ldn(int pos) const { return node[pos + 1]; }
    void stn(int pos, unsigned v) const { node[pos + 1] = v; }
};
struct OneShot {                                   // stand-alone run of one vector: what the attempt must produce
    const unsigned char *sym;
    bool given;
    FanoResult res;
    const unsigned char *next(unsigned &stop) {
        stop = 0;
        if (given) return nullptr;
        given = true;
        return sym;
    }
    bool abandon() const { return false; }
    void finish(const FanoResult &r) { res = r; }
    void period(bool) {}
};

inline unsigned mix(unsigned a, unsigned b, unsigned c) {
    unsigned long long z = ((unsigned long long)a << 40) ^ ((unsigned long long)b << 20) ^ c;
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return (unsigned)((z ^ (z >> 31)) >> 16);
}

struct Sim {
    FanoQueue q;
    std::vector<FanoQueueEntry> ring;
    const unsigned char *vecs;                         // [nvec][162] soft symbols
    int nvec, delta, nsm;
    unsigned maxcycles, seed;
    std::vector<FanoResult> alone;                     // stand-alone result of every vector
    std::mutex reserve_mu;                             // (harness only: producers wait for room instead of overflowing the tiny ring)
    std::atomic<int> workers_alive{0}, workers_started{0}, failures{0};
    std::atomic<long long> last_progress_ms{0};
    int first_failure = 0;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();

    long long now_ms() const { return std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count(); }
    void progress() { last_progress_ms.store(now_ms()); }
    bool stalled(long long limit_ms) const { return now_ms() - last_progress_ms.load() > limit_ms; }
    void fail(int code) {
        if (failures.fetch_add(1) == 0) first_failure = code;
    }
};

// one worker warp (one lane): the pool protocol of k_fano_workers around the real decoder loop
void worker(Sim *s, unsigned smid) {
    FanoQueue *q = &s->q;
    int *active = &q->active;
    const volatile int *pool = &q->pool;
    FanoQueueFeed feed{q, nullptr, 0, false, 0u, 0u, 0u, 0u};
    const int per_sm = *(volatile int *)&q->per_sm;
    int mine = 1;
    if (per_sm > 0) {                                          // (this SM has its share of workers: leave)
        mine = atomicAdd(&q->sm_workers[smid], 1) < per_sm;
        if (!mine) atomicSub(&q->sm_workers[smid], 1);
    }
    if (mine) {
        mine = atomicAdd(active, 1) < *pool;
        if (!mine) {
            atomicSub(active, 1);
            if (per_sm > 0) atomicSub(&q->sm_workers[smid], 1);
        }
    }
    if (mine) {
        std::vector<uint2> lvl(FANO_LEVELS, uint2{0xdeadbeefu, 0xdeadbeefu});
        std::vector<unsigned> node(FANO_NODES + 1, 0xdeadbeefu);
        HostMem mem{lvl.data(), node.data()};
        while (mine) {
            fano_run<false>(feed, mem, &g_mettab[0][0], s->delta, s->maxcycles);
            mine = 0;
            atomicSub(active, 1);
            __threadfence();
            const unsigned tail = *(volatile unsigned *)&q->tail;
            const bool pending = (*(volatile unsigned *)&q->head0 != tail || *(volatile unsigned *)&q->head1 != tail);
            if (pending) {
                mine = atomicAdd(active, 1) < *pool;
                if (!mine) atomicSub(active, 1);
            }
        }
        // FANO_QUEUE_SLOW_EXIT_US (experiment, not used by the tests): the SM place is given up this long after the last look at
        // the queue.  With 300 us on a one- or two-SM model, workers started for the last candidates find "their" SM full of
        // workers that are on their way out and leave, and the candidates are stranded (the watchdog fires) -- the window the
        // comment above k_fano_workers describes; without the delay 3000 such runs strand nothing.
        static const long slow_exit_us = [] { const char *e = getenv("FANO_QUEUE_SLOW_EXIT_US"); return e ? atol(e) : 0L; }();
        if (slow_exit_us > 0) {
            struct timespec ts = {0, slow_exit_us * 1000L};
            nanosleep(&ts, nullptr);
        }
        if (per_sm > 0) atomicSub(&q->sm_workers[smid], 1);
        atomicAdd(&q->st_attempts, (unsigned long long)feed.attempts);
        atomicAdd(&q->st_dropped, (unsigned long long)feed.dropped);
        atomicAdd(&q->st_warps, 1ull);
    }
    s->workers_alive.fetch_sub(1);
}

void start_workers(Sim *s, int n, unsigned salt) {
    for (int i = 0; i < n; i++) {
        s->workers_alive.fetch_add(1);
        const int id = s->workers_started.fetch_add(1);
        std::thread(worker, s, mix(s->seed, salt, (unsigned)id) % (unsigned)s->nsm).detach();
    }
}

// one context: `ncap` captures, each with one scratch record; parks `ncand` candidates in bursts and checks every hand-back
// (quick_every: every quick_every-th candidate is a quick-mode one -- attempt 0 only -- parked on the context's second set of
// records, as wspr_ctx_decode does: a record is only ever armed with ONE number of attempts, see lease_scratch)
void producer(Sim *s, int ctx, int ncap, int ncand, int burst_max, int quick_every, long long watchdog_ms) {
    FanoQueue *q = &s->q;
    // (the records outlive this function on purpose: a worker may still look at one through an old ring entry)
    std::vector<ChainScratch> &scratch = *new std::vector<ChainScratch>(2 * (size_t)ncap);
    std::vector<int> which(ncap, 0);                   // 0: the capture's normal record is in use, 1: its quick-mode record
    std::vector<Job> &jobs = *new std::vector<Job>(ncap);
    std::vector<int> &phase = *new std::vector<int>(ncap, PH_READY);
    std::vector<int> cand_of(ncap, -1);
    int *stats = new int[8](), &host_done = *new int(0), parked = 0, returned = 0;
    std::vector<int> defer;
    while (returned < ncand) {
        // ---- hand-backs (the resolve step of a round) ----
        for (int c = 0; c < ncap; c++) {
            if (cand_of[c] < 0 || *(volatile int *)&phase[c] != PH_RESOLVE) continue;
            __threadfence();
            const ChainScratch &cs = scratch[c + which[c] * ncap];
            const Job &job = jobs[c];
            const unsigned id = (unsigned)cand_of[c];
            const int na = cs.nattempts;
            int winner = -1, last_run = -1;
            for (int y = 0; y < na && winner < 0; y++) {
                if (!cs.gate[y]) continue;
                const unsigned v = mix(s->seed + (unsigned)ctx, id, (unsigned)y) % (unsigned)s->nvec;
                last_run = y;
                if (s->alone[v].rc == 0) winner = y;
            }
            if (cs.done != na) s->fail(10);                                    // every attempt accounted for, once
            if ((winner >= 0) != (job.decoded != 0)) s->fail(11);
            if (winner >= 0) {
                const unsigned v = mix(s->seed + (unsigned)ctx, id, (unsigned)winner) % (unsigned)s->nvec;
                if (job.idt != winner) s->fail(12);                             // the lowest attempt that decodes
                if (memcmp(job.dec, s->alone[v].data, 10) != 0) s->fail(13);
                if (job.cycles != s->alone[v].cycles) s->fail(14);
                if (cs.best != winner) s->fail(15);
            } else if (last_run >= 0) {
                const unsigned v = mix(s->seed + (unsigned)ctx, id, (unsigned)last_run) % (unsigned)s->nvec;
                if (job.cycles != s->alone[v].cycles) s->fail(16);              // cycles of the last attempt the serial loop runs
            }
            for (int y = 0; y < (winner >= 0 ? winner : na); y++)               // every gated attempt before the winner ran to its end and failed
                if (cs.gate[y] && (cs.unfinished[y] || cs.ok[y])) s->fail(17);
            cand_of[c] = -1;
            phase[c] = PH_READY;
            returned++;
            s->progress();
        }
        // (host_done trails the phase flip by a few instructions of the settling lane: it is checked at the end)
        if (s->failures.load() || s->stalled(watchdog_ms)) {
            if (!s->failures.load()) s->fail(1);                                // nothing has moved for watchdog_ms: stranded work
            return;
        }
        // ---- park a burst of new candidates on free captures (k_jitter_soft + k_fano_enqueue + the worker launch) ----
        defer.clear();
        const int want = 1 + (int)(mix(s->seed, (unsigned)ctx, (unsigned)parked) % (unsigned)burst_max);
        for (int c = 0; c < ncap && (int)defer.size() < want && parked + (int)defer.size() < ncand; c++)
            if (cand_of[c] < 0) defer.push_back(c);
        if (defer.empty()) {
            sched_yield();
            continue;
        }
        int attempts = 0;
        for (int c : defer) {
            const unsigned id = (unsigned)(parked++);
            cand_of[c] = (int)id;
            const int na = (quick_every > 0 && id % (unsigned)quick_every == 3u % (unsigned)quick_every) ? 1 : NJIT;
            which[c] = na == 1;
            ChainScratch &cs = scratch[c + which[c] * ncap];
            Job &job = jobs[c];
            for (int y = 0; y < na; y++) {
                const unsigned v = mix(s->seed + (unsigned)ctx, id, (unsigned)y) % (unsigned)s->nvec;
                memcpy(cs.sym[y], s->vecs + 162 * (size_t)v, 162);
                cs.gate[y] = (mix(s->seed ^ 0x55u, id * 64u + (unsigned)y, (unsigned)ctx) % 8u) != 0u;
                cs.ok[y] = cs.unfinished[y] = 0;
            }
            job.decoded = 0;
            job.idt = -1;
            job.cycles = 0;
            memset(job.dec, 0, sizeof job.dec);
            cs.best = NJIT;
            cs.done = 0;
            cs.nattempts = na;
            cs.job = &job;
            cs.phase = &phase[c];
            cs.stats = stats;
            cs.host_done = &host_done;
            phase[c] = PH_WAIT;
            attempts += na;
        }
        const int n = (int)defer.size();
        unsigned base;
        for (;;) {                                                             // wait for room in the (tiny) ring, then reserve
            {
                std::lock_guard<std::mutex> g(s->reserve_mu);
                const unsigned h0 = *(volatile unsigned *)&q->head0, h1 = *(volatile unsigned *)&q->head1;
                const unsigned oldest = (int)(h0 - h1) < 0 ? h0 : h1;
                // (half the ring stays free: a worker that has popped an entry and is preempted before it reads it must not
                // find the slot rewritten -- at the library's 262 144 entries that takes a quarter of a million candidates)
                if (*(volatile unsigned *)&q->tail + (unsigned)n - oldest <= (q->mask + 1u) / 2u) {
                    base = atomicAdd(&q->tail, (unsigned)n);
                    break;
                }
            }
            if (s->failures.load() || s->stalled(watchdog_ms)) {
                if (!s->failures.load()) s->fail(2);                            // the ring never drains
                return;
            }
            // somebody has to be alive to move the cursors on: the real scheduler launches workers with every enqueue
            if (s->workers_alive.load() == 0) start_workers(s, 1, 0x777u);
            sched_yield();
        }
        for (int i = 0; i < n; i++) {                                           // k_fano_enqueue
            ChainScratch *cs = &scratch[defer[i] + which[defer[i]] * ncap];
            *(volatile int *)&cs->next = 1;
            FanoQueueEntry &x = q->ring[(base + (unsigned)i) & q->mask];
            *(ChainScratch *volatile *)&x.cs = cs;
            __threadfence();
            *(volatile unsigned *)&x.seq = base + (unsigned)i + 1u;
        }
        start_workers(s, std::min(q->pool, attempts), (unsigned)parked);        // one lane per worker here
    }
    for (int spin = 0; *(volatile int *)&host_done != ncand && spin < 1000000; spin++) sched_yield();
    if (stats[0] + stats[1] + stats[2] != ncand) s->fail(19);
    if (*(volatile int *)&host_done != ncand) s->fail(20);
}
}  // namespace

// returns 0, or the code of the first failed check (1 / 2: the watchdog fired); out[0..4] = attempts run, attempts
// skipped or abandoned, worker threads started, worker threads that found a place in the pool, vectors that decode alone
extern "C" int fano_queue_sim(const unsigned char *vecs, int nvec, int nctx, int ncap, int ncand, int burst_max, int ring_log2,
                              int pool, int per_sm, int nsm, int delta, unsigned maxcycles, unsigned seed, int watchdog_ms,
                              int quick_every, int chaos_one_in, long long *out) {
    g_chaos = chaos_one_in;
    Sim *sp = new Sim;                                 // (left behind on purpose if a worker thread never ends)
    Sim &s = *sp;
    memset(&s.q, 0, sizeof s.q);
    s.ring.assign((size_t)1 << ring_log2, FanoQueueEntry{nullptr, 0u, 0u});
    s.q.ring = s.ring.data();
    s.q.mask = (1u << ring_log2) - 1u;
    s.q.pool = pool;
    s.q.per_sm = per_sm;
    s.q.ovf_backlog = 1 << 30;
    s.vecs = vecs;
    s.nvec = nvec;
    s.delta = delta;
    s.maxcycles = maxcycles;
    s.seed = seed;
    s.nsm = std::max(1, std::min(nsm, 256));
    s.alone.resize(nvec);
    {
        std::vector<uint2> lvl(FANO_LEVELS, uint2{0u, 0u});
        std::vector<unsigned> node(FANO_NODES + 1, 0u);
        HostMem mem{lvl.data(), node.data()};
        for (int v = 0; v < nvec; v++) {
            OneShot one{vecs + 162 * (size_t)v, false, FanoResult{}};
            fano_run<false>(one, mem, &g_mettab[0][0], delta, maxcycles);
            s.alone[v] = one.res;
        }
    }
    s.progress();
    std::vector<std::thread> producers;
    for (int c = 0; c < nctx; c++) producers.emplace_back(producer, &s, c, ncap, ncand, burst_max, quick_every, (long long)watchdog_ms);
    for (auto &t : producers) t.join();
    // the workers leave by themselves once the queue is empty
    const long long t_end = s.now_ms() + watchdog_ms;
    while (s.workers_alive.load() > 0 && s.now_ms() < t_end) std::this_thread::sleep_for(std::chrono::milliseconds(1));
    const bool stuck = s.workers_alive.load() > 0;
    if (stuck) s.fail(3);                                                       // a worker never leaves
    if (!s.failures.load()) {
        if (s.q.head0 != s.q.tail || s.q.head1 != s.q.tail) s.fail(30);          // nothing left in the queue
        if (s.q.active != 0) s.fail(31);
        for (int i = 0; i < 256; i++)
            if (s.q.sm_workers[i] != 0) s.fail(32);
        if (s.q.tail != (unsigned)(nctx * ncand)) s.fail(33);
    }
    if (out) {
        out[0] = (long long)s.q.st_attempts;
        out[1] = (long long)s.q.st_dropped;
        out[2] = s.workers_started.load();
        out[3] = (long long)s.q.st_warps;
        int ok = 0;
        for (int v = 0; v < nvec; v++) ok += s.alone[v].rc == 0;
        out[4] = ok;
    }
    const int rc = s.failures.load() ? s.first_failure : 0;
    if (!stuck) delete sp;
    return rc;
}
