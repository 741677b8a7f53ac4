"""The per-device Fano service -- the queue of parked candidates, its two cursors, the claim counters, the hand-back and the
come-and-go protocol of the worker pool (rtlsdr_wsprd_b200/csrc/wspr_kernels.cu) -- run on the HOST under real concurrency.

fano_settle and FanoQueueFeed are cut out of wspr_kernels.cu verbatim and compiled with g++ (tools/fano_queue_host_check.cpp
maps the device atomics to the compiler's); host threads stand in for worker warps (each one lane wide, running the real
decoder loop of wspr_fano.cuh), producer threads for the contexts of a process.  The ring is small, so it wraps many times;
records are re-armed as soon as their capture is handed back; workers leave when they find nothing and are started with
every enqueue.  For every parked candidate the outcome must be the one the reference's sequential jitter loop
(wsprd.c:741-766) produces -- the lowest gated attempt that decodes, its bytes, its cycle count -- and it must be handed
back exactly once; at the end the queue is empty and no worker is left alive.  A watchdog turns a stranded candidate or a
worker that never leaves into a failure instead of a hang.

This model found two defects of the first version of the protocol, both fixed where the comments in wspr_kernels.cu /
wspr_decode.cu say so: a lane could pop an attempt-0 entry that did not exist yet (tail read before head0) and wait for it
with its whole warp -- a device that never goes idle if no further candidate comes --, and a record last armed for a
quick-mode candidate (1 attempt) looked claimable through an old ring entry while it was being re-armed for 43."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import helpers as H
from test_fano_host import vectors

CSRC = os.path.join(H.ROOT, "rtlsdr_wsprd_b200", "csrc")


def _cuda_include():
    for d in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if d and os.path.exists(os.path.join(d, "include", "cuda_runtime.h")):
            return os.path.join(d, "include")
    return None


def _build(tmp, name, extracted, inc):
    d = tmp / name
    d.mkdir()
    (d / "fano_queue_extracted.inc").write_text(extracted)
    out = str(d / "libfano_queue_host.so")
    cmd = ["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", "-I" + inc, "-I" + CSRC, "-I" + str(d), "-o", out,
           os.path.join(H.ROOT, "tools", "fano_queue_host_check.cpp")]
    try:
        subprocess.run(cmd, check=True, capture_output=True, text=True)
    except FileNotFoundError:
        pytest.skip("g++ not available")
    except subprocess.CalledProcessError as e:
        pytest.fail("host build of the queue code failed:\n" + e.stderr[-3000:])
    lib = C.CDLL(out)
    lib.fano_queue_sim.argtypes = [C.c_void_p] + [C.c_int] * 10 + [C.c_uint, C.c_uint] + [C.c_int] * 3 + [C.c_void_p]
    lib.fano_queue_sim.restype = C.c_int
    return lib


@pytest.fixture(scope="module")
def queue_sims(tmp_path_factory):
    """(the code as it is, the same code with a yield between the reads of `tail` and `head0` in FanoQueueFeed::next)"""
    inc = _cuda_include()
    if inc is None:
        pytest.skip("CUDA headers not available (vector types of wspr_kernels.cuh)")
    tmp = tmp_path_factory.mktemp("fano_queue")
    src = open(os.path.join(CSRC, "wspr_kernels.cu")).read()
    a = src.index("__device__ void fano_settle(ChainScratch *cs, int count) {")
    b = src.index("// The pool: at most q->pool worker warps are alive at any time.")
    code = src[a:b]
    assert "struct FanoQueueFeed" in code
    read_head0 = "            unsigned h = *(volatile unsigned *)&q->head0;\n"
    assert code.count(read_head0) == 1
    return _build(tmp, "verbatim", code, inc), _build(tmp, "widened", code.replace(read_head0, "            sched_yield();\n" + read_head0), inc)


def simulate(lib, vecs, nctx, ncap, ncand, burst, ring_log2, pool, per_sm, nsm, maxcycles, seed, quick_every=0, chaos=0):
    out = np.zeros(8, np.int64)
    rc = lib.fano_queue_sim(vecs.ctypes.data, len(vecs), nctx, ncap, ncand, burst, ring_log2, pool, per_sm, nsm, 60, maxcycles,
                            seed, 30000, quick_every, chaos, out.ctypes.data)
    return rc, dict(attempts_run=int(out[0]), attempts_dropped=int(out[1]), workers_started=int(out[2]),
                    workers_in_pool=int(out[3]), vectors_that_decode=int(out[4]))


@pytest.fixture(scope="module")
def corpora():
    easy = np.ascontiguousarray(np.stack(vectors(64, 7)))                  # clean to hopeless: about a third decode
    hard = [v for i, v in enumerate(vectors(256, 11)) if i % 8 >= 5]       # the noisy end only: every attempt runs to its time-out
    return easy, np.ascontiguousarray(np.stack(hard[:64]))


# (contexts, captures per context, candidates per context, largest burst, log2 ring entries, pool, per SM, SMs, cycles per bit)
SHAPES = [(3, 24, 1500, 8, 10, 6, 2, 4, 30),
          (1, 8, 1500, 8, 8, 2, 0, 1, 30),
          (4, 40, 1500, 16, 10, 16, 2, 8, 100),
          (9, 16, 700, 4, 10, 8, 2, 8, 50)]


@pytest.mark.parametrize("shape", SHAPES)
def test_every_candidate_settles_once_with_the_sequential_outcome(queue_sims, corpora, shape):
    nctx, ncap, ncand, burst, ring_log2, pool, per_sm, nsm, maxc = shape
    for seed, (lib, chaos) in enumerate(((queue_sims[0], 0), (queue_sims[1], 3))):
        rc, st = simulate(lib, corpora[0], nctx, ncap, ncand, burst, ring_log2, pool, per_sm, nsm, maxc, seed, chaos=chaos)
        assert rc == 0, (rc, st, chaos)
        assert nctx * ncand > (1 << ring_log2), "the ring must wrap"
        assert 0 < st["vectors_that_decode"] < 64 and st["attempts_run"] > 0 and st["attempts_dropped"] > 0


def test_hopeless_candidates_run_all_their_attempts(queue_sims, corpora):
    rc, st = simulate(queue_sims[0], corpora[1], 3, 24, 500, 8, 9, 6, 2, 4, 30, 1)
    assert rc == 0, (rc, st)
    assert st["vectors_that_decode"] == 0
    # nothing ever decodes, so nothing is abandoned: every gated attempt runs to its end (7 in 8 are gated)
    assert st["attempts_run"] + st["attempts_dropped"] == 3 * 500 * 43


def test_quick_mode_candidates_on_their_own_records(queue_sims, corpora):
    """Every fifth candidate is a quick-mode one (attempt 0 only).  wspr_ctx_decode parks those on a second set of records:
    a record is only ever armed with one number of attempts, so it never looks claimable while it is idle."""
    for vecs in corpora:
        for seed, (lib, chaos) in enumerate(((queue_sims[0], 0), (queue_sims[1], 3))):
            rc, st = simulate(lib, vecs, 3, 24, 1200, 8, 10, 6, 2, 4, 30, seed, quick_every=5, chaos=chaos)
            assert rc == 0, (rc, st, chaos)
