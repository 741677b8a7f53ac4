"""The GPU parity tests on the CPU: the library's own .cu / .cuh sources are transpiled (tools/cuda_emu/transpile.py rewrites
kernel launches, inline PTX and extern __shared__; nothing else), compiled by g++ against stand-in CUDA headers and run under
a SIMT emulation (tools/cuda_emu: one cooperative fiber per CUDA thread, real __syncthreads / warp-collective semantics,
poisoned fresh memory).  A subset of tests/test_gpu_parity.py is then run against that build through WSPR_B200_LIB, in a
subprocess (this process may already have loaded the CUDA build).  It pins the logic and the exact-order arithmetic of the
kernel and scheduler sources without a GPU; it is not a fallback of the product, which refuses to run without CUDA."""
import os
import subprocess
import sys

import pytest

import helpers as H

SUBSET = ("reference_fixture or drifting_and_edge or degenerate or short_capture "
          "or persistent_hashtable_option or sync_and_demodulate_abi or subtract_signal2_abi "
          "or subtract_signal_abi or stage_spectrogram or frontend_against_reference_golden or frontend_ragged "
          "or streaming_frontend or one_shot_batch_entry or quick_and_normal or candidate_loop_breaks")


@pytest.fixture(scope="module")
def emulated_library(tmp_path_factory):
    sys.path.insert(0, os.path.join(H.ROOT, "tools", "cuda_emu"))
    try:
        import build as emu_build
    finally:
        sys.path.pop(0)
    try:
        return emu_build.build(str(tmp_path_factory.mktemp("cuda_emu")))
    except FileNotFoundError:
        pytest.skip("g++ not available")


def test_gpu_parity_subset_on_the_emulated_build(emulated_library):
    # (EMU_DEFER_WORKERS: the worker pool's launches run late, so parked captures stay parked across rounds and the host's
    # waiting / lingering / relaunching paths run as they do next to a GPU; results must not depend on it)
    env = dict(os.environ, WSPR_B200_LIB=emulated_library, EMU_DEFER_WORKERS="2")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(H.ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q", "-x",
                        "-p", "no:cacheprovider", "-k", SUBSET], capture_output=True, text=True, env=env, cwd=H.ROOT, timeout=3000)
    tail = r.stdout[-3000:] + r.stderr[-1500:]
    assert r.returncode == 0, tail
    last = [x for x in r.stdout.splitlines() if " passed" in x][-1]
    assert "failed" not in last and int(last.split()[0]) >= 22, tail


def test_emulated_build_exports_the_c_abi(emulated_library):
    import ctypes as C
    import re
    lib = C.CDLL(emulated_library)
    with open(os.path.join(H.ROOT, "include", "wspr_b200.h")) as f:
        names = set(re.findall(r"\b(wspr_[a-z0-9_]+)\s*\(", f.read()))
    assert len(names) > 25
    for n in sorted(names):
        assert hasattr(lib, n), n
