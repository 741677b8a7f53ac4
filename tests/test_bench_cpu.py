"""bench.py without a GPU: the reference arm (the one leg that runs entirely on the host cores), the pieces both arms
share (corpus planes in shared memory, generator and CPU-decoder pools, the spot-list comparison) and the rule that the
product arm has no CPU fallback."""
import importlib.util
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as H

BENCH = os.path.join(H.ROOT, "bench.py")


def run_bench(args, env=None, timeout=600):
    e = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE"):
        e.pop(k, None)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=timeout, env=e, cwd=H.ROOT)


@pytest.fixture(scope="module")
def bench_module():
    spec = importlib.util.spec_from_file_location("bench_under_test", BENCH)
    m = importlib.util.module_from_spec(spec)
    sys.modules["bench_under_test"] = m              # (the pools pickle module-level functions by name)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("anon", ["0", "1"])
def test_reference_arm_line(anon):
    """`--impl reference`: one JSON line, the contract's keys, the reference's own code on every host core; with the corpus
    in POSIX shared memory and (BENCH_ANON_SHM=1: what bench.py falls back to when /dev/shm is too small) in anonymous
    shared mappings inherited through fork."""
    r = run_bench(["--impl", "reference", "--steps", "1", "--warmup", "1", "--workload", "config2"], {"BENCH_ANON_SHM": anon})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [x for x in r.stdout.splitlines() if x.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "captures/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["config"]["workload"].startswith("config2") and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = run_bench(["--impl", "reference", "--steps", "1", "--warmup", "0"], {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run_bench(["--units", "16", "--steps", "1", "--cpu-sample", "0", "--no-frontend"])
    assert r.returncode != 0
    assert "CUDA" in r.stderr and "no CPU fallback" in r.stderr
    assert not [x for x in r.stdout.splitlines() if x.startswith("{")]


def test_corpus_planes_pools_and_comparison(bench_module):
    """Generator pool -> shared planes -> CPU decoder pool, in both sharing modes, give identical corpora and identical
    spot lists; compare_spot_lists counts what it should."""
    b = bench_module
    results = []
    for anon in ("0", "1"):
        os.environ["BENCH_ANON_SHM"] = anon
        try:
            planes = b.SharedPlanes(6)
            assert (planes.names is None) == (anon == "1")
            b.generate_corpus(2, 100, planes, 2)
            cpu = b.CpuPool(planes, 2)
            res, wall = cpu.decode(0, 6, chunk=2)
            cpu.close()
            results.append((planes.I.copy(), planes.Q.copy(), res))
            planes.close()
        finally:
            os.environ.pop("BENCH_ANON_SHM", None)
    assert np.array_equal(results[0][0], results[1][0]) and np.array_equal(results[0][1], results[1][1])
    assert results[0][2] == results[1][2] and len(results[0][2]) == 6
    assert sum(len(x) for x in results[0][2]) >= 4                      # config 2: one -20 dB signal per capture, nearly all decode
    ref = results[0][2]
    same = b.compare_spot_lists(ref, ref)
    assert same["captures_checked"] == 6 and same["identical_spot_lists"] == 6 and same["hard_identical"] == 6
    assert same["spots_missing"] == same["spots_extra"] == same["spots_field_mismatched"] == 0
    k = next(i for i, x in enumerate(ref) if x)
    broken = [list(x) for x in ref]
    first = list(broken[k][0])
    first[5] = first[5] + 1.0                                            # snr of one spot: a field mismatch, not a hard one
    broken[k][0] = tuple(first)
    d = b.compare_spot_lists(ref, broken)
    assert d["identical_spot_lists"] == 5 and d["hard_identical"] == 6 and d["spots_field_mismatched"] == 1
    broken = [list(x) for x in ref]
    del broken[k][0]
    d = b.compare_spot_lists(ref, broken)
    assert d["identical_spot_lists"] == 5 and d["hard_identical"] == 5 and d["spots_missing"] == 1 and d["spots_extra"] == 0


def test_raw_stream_synthesiser_is_the_same_on_numpy_and_torch():
    """Config 4 generates its streams on the device (torch) and regenerates the parity streams on the host (numpy): the
    integer recipe must give the same bytes (here: torch on the CPU)."""
    import torch
    from rtlsdr_wsprd_b200 import corpus
    plan = corpus.raw_stream_plan(3, H.channel_symbols)
    n_iq = int(plan["start"]) + 6400 * 300 + 17          # reaches well into the signal: noise, tone and symbol changes
    a = corpus.synth_raw_stream(np, plan, n_iq)
    out = torch.empty(2 * n_iq, dtype=torch.uint8)
    bt = corpus.synth_raw_stream(torch, plan, n_iq, out=out, device="cpu").numpy()
    assert a.dtype == np.uint8 and a.shape == (2 * n_iq,)
    assert np.array_equal(a, bt[: 2 * n_iq])
    assert 100 < a.mean() < 155 and a.std() > 10                         # noise of sigma 20 LSB around 127.5
