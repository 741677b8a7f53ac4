"""Shared test helpers (test infrastructure: may use oracle/)."""
import ctypes as C
import os

import numpy as np

from oracle import pyoracle as po
from rtlsdr_wsprd_b200 import corpus

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
FIELDS = ["freq", "sync", "snr", "dt", "drift", "jitter", "message", "call", "loc", "pwr", "cycles"]
HARD_FIELDS = ["message", "call", "loc", "pwr"]          # BASELINE: bit-for-bit


def channel_symbols(msg, lib=None):
    """162 channel symbols of a message through get_wspr_channel_symbols (wsprsim_utils.c:163-316) of `lib`."""
    lib = lib or po.oracle()
    sym = (C.c_ubyte * 162)()
    ht = C.create_string_buffer(32768 * 13)
    lt = C.create_string_buffer(32768 * 5)
    ok = lib.get_wspr_channel_symbols(C.create_string_buffer(msg.encode(), 32), ht, lt, sym)
    assert ok == 1, msg
    return np.frombuffer(bytes(sym), np.uint8).copy()


def make_corpus(config, count, start=0):
    return corpus.make_corpus(config, count, channel_symbols, start=start)


def results_equal(a, b, fields=FIELDS):
    """Field-wise equality of two RESULT_DTYPE arrays (struct padding and bytes after a NUL are not compared)."""
    if len(a) != len(b):
        return False
    return all(np.array_equal(a[f], b[f]) for f in fields)


def diff_results(a, b):
    out = []
    if len(a) != len(b):
        out.append("count %d vs %d" % (len(a), len(b)))
    for k in range(min(len(a), len(b))):
        for f in FIELDS:
            if a[k][f] != b[k][f]:
                out.append("spot %d %s: %r vs %r" % (k, f, a[k][f], b[k][f]))
    return out


def spot_lines(r):
    return [po.spot_line(x) for x in r]
