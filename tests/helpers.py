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


def symbols_from_packed(n, m, lib=None):
    """162 channel symbols from the packed 28-bit callsign field n and 22-bit grid/power field m, the tail of
    get_wspr_channel_symbols (wsprsim_utils.c:262-316): 50 bits -> 11 bytes -> convolutional code -> interleave -> 2 * bit + sync.
    For messages the text interface refuses to encode although the decoder produces them."""
    lib = lib or po.oracle()
    data = (C.c_ubyte * 11)()
    data[0], data[1], data[2] = (n >> 20) & 255, (n >> 12) & 255, (n >> 4) & 255
    data[3], data[4], data[5], data[6] = ((n & 15) << 4) | ((m >> 18) & 15), (m >> 10) & 255, (m >> 2) & 255, (m & 3) << 6
    bits = (C.c_ubyte * 176)()
    lib.encode(bits, data, 11)
    chan = (C.c_ubyte * 162)(*bits[:162])
    lib.interleave(chan)
    sync = channel_symbols("K1JT FN20 20", lib) & 1              # the sync vector is the low bit of every channel symbol
    return (2 * np.frombuffer(bytes(chan), np.uint8) + sync).astype(np.uint8)


def break_captures(lib=None):
    """Two captures that make the reference leave its candidate loop early (wsprd.c:781-793): the strongest signal, examined
    first in pass 0, decodes to a message that (a) get_wspr_channel_symbols cannot encode again for the subtraction -- a
    type-1 message with a three-character callsign -- or (b) carries the locator 'A000AA'; two ordinary signals follow and
    are never reached, and with no unique decode pass 1 does not run (wsprd.c:521-522): the result is empty."""
    lib = C.CDLL((lib or po.oracle())._name)                     # (a private handle: return types are set on it below)
    lib.pack_call.restype = C.c_ulong
    lib.pack_grid4_power.restype = C.c_ulong
    lib.get_locator_character_code.restype = C.c_ubyte
    grid = (C.c_char * 5)(*[bytes([lib.get_locator_character_code(C.c_char(ch.encode()))]) for ch in "FN20"])
    special = {"three-character callsign": symbols_from_packed(int(lib.pack_call(b"A1A")), int(lib.pack_grid4_power(grid, 20)), lib),
               "locator A000AA": channel_symbols("<K1JT> A000AA 23", lib)}
    out = []
    for k, (name, sym) in enumerate(special.items()):
        plan = [dict(message="SPECIAL", f0=-40.0, dt0=0.0, snr=-5.0), dict(message="K1JT FN20 20", f0=20.0, dt0=0.2, snr=-12.0),
                dict(message="W1AW FN31 30", f0=70.0, dt0=-0.3, snr=-14.0)]
        out.append((name, corpus.make_capture(91, 1 + k, plan, lambda msg, s=sym: s if msg == "SPECIAL" else channel_symbols(msg, lib))))
    return out


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


# a hashtable.txt found on entry: existing entries, one the first call overwrites, an out-of-range one, garbage
HASHTABLE_SEED_FILE = "   17 ZZ9ZZZ AA00\n40000 BAD\n  junk\n 5970 OLDCALL\n"


def spots_match_golden(r, gold_spots, line_fn):
    """Every field of a RESULT_DTYPE array against the JSON form written by tools/make_golden*.py (floats as hex)."""
    if len(r) != len(gold_spots):
        return False
    for x, y in zip(r, gold_spots):
        if (x["message"].decode(), x["call"].decode(), x["loc"].decode(), x["pwr"].decode()) != \
                (y["message"], y["call"], y["loc"], y["pwr"]):
            return False
        if (float(x["freq"]).hex(), float(x["snr"]).hex(), float(x["dt"]).hex(), float(x["sync"]).hex()) != \
                (y["freq"], y["snr"], y["dt"], y["sync"]):
            return False
        if (float(x["drift"]), int(x["jitter"]), int(x["cycles"]), line_fn(x)) != (y["drift"], y["jitter"], y["cycles"], y["line"]):
            return False
    return True


def hashtable_scenario():
    """Three captures for the persistent-hashtable option (reference -H, wsprd.c:481-494,842-852): A teaches the table a
    type-1 and a type-2 callsign, B refers to them (and to an unknown one) by hash in type-3 messages, then A again."""
    def cap(idx, msgs):
        sig = [dict(message=m, f0=-80.0 + 35.0 * k, dt0=0.1 * k, snr=-12.0) for k, m in enumerate(msgs)]
        return corpus.make_capture(77, idx, sig, channel_symbols)
    a = cap(0, ["K1JT FN20 20", "PJ4/K1ABC 37", "W1AW FN31 30"])
    b = cap(1, ["<PJ4/K1ABC> FK52UD 37", "<K1JT> FN20AB 20", "<G4JNT> IO90AA 23", "VA2GKA FN35 10"])
    return [a, b, a]


def run_hashtable_scenario(decode_one, seed_file=None):
    """decode_one(i, q) -> results, called in a fresh scratch CWD; returns [(results, hashtable.txt text)] per capture."""
    import tempfile
    out, old = [], os.getcwd()
    with tempfile.TemporaryDirectory(prefix="wspr_ht_") as d:
        os.chdir(d)
        try:
            if seed_file is not None:
                with open("hashtable.txt", "w") as f:
                    f.write(seed_file)
            for i, q in hashtable_scenario():
                r = decode_one(i.copy(), q.copy())
                out.append((r, open("hashtable.txt").read()))
        finally:
            os.chdir(old)
    return out
