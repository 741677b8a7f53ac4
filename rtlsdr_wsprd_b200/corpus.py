"""Seeded synthetic WSPR corpus, the signal model of the reference's own self-test
(decoderSelfTest, rtlsdr_wsprd.c:729-760): continuous-phase 4-FSK, 162 symbols of 256 samples at 375 sps,
tone spacing 375/256 Hz, nominal start 2.0 s into the capture, additive white Gaussian noise.

SNR is quoted in the WSPR convention (2500 Hz reference bandwidth): with complex noise variance 1 per sample
(0.5 per component) the tone amplitude is A = sqrt(10^(snr/10) * 2500 / 375).

Captures are keyed by (config, index) through numpy's counter-based Philox generator, so any capture can be
regenerated independently on any rank.  The arrays leave here peak-normalised to 0.5 exactly like the
reference's hand-off (rtlsdr_wsprd.c:291-305).
"""
import numpy as np

NSAMP = 45000
NSYM = 162
SPS = 256
FS = 375.0
DF = 375.0 / 256.0

# fixed table of (callsign, grid) used by the multi-signal configurations
STATIONS = [("K1JT", "FN20"), ("VA2GKA", "FN35"), ("W1AW", "FN31"), ("G4JNT", "IO90"), ("DL1ABC", "JO62"),
            ("JA1XYZ", "PM95"), ("VK2DEF", "QF56"), ("ZL3GHI", "RE66"), ("PY2JKL", "GG66"), ("K9AN", "EN50"),
            ("OH2MNO", "KP20"), ("EA4PQR", "IN80")]
POWERS = [0, 3, 7, 10, 13, 17, 20, 23, 27, 30, 33, 37, 40, 43, 47, 50, 53, 57, 60]


def rng_for(config_id, index):
    return np.random.Generator(np.random.Philox(key=[np.uint64(config_id), np.uint64(index)]))


def normalise_half(i, q):
    """Peak normalisation of the hand-off: scale = (float)(0.5 / (double)max), multiply in float."""
    m = np.float32(1e-24)
    if i.size:
        m = max(m, np.float32(np.max(np.abs(i))), np.float32(np.max(np.abs(q))))
    scale = np.float32(0.5 / float(m))
    return (i * scale).astype(np.float32), (q * scale).astype(np.float32)


def add_signal(i_acc, q_acc, symbols, f0, dt0, amp, drift=0.0):
    """Accumulate one WSPR transmission into float64 accumulators. symbols: 162 values in 0..3."""
    sym = np.asarray(symbols, dtype=np.float64)
    k = np.arange(NSYM * SPS)
    s_idx = k // SPS
    # linear drift of +-drift/2 Hz over the transmission, 0 at its centre (wsprd.c:642-644)
    f = f0 + (sym[s_idx] - 1.5) * DF + 0.5 * drift * (s_idx - NSYM / 2.0) / (NSYM / 2.0)
    dphi = 2.0 * np.pi * f / FS
    phi = np.concatenate(([0.0], np.cumsum(dphi[:-1])))
    start = int(np.floor((2.0 + dt0) * FS))
    lo, hi = max(start, 0), min(start + NSYM * SPS, NSAMP)
    if hi > lo:
        sl = slice(lo - start, hi - start)
        i_acc[lo:hi] += amp * np.cos(phi[sl])
        q_acc[lo:hi] += amp * np.sin(phi[sl])


def snr_to_amp(snr_db):
    return float(np.sqrt(10.0 ** (snr_db / 10.0) * 2500.0 / FS))


def make_capture(config_id, index, signals, symbols_fn, noise=True):
    """signals: list of dicts(message, f0, dt0, snr[, drift]); symbols_fn(message)->uint8[162].
    Returns normalised (I, Q) float32[45000]."""
    rng = rng_for(config_id, index)
    if noise:
        n = rng.standard_normal((2, NSAMP)) * np.sqrt(0.5)
        i_acc, q_acc = n[0].copy(), n[1].copy()
    else:
        i_acc, q_acc = np.zeros(NSAMP), np.zeros(NSAMP)
    for s in signals:
        add_signal(i_acc, q_acc, symbols_fn(s["message"]), s["f0"], s["dt0"], snr_to_amp(s["snr"]), s.get("drift", 0.0))
    return normalise_half(i_acc.astype(np.float32), q_acc.astype(np.float32))


def single_signal_plan(index, snr=-20.0, config_id=2):
    """BASELINE config 2: one signal per capture at a fixed SNR."""
    rng = rng_for(config_id + 1000, index)
    call, grid = STATIONS[int(rng.integers(len(STATIONS)))]
    pwr = POWERS[int(rng.integers(len(POWERS)))]
    return [dict(message="%s %s %d" % (call, grid, pwr), f0=float(rng.uniform(-100, 100)),
                 dt0=float(rng.uniform(-1, 1)), snr=snr)]


def ten_signal_plan(index, config_id=3, snrs=None):
    """BASELINE config 3: ten overlapping signals, SNRs -28..-10 dB in 2 dB steps randomly permuted, ten 20 Hz
    frequency slots across +-100 Hz with +-2 Hz jitter, dt0 in +-0.5 s."""
    rng = rng_for(config_id + 1000, index)
    snrs = np.array(snrs if snrs is not None else np.arange(-28.0, -9.0, 2.0))
    snrs = rng.permutation(snrs)
    who = rng.permutation(len(STATIONS))[: len(snrs)]
    plan = []
    for s in range(len(snrs)):
        call, grid = STATIONS[int(who[s])]
        pwr = POWERS[int(rng.integers(len(POWERS)))]
        f0 = -90.0 + 20.0 * s + float(rng.uniform(-2, 2))
        plan.append(dict(message="%s %s %d" % (call, grid, pwr), f0=f0, dt0=float(rng.uniform(-0.5, 0.5)),
                         snr=float(snrs[s])))
    return plan


def make_corpus(config, count, symbols_fn, start=0):
    """config in {2, 3}: returns (I, Q) float32[count, 45000] and the list of plans."""
    I = np.zeros((count, NSAMP), np.float32)
    Q = np.zeros((count, NSAMP), np.float32)
    plans = []
    for c in range(count):
        plan = single_signal_plan(start + c) if config == 2 else ten_signal_plan(start + c)
        I[c], Q[c] = make_capture(config, start + c, plan, symbols_fn)
        plans.append(plan)
    return I, Q, plans


def make_raw_stream(config_id, index, n_iq, f0=50.0, snr=-10.0, symbols=None, sigma=20.0, dt0=0.0):
    """BASELINE config 4: raw RTL-SDR style stream, interleaved u8 (I,Q) at 2.4 Msps.  The front end mixes by
    +fs/4 (rtlsdr_wsprd.c:171-182), so a tone at baseband offset (-600000 + f) lands at f after decimation.
    Returns uint8[2*n_iq].  Generated in chunks to bound memory."""
    rng = rng_for(config_id, index)
    out = np.empty(2 * n_iq, np.uint8)
    fs = 2400000.0
    amp = float(np.sqrt(10.0 ** (snr / 10.0) * 2.0 * sigma * sigma * 2500.0 / fs))
    chunk = 1 << 22
    sym = None if symbols is None else np.asarray(symbols, dtype=np.float64)
    phi0 = 0.0
    for lo in range(0, n_iq, chunk):
        hi = min(lo + chunk, n_iq)
        n = np.arange(lo, hi, dtype=np.float64)
        x = rng.standard_normal((2, hi - lo)) * sigma
        if sym is not None:
            t = n / fs - (2.0 + dt0)
            si = np.floor(t * FS / SPS).astype(np.int64)
            on = (si >= 0) & (si < NSYM)
            tone = np.where(on, sym[np.clip(si, 0, NSYM - 1)], 0.0)
            f = -600000.0 + f0 + (tone - 1.5) * DF
            dphi = 2.0 * np.pi * f / fs
            phi = phi0 + np.concatenate(([0.0], np.cumsum(dphi[:-1])))
            phi0 = float(phi[-1] + dphi[-1])
            x[0] += np.where(on, amp * np.cos(phi), 0.0)
            x[1] += np.where(on, amp * np.sin(phi), 0.0)
        v = np.clip(np.rint(127.5 + x), 0, 255).astype(np.uint8)
        out[2 * lo:2 * hi:2] = v[0]
        out[2 * lo + 1:2 * hi:2] = v[1]
    return out


# ---- BASELINE config 4 at full size: raw streams synthesised from a counter-based INTEGER generator -------------------
# 256 streams x 288e6 IQ pairs are 147 GB: they are generated on the device, and the parity streams are regenerated on the
# host.  Every step below is 64-bit integer arithmetic with wrap-around, written against an array module `xp` (numpy or
# torch), so that both produce the same bytes: noise = sum of four hash bytes per component (Irwin-Hall, sigma = 20.06 LSB),
# signal = a 32-bit phase accumulator through a 1024-entry cosine table in Q14.
RAW_FS = 2400000.0
RAW_SYMBOL = 256 * 6400                              # raw samples per WSPR symbol (375 sps x 6400)
_M64 = (1 << 64) - 1


def _s64(v):
    """python int -> the signed 64-bit value with the same bit pattern"""
    v &= _M64
    return v - (1 << 64) if v >> 63 else v


def raw_stream_plan(index, symbols_fn, snr=-10.0, sigma_lsb=20.06):
    """One signal per stream: tone at -600000 + f0 Hz so that the fs/4 mixer lands it at f0."""
    rng = rng_for(4 + 1000, index)
    call, grid = STATIONS[int(rng.integers(len(STATIONS)))]
    pwr = POWERS[int(rng.integers(len(POWERS)))]
    msg = "%s %s %d" % (call, grid, pwr)
    f0, dt0 = float(rng.uniform(-100, 100)), float(rng.uniform(-0.5, 0.5))
    sym = np.asarray(symbols_fn(msg), dtype=np.int64)
    amp = float(np.sqrt(10.0 ** (snr / 10.0) * 2.0 * sigma_lsb * sigma_lsb * 2500.0 / RAW_FS))
    inc4 = [int(round((-600000.0 + f0 + (s - 1.5) * DF) / RAW_FS * 2.0 ** 32)) & 0xffffffff for s in range(4)]
    inc = [inc4[int(s)] for s in sym]
    base, acc = [], 0
    for k in range(NSYM):                            # phase at the first sample of symbol k
        base.append(acc)
        acc = (acc + inc[k] * RAW_SYMBOL) & 0xffffffff
    return dict(index=index, message=msg, f0=f0, dt0=dt0, snr=snr, start=int(np.floor((2.0 + dt0) * RAW_FS)),
                amp_q14=int(round(amp * 16384.0)), inc=inc, base=base, key=_s64(0x9E3779B97F4A7C15 * (index + 1)))


_COS_Q14 = np.round(np.cos(2.0 * np.pi * np.arange(1024) / 1024.0) * 16384.0).astype(np.int64)


def synth_raw_stream(xp, plan, n_iq, out=None, device=None, block=1 << 23):
    """uint8[2 * n_iq] interleaved (I, Q) of plan's stream.  xp = numpy (host) or torch (out: a uint8 device tensor view)."""
    is_np = xp is np
    if is_np:
        res = np.empty(2 * n_iq, np.uint8) if out is None else out
        mk = lambda a: np.asarray(a, dtype=np.int64)
        arange = lambda lo, hi: np.arange(lo, hi, dtype=np.int64)
        clamp = lambda a, lo, hi: np.clip(a, lo, hi)
        old = np.seterr(over="ignore")
    else:
        res = out
        mk = lambda a: xp.as_tensor(np.asarray(a, dtype=np.int64), device=device)
        arange = lambda lo, hi: xp.arange(lo, hi, dtype=xp.int64, device=device)
        clamp = lambda a, lo, hi: xp.clamp(a, lo, hi)
    inc_t, base_t, cos_t = mk(plan["inc"] + [0]), mk(plan["base"] + [0]), mk(_COS_Q14)
    c1, c2 = _s64(0xBF58476D1CE4E5B9), _s64(0x94D049BB133111EB)

    def lsr(x, s):                                   # logical shift right of a signed 64-bit array
        return (x >> s) & ((1 << (64 - s)) - 1)

    for lo in range(0, n_iq, block):
        hi = min(lo + block, n_iq)
        n = arange(lo, hi)
        x = n + plan["key"]                          # splitmix64 finaliser of (sample index + stream key)
        x = (x ^ lsr(x, 30)) * c1
        x = (x ^ lsr(x, 27)) * c2
        x = x ^ lsr(x, 31)
        ni = ((x & 255) + ((x >> 8) & 255) + ((x >> 16) & 255) + ((x >> 24) & 255) - 510) * 139
        nq = (((x >> 32) & 255) + ((x >> 40) & 255) + ((x >> 48) & 255) + (lsr(x, 56) & 255) - 510) * 139
        rel = n - plan["start"]
        k = rel // RAW_SYMBOL if is_np else xp.div(rel, RAW_SYMBOL, rounding_mode="floor")
        on = (k >= 0) & (k < NSYM)
        kc = clamp(k, 0, NSYM - 1)
        ph = (base_t[kc] + inc_t[kc] * (rel - kc * RAW_SYMBOL)) & 0xffffffff
        amp = on * plan["amp_q14"] if is_np else on.to(xp.int64) * plan["amp_q14"]
        si = (amp * cos_t[(ph >> 22) & 1023]) >> 18
        sq = (amp * cos_t[((ph >> 22) + 768) & 1023]) >> 18         # sin = cos(phase - 90 deg)
        vi = clamp((130560 + 512 + ni + si) >> 10, 0, 255)
        vq = clamp((130560 + 512 + nq + sq) >> 10, 0, 255)
        if is_np:
            res[2 * lo:2 * hi:2] = vi.astype(np.uint8)
            res[2 * lo + 1:2 * hi:2] = vq.astype(np.uint8)
        else:
            res[2 * lo:2 * hi:2] = vi.to(xp.uint8)
            res[2 * lo + 1:2 * hi:2] = vq.to(xp.uint8)
    if is_np:
        np.seterr(**old)
    return res
