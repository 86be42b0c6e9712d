"""Text sink: the device formatter must print the bytes the
reference's own sink prints (CompressionWiggleIterator + TeeWiggleIterator +
printBlock, src/unaryOps.c:235-263, src/wigWriter.c:55-200) -- compared byte for byte.

CPU: the C restatement (oracle/wt_oracle.c: wto_write) against the unmodified reference
sink compiled into oracle/_ref and against the committed golden texts.
GPU: wb200_format_host against the restatement (and the reference sink when present)."""
import hashlib
import json
import os

import numpy as np
import pytest

from util import HERE

NAMES = ("chr1", "chr10", "chrX")
VALUE_POOL = np.array([0.5, 0.5, 0.5000004, 0.5000009, 0.4999991, 1.25, np.nan, -3.0000005, 1e-7, -1e-7, 0.0, -0.0,
                       123456.7890125, 0.0000005, 0.0000015, 0.0000025, 2.5e-7, 1e15 + 0.5, -1e18, np.inf, -np.inf,
                       1 / 3, 2 / 3, 1e-300, 5e-324, 9007199254740993.0, 0.9999995, 0.99999949999, 1234.5678905])


def make_segments(n, seed, nchrom=3, pool=VALUE_POOL, runs=True):
    rng = np.random.default_rng(seed)
    lens = rng.choice([1, 1, 1, 2, 3, 5, 6, 7, 30], n)
    gaps = rng.integers(0, 3, n) * rng.integers(0, 2, n)
    chrom = np.sort(rng.integers(0, nchrom, n)).astype(np.int32)
    start = np.empty(n, np.int64)
    pos, prev = 1, -1
    for i in range(n):
        if chrom[i] != prev:
            pos, prev = 1, chrom[i]
        pos += gaps[i]
        start[i] = pos
        pos += lens[i]
    value = pool[rng.integers(0, len(pool), n)]
    if runs:   # long runs of near-equal values: chains of the compress rule
        k = 0
        while k < n:
            m = int(rng.integers(1, 80))
            if rng.random() < 0.3:
                value[k:k + m] = 7.0 + rng.integers(-9, 10, min(m, n - k)) * 1e-7
            k += m
    return chrom, start.astype(np.int32), (start + lens).astype(np.int32), value


CASES = [(300, 1), (25000, 2), (25000, 3), (61000, 4)]


def golden_digest():
    with open(os.path.join(HERE, "golden", "writer_digests.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("n,seed", CASES)
@pytest.mark.parametrize("bedgraph", [False, True])
def test_port_matches_golden(port, n, seed, bedgraph):
    c, s, e, v = make_segments(n, seed)
    text = port.write(c, s, e, v, NAMES, bedgraph)
    want = golden_digest()["%d/%d/%d" % (n, seed, int(bedgraph))]
    assert hashlib.sha256(text).hexdigest() == want["sha256"] and len(text) == want["bytes"]


@pytest.mark.parametrize("n,seed", CASES + [(12000, 11), (12000, 12)])
@pytest.mark.parametrize("bedgraph", [False, True])
def test_port_matches_reference_sink(port, reference, n, seed, bedgraph):
    c, s, e, v = make_segments(n, seed)
    assert port.write(c, s, e, v, NAMES, bedgraph) == reference.write(c, s, e, v, NAMES, bedgraph)


def test_reference_fixture_text(port):
    """`mean fixedStep.wig variableStep.wig` through the default sink: the text SURVEY 8(d)
    config 1 quotes from the reference CLI."""
    start = np.arange(1, 11, dtype=np.int32)
    value = np.array([.5, 1.5, 1, 3, 2, 4.5, 3, 6, 4, 4.5])
    text = port.write(np.zeros(10, np.int32), start, start + 1, value, ("chr1",), False)
    assert text == b"fixedStep chrom=chr1 start=1 step=1\n" + b"".join(b"%.6f\n" % x for x in value)


# ---- GPU --------------------------------------------------------------------------------------
class _Seg:
    def __init__(self, c, s, e, v):
        self.chrom, self.start, self.end, self.value = c, s, e, v


def first_diff(a, b):
    m = min(len(a), len(b))
    k = next((i for i in range(m) if a[i] != b[i]), m)
    return "lengths %d / %d, first difference at byte %d: %r vs %r" % (len(a), len(b), k, a[max(0, k - 60):k + 40], b[max(0, k - 60):k + 40])


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed", CASES + [(12000, 11), (200000, 13)])
@pytest.mark.parametrize("bedgraph", [False, True])
def test_device_sink_matches_oracle(engine, port, n, seed, bedgraph):
    c, s, e, v = make_segments(n, seed)
    finite = np.where(np.abs(v) >= 2.0 ** 63, 1.0, v)      # infinities stay; the formatter's documented limit is 2^63
    finite = np.where(np.isinf(v), v, finite)
    got = engine.format_segments(_Seg(c, s, e, finite), NAMES, bedgraph)
    want = port.write(c, s, e, finite, NAMES, bedgraph)
    assert got == want, first_diff(got, want)


@pytest.mark.gpu
def test_device_sink_decimal_rounding(engine, port):
    """"%lf": round-half-even on the exact binary value, carries into the integer part, signs."""
    rng = np.random.default_rng(5)
    v = np.concatenate([
        (rng.integers(0, 1 << 20, 4000) + 0.5) / (1 << rng.integers(1, 30, 4000)).astype(np.float64),     # exact ties and near-ties
        rng.integers(0, 1 << 30, 2000) * 2.0 ** -rng.integers(0, 80, 2000),
        np.ldexp(rng.random(2000), rng.integers(-80, 62, 2000)) * rng.choice([-1, 1], 2000),
        np.array([0.9999995, 0.9999994999999999, 1.9999995, 999999.9999995, 0.0000005, 0.00000049999999, 0.0000015,
                  -0.0000005, 4503599627370496.5, 9223372036854774784.0, -9223372036854774784.0, 2.0 ** -1074, 2.0 ** -64,
                  2.0 ** -65 * 3, 0.5 ** 20]),
    ])
    n = len(v)
    start = np.arange(1, 1 + 10 * n, 10, dtype=np.int32)
    for bedgraph in (True, False):
        got = engine.format_segments(_Seg(np.zeros(n, np.int32), start, start + 7, v), ("chr7",), bedgraph)
        want = port.write(np.zeros(n, np.int32), start, start + 7, v, ("chr7",), bedgraph)
        assert got == want, first_diff(got, want)


@pytest.mark.gpu
def test_device_sink_long_chain_and_blocks(engine, port):
    """One chain of near-equal neighbours longer than a resolve tile, drifting past the
    tolerance several times (heads must be chosen against the run's first value), and
    record counts straddling the 10000-record block reset."""
    n = 50000
    start = np.arange(1, n + 1, dtype=np.int32)
    v = 3.0 + np.arange(n) * 4e-8                      # neighbours 4e-8 apart: a new head every 25 steps
    v[30000:30040] = np.nan
    c = np.zeros(n, np.int32)
    for bedgraph in (False, True):
        got = engine.format_segments(_Seg(c, start, start + 1, v), ("chrQ",), bedgraph)
        want = port.write(c, start, start + 1, v, ("chrQ",), bedgraph)
        assert got == want, first_diff(got, want)


@pytest.mark.gpu
def test_device_sink_rejects_huge_values(engine):
    from wiggletools_b200.api import WB200Error
    s = np.array([1], np.int32)
    with pytest.raises(WB200Error):
        engine.format_segments(_Seg(np.zeros(1, np.int32), s, s + 1, np.array([1e19])), ("chr1",), True)


@pytest.mark.gpu
def test_device_sink_after_reduction(engine, port):
    """mean over random tracks, then the sink: the CLI's `write - mean ...` end to end."""
    from oracle.pyoracle import Track as OTrack
    from wiggletools_b200.api import Track
    rng = np.random.default_rng(9)
    tracks, otracks = [], []
    for _ in range(5):
        st = np.cumsum(rng.integers(1, 9, 4000)).astype(np.int32)
        en = np.minimum(st + rng.integers(1, 6, 4000), np.r_[st[1:], st[-1] + 9]).astype(np.int32)
        val = rng.integers(0, 4, 4000) / 2.0
        tracks.append(Track(np.zeros(4000, np.int32), st, en, val))
        otracks.append(OTrack(np.zeros(4000, np.int32), st, en, val))
    seg = engine.reduce("mean", tracks, ("chr1",))
    ref = port.reduce("mean", otracks, ("chr1",))
    for bedgraph in (False, True):
        got = engine.format_segments(seg, ("chr1",), bedgraph)
        assert got == port.write(ref.chrom, ref.start, ref.end, ref.value, ("chr1",), bedgraph)


@pytest.mark.gpu
def test_cli_write_through_device_sink(tmp_path):
    """`write f mean ...`, `write_bg f mean ...`, plain `mean ...` (stdout) and a Tee in the middle of
    a pipeline: the reference CLI linked against the library (device sink routed in by
    integration/dropin_binding.c) must print the bytes of the unmodified reference CLI."""
    import subprocess
    ref = os.path.join(HERE, "..", "oracle", "_ref", "wiggletools")
    exe = os.path.join(HERE, "..", "integration", "_build", "wiggletools_b200")
    if not (os.path.exists(ref) and os.path.exists(exe)):
        pytest.skip("oracle/_ref CLIs not built (needs /root/reference at build time)")
    rng = np.random.default_rng(21)
    files = []
    for t in range(4):
        path = tmp_path / ("t%d.bg" % t)
        with open(path, "w") as f:
            for chrom in ("chr1", "chr10", "chr2"):
                pos = 1
                for _ in range(6000):
                    pos += int(rng.integers(0, 4)) * int(rng.integers(0, 2))
                    ln = int(rng.choice([1, 1, 2, 3, 6, 9]))
                    f.write("%s\t%d\t%d\t%g\n" % (chrom, pos - 1, pos - 1 + ln, rng.integers(0, 5) / 4.0))
                    pos += ln
        files.append(path.name)

    def run(binary, args, out=None):
        r = subprocess.run([binary] + args, capture_output=True, cwd=tmp_path, check=True)
        return open(tmp_path / out, "rb").read() if out else r.stdout

    for verb in ("write", "write_bg"):
        want = run(ref, [verb, "ref_%s.out" % verb, "mean"] + files, "ref_%s.out" % verb)      # the CLI refuses to overwrite
        got = run(exe, [verb, "gpu_%s.out" % verb, "mean"] + files, "gpu_%s.out" % verb)
        assert len(want) > 100000 and got == want, verb + ": " + first_diff(got, want)
    want = run(ref, ["sum"] + files)
    got = run(exe, ["sum"] + files)
    assert got == want, "stdout: " + first_diff(got, want)
    # a Tee whose records are consumed downstream stays the reference's Tee
    want = run(ref, ["AUC", "write", "ref2.out", "mean"] + files) + run(ref, ["do", "unit", "write_bg", "ref3.out", "mean"] + files)
    got = run(exe, ["AUC", "write", "gpu2.out", "mean"] + files) + run(exe, ["do", "unit", "write_bg", "gpu3.out", "mean"] + files)
    assert got == want
    assert open(tmp_path / "gpu2.out", "rb").read() == open(tmp_path / "ref2.out", "rb").read()
    assert open(tmp_path / "gpu3.out", "rb").read() == open(tmp_path / "ref3.out", "rb").read()


def test_formatter_host_build_against_printf():
    """The device formatter's "%lf" is plain integer arithmetic; the same source function is also
    compiled for the host (wb200_debug_format_value).  Fuzz it against the C library's printf:
    random doubles of every magnitude, exact ties of the 6th decimal, subnormals, signs, NaN / Inf."""
    import ctypes as C
    from wiggletools_b200 import _lib
    lib = _lib.load()
    libc = C.CDLL(None)
    libc.snprintf.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_double]
    rng = np.random.default_rng(123)
    parts = [
        np.ldexp(rng.random(200000), rng.integers(-1080, 63, 200000)) * rng.choice([-1.0, 1.0], 200000),
        (rng.integers(0, 1 << 40, 100000) * 2 + 1) / 2.0 ** rng.integers(1, 60, 100000),            # odd / 2^k: many exact ties
        rng.integers(0, 10 ** 9, 100000) / 1e6 + rng.choice([0.0, 5e-7, -5e-7, 4.999999e-7], 100000),
        rng.integers(-(1 << 62), 1 << 62, 50000).astype(np.float64),
        np.frombuffer(rng.integers(0, 1 << 63, 100000, dtype=np.int64).tobytes(), dtype=np.float64),  # random bit patterns
        np.array([0.0, -0.0, 5e-324, -5e-324, 2.0 ** -1074, 0.9999995, 0.99999949999999994, 1e-7, 4.5e-7, 5e-7, 1.5e-6, 2.5e-6,
                  9223372036854774784.0, -9223372036854774784.0, 4503599627370495.5, np.inf, -np.inf, np.nan, -np.nan,
                  123456789.1234565, 0.1234565, 0.1234575, 1e22, 999999.9999995]),
    ]
    buf, ref = C.create_string_buffer(64), C.create_string_buffer(512)
    checked = 0
    for v in np.concatenate(parts):
        n = lib.wb200_debug_format_value(float(v), buf)
        if abs(v) >= 2.0 ** 63 and np.isfinite(v):
            assert n == -1
            continue
        libc.snprintf(ref, 512, b"%lf", float(v))
        assert n == len(ref.value) and buf.value == ref.value, (float(v).hex(), buf.value, ref.value)
        checked += 1
    assert checked > 400000
