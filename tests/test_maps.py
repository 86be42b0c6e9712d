"""Element-wise unary operators on the inputs: `mean map ln a b c`,
`sum scale 2 x y` ... (src/unaryOps.c:650-960 behind src/commandParser.c:115-211).

CPU: the restatement (oracle/wt_oracle.c: wto_map) against the reference's own unary iterators
compiled into oracle/_ref, records and derived defaults bit for bit.
GPU: wb200_set_input_maps + wb200_reduce_host against the oracle reducing oracle-mapped tracks:
coordinates bit for bit (log drops records), values within 1e-10 relative (device log / exp /
pow are within 2 ulp of the host's libm)."""
import numpy as np
import pytest

from oracle.pyoracle import Track as OTrack
from util import RTOL, assert_segments

NAMES = ["chr1", "chr10", "chr2"]
CHAINS = [
    [("scale", 2.0)], [("ln",)], [("log", 10.0)], [("offset", -2.25), ("abs",), ("pow", 0.5)], [("exp",), ("ln",)],
    [("pow", -1.0)], [("default", 3.0), ("scale", 0.1), ("log", 2.0), ("offset", 1.0)], [("scale", -1.0), ("ln",)],
    [("pow", 2.0), ("default", -1.0), ("ln",), ("exp",)], [("abs",), ("offset", 0.5), ("log", 2.0), ("scale", 3.0), ("pow", 3.0)],
]


def random_track(rng, n, nchrom=3, default=0.0, cls=OTrack):
    chrom = np.sort(rng.integers(0, nchrom, n)).astype(np.int32)
    start = np.empty(n, np.int32)
    pos, prev = 1, -1
    lens = rng.integers(1, 12, n)
    gaps = rng.integers(0, 9, n)
    for i in range(n):
        if chrom[i] != prev:
            pos, prev = 1, chrom[i]
        pos += gaps[i]
        start[i] = pos
        pos += lens[i]
    value = rng.choice([0.0, -1.5, 2.25, np.nan, 1e-3, 7.0, -0.0, 100.0, 0.3, 64.0, 1.0], n)
    return cls(chrom, start, (start + lens).astype(np.int32), value, default=default)


@pytest.mark.parametrize("default", [0.0, 2.5, float("nan"), -1.0])
def test_port_matches_reference_unary_iterators(port, reference, default):
    rng = np.random.default_rng(3)
    t = random_track(rng, 4000, default=default)
    for chain in CHAINS:
        a, b = port.map(t, chain, NAMES), reference.map(t, chain, NAMES)
        assert len(a) == len(b) and np.array_equal(a.chrom, b.chrom) and np.array_equal(a.start, b.start) and np.array_equal(a.end, b.end), chain
        assert np.array_equal(a.value, b.value, equal_nan=True), chain
        assert a.default == b.default or (np.isnan(a.default) and np.isnan(b.default)), (chain, a.default, b.default)


def test_reference_identities(port):
    """test/test.py:76-88: `sum x x == scale 2 x`, `pow 2 x == mult x x`, `ln x == log e x` up to the
    reference's own E constant."""
    rng = np.random.default_rng(8)
    t = random_track(rng, 3000)
    two = port.map(t, [("scale", 2.0)], NAMES)
    s = port.reduce("sum", [t, t], NAMES)
    assert np.array_equal(s.start, two.start) and np.array_equal(s.value, two.value, equal_nan=True)
    sq = port.map(t, [("pow", 2.0)], NAMES)
    m = port.reduce("mult", [t, t], NAMES)
    assert np.allclose(m.value, sq.value, rtol=1e-15, atol=0, equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("op", ["mean", "sum", "max", "stddev"])
def test_device_maps_match_oracle(engine, port, op):
    from wiggletools_b200 import api
    rng = np.random.default_rng(5)
    n_tracks = 7
    otracks = [random_track(rng, 3000, default=d) for d in (0.0, 0.0, 2.5, 0.0, float("nan") if op == "max" else 1.0, 0.0, 0.0)]
    gtracks = [api.Track(t.chrom, t.start, t.end, t.value, default=t.default) for t in otracks]
    for k in range(len(CHAINS)):
        maps = [CHAINS[(k + i) % len(CHAINS)] if i % 3 != 2 else [] for i in range(n_tracks)]   # every third track unmapped
        mapped = [port.map(t, ch, NAMES) if ch else t for t, ch in zip(otracks, maps)]
        want = port.reduce(op, mapped, NAMES)
        got = engine.reduce(op, gtracks, NAMES, maps=maps)
        assert_segments(got, want.chrom, want.start, want.end, want.value, rtol=RTOL, what="%s/chain%d" % (op, k))
        assert got.default == want.default or (np.isnan(got.default) and np.isnan(want.default)), (op, k, got.default, want.default)


@pytest.mark.gpu
def test_device_maps_exact_ops_are_bit_identical(engine, port):
    """scale / offset / abs involve one correctly rounded operation: bit-identical, and
    `sum map scale 2 x` == `scale 2 sum x` on the device as well."""
    from wiggletools_b200 import api
    rng = np.random.default_rng(6)
    otracks = [random_track(rng, 5000) for _ in range(5)]
    gtracks = [api.Track(t.chrom, t.start, t.end, t.value) for t in otracks]
    chain = [("scale", 0.75), ("offset", -3.0), ("abs",)]
    want = port.reduce("max", [port.map(t, chain, NAMES) for t in otracks], NAMES)
    got = engine.reduce("max", gtracks, NAMES, maps=[chain] * 5)
    assert_segments(got, want.chrom, want.start, want.end, want.value, exact=True, what="max of exact maps")
    a = engine.reduce("sum", gtracks, NAMES, maps=[[("scale", 2.0)]] * 5)
    b = engine.reduce("sum", gtracks, NAMES)
    assert np.array_equal(a.start, b.start) and np.array_equal(a.value, 2 * b.value, equal_nan=True)


@pytest.mark.gpu
def test_device_maps_everything_dropped_and_errors(engine):
    from wiggletools_b200 import api
    s = np.arange(1, 2001, dtype=np.int32)
    t = api.Track(np.zeros(2000, np.int32), s, s + 1, -np.ones(2000))
    got = engine.reduce("mean", [t, t], ["chr1"], maps=[[("ln",)], [("ln",)]])       # every record has value <= 0
    assert len(got) == 0
    with pytest.raises(api.WB200Error):
        engine.reduce("mean", [t], ["chr1"], maps=[[("scale", 1.0)] * 9])             # more than 8 maps on a track


@pytest.mark.gpu
def test_cli_maps_on_the_device(tmp_path):
    """`mean map ln ...`, `sum map scale ...`, nested maps, a map feeding a host-side operator: the
    reference CLI linked against the library (unary constructors routed to wb200_mapped_iterator
    by integration/dropin_binding.c) prints what the unmodified reference CLI prints."""
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    ref = os.path.join(here, "..", "oracle", "_ref", "wiggletools")
    exe = os.path.join(here, "..", "integration", "_build", "wiggletools_b200")
    if not (os.path.exists(ref) and os.path.exists(exe)):
        pytest.skip("oracle/_ref CLIs not built (needs /root/reference at build time)")
    rng = np.random.default_rng(31)
    files = []
    for t in range(4):
        path = tmp_path / ("t%d.bg" % t)
        with open(path, "w") as f:
            for chrom in ("chr1", "chr10", "chr2"):
                pos = 1
                for _ in range(3000):
                    pos += int(rng.integers(0, 4)) * int(rng.integers(0, 2))
                    ln = int(rng.choice([1, 2, 3, 6, 9]))
                    f.write("%s\t%d\t%d\t%g\n" % (chrom, pos - 1, pos - 1 + ln, rng.integers(-1, 6) / 4.0))
                    pos += ln
        files.append(path.name)
    programs = [
        ["mean", "map", "ln"] + files,
        ["write_bg", "-", "sum", "map", "scale", "0.5"] + files,
        ["mean", "map", "offset", "-0.5", "map", "abs"] + files,
        ["max", "map", "pow", "2"] + files,
        ["AUC", "mean", "map", "log", "2", "map", "offset", "1.5"] + files,
        ["write_bg", "-", "sum", "scale", "3", files[0], "ln", files[1], "default", "2", files[2], files[3]],
        ["write_bg", "-", "unit", "sum", "map", "exp"] + files[:2],
        ["write_bg", "-", "gt", "0.3", "scale", "2", files[0]],                    # a map consumed by a host-side operator
    ]
    env = dict(os.environ, WB200_VERBOSE="1")
    for k, prog in enumerate(programs):
        want = subprocess.run([ref] + prog, capture_output=True, cwd=tmp_path, check=True).stdout
        run = subprocess.run([exe] + prog, capture_output=True, cwd=tmp_path, check=True, env=env)
        assert len(want) > 0 and run.stdout == want, " ".join(prog)
        if k < 7:     # every program but the last feeds its maps to a GPU reduction: they must have run there
            assert b"inputs run on the device" in run.stderr, " ".join(prog)


def test_map_default_matches_the_constructors(port):
    """wb200_map_default (host code, no device needed) against the oracle's derived defaults, which are
    pinned on the reference's constructors above: every operator, defaults that hit each branch."""
    from wiggletools_b200 import _lib
    import math
    lib = _lib.load()
    empty = lambda d: OTrack(np.empty(0, np.int32), np.empty(0, np.int32), np.empty(0, np.int32), np.empty(0), default=d)
    cases = [("scale", 1, 0.3), ("offset", 2, -1.7), ("ln", 3, None), ("log", 3, 10.0), ("exp", 4, None), ("pow", 5, 2.0), ("pow", 5, -0.5), ("abs", 6, None)]
    for d in (0.0, 2.5, -3.25, 1e-3, float("nan"), 1e30):
        for name, code, arg in cases:
            want = port.map(empty(d), [(name,) if arg is None else (name, arg)], NAMES).default
            lowered = 1.0 if name in ("ln", "exp") else 0.0 if name == "abs" else math.log(arg) if name == "log" else arg
            got = lib.wb200_map_default(code, lowered, d)
            assert got == want or (math.isnan(got) and math.isnan(want)), (name, arg, d, got, want)
