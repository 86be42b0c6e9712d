"""BASELINE.json `configs`, each at its own track count on a genome small enough for the CPU
oracle: same generator as bench.py, results against the oracle -- coordinates
bit for bit, values bit for bit where the algorithm guarantees it, else within 1e-10 relative
(the tolerance BASELINE.json's north_star states)."""
import math

import numpy as np
import pytest

from wiggletools_b200 import api
from wiggletools_b200.synth import make_tracks
from oracle.pyoracle import Track as OTrack
from util import RTOL, assert_segments

pytestmark = pytest.mark.gpu

NAMES = ["chr1", "chr10", "chr2", "chrX"]          # strcmp order: chr1 < chr10 < chr2 < chrX


def both(tr):
    o = [OTrack(t["chrom"], t["start"], t["end"], t["value"]) for t in tr]
    g = [api.Track(t["chrom"], t["start"], t["end"], t["value"]) for t in tr]
    return o, g


@pytest.mark.parametrize("align", [1, 50])
def test_config1_mean_64_tracks(engine, port, align):
    """configs[1]: mean over 64 tracks; align=50 is the shared-breakpoint variant (binned coverage:
    far fewer, far wider-fan-in segments)."""
    tr = make_tracks(64, [40_000, 9_000, 25_000], seed=11, align=align, max_len=400 // align, max_gap=200 // align)
    o, g = both(tr)
    want = port.reduce("mean", o, NAMES)
    got = engine.reduce("mean", g, NAMES)
    assert_segments(got, want.chrom, want.start, want.end, want.value, exact=True, what="mean/64/align%d" % align)
    assert engine.last_stats[-1]["sweep_kernel"] == 1            # the headline kernel is the one that ran
    assert math.isclose(got.auc, port.auc("mean", o, NAMES), rel_tol=RTOL)


def test_config2_var_stddev_512_tracks(engine, port):
    """configs[2]: var + stddev over 512 tracks; about a fifth of the (segment, track) pairs are
    absent, which is where var (absent tracks skipped in the error pass) and stddev differ."""
    tr = make_tracks(512, [12_000, 5_000], seed=12)
    o, g = both(tr)
    res = {}
    for op in ("var", "stddev"):
        want = port.reduce(op, o, NAMES)
        got = engine.reduce(op, g, NAMES)
        assert_segments(got, want.chrom, want.start, want.end, want.value, rtol=RTOL, what=op + "/512")
        assert engine.last_stats[-1]["fast_path"] == 2            # exact-moments path
        res[op] = got
    partial = res["var"].value != res["stddev"].value ** 2
    assert partial.any()                                          # quirk Q2 is exercised


@pytest.mark.parametrize("levels", [100000, 8])
def test_config3_median_256_tracks(engine, port, levels):
    """configs[3]: median over 256 tracks (upper median, no averaging); levels=8 is the heavy-ties
    variant."""
    tr = make_tracks(256, [9_000, 4_000], seed=13, vmax=levels)
    o, g = both(tr)
    want = port.reduce("median", o, NAMES)
    got = engine.reduce("median", g, NAMES)
    assert_segments(got, want.chrom, want.start, want.end, want.value, exact=True, what="median/256/%d" % levels)


def test_config4_mean_512_tracks_sharded(engine, port):
    """configs[4]: mean over 512 tracks sharded by chromosome: the concatenation of the per-shard
    outputs in strcmp order equals the single-stream output, and the summed AUC partials equal the
    oracle's sequential AUC within 1e-10.  Host tracks are slices of one array per chromosome
    (the SoA layout a decoder hands over), so the batch entry stages them as one copy per array."""
    lengths = [9_000, 3_000, 7_000, 5_000]
    tr = make_tracks(512, lengths, seed=14)
    o, g = both(tr)
    want = port.reduce("mean", o, NAMES)
    shards = []
    for c in range(len(lengths)):
        parts = [(t["start"][t["chrom"] == c], t["end"][t["chrom"] == c], t["value"][t["chrom"] == c]) for t in tr]
        S = np.concatenate([p[0] for p in parts]); E = np.concatenate([p[1] for p in parts]); V = np.concatenate([p[2] for p in parts])
        off = np.cumsum([0] + [len(p[0]) for p in parts])
        shards.append([(S[off[i]:off[i + 1]], E[off[i]:off[i + 1]], V[off[i]:off[i + 1]], 0.0) for i in range(512)])
    res = engine.reduce_shards("mean", shards)
    order = sorted(range(len(lengths)), key=lambda c: NAMES[c].encode())
    cat = lambda k: np.concatenate([res[c][k] for c in order])
    chrom = np.concatenate([np.full(len(res[c][0]), c, np.int32) for c in order])
    got = api.Segments(chrom, cat(0), cat(1), cat(2))
    assert_segments(got, want.chrom, want.start, want.end, want.value, exact=True, what="mean/512/sharded")
    auc = sum(res[c][4] for c in order)
    assert math.isclose(auc, port.auc("mean", o, NAMES), rel_tol=RTOL)
    # two "ranks": the LPT split bench.py uses; each rank's shards in one call, scalars summed
    from wiggletools_b200.shard import lpt_assign
    owner = lpt_assign(lengths, 2)
    auc2 = 0.0
    for r in range(2):
        mine = [c for c in range(len(lengths)) if owner[c] == r]
        part = engine.reduce_shards("mean", [shards[c] for c in mine])
        for c, p in zip(mine, part):
            assert np.array_equal(p[0], res[c][0]) and np.array_equal(p[2], res[c][2])
            auc2 += p[4]
    assert math.isclose(auc2, auc, rel_tol=1e-12)


@pytest.mark.parametrize("n", [513, 1023])
def test_order_statistics_up_to_the_track_limit(engine, port, n):
    """the shared-memory check rejected median above 512 tracks on sparse data although a
    smaller window fits.  The window now shrinks until the kernel that actually runs fits: median and
    max over 513 and 1023 tracks (the documented per-set limit), sparse enough for 4096-bp windows."""
    tr = make_tracks(n, [6_000, 2_500], seed=21, max_len=300, max_gap=3000, gap_prob=0.9)
    o, g = both(tr)
    for op in ("median", "max"):
        want = port.reduce(op, o, NAMES[:2])
        got = engine.reduce(op, g, NAMES[:2])
        assert_segments(got, want.chrom, want.start, want.end, want.value, exact=True, what="%s/%d" % (op, n))


def test_wilcoxon_track_limit(engine, port):
    """wilcoxon sorts the two sets' values per segment in per-warp shared-memory rows: 1024 tracks in
    total fit (512 + 512), more is refused with WB200_EUNSUPPORTED instead of failing at launch."""
    tr = make_tracks(1024, [1_500], seed=22, max_len=200, max_gap=2000, gap_prob=0.9)
    o, g = both(tr)
    want = port.setcompare("wilcoxon", o[:512], o[512:], NAMES[:1])
    got = engine.setcompare("wilcoxon", g[:512], g[512:], NAMES[:1])
    assert_segments(got, want.chrom, want.start, want.end, want.value, what="wilcoxon/512+512")
    tr = make_tracks(1200, [800], seed=23, max_len=200, max_gap=2000, gap_prob=0.9)
    _, g = both(tr)
    with pytest.raises(api.WB200Error, match="1024 tracks"):
        engine.setcompare("wilcoxon", g[:600], g[600:], NAMES[:1])
