"""`apply`: only the ORACLE exists so far -- the restatement of the per-region
integrators (oracle/wt_oracle.c: wto_apply) pinned on the reference's golden vector
(test/expected/regional_means.txt through test/test.py:101) and bit for bit on the reference's own
ApplyMultiplexer compiled into oracle/_ref.  The device implementation is round-2 work; these tests
are what it will be held against."""
import numpy as np
import pytest

from oracle.pyoracle import Track

NAMES = ("chr1", "chr10", "chr2")
STATS = ["AUC", "meanI", "varI", "stddevI", "CVI", "maxI", "minI", "spanI"]


def test_reference_golden_regional_means(port):
    """`apply_paste out meanI overlapping.bed fixedStep.wig` (strict is the CLI's default,
    src/commandParser.c:919): 3.5, 5.0, nan for chr1:2-6, chr1:3-8, chr2:1-4 (0-based BED)."""
    d = Track(np.zeros(10, np.int32), np.arange(1, 11, dtype=np.int32), np.arange(2, 12, dtype=np.int32), np.arange(10.0))
    regions = Track(np.array([0, 0, 1], np.int32), np.array([3, 4, 2], np.int32), np.array([7, 9, 5], np.int32), np.ones(3))
    got = port.apply(d, regions, ["meanI"], ("chr1", "chr2"), strict=True).ravel()
    assert got[0] == 3.5 and got[1] == 5.0 and np.isnan(got[2])
    loose = port.apply(d, regions, ["meanI", "spanI"], ("chr1", "chr2"), strict=False)
    assert loose[2, 0] == 0.0 and loose[2, 1] == 3.0           # fillIn: the dataset's default over the whole region


def random_sorted(rng, n, nchrom, maxlen, maxgap, overlap=False):
    chrom = np.sort(rng.integers(0, nchrom, n)).astype(np.int32)
    start = np.empty(n, np.int32)
    lens = rng.integers(1, maxlen, n)
    gaps = rng.integers(0, maxgap, n)
    pos, prev = 1, -1
    for i in range(n):
        if chrom[i] != prev:
            pos, prev = 1, chrom[i]
        pos += gaps[i]
        start[i] = pos
        pos += 1 if overlap else lens[i]        # overlapping: regions only advance by their gap + 1 (nested and staggered)
    return chrom, start, (start + lens).astype(np.int32)


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("strict", [False, True])
def test_port_matches_reference_apply(port, reference, seed, strict):
    rng = np.random.default_rng(seed)
    c, s, e = random_sorted(rng, 4000, 3, 12, 9)
    default = [0.0, 1.3, 0.0, 2.0, 0.0, 0.7][seed]             # 1.3 / 0.7: rounded through float by the buffer's constructor
    dataset = Track(c, s, e, rng.integers(0, 50, len(s)) / 4.0, default=default)
    rc, rs, re_ = random_sorted(rng, 300, 3, 150, 60, overlap=bool(seed % 2))
    regions = Track(rc, rs, re_, np.ones(len(rs)))
    a = port.apply(dataset, regions, STATS, NAMES, strict)
    b = reference.apply(dataset, regions, STATS, NAMES, strict)
    assert np.array_equal(a, b, equal_nan=True), np.argwhere(~((a == b) | (np.isnan(a) & np.isnan(b))))[:5]
