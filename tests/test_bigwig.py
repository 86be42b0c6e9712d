"""Host BigWig batch decoder: CPU tests.  The two .bw files in
tests/golden/ are the reference's own fixtures (test/fixedStep.bw, variableStep.bw,
written by the UCSC tools and read by the reference through libBigWig); their .wig
twins hold the same values (test/test.py:28,52)."""
import os

import numpy as np
import pytest

from wiggletools_b200.bigwig import BigWig
from bwwrite import write_bigwig

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_stretch_records(intervals, length, stretch=10000):
    """What the reference's reader emits for 0-based half-open intervals of one
    chromosome: src/bigWiggleReader.c:73-83 queries [start, start+stretch) for
    start = 1, 1+stretch, ... while start < length; :36-50 shifts to 1-based and
    boxes each overlapping interval into the queried stretch."""
    out = []
    start = 1
    while start < length:
        stop = start + stretch
        for s, e, v in intervals:
            s1, e1 = s + 1, e + 1
            if s1 < stop and e1 > start:          # overlaps the 0-based query [start-1, stop-1)
                out.append((max(s1, start), min(e1, stop), v))
        start += stretch
    return out


def test_reference_fixtures_decode_to_their_wig_twins():
    b = BigWig(os.path.join(HERE, "golden", "fixedStep.bw"))
    assert b.chroms == [("chr1", 15)]
    s, e, v = b.decode(0)
    assert list(s) == list(range(1, 11)) and list(e) == list(range(2, 12)) and list(v) == list(range(10))
    b = BigWig(os.path.join(HERE, "golden", "variableStep.bw"))
    s, e, v = b.decode(0)
    assert list(s) == [1, 2, 4, 6, 8] and list(e) == [2, 3, 5, 7, 9] and list(v) == [1, 2, 3, 4, 5]


def test_not_a_bigwig(tmp_path):
    p = tmp_path / "x.bw"
    p.write_bytes(b"fixedStep chrom=chr1 start=1 step=1\n1\n")
    with pytest.raises(IOError, match="not in BigWig format"):
        BigWig(p)


@pytest.mark.parametrize("compress", [True, False])
def test_written_files_all_section_types(tmp_path, compress):
    rng = np.random.default_rng(5)
    chroms = [("chr2", 65000), ("chr10", 30001), ("chr1", 123457)]   # ids 0,1,2; strcmp order chr1 < chr10 < chr2
    per = {0: [], 1: [], 2: []}
    sections = []
    for cid, (name, length) in enumerate(chroms):
        pos = 0
        while pos < length - 2500:
            kind = int(rng.integers(1, 4))
            n = int(rng.integers(1, 60))
            if kind == 1:
                items = []
                for _ in range(n):
                    pos += int(rng.integers(0, 30))
                    ln = int(rng.integers(1, 25000 if rng.random() < 0.02 else 40))
                    if pos + ln >= length:
                        break
                    items.append((pos, pos + ln, float(np.float32(rng.normal()))))
                    pos += ln
                if not items:
                    break
                sections.append((cid, 1, items, 0, 0))
                per[cid] += items
            elif kind == 2:
                span = int(rng.integers(1, 6))
                items = []
                for _ in range(n):
                    pos += int(rng.integers(0, 20))
                    if pos + span >= length:
                        break
                    items.append((pos, float(np.float32(rng.normal()))))
                    pos += span
                if not items:
                    break
                sections.append((cid, 2, items, span, 0))
                per[cid] += [(s, s + span, v) for s, v in items]
            else:
                span = int(rng.integers(1, 4)); step = span + int(rng.integers(0, 5))
                if pos + n * step + span >= length:
                    break
                items = [(pos + k * step, float(np.float32(rng.normal()))) for k in range(n)]
                sections.append((cid, 3, items, span, step))
                per[cid] += [(s, s + span, v) for s, v in items]
                pos += n * step + span
    sections.sort(key=lambda t: (t[0], t[2][0][0]))
    path = tmp_path / "t.bw"
    write_bigwig(path, chroms, sections, compress=compress, rtree_block=5)
    names = sorted(n for n, _ in chroms)
    for stretch in (0, 10000):
        b = BigWig(path, stretch=stretch)
        assert [c[0] for c in b.chroms] == names
        for i, (name, length) in enumerate(b.chroms):
            cid = [n for n, _ in chroms].index(name)
            s, e, v = b.decode(i)
            if stretch == 0:
                want = [(a + 1, c + 1, x) for a, c, x in per[cid]]
            else:
                want = reference_stretch_records(per[cid], length, stretch)
                # the reference emits stretch by stretch; so does a sorted split
                want.sort(key=lambda t: t[0])
            assert len(s) == len(want), (name, stretch, len(s), len(want))
            assert list(s) == [w[0] for w in want] and list(e) == [w[1] for w in want]
            assert np.array_equal(v, np.array([w[2] for w in want], dtype=np.float64))


def test_last_base_quirk(tmp_path):
    """(length-1) % 10000 == 0: the reference never queries the last stretch start (Q9)."""
    chroms = [("chrA", 10001)]
    items = [(0, 9999, 1.0), (9999, 10001, 2.0)]
    write_bigwig(tmp_path / "q.bw", chroms, [(0, 1, items, 0, 0)])
    s, e, v = BigWig(tmp_path / "q.bw").decode(0)
    want = reference_stretch_records(items, 10001)
    assert list(zip(s, e, v)) == want == [(1, 10000, 1.0), (10000, 10001, 2.0)]
