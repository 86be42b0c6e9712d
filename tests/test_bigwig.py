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


def mixed_file(tmp_path, compress, seed=5, chroms=(("chr2", 65000), ("chr10", 30001), ("chr1", 123457))):
    """A BigWig with all three section types, long intervals crossing stretch boundaries and several
    chromosomes (ids in list order; strcmp order differs).  Returns (path, chroms, per-chromosome items)."""
    rng = np.random.default_rng(seed)
    chroms = list(chroms)
    per = {k: [] for k in range(len(chroms))}
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
    path = tmp_path / ("t%d.bw" % seed)
    write_bigwig(path, chroms, sections, compress=compress, rtree_block=5)
    return path, chroms, per


@pytest.mark.parametrize("compress", [True, False])
def test_written_files_all_section_types(tmp_path, compress):
    path, chroms, per = mixed_file(tmp_path, compress)
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


def reference_reader():
    from oracle.pyoracle import reference_read_bigwig
    if not os.path.exists(os.path.join(HERE, "..", "oracle", "_ref", "libwtrefbw.so")):
        pytest.skip("oracle/_ref/libwtrefbw.so not built (`make -C oracle refbw`, needs /root/reference)")
    return reference_read_bigwig


@pytest.mark.parametrize("seed", [5, 6, 7])
def test_decoder_matches_the_references_own_reader(tmp_path, seed):
    """The decoder's emulation of the reference's reader (10000-bp stretch queries, boxing, visiting order,
    last-base quirk: Q9) against that reader itself: src/bigWiggleReader.c, unmodified, compiled over a
    libBigWig-interface shim that only hands out the stored records (oracle/bigwig_shim.c)."""
    read = reference_reader()
    chroms = (("chr2", 65000), ("chr10", 30001), ("chr1", 123457), ("chrM", 10001), ("chr21", 20001), ("chrX", 9999))
    path, chroms, _ = mixed_file(tmp_path, True, seed=seed, chroms=chroms)
    b = BigWig(path, stretch=10000)
    names = [c[0] for c in b.chroms]                                   # strcmp order
    want = read(path, names)
    got_c, got_s, got_e, got_v = [], [], [], []
    for i in range(len(names)):
        s, e, v = b.decode(i)
        got_c.append(np.full(len(s), i, np.int32)); got_s.append(s); got_e.append(e); got_v.append(v)
    got_c, got_s, got_e, got_v = map(np.concatenate, (got_c, got_s, got_e, got_v))
    assert len(want) == len(got_s) and len(got_s) > 1000
    assert np.array_equal(want.chrom, got_c) and np.array_equal(want.start, got_s) and np.array_equal(want.end, got_e)
    assert np.array_equal(want.value, got_v)


def test_reference_reader_on_the_reference_fixtures():
    read = reference_reader()
    for name, values, starts in (("fixedStep.bw", np.arange(10.0), np.arange(1, 11)), ("variableStep.bw", np.arange(1.0, 6.0), [1, 2, 4, 6, 8])):
        t = read(os.path.join(HERE, "golden", name), ["chr1"])
        assert list(t.start) == list(starts) and list(t.value) == list(values)


def test_config0_with_the_reference_itself(tmp_path):
    """BASELINE.json configs[0], literally, run by the REFERENCE (its own BigWig reader over the shim, its
    own multiplexer, reducers and writer): `wiggletools mean test/fixedStep.bw test/variableStep.bw`, plus
    the reference's goldens on the BigWig twins (test/expected/pearson.txt, test/test.py:95-98 style AUC)."""
    import shutil
    import subprocess
    exe = os.path.join(HERE, "..", "oracle", "_ref", "wiggletools_bw")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/wiggletools_bw not built (`make -C oracle refbw`, needs /root/reference)")
    for f in ("fixedStep.bw", "variableStep.bw"):
        shutil.copy(os.path.join(HERE, "golden", f), tmp_path / f)
    run = lambda *a: subprocess.run([exe] + list(a), capture_output=True, text=True, check=True, cwd=tmp_path).stdout
    out = run("mean", "fixedStep.bw", "variableStep.bw").split("\n")
    assert out[0] == "fixedStep chrom=chr1 start=1 step=1"
    assert [float(x) for x in out[1:11]] == [.5, 1.5, 1, 3, 2, 4.5, 3, 6, 4, 4.5]
    assert run("pearson", "fixedStep.bw", "variableStep.bw").strip() == "-0.028968"
    assert run("AUC", "mean", "fixedStep.bw", "variableStep.bw").strip() == "30.000000"


def test_corrupt_files_fail_cleanly(tmp_path):
    """Truncated files and corrupted offsets / counts / section types make the decoder report
    'corrupt BigWig ...' instead of reading past the buffer (the parser trusted every
    offset in the file; libBigWig fails cleanly on the same inputs)."""
    import struct
    path, chroms, per = mixed_file(tmp_path, compress=True)
    good = open(path, "rb").read()

    def outcome(data):
        p = tmp_path / "bad.bw"
        p.write_bytes(data)
        try:
            b = BigWig(p)
        except IOError as e:
            return str(e)
        try:
            for k in range(len(b.chroms)):
                b.decode(k)
        except IOError as e:
            return str(e)
        return "ok"

    assert outcome(good) == "ok"
    # truncations at every structural boundary and at many arbitrary lengths
    for cut in [65, 100, 200, len(good) // 3, len(good) // 2, len(good) - 200, len(good) - 1]:
        msg = outcome(good[:cut])
        assert msg == "ok" or "corrupt" in msg or "not in BigWig" in msg, (cut, msg)
    chrom_tree, = struct.unpack_from("<Q", good, 8)
    index_off, = struct.unpack_from("<Q", good, 24)
    # header offsets pointing outside the file
    for field in (8, 24):
        bad = bytearray(good)
        struct.pack_into("<Q", bad, field, len(good) + 12345)
        assert "corrupt" in outcome(bytes(bad))
        struct.pack_into("<Q", bad, field, 2 ** 63)
        assert "corrupt" in outcome(bytes(bad))
    # a chromosome-tree node claiming 65535 entries, an absurd key size
    bad = bytearray(good); struct.pack_into("<H", bad, chrom_tree + 32 + 2, 65535)
    assert "corrupt" in outcome(bytes(bad))
    bad = bytearray(good); struct.pack_into("<I", bad, chrom_tree + 8, 0x7fffffff)
    assert "corrupt" in outcome(bytes(bad))
    # R-tree root: child / block offsets outside the file, a node count that runs off the end
    bad = bytearray(good); struct.pack_into("<H", bad, index_off + 48 + 2, 65535)
    assert "corrupt" in outcome(bytes(bad))
    root_leaf = good[index_off + 48] != 0
    bad = bytearray(good); struct.pack_into("<Q", bad, index_off + 48 + 4 + 16, len(good) * 4)
    assert "corrupt" in outcome(bytes(bad))
    if root_leaf:
        bad = bytearray(good); struct.pack_into("<Q", bad, index_off + 48 + 4 + 24, 2 ** 62)     # block size
        assert "corrupt" in outcome(bytes(bad))
    # an uncompressed file with an unknown section type
    path2, _, _ = mixed_file(tmp_path, compress=False, seed=9)
    raw = bytearray(open(path2, "rb").read())
    index2, = struct.unpack_from("<Q", raw, 24)
    assert raw[index2 + 48] != 0 or True
    # find the first data section through the header's full_data_offset (8 bytes count, then sections)
    data_off, = struct.unpack_from("<Q", raw, 16)
    raw[data_off + 8 + 20] = 9
    assert "corrupt" in outcome(bytes(raw))
    # random single-byte damage must never crash the process
    rng = np.random.default_rng(1)
    for _ in range(300):
        bad = bytearray(good)
        k = int(rng.integers(0, len(bad)))
        bad[k] ^= int(rng.integers(1, 256))
        outcome(bytes(bad))


def test_decoder_against_independent_python_reader(tmp_path):
    """Second opinion (the decoder's parity was partly circular): tests/bwread.py is a
    pure-Python struct/zlib BigWig reader that shares no code with csrc/bigwig.cpp.  Both must read
    the same records from (a) the reference's own fixtures, written by the UCSC tools, and (b)
    multi-section files with two- and three-level R-trees, compressed and not."""
    from bwread import read_bigwig
    paths = [os.path.join(HERE, "golden", "fixedStep.bw"), os.path.join(HERE, "golden", "variableStep.bw")]
    for compress in (True, False):
        paths.append(mixed_file(tmp_path, compress, seed=21 + int(compress))[0])
    # a three-level R-tree: many small sections, fan-out 3
    rng = np.random.default_rng(8)
    chroms = [("chrB", 400000), ("chrA", 90000)]
    sections, pos = [], {0: 0, 1: 0}
    for cid in (0, 1):
        while pos[cid] < chroms[cid][1] - 600:
            items = []
            for _ in range(int(rng.integers(1, 9))):
                pos[cid] += int(rng.integers(0, 40)); ln = int(rng.integers(1, 50))
                items.append((pos[cid], pos[cid] + ln, float(np.float32(rng.normal()))))
                pos[cid] += ln
            sections.append((cid, 1, items, 0, 0))
    deep = tmp_path / "deep.bw"
    write_bigwig(deep, chroms, sections, compress=True, rtree_block=3, rtree_levels=3)
    paths.append(deep)
    for path in paths:
        want = read_bigwig(path)
        b = BigWig(path, stretch=0)
        assert sorted(want) == [c[0] for c in b.chroms]
        for i, (name, length) in enumerate(b.chroms):
            assert length == want[name][0]
            s, e, v = b.decode(i)
            items = want[name][1]
            assert len(s) == len(items), (path, name)
            assert list(s) == [it[0] + 1 for it in items] and list(e) == [it[1] + 1 for it in items]      # 1-based, as the reference hands them on
            assert np.array_equal(v, np.array([it[2] for it in items], dtype=np.float64))
    assert len(read_bigwig(deep)["chrB"][1]) > 3000
