"""TEST INFRASTRUCTURE: an independent, pure-Python BigWig (v4) reader -- struct + zlib only, written
from the UCSC bigWig layout, sharing no code with the product's decoder (csrc/bigwig.cpp).  It is the
second opinion on header parsing, the chromosome B+ tree, the R-tree walk (any depth), section
decoding and inflate: tests compare the product against it on the reference's UCSC-written fixtures
and on multi-section, multi-level-R-tree files."""
import struct
import zlib


def read_bigwig(path):
    """{chrom name: (length, [(start0, end0, value), ...])} with 0-based half-open coordinates, in file order."""
    data = open(path, "rb").read()
    magic, version, zooms, chrom_tree, full_data, full_index = struct.unpack_from("<IHHQQQ", data, 0)
    assert magic == 0x888FFC26, "not a bigWig"
    uncompress_buf, = struct.unpack_from("<I", data, 52)
    # chromosome B+ tree
    tmagic, block, key_size, val_size, count, _ = struct.unpack_from("<IIIIQQ", data, chrom_tree)
    assert tmagic == 0x78CA8C91 and val_size == 8
    chroms = {}

    def walk_tree(off):
        leaf, _, n = struct.unpack_from("<BBH", data, off)
        off += 4
        for _ in range(n):
            key = data[off:off + key_size].split(b"\0")[0].decode()
            if leaf:
                cid, length = struct.unpack_from("<II", data, off + key_size)
                chroms[cid] = (key, length)
            else:
                child, = struct.unpack_from("<Q", data, off + key_size)
                walk_tree(child)
            off += key_size + 8

    walk_tree(chrom_tree + 32)
    # R-tree: every leaf item
    rmagic, = struct.unpack_from("<I", data, full_index)
    assert rmagic == 0x2468ACE0
    blocks = []

    def walk_rtree(off):
        leaf, _, n = struct.unpack_from("<BBH", data, off)
        off += 4
        for _ in range(n):
            if leaf:
                c0, s0, c1, e1, doff, dsize = struct.unpack_from("<IIIIQQ", data, off)
                blocks.append((doff, dsize))
                off += 32
            else:
                c0, s0, c1, e1, child = struct.unpack_from("<IIIIQ", data, off)
                walk_rtree(child)
                off += 24

    walk_rtree(full_index + 48)
    out = {cid: [] for cid in chroms}
    for doff, dsize in sorted(blocks):
        raw = data[doff:doff + dsize]
        if uncompress_buf:
            raw = zlib.decompress(raw)
        cid, cstart, cend, step, span, kind, _, n = struct.unpack_from("<IIIIIBBH", raw, 0)
        p = 24
        for k in range(n):
            if kind == 1:
                s, e, v = struct.unpack_from("<IIf", raw, p); p += 12
            elif kind == 2:
                s, v = struct.unpack_from("<If", raw, p); p += 8
                e = s + span
            elif kind == 3:
                v, = struct.unpack_from("<f", raw, p); p += 4
                s = cstart + k * step
                e = s + span
            else:
                raise ValueError("unknown section type %d" % kind)
            out[cid].append((s, e, v))
    return {chroms[cid][0]: (chroms[cid][1], items) for cid, items in out.items()}
