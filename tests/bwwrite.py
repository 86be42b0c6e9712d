"""TEST INFRASTRUCTURE: a minimal BigWig (v4) writer, so that the decoder can be
exercised on multi-chromosome, multi-block files of all three section types.
Layout follows the UCSC bigWig specification (header, chromosome B+ tree, data
sections, R-tree index); zoom levels are omitted (readers of the full-resolution
data, like the reference's, do not need them)."""
import struct
import zlib


def _section(chrom_id, kind, items, span=1, step=1):
    """kind 1: (start,end,value) bedGraph; 2: (start,value) variableStep; 3: fixedStep values from items[0][0]."""
    if kind == 1 and hasattr(items, "dtype"):        # numpy structured array (start u4, end u4, value f4): large files
        cs, ce = int(items["s"][0]), int(items["e"][-1])
        body = items.tobytes()
    elif kind == 1:
        cs, ce = items[0][0], items[-1][1]
        body = b"".join(struct.pack("<IIf", s, e, v) for s, e, v in items)
    elif kind == 2:
        cs, ce = items[0][0], items[-1][0] + span
        body = b"".join(struct.pack("<If", s, v) for s, v in items)
    else:
        cs = items[0][0]
        ce = cs + (len(items) - 1) * step + span
        body = b"".join(struct.pack("<f", v) for _, v in items)
    head = struct.pack("<IIIIIBBH", chrom_id, cs, ce, step if kind == 3 else 0, span if kind != 1 else 0, kind, 0, len(items))
    return head + body, cs, ce


def write_bigwig(path, chroms, sections, compress=True, rtree_block=4, rtree_levels=2):
    """chroms: list of (name, length) -- ids are positions in this list;
    sections: list of (chrom_id, kind, items, span, step), sorted by (chrom_id, start).
    rtree_levels=3 puts a layer of inner nodes between the root and the leaf nodes."""
    key_size = max(len(n) for n, _ in chroms) + 1
    out = bytearray(64)
    # chromosome B+ tree: one leaf node
    chrom_tree = len(out)
    out += struct.pack("<IIIIQQ", 0x78CA8C91, len(chroms), key_size, 8, len(chroms), 0)
    out += struct.pack("<BBH", 1, 0, len(chroms))
    for cid, (name, length) in sorted(enumerate(chroms), key=lambda t: t[1][0]):
        out += name.encode().ljust(key_size, b"\0") + struct.pack("<II", cid, length)
    # data
    full_data = len(out)
    out += struct.pack("<Q", len(sections))
    leaves = []
    max_raw = 0
    for cid, kind, items, span, step in sections:
        raw, cs, ce = _section(cid, kind, items, span, step)
        max_raw = max(max_raw, len(raw))
        blob = zlib.compress(raw) if compress else raw
        leaves.append((cid, cs, cid, ce, len(out), len(blob)))
        out += blob
    # R-tree: leaf nodes of rtree_block items under one root
    full_index = len(out)
    header_at = len(out)
    out += bytes(48)
    groups = [leaves[i:i + rtree_block] for i in range(0, len(leaves), rtree_block)] or [[]]
    if len(groups) == 1:
        out += struct.pack("<BBH", 1, 0, len(groups[0]))
        for l in groups[0]:
            out += struct.pack("<IIIIQQ", *l)
    elif rtree_levels >= 3:
        # root -> inner nodes (rtree_block leaf nodes each) -> leaf nodes
        supers = [groups[i:i + rtree_block] for i in range(0, len(groups), rtree_block)]
        root_at = len(out)
        out += struct.pack("<BBH", 0, 0, len(supers)) + bytes(24 * len(supers))
        for si, sg in enumerate(supers):
            inner_at = len(out)
            out += struct.pack("<BBH", 0, 0, len(sg)) + bytes(24 * len(sg))
            for gi, g in enumerate(sg):
                child = len(out)
                out += struct.pack("<BBH", 1, 0, len(g))
                for l in g:
                    out += struct.pack("<IIIIQQ", *l)
                struct.pack_into("<IIIIQ", out, inner_at + 4 + 24 * gi, g[0][0], g[0][1], g[-1][2], max(x[3] for x in g), child)
            flat = [l for g in sg for l in g]
            struct.pack_into("<IIIIQ", out, root_at + 4 + 24 * si, flat[0][0], flat[0][1], flat[-1][2], max(x[3] for x in flat), inner_at)
    else:
        root_at = len(out)
        out += struct.pack("<BBH", 0, 0, len(groups)) + bytes(24 * len(groups))
        for gi, g in enumerate(groups):
            child = len(out)
            out += struct.pack("<BBH", 1, 0, len(g))
            for l in g:
                out += struct.pack("<IIIIQQ", *l)
            struct.pack_into("<IIIIQ", out, root_at + 4 + 24 * gi, g[0][0], g[0][1], g[-1][2], max(x[3] for x in g), child)
    lo = leaves[0] if leaves else (0, 0, 0, 0)
    hi = leaves[-1] if leaves else (0, 0, 0, 0)
    struct.pack_into("<IIQIIIIQII", out, header_at, 0x2468ACE0, rtree_block, len(leaves), lo[0], lo[1], hi[2], hi[3], len(out), 1, 0)
    struct.pack_into("<IHHQQQHHQQIQ", out, 0, 0x888FFC26, 4, 0, chrom_tree, full_data, full_index, 0, 0, 0, 0,
                     max_raw if compress else 0, 0)
    with open(path, "wb") as f:
        f.write(out)
