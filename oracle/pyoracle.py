"""TEST INFRASTRUCTURE ONLY: ctypes access to the CPU oracle.

Two back ends with the same call signatures:
  * ``Oracle("port")``      -> oracle/liboracle.so      (C restatement, wt_oracle.c)
  * ``Oracle("reference")`` -> oracle/_ref/libwtref.so  (the unmodified reference
    hot path compiled from /root/reference/src by oracle/Makefile)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module; the product package
(wiggletools_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

OPS = {
    "sum": 0, "mult": 1, "product": 1, "mean": 2, "var": 3, "stddev": 4, "entropy": 5,
    "CV": 6, "cv": 6, "min": 7, "max": 8, "median": 9, "fillIn": 10, "trimFill": 11, "ttest": 16, "wilcoxon": 17, "mwu": 17,
}


class CTrack(C.Structure):
    _fields_ = [("chrom", C.c_void_p), ("start", C.c_void_p), ("end", C.c_void_p),
                ("value", C.c_void_p), ("n", C.c_int64), ("default_value", C.c_double)]


class COut(C.Structure):
    _fields_ = [("chrom", C.c_void_p), ("start", C.c_void_p), ("end", C.c_void_p),
                ("value", C.c_void_p), ("cap", C.c_int64)]


class Track:
    """One genome track: parallel arrays sorted by (chrom index, start)."""

    def __init__(self, chrom, start, end, value, default=0.0):
        self.chrom = np.ascontiguousarray(chrom, dtype=np.int32)
        self.start = np.ascontiguousarray(start, dtype=np.int32)
        self.end = np.ascontiguousarray(end, dtype=np.int32)
        self.value = np.ascontiguousarray(value, dtype=np.float64)
        self.default = float(default)
        assert len(self.chrom) == len(self.start) == len(self.end) == len(self.value)

    def __len__(self):
        return len(self.start)

    def c(self):
        return CTrack(self.chrom.ctypes.data, self.start.ctypes.data, self.end.ctypes.data,
                      self.value.ctypes.data, len(self), self.default)


def build(ref=None):
    """(Re)build liboracle.so, and _ref/ when the reference tree is present."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref is None:
        ref = os.path.isdir("/root/reference/src")
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def reference_read_bigwig(path, names, cap=4_000_000):
    """Records of a BigWig file as the reference's OWN reader emits them: src/bigWiggleReader.c,
    unmodified, compiled into oracle/_ref/libwtrefbw.so over a libBigWig-interface shim
    (oracle/bigwig_shim.c).  Chromosome indices refer to `names`."""
    path_so = os.path.join(HERE, "_ref", "libwtrefbw.so")
    if not os.path.exists(path_so):
        raise FileNotFoundError(path_so + " (run `make -C oracle refbw`)")
    lib = C.CDLL(path_so)
    lib.wtref_read_bigwig.restype = C.c_int64
    lib.wtref_read_bigwig.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p]
    chrom = np.empty(cap, np.int32); start = np.empty(cap, np.int32); end = np.empty(cap, np.int32); value = np.empty(cap, np.float64)
    o = COut(chrom.ctypes.data, start.ctypes.data, end.ctypes.data, value.ctypes.data, cap)
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    n = lib.wtref_read_bigwig(str(path).encode(), len(names), C.addressof(arr), C.addressof(o))
    assert n <= cap
    return Track(chrom[:n].copy(), start[:n].copy(), end[:n].copy(), value[:n].copy())


class Segments:
    def __init__(self, chrom, start, end, value, default=None):
        self.chrom, self.start, self.end, self.value, self.default = chrom, start, end, value, default

    def __len__(self):
        return len(self.start)


class Oracle:
    def __init__(self, kind="port"):
        self.kind = kind
        if kind == "port":
            path, self.p = os.path.join(HERE, "liboracle.so"), "wto_"
        elif kind == "reference":
            path, self.p = os.path.join(HERE, "_ref", "libwtref.so"), "wtref_"
        elif kind == "dropin":
            # ref_harness.c bound to libwiggletools_b200.so's drop-in API: the
            # thing under test driven through the reference's own call sequence
            path, self.p = os.path.join(HERE, "_ref", "libwtdropin.so"), "wtref_"
        else:
            raise ValueError(kind)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle` / `make -C oracle ref`)")
        self.lib = C.CDLL(path)
        f = self._f("reduce")
        f.restype = C.c_int64
        f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        f = self._f("setcompare")
        f.restype = C.c_int64
        f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        f = self._f("auc")
        f.restype = C.c_double
        f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        f = self._f("pearson")
        f.restype = C.c_double
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        f = self._f("ndpearson")
        f.restype = C.c_double
        f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        f = self._f("rows")
        f.restype = C.c_int64
        f.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        if kind != "dropin":
            f = self._f("tdist_Q")
            f.restype = C.c_double
            f.argtypes = [C.c_double, C.c_double]
            f = self._f("map")
            f.restype = C.c_int64
            if kind == "port":
                f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            else:
                f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
            f = self._f("apply")
            f.restype = C.c_int64
            f = self._f("write")
            f.restype = C.c_int64
            f.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_char_p]

    def _f(self, name):
        return getattr(self.lib, self.p + name)

    @staticmethod
    def _names(names):
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        return arr

    @staticmethod
    def _tracks(tracks):
        arr = (CTrack * len(tracks))(*[t.c() for t in tracks])
        return arr

    @staticmethod
    def _out(cap):
        chrom = np.empty(cap, np.int32)
        start = np.empty(cap, np.int32)
        end = np.empty(cap, np.int32)
        value = np.empty(cap, np.float64)
        o = COut(chrom.ctypes.data, start.ctypes.data, end.ctypes.data, value.ctypes.data, cap)
        return o, (chrom, start, end, value)

    @staticmethod
    def _cap(tracks):
        return 2 * sum(len(t) for t in tracks) + 16

    def reduce(self, op, tracks, names=("chr1",), strict=False):
        cap = self._cap(tracks)
        o, arrs = self._out(cap)
        ct, cn = self._tracks(tracks), self._names(names)
        dflt = C.c_double(0)
        n = self._f("reduce")(OPS[op], len(tracks), C.addressof(ct), int(strict), len(names),
                              C.addressof(cn), C.addressof(o), C.addressof(dflt))
        assert n <= cap
        return Segments(*[a[:n].copy() for a in arrs], default=dflt.value)

    def setcompare(self, op, set1, set2, names=("chr1",)):
        cap = self._cap(list(set1) + list(set2))
        o, arrs = self._out(cap)
        c1, c2, cn = self._tracks(set1), self._tracks(set2), self._names(names)
        n = self._f("setcompare")(OPS[op], len(set1), C.addressof(c1), len(set2), C.addressof(c2),
                                  len(names), C.addressof(cn), C.addressof(o))
        assert n <= cap
        return Segments(*[a[:n].copy() for a in arrs])

    def auc(self, op, tracks, names=("chr1",), strict=False):
        ct, cn = self._tracks(tracks), self._names(names)
        return self._f("auc")(OPS[op], len(tracks), C.addressof(ct), int(strict), len(names), C.addressof(cn))

    def pearson(self, x, y, names=("chr1",)):
        cx, cy, cn = x.c(), y.c(), self._names(names)
        return self._f("pearson")(C.addressof(cx), C.addressof(cy), len(names), C.addressof(cn))

    def ndpearson(self, set1, set2, names=("chr1",)):
        c1, c2, cn = self._tracks(set1), self._tracks(set2), self._names(names)
        return self._f("ndpearson")(len(set1), C.addressof(c1), C.addressof(c2), len(names), C.addressof(cn))

    def rows(self, tracks, names=("chr1",), strict=False):
        cap = self._cap(tracks)
        o, arrs = self._out(cap)
        vals = np.empty((cap, len(tracks)), np.float64)
        inplay = np.empty((cap, len(tracks)), np.int8)
        ct, cn = self._tracks(tracks), self._names(names)
        n = self._f("rows")(len(tracks), C.addressof(ct), int(strict), len(names), C.addressof(cn),
                            C.addressof(o), vals.ctypes.data, inplay.ctypes.data)
        assert n <= cap
        seg = Segments(*[a[:n].copy() for a in arrs])
        return seg, vals[:n].copy(), inplay[:n].copy()

    def tdist_Q(self, t, nu):
        return self._f("tdist_Q")(t, nu)

    MAP_CODES = {"scale": 1, "offset": 2, "log": 3, "ln": 3, "exp": 4, "pow": 5, "abs": 6, "default": 7}

    def map(self, track, chain, names=("chr1",)):
        """A chain of unary operators over one track, innermost first:
        [("scale", 2.0), ("ln",), ("log", 10.0), ("pow", 0.5), ("abs",), ("offset", 1), ("exp",), ("default", 3)]."""
        codes = np.array([self.MAP_CODES[c[0]] for c in chain], np.int32)
        args = np.array([0.0 if c[0] in ("ln", "exp", "abs") else float(c[1]) for c in chain], np.float64)
        cap = len(track) + 1
        o, arrs = self._out(cap)
        ct = track.c()
        dflt = C.c_double(0)
        if self.kind == "port":
            n = self._f("map")(C.addressof(ct), len(chain), codes.ctypes.data, args.ctypes.data, C.addressof(o), C.addressof(dflt))
        else:
            cn = self._names(names)
            n = self._f("map")(C.addressof(ct), len(chain), codes.ctypes.data, args.ctypes.data, len(names), C.addressof(cn),
                               C.addressof(o), C.addressof(dflt))
        return Track(arrs[0][:n].copy(), arrs[1][:n].copy(), arrs[2][:n].copy(), arrs[3][:n].copy(), default=dflt.value)

    STAT_CODES = {"AUC": 1, "meanI": 2, "varI": 3, "stddevI": 4, "CVI": 5, "maxI": 6, "minI": 7, "spanI": 8}

    def apply(self, dataset, regions, stats, names=("chr1",), strict=False):
        """`apply` of a list of statistics over the regions of a sorted BED-like track:
        array [len(regions), len(stats)]."""
        codes = np.array([self.STAT_CODES[s] for s in stats], np.int32)
        out = np.full((len(regions), len(stats)), np.nan)
        if self.kind == "port":
            n = self._f("apply")(C.c_int64(len(dataset)), C.c_void_p(dataset.chrom.ctypes.data), C.c_void_p(dataset.start.ctypes.data),
                                 C.c_void_p(dataset.end.ctypes.data), C.c_void_p(dataset.value.ctypes.data), C.c_double(dataset.default),
                                 C.c_int64(len(regions)), C.c_void_p(regions.chrom.ctypes.data), C.c_void_p(regions.start.ctypes.data),
                                 C.c_void_p(regions.end.ctypes.data), C.c_int(len(stats)), C.c_void_p(codes.ctypes.data),
                                 C.c_int(int(strict)), C.c_void_p(out.ctypes.data))
        else:
            cd, cr, cn = dataset.c(), regions.c(), self._names(names)
            n = self._f("apply")(C.c_void_p(C.addressof(cd)), C.c_void_p(C.addressof(cr)), C.c_int(len(stats)), C.c_void_p(codes.ctypes.data),
                                 C.c_int(int(strict)), C.c_int(len(names)), C.c_void_p(C.addressof(cn)), C.c_void_p(out.ctypes.data))
        assert n == len(regions), (n, len(regions))
        return out

    def write(self, chrom, start, end, value, names=("chr1",), bedgraph=False):
        """Bytes the default sink prints for the segment list (`write` / `write_bg`)."""
        import tempfile
        chrom = np.ascontiguousarray(chrom, np.int32); start = np.ascontiguousarray(start, np.int32)
        end = np.ascontiguousarray(end, np.int32); value = np.ascontiguousarray(value, np.float64)
        cn = self._names(names)
        with tempfile.NamedTemporaryFile(suffix=".wig") as tmp:
            rc = self._f("write")(len(start), chrom.ctypes.data, start.ctypes.data, end.ctypes.data, value.ctypes.data,
                                  len(names), C.addressof(cn), int(bedgraph), tmp.name.encode())
            assert rc >= 0
            with open(tmp.name, "rb") as f:
                return f.read()
