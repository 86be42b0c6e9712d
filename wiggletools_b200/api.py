"""Host-side mirror (Python) of the WiggleTools operator interface for the
merge+reduce hot path, on top of the C ABI (include/wiggletools_b200.h).

Reference interface mirrored (file:line under /root/reference/src):
  newMultiplexer(iters, count, strict)      multiplexer.c:157-169 -> Multiplexer(tracks, strict)
  newMultiset(multis, 2)                    multiSet.c:115-128    -> Multiset(mux_a, mux_b)
  Sum/Product/Mean/.../MedianReduction      reducers.c            -> SumReduction(mux) ...
  TTestReduction / MWUReduction             setComparisons.c      -> TTestReduction(mset) ...
  AUCIntegrator                             statistics.c:103-127  -> AUCIntegrator(reduction)
  consumer loop while(!wi->done){...;pop}   wiggleIterator.c:57-65 -> WiggleIterator.pop()/.done

All arithmetic happens in the CUDA library; this module only slices tracks by
chromosome (the reference visits chromosomes in strcmp order, multiplexer.c:50-74)
and wraps results.  Coordinates are 1-based half-open, as inside the reference.
"""
import ctypes as C

import numpy as np

from . import _lib

OPS = {
    "sum": 0, "mult": 1, "product": 1, "mean": 2, "var": 3, "stddev": 4, "entropy": 5,
    "CV": 6, "cv": 6, "min": 7, "max": 8, "median": 9, "fillIn": 10, "trimFill": 11, "ttest": 16, "wilcoxon": 17, "mwu": 17,
}


class WB200Error(RuntimeError):
    pass


class Track:
    """One genome track: parallel arrays sorted by chromosome (strcmp order of the
    names they index) then start; intervals do not overlap within a track."""

    def __init__(self, chrom, start, end, value, default=0.0):
        self.chrom = np.ascontiguousarray(chrom, dtype=np.int32)
        self.start = np.ascontiguousarray(start, dtype=np.int32)
        self.end = np.ascontiguousarray(end, dtype=np.int32)
        self.value = np.ascontiguousarray(value, dtype=np.float64)
        self.default = float(default)
        n = len(self.start)
        if not (len(self.chrom) == len(self.end) == len(self.value) == n):
            raise ValueError("track arrays differ in length")
        # chromosome runs: {chrom index: (lo, hi)}
        self.runs = {}
        if n:
            cut = np.flatnonzero(np.diff(self.chrom)) + 1
            lo = np.concatenate(([0], cut))
            hi = np.concatenate((cut, [n]))
            for a, b in zip(lo, hi):
                c = int(self.chrom[a])
                if c in self.runs:
                    raise ValueError("track is not sorted by chromosome")
                self.runs[c] = (int(a), int(b))

    def __len__(self):
        return len(self.start)


class Segments:
    """Result records: chrom index, start, end (exclusive), value."""

    def __init__(self, chrom, start, end, value, default=float("nan"), auc=0.0):
        self.chrom, self.start, self.end, self.value = chrom, start, end, value
        self.default, self.auc = default, auc

    def __len__(self):
        return len(self.start)


class Engine:
    """One CUDA context of the engine (one per process/GPU)."""

    def __init__(self, device=0, window=0):
        self.lib = _lib.load()
        self.ctx = self.lib.wb200_create(device)
        if not self.ctx:
            raise WB200Error(self.lib.wb200_last_error(None).decode())
        if window:
            self._check(self.lib.wb200_set_window(self.ctx, window))
        self.last_stats = []

    def close(self):
        if self.ctx:
            self.lib.wb200_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise WB200Error("wb200 error %d: %s" % (rc, self.lib.wb200_last_error(self.ctx).decode()))

    def stats(self):
        s = self.lib.wb200_last_stats(self.ctx).contents
        return {k: getattr(s, k) for k, _ in s._fields_}

    # -- element-wise maps (unaryOps.c; `map` in the CLI grammar) ---------------
    MAP_CODES = {"scale": 1, "offset": 2, "log": 3, "ln": 3, "exp": 4, "pow": 5, "abs": 6}

    def _lower_chain(self, chain, default):
        """CLI-level chain [("scale", 2), ("ln",), ("log", 10), ("default", 1) ...] -> (C maps,
        mapped default): arguments become what the reference's constructors store
        (log base -> log(base), ln / exp -> 1) and the default follows their rules."""
        import math
        maps = []
        for c in chain:
            if c[0] == "default":                      # DefaultValueWiggleIterator: records untouched
                default = float(c[1])
                continue
            code = self.MAP_CODES[c[0]]
            arg = 1.0 if c[0] in ("ln", "exp") else 0.0 if c[0] == "abs" else math.log(float(c[1])) if c[0] == "log" else float(c[1])
            default = self.lib.wb200_map_default(code, arg, default)
            maps.append(_lib.CMap(code, 0, arg))
        return maps, default

    # -- one chromosome, host arrays ---------------------------------------
    def reduce_chrom(self, op, tracks, nset1=None, strict=False, maps=None):
        """tracks: list of (start, end, value, default) numpy arrays of ONE chromosome;
        maps: optional list (one per track) of unary-operator chains applied on the device."""
        n = len(tracks)
        arr = (_lib.CTrack * n)()
        keep = []
        lowered = [self._lower_chain(maps[i], float(tracks[i][3])) for i in range(n)] if maps else None
        for i, (s, e, v, d) in enumerate(tracks):
            s = np.ascontiguousarray(s, np.int32); e = np.ascontiguousarray(e, np.int32)
            v = np.ascontiguousarray(v, np.float64)
            keep.append((s, e, v))
            arr[i] = _lib.CTrack(s.ctypes.data, e.ctypes.data, v.ctypes.data, len(s), float(lowered[i][1] if lowered else d))
        if lowered:
            chains = [(_lib.CMap * max(1, len(m)))(*m) for m, _ in lowered]
            ptrs = (C.c_void_p * n)(*[C.addressof(c) for c in chains])
            counts = (C.c_int * n)(*[len(m) for m, _ in lowered])
            self._check(self.lib.wb200_set_input_maps(self.ctx, n, C.addressof(ptrs), C.addressof(counts)))
        out = _lib.CSegments()
        code = OPS[op] if isinstance(op, str) else int(op)
        self._check(self.lib.wb200_reduce_host(self.ctx, code, n, n if nset1 is None else nset1,
                                               int(bool(strict)), C.addressof(arr), C.addressof(out)))
        m = out.n
        if m:
            start = np.ctypeslib.as_array(C.cast(out.start, C.POINTER(C.c_int32)), (m,)).copy()
            end = np.ctypeslib.as_array(C.cast(out.end, C.POINTER(C.c_int32)), (m,)).copy()
            value = np.ctypeslib.as_array(C.cast(out.value, C.POINTER(C.c_double)), (m,)).copy()
        else:
            start = np.empty(0, np.int32); end = np.empty(0, np.int32); value = np.empty(0, np.float64)
        self.last_stats.append(self.stats())
        return start, end, value, out.default_value, out.auc

    # -- several chromosomes, pipelined ----------------------------------------
    def reduce_shards(self, op, shards, nset1=None, strict=False):
        """shards: list (one per chromosome) of lists of (start, end, value, default).
        One wb200_reduce_host_batch call: staging, kernels and read-back of
        consecutive chromosomes overlap."""
        nsh, n = len(shards), len(shards[0])
        arr = (_lib.CTrack * (nsh * n))()
        keep = []
        for c, per in enumerate(shards):
            for i, (s, e, v, d) in enumerate(per):
                s = np.ascontiguousarray(s, np.int32); e = np.ascontiguousarray(e, np.int32)
                v = np.ascontiguousarray(v, np.float64)
                keep.append((s, e, v))
                arr[c * n + i] = _lib.CTrack(s.ctypes.data, e.ctypes.data, v.ctypes.data, len(s), float(d))
        out = (_lib.CSegments * nsh)()
        code = OPS[op] if isinstance(op, str) else int(op)
        self._check(self.lib.wb200_reduce_host_batch(self.ctx, code, nsh, n, n if nset1 is None else nset1,
                                                     int(bool(strict)), C.addressof(arr), C.addressof(out)))
        res = []
        for c in range(nsh):
            m = out[c].n
            if m:
                start = np.ctypeslib.as_array(C.cast(out[c].start, C.POINTER(C.c_int32)), (m,)).copy()
                end = np.ctypeslib.as_array(C.cast(out[c].end, C.POINTER(C.c_int32)), (m,)).copy()
                value = np.ctypeslib.as_array(C.cast(out[c].value, C.POINTER(C.c_double)), (m,)).copy()
            else:
                start = np.empty(0, np.int32); end = np.empty(0, np.int32); value = np.empty(0, np.float64)
            res.append((start, end, value, out[c].default_value, out[c].auc))
        self.last_stats = [self.stats()]
        return res

    # -- whole genome -------------------------------------------------------
    def reduce(self, op, tracks, names=("chr1",), strict=False, nset1=None, maps=None):
        """<op> over N genome tracks: segments of every chromosome, in the
        reference's visiting order (ascending strcmp of the names).  maps: optional
        unary-operator chain per track (`<op> map ln a b c`), applied on the device."""
        self.last_stats = []
        present = sorted({c for t in tracks for c in t.runs}, key=lambda c: names[c].encode())
        if maps is not None:
            parts, auc, dflt = [], 0.0, float("nan")
            for c in present or [None]:
                per = []
                for t in tracks:
                    lo, hi = t.runs.get(c, (0, 0)) if c is not None else (0, 0)
                    per.append((t.start[lo:hi], t.end[lo:hi], t.value[lo:hi], t.default))
                s_, e_, v_, dflt, a_ = self.reduce_chrom(op, per, nset1=nset1, strict=strict, maps=maps)
                parts.append((np.full(len(s_), 0 if c is None else c, np.int32), s_, e_, v_))
                auc += a_
            return Segments(*[np.concatenate([p[k] for p in parts]) for k in range(4)], dflt, auc)
        if not present:
            # defaults only: constructor semantics still apply
            per = [(np.empty(0, np.int32), np.empty(0, np.int32), np.empty(0), t.default) for t in tracks]
            _, _, _, dflt, _ = self.reduce_chrom(op, per, nset1=nset1, strict=strict)
            return Segments(np.empty(0, np.int32), np.empty(0, np.int32), np.empty(0, np.int32), np.empty(0), dflt, 0.0)
        shards = []
        for c in present:
            per = []
            for t in tracks:
                lo, hi = t.runs.get(c, (0, 0))
                per.append((t.start[lo:hi], t.end[lo:hi], t.value[lo:hi], t.default))
            shards.append(per)
        res = self.reduce_shards(op, shards, nset1=nset1, strict=strict)
        cs = [np.full(len(r[0]), c, np.int32) for c, r in zip(present, res)]
        auc = 0.0
        for r in res:
            auc += r[4]   # chromosome partials are folded in visiting order
        return Segments(np.concatenate(cs), np.concatenate([r[0] for r in res]), np.concatenate([r[1] for r in res]),
                        np.concatenate([r[2] for r in res]), res[-1][3], auc)

    def setcompare(self, op, set1, set2, names=("chr1",)):
        return self.reduce(op, list(set1) + list(set2), names, nset1=len(set1))

    def pearson(self, x, y, names=("chr1",)):
        """Pearson correlation of two tracks over the genome (PearsonIntegrator,
        src/statistics.c:414-465): per-chromosome moments from the GPU, folded in
        visiting order with the pairwise form of the reference's update."""
        present = sorted(set(x.runs) | set(y.runs), key=lambda c: names[c].encode())
        total = (C.c_double * 6)(0, 0, 0, 0, 0, 0)
        for c in present:
            keep, ct = [], []
            for t in (x, y):
                lo, hi = t.runs.get(c, (0, 0))
                s, e, v = t.start[lo:hi], t.end[lo:hi], t.value[lo:hi]
                keep.append((s, e, v))
                ct.append(_lib.CTrack(s.ctypes.data, e.ctypes.data, v.ctypes.data, len(s), t.default))
            m = (C.c_double * 6)()
            self._check(self.lib.wb200_pearson_host(self.ctx, C.addressof(ct[0]), C.addressof(ct[1]), C.addressof(m)))
            self.lib.wb200_pearson_merge(C.addressof(total), C.addressof(m))
        return self.lib.wb200_pearson_result(C.addressof(total))

    def ndpearson(self, set1, set2, names=("chr1",)):
        """N-dimensional Pearson of two equally sized sets (NDPearsonIntegrator,
        src/statistics.c:490-556; the reference reads rank = 2 dimensions, :551)."""
        tracks = list(set1) + list(set2)
        n, n1 = len(tracks), len(set1)
        present = sorted({c for t in tracks for c in t.runs}, key=lambda c: names[c].encode())
        total = [(C.c_double * 6)(0, 0, 0, 0, 0, 0) for _ in range(2)]
        for c in present:
            keep, arr = [], (_lib.CTrack * n)()
            for i, t in enumerate(tracks):
                lo, hi = t.runs.get(c, (0, 0))
                s, e, v = t.start[lo:hi], t.end[lo:hi], t.value[lo:hi]
                keep.append((s, e, v))
                arr[i] = _lib.CTrack(s.ctypes.data, e.ctypes.data, v.ctypes.data, len(s), t.default)
            for d in range(2):
                m = (C.c_double * 6)()
                self._check(self.lib.wb200_pearson_pair_host(self.ctx, n, C.addressof(arr), d, n1 + d, C.addressof(m)))
                self.lib.wb200_pearson_merge(C.addressof(total[d]), C.addressof(m))
        txx, txy, tyy = (sum(t[k] for t in total) for k in (3, 4, 5))
        return txy / (txx * tyy) ** 0.5 if txx * tyy != 0 else float("nan")

    def auc(self, op, tracks, names=("chr1",), strict=False):
        return self.reduce(op, tracks, names, strict).auc

    # -- regional integrators (apply.c / statistics.c) ---------------------------
    STATS = {"AUC": 1, "meanI": 2, "varI": 3, "stddevI": 4, "CVI": 5, "maxI": 6, "minI": 7, "spanI": 8}

    def apply(self, dataset, regions, stats, strict=False):
        """`apply` of a list of integrators (AUC, meanI, varI, stddevI, CVI, maxI, minI, spanI;
        statistics.c:62-326) over the regions of a sorted BED-like Track, on the segments of
        `dataset` (a Track or Segments; its `default` fills the bases it does not cover unless
        strict) -- ApplyMultiplexer (apply.c:340-353), one wb200_apply_host call per chromosome.
        Returns an array [len(regions), len(stats)]."""
        codes = np.array([self.STATS[s_] for s_ in stats], np.int32)
        out = np.full((len(regions.start), len(stats)), np.nan)
        dchrom, rchrom = np.asarray(dataset.chrom), np.asarray(regions.chrom)
        default = float(getattr(dataset, "default", 0.0))
        for c in np.unique(rchrom):
            ridx = np.flatnonzero(rchrom == c)
            didx = np.flatnonzero(dchrom == c)
            s = np.ascontiguousarray(np.asarray(dataset.start)[didx], np.int32); e = np.ascontiguousarray(np.asarray(dataset.end)[didx], np.int32)
            v = np.ascontiguousarray(np.asarray(dataset.value)[didx], np.float64)
            rs = np.ascontiguousarray(np.asarray(regions.start)[ridx], np.int32); re_ = np.ascontiguousarray(np.asarray(regions.end)[ridx], np.int32)
            res = np.empty((len(ridx), len(stats)), np.float64)
            self._check(self.lib.wb200_apply_host(self.ctx, s.ctypes.data, e.ctypes.data, v.ctypes.data, len(s), default,
                                                  len(rs), rs.ctypes.data, re_.ctypes.data, len(stats), codes.ctypes.data,
                                                  int(bool(strict)), res.ctypes.data))
            out[ridx] = res
        return out

    # -- text sink (wigWriter.c: toFile / toStdout) ------------------------------
    def format_segments(self, seg, names=("chr1",), bedgraph=False):
        """Bytes the reference's default sink prints for a Segments list: `write`
        (compress + fixedStep / bedGraph lines) or `write_bg`.  One
        wb200_format_host call per chromosome, sink state carried across them."""
        state = _lib.CSinkState(0, 0, 0)
        chrom = np.asarray(seg.chrom)
        out = []
        # chromosomes in the order the records arrive
        bounds = np.flatnonzero(np.diff(chrom)) + 1 if len(chrom) else np.empty(0, np.int64)
        lo = 0
        for hi in list(bounds) + [len(chrom)]:
            if hi == lo:
                continue
            s = np.ascontiguousarray(seg.start[lo:hi], np.int32); e = np.ascontiguousarray(seg.end[lo:hi], np.int32)
            v = np.ascontiguousarray(seg.value[lo:hi], np.float64)
            text, nbytes = C.c_char_p(), C.c_int64(0)
            self._check(self.lib.wb200_format_host(self.ctx, 1 if bedgraph else 0, names[int(chrom[lo])].encode(),
                                                   s.ctypes.data, e.ctypes.data, v.ctypes.data, hi - lo,
                                                   C.addressof(state), C.addressof(text), C.addressof(nbytes)))
            out.append(C.string_at(text, nbytes.value))
            lo = hi
        return b"".join(out)


_default_engine = None


def default_engine():
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine(0)
    return _default_engine


# ---------------------------------------------------------------------------
# Mirror of the reference's object protocol
# ---------------------------------------------------------------------------
class Multiplexer:
    """newMultiplexer(iters, count, strict): multiplexer.c:157-169."""

    def __init__(self, tracks, strict=False, names=("chr1",)):
        self.tracks, self.strict, self.names = list(tracks), bool(strict), tuple(names)
        self.count = len(self.tracks)


class Multiset:
    """newMultiset(multis, count): multiSet.c:115-128 (two sets)."""

    def __init__(self, *multis):
        if len(multis) != 2:
            raise WB200Error("set comparisons take exactly two multiplexers")
        self.multis = multis
        self.count = 2


class WiggleIterator:
    """The consumer protocol of src/wiggleIterator.c:20-65: the constructor pops
    once, so the first record is readable on return; loop ``while not wi.done``."""

    def __init__(self, seg, names):
        self._seg, self._names, self._k = seg, names, 0
        self.default_value = seg.default
        self.chrom = None
        self.start = self.finish = 0
        self.value = float("nan")
        self.done = False
        self.pop()

    def pop(self):
        if self.done:
            return
        if self._k >= len(self._seg):
            self.done = True
            return
        k = self._k
        self.chrom = self._names[self._seg.chrom[k]]
        self.start = int(self._seg.start[k]); self.finish = int(self._seg.end[k])
        self.value = float(self._seg.value[k])
        self._k += 1


def _reduction(op):
    def ctor(multi, engine=None):
        eng = engine or default_engine()
        if isinstance(multi, Multiset):
            a, b = multi.multis
            seg = eng.setcompare(op, a.tracks, b.tracks, a.names)
            return WiggleIterator(seg, a.names)
        seg = eng.reduce(op, multi.tracks, multi.names, strict=multi.strict)
        return WiggleIterator(seg, multi.names)
    return ctor


SumReduction = _reduction("sum")
ProductReduction = _reduction("mult")
MeanReduction = _reduction("mean")
VarianceReduction = _reduction("var")
StdDevReduction = _reduction("stddev")
EntropyReduction = _reduction("entropy")
CVReduction = _reduction("cv")
MinReduction = _reduction("min")
MaxReduction = _reduction("max")
MedianReduction = _reduction("median")
TTestReduction = _reduction("ttest")
MWUReduction = _reduction("wilcoxon")
