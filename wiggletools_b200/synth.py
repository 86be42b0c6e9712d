"""Synthetic genome tracks (numpy, host side) for tests and the benchmark.

Shapes follow SURVEY.md section 8(d): interval length U[1,max_len], a gap
U[1,max_gap] before an interval with probability ``gap_prob``, values that are
multiples of 1/64 (float32-representable, so fp64 sums are order-exact).
Coordinates are 1-based half-open [start, end) like the reference's in-memory
convention (src/wigReader.c:123-124).
"""
import numpy as np

# hg38-like chromosome lengths (bp); visit order is strcmp order of the names.
HG38 = {
    "chr1": 248956422, "chr2": 242193529, "chr3": 198295559, "chr4": 190214555,
    "chr5": 181538259, "chr6": 170805979, "chr7": 159345973, "chr8": 145138636,
    "chr9": 138394717, "chr10": 133797422, "chr11": 135086622, "chr12": 133275309,
    "chr13": 114364328, "chr14": 107043718, "chr15": 101991189, "chr16": 90338345,
    "chr17": 83257441, "chr18": 80373285, "chr19": 58617616, "chr20": 64444167,
    "chr21": 46709983, "chr22": 50818468, "chrX": 156040895, "chrY": 57227415,
}


def sorted_chroms(lengths=None):
    """Chromosome names in the reference's visiting order (strcmp ascending,
    src/multiplexer.c:56,70) and their lengths."""
    lengths = HG38 if lengths is None else lengths
    names = sorted(lengths, key=lambda s: s.encode())
    return names, [lengths[n] for n in names]


def make_chrom_track(rng, length, max_len=400, max_gap=200, gap_prob=0.5, vmax=100000,
                     align=1, nan_prob=0.0, neg=False):
    """One track on one chromosome of ``length`` bp: (start, end, value) arrays."""
    mean_step = (max_len + 1) / 2 + gap_prob * (max_gap + 1) / 2
    n = int(length / (mean_step * align) * 1.05) + 16
    lens = rng.integers(1, max_len + 1, n, dtype=np.int64) * align
    gaps = np.where(rng.random(n) < gap_prob, rng.integers(1, max_gap + 1, n, dtype=np.int64), 0) * align
    start = 1 + np.cumsum(gaps + np.concatenate(([0], lens[:-1])))
    end = start + lens
    keep = end <= length + 1
    start, end = start[keep].astype(np.int32), end[keep].astype(np.int32)
    value = rng.integers(0, vmax, len(start)).astype(np.float64) / 64.0
    if neg:
        value -= vmax / 128.0
    if nan_prob > 0:
        value[rng.random(len(start)) < nan_prob] = np.nan
    return start, end, value


def make_tracks(n_tracks, chrom_lengths, seed=0, **kw):
    """List of per-track dicts {chrom,start,end,value} over several chromosomes
    (chromosome index = position in ``chrom_lengths``)."""
    tracks = []
    for t in range(n_tracks):
        rng = np.random.default_rng([seed, t])
        cs, ss, es, vs = [], [], [], []
        for ci, length in enumerate(chrom_lengths):
            s, e, v = make_chrom_track(rng, length, **kw)
            cs.append(np.full(len(s), ci, np.int32))
            ss.append(s); es.append(e); vs.append(v)
        tracks.append(dict(chrom=np.concatenate(cs), start=np.concatenate(ss),
                           end=np.concatenate(es), value=np.concatenate(vs)))
    return tracks
