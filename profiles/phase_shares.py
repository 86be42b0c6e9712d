#!/usr/bin/env python
"""Per-phase shares of executed warp-instructions and stall samples of the sweep kernel:
joins `ncu --page source --csv` with `nvdisasm -g -c` line info (instruction order) and sums by
source-line range.  Lines outside fastsum*.cu (inlined helpers) inherit the phase of the
instruction before them.
usage: phase_shares.py <ncu_source.csv> <nvdisasm.sass> <kernel-substr> <file-prefix> lo-hi:name ..."""
import csv, re, sys
ncu_csv, sass, kern, fpre = sys.argv[1:5]
phases = []
for spec in sys.argv[5:]:
    rng, name = spec.split(":")
    lo, hi = rng.split("-")
    phases.append((int(lo), int(hi), name))
lines = open(sass).read().split("\n")
in_fn = False; cur = None; seq = []
for ln in lines:
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
        in_fn = kern in ln; cur = None; continue
    if not in_fn: continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln): seq.append(cur)
allrows = list(csv.reader(open(ncu_csv)))
starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"] + [len(allrows)]
for a, b in zip(starts[:-1], starts[1:]):
    if "fastsum" in allrows[a][1]: rows = allrows[a:b]
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > 8]
assert len(seq) == len(data), (len(seq), len(data))
agg = {}; tot = [0, 0, 0]; prev = "?"
for k, r in enumerate(data):
    cur = seq[k]
    ie = int(r[ix["Instructions Executed"]] or 0); sm = int(r[ix["# Samples"]] or 0); ti = int(r[ix["Thread Instructions Executed"]] or 0)
    ph = prev
    if cur and cur[0].startswith(fpre):
        for lo, hi, name in phases:
            if lo <= cur[1] <= hi: ph = name
        prev = ph
    a = agg.setdefault(ph, [0, 0, 0]); a[0] += ie; a[1] += sm; a[2] += ti
    tot[0] += ie; tot[1] += sm; tot[2] += ti
for ph, (ie, sm, ti) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-12s inst %5.1f%%  (%9d)  samples %5.1f%%  active lanes %4.1f/32" % (ph, 100 * ie / tot[0], ie, 100 * sm / tot[1], ti / max(1, ie)))
print("total warp-instructions %d, samples %d" % (tot[0], tot[1]))
