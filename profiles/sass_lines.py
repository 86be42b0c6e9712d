#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS dump with `nvdisasm -g -c` line info
(by instruction order) and print per-source-line instruction / stall-sample
totals.  usage: sass_lines.py <ncu_sass.csv> <nvdisasm.sass> <kernel-substr> [topN]"""
import csv, re, sys
ncu_csv, sass, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# nvdisasm: find function section for kernel, collect (line) per instruction in order
lines = open(sass).read().split("\n")
in_fn = False; cur = None; seq = []
for ln in lines:
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
        in_fn = kern in ln
        cur = None
        continue
    if not in_fn: continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1), int(m.group(2)), "inlined" in m.group(3))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        seq.append(cur)
allrows = list(csv.reader(open(ncu_csv)))
starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"] + [len(allrows)]
rows = None
want = re.sub(r"ILi\d+.*", "", kern)
for a, b in zip(starts[:-1], starts[1:]):
    if want.replace("_kernel", "") in allrows[a][1]:
        rows = allrows[a:b]
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > 8]
print("sass instr: nvdisasm %d, ncu %d" % (len(seq), len(data)), file=sys.stderr)
agg = {}
tot_i = tot_s = 0
for k, r in enumerate(data):
    key = seq[k][:2] if k < len(seq) and seq[k] else ("?", 0)
    ie = int(r[ix["Instructions Executed"]] or 0); sm = int(r[ix["# Samples"]] or 0)
    a = agg.setdefault(key, [0, 0]); a[0] += ie; a[1] += sm
    tot_i += ie; tot_s += sm
src = {}
def text(f, n):
    try:
        if f not in src:
            import glob
            cand = glob.glob("/root/repo/wiggletools_b200/csrc/" + f)
            src[f] = open(cand[0]).read().split("\n") if cand else []
        return src[f][n - 1].strip()[:100]
    except Exception:
        return ""
print("total warp-instr %d, samples %d" % (tot_i, tot_s))
for key, (ie, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%6.2f%% inst %6.2f%% samp  %s:%d  %s" % (100.0 * ie / max(1, tot_i), 100.0 * sm / max(1, tot_s), key[0], key[1], text(*key)))
