#!/bin/bash
# Turn what profiles/collect.sh left in gpurun_out/ into the tracked summaries under profiles/ (run here, no GPU needed).
#   profiles/summarise.sh [tag-in-gpurun_out=r02f] [prefix-in-profiles=r02]
cd "$(dirname "$0")/.."
tag=${1:-r02f}; out=${2:-r02}
cp gpurun_out/${tag}_bench_full.json profiles/${out}_bench_full.json
cp gpurun_out/${tag}_bench_reference_arm.json profiles/${out}_bench_reference_arm.json
cp gpurun_out/${tag}_launches.csv profiles/${out}_launches.csv
python profiles/ncu_summary.py gpurun_out/${tag}_prof.ncu-rep > profiles/${out}_ncu_full_summary.txt
[ -f gpurun_out/${tag}_prof_median256.ncu-rep ] && python profiles/ncu_summary.py gpurun_out/${tag}_prof_median256.ncu-rep > profiles/${out}_ncu_full_summary_median256.txt
[ -f gpurun_out/${tag}_prof_var512.ncu-rep ] && python profiles/ncu_summary.py gpurun_out/${tag}_prof_var512.ncu-rep > profiles/${out}_ncu_full_summary_var512.txt
python - "$tag" "$out" <<'PY'
import csv, json, sys
tag, out = sys.argv[1:3]
rows = list(csv.reader(open('gpurun_out/%s_launches.csv' % tag)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
ix = {h: i for i, h in enumerate(rows[hi])}
agg = {}
for r in rows[hi + 1:]:
    if len(r) < len(ix) or r[ix["Metric Name"]] != "gpu__time_duration.sum": continue
    v = float(r[ix["Metric Value"]]); u = r[ix["Metric Unit"]]
    v = v / 1000 if u.startswith("n") else v * 1000 if u.startswith("m") else v
    a = agg.setdefault(r[ix["Kernel Name"]].split("(")[0].replace("void ", ""), [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
d = json.loads([l for l in open('gpurun_out/%s_bench_full.json' % tag) if l.startswith('{')][-1])
r = d["roofline"]
with open('profiles/%s_launch_shares.txt' % out, 'w') as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none, quarter-scale genome (WB200_BENCH_SCALE=0.25), 3 steps x 24 chromosomes\n")
    f.write("(per-launch times are cold-cache and serialised: compare shares, not absolutes)\n")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write("%-40s launches %3d  total %9.1f us  share %5.1f %%\n" % (k, a[0], a[1], 100 * a[1] / tot))
    e, c = r["emit_ms_per_step"], r["cursor_ms_per_step"]
    f.write("bench (full scale, CUDA events, %s_bench_full.json): sweep %.2f ms + cursor %.2f ms of %.2f ms per step = %.1f %% / %.1f %%\n"
            % (out, e, c, d["ms_per_step"], 100 * e / d["ms_per_step"], 100 * c / d["ms_per_step"]))
print(open('profiles/%s_launch_shares.txt' % out).read())
PY
# SASS evidence of the shipped library
python - "$out" <<'PY'
import re, subprocess, sys, collections
out = sys.argv[1]
sass = subprocess.run(["cuobjdump", "-sass", "wiggletools_b200/libwiggletools_b200.so"], capture_output=True, text=True).stdout
fn = None; per = collections.OrderedDict(); lines = {}
for ln in sass.split("\n"):
    m = re.search(r"Function : (\S+)", ln)
    if m: fn = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]+)", ln)
    if fn and m and "fastsum_kernel" in fn:
        c = per.setdefault(fn, collections.Counter()); c["_n"] += 1
        op = m.group(1)
        if op.startswith(("UBLKCP", "SYNCS", "ATOMS", "LDG.E.128", "CCTL.E.PF2")): c[op] += 1
        if op.startswith(("UBLKCP", "SYNCS")) and len(lines.setdefault(fn, [])) < 12: lines[fn].append(ln.split("/*", 2)[0] + ln[ln.index("/*"):].split(";")[0] + ";")
with open("profiles/%s_fastsum_sass.txt" % out, "w") as f:
    f.write("# SASS evidence for the sweep kernel (cuobjdump -sass of wiggletools_b200/libwiggletools_b200.so, fastsum.cu; sm_100a)\n"
            "# template arguments: <VLOG, LIMBS, COV16, MOM>.  Bulk-async (TMA engine) copies: UBLKCP.S.G = cp.async.bulk.shared::cluster.global\n"
            "# (the two cursor rows of a window, and the row of the window prefetched for); SYNCS.* = the mbarrier operations that signal them\n"
            "# (arrive.expect_tx / try_wait.parity).  128-bit interval loads: LDG.E.128; L2 prefetch: CCTL.E.PF2; shared atomics: ATOMS.ADD.\n\n")
    for fn, c in per.items():
        f.write("== %s\n   instructions %d; %s\n\n" % (fn, c["_n"], ", ".join("%s x%d" % (k, v) for k, v in sorted(c.items()) if k != "_n")))
    first = next(iter(lines))
    f.write("\n".join(lines[first]) + "\n")
print(open("profiles/%s_fastsum_sass.txt" % out).read()[:1500])
PY
