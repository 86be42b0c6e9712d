#!/bin/bash
# Development: A/B the library variants under build/variants on one GPU box (kernel-resident timing only).
#   profiles/kexp.sh [scale] -> gpurun_out/kexp.jsonl (one bench line per variant, tagged)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export WB200_BENCH_SCALE=${1:-0.25} WB200_BENCH_CHECKSUM=1 WB200_BENCH_EXTRAS=0
out=${KEXP_OUT:-gpurun_out/kexp.jsonl}
: > $out
run() {
  tag=$1; shift
  line=$(env "$@" python bench.py --no-e2e --no-cpu --steps 5 --warmup 3 2>gpurun_out/kexp_err.log | tail -1)
  echo "{\"tag\": \"$tag\", \"line\": ${line:-null}}" >> $out
  echo "$tag $(echo "$line" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d["roofline"]; print("ms/step %.3f emit %.3f cursor %.3f frac %.4f S %d checksum %s slow %d" % (d["ms_per_step"], r["emit_ms_per_step"], r["cursor_ms_per_step"], r["frac"], d["segments"], d.get("checksum"), d["slow_segments"]), {k: round(v) for k, v in d.get("phase_cycles_per_window", {}).items()})
except Exception as e: print("FAILED", e)')"
}
if [ -n "$KEXP_EXTRA" ]; then eval "$KEXP_EXTRA"; fi
for lib in build/variants/*.so; do
  run $(basename $lib .so) WB200_LIB=$PWD/$lib
done
