#!/bin/bash
# Round evidence, one GPU box: bench lines, ncu launch list, ncu --set full of the dominant kernels.
# Outputs land in gpurun_out/ (scratch); the summaries are copied into profiles/ by hand.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r01c}
# 1. the line the driver records, and the reference arm
python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_full.json 2> gpurun_out/${tag}_bench_full.err
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err
# 2. other track counts (resident only)
WB200_BENCH_TRACKS=8 python bench.py --no-e2e --no-cpu --steps 5 --warmup 3 > gpurun_out/${tag}_bench_n8.json 2>/dev/null
WB200_BENCH_TRACKS=512 WB200_BENCH_SCALE=0.25 python bench.py --no-e2e --no-cpu --steps 5 --warmup 3 > gpurun_out/${tag}_bench_n512_quarter.json 2>/dev/null
# 3. launch list of one quarter-scale step (cold-cache, serialised: compare shares)
WB200_BENCH_SCALE=0.25 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'fastsum|cursor|sample_maxabs|reduce_auc|fixup|emit_kernel' -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --no-e2e --no-cpu --steps 2 --warmup 1 > /dev/null 2>&1
# 4. full capture of the two kernels of the path (one launch each, quarter-scale chromosome)
WB200_BENCH_SCALE=0.25 ncu --set full --clock-control none --import-source on -k regex:'fastsum_kernel|cursor_kernel' -s 60 -c 2 \
  -o gpurun_out/${tag}_prof python bench.py --no-e2e --no-cpu --steps 1 --warmup 1 > /dev/null 2>&1
ls -la gpurun_out | tail -12
