#!/bin/bash
# Round evidence, one GPU box: the bench line, ncu launch list, ncu --set full of the dominant kernels, BigWig end to end.
# Outputs land in gpurun_out/ (scratch); the summaries are copied into profiles/ by hand.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r02}
# 1. the line the driver records (headline + extras, parity-gated), and the reference arm
python bench.py --steps 10 --warmup 5 > gpurun_out/${tag}_bench_full.json 2> gpurun_out/${tag}_bench_full.err
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err
# 2. launch list of one quarter-scale step (cold-cache, serialised: compare shares)
WB200_BENCH_EXTRAS=0 WB200_BENCH_SCALE=0.25 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'fastsum|cursor|sample_maxabs|reduce_auc|fixup|emit_kernel' -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --no-e2e --no-cpu --steps 2 --warmup 1 > /dev/null 2>&1
# 3. full capture of the two kernels of the path (one launch each, quarter-scale chromosome)
WB200_BENCH_EXTRAS=0 WB200_BENCH_SCALE=0.25 ncu --set full --clock-control none --import-source on -k regex:'fastsum_kernel|cursor_kernel' -s 60 -c 2 \
  -o gpurun_out/${tag}_prof python bench.py --no-e2e --no-cpu --steps 1 --warmup 1 > /dev/null 2>&1
# 3b. the same for the other two kernels a bench line reports on: the selection kernel (median over 256 tracks) and the
#     moments variant of the sweep kernel (var over 512 tracks)
WB200_BENCH_TRACKS=256 WB200_BENCH_OP=median WB200_BENCH_SCALE=0.25 ncu --set full --clock-control none --import-source on -k regex:'emit_kernel' -s 30 -c 1 \
  -o gpurun_out/${tag}_prof_median256 python bench.py --no-e2e --no-cpu --steps 1 --warmup 1 > /dev/null 2>&1
WB200_BENCH_TRACKS=512 WB200_BENCH_OP=var WB200_BENCH_SCALE=0.25 ncu --set full --clock-control none --import-source on -k regex:'fastsum_kernel' -s 30 -c 1 \
  -o gpurun_out/${tag}_prof_var512 python bench.py --no-e2e --no-cpu --steps 1 --warmup 1 > /dev/null 2>&1
# 4. BigWig files -> text, end to end: the reference's own pthread pipeline beside the drop-in CLI
: > gpurun_out/${tag}_bw_e2e.jsonl
python profiles/bw_e2e_bench.py 8 20 >> gpurun_out/${tag}_bw_e2e.jsonl 2>gpurun_out/${tag}_bw_e2e.err
python profiles/bw_e2e_bench.py 64 10 >> gpurun_out/${tag}_bw_e2e.jsonl 2>>gpurun_out/${tag}_bw_e2e.err
ls -la gpurun_out | tail -12
