#!/bin/bash
# Development: build alternative configurations of the library for one-call A/B runs on the GPU box.
#   profiles/build_variants.sh "PREF:MINB[:extra nvcc flags]" ...   -> gpurun_out is scratch; variants go to build/variants/
set -e
cd "$(dirname "$0")/../wiggletools_b200/csrc"
mkdir -p ../../build/variants
for v in "$@"; do
  IFS=: read pref minb extra <<< "$v"
  out=../../build/variants/libwb_p${pref}_m${minb}$(echo "$extra" | tr -d ' =-' | tr -c 'A-Za-z0-9_\n' '_').so
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -DWB_PREF=$pref -DWB_FAST_MINB=$minb $extra \
    -Xcompiler -fPIC,-Wall,-Wno-unused-function -shared -o $out engine.cu fastsum.cu writer.cu maps.cu capi.cu dropin.cpp bigwig.cpp -lcudart_static -lpthread -ldl -lrt -lz &
done
wait
ls -la ../../build/variants
