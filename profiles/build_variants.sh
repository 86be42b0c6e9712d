#!/bin/bash
# Development: build alternative configurations of the sweep kernel for one-call A/B runs on the GPU box.
#   profiles/build_variants.sh "tag:[-Dflags ...][:alternative fastsum source]" ...   -> build/variants/lib<tag>.so
# Only fastsum.cu is recompiled per variant; the other objects are built once.
set -e
cd "$(dirname "$0")/../wiggletools_b200/csrc"
V=../../build/variants; O=../../build/obj
mkdir -p $V $O
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-Wall,-Wno-unused-function"
for f in engine.cu writer.cu maps.cu capi.cu dropin.cpp bigwig.cpp apply.cu multi.cpp; do
  [ -f $f ] || continue
  [ $O/$f.o -nt $f ] && [ $O/$f.o -nt engine.cuh ] || $NV -c $f -o $O/$f.o &
done
wait
for v in "$@"; do
  IFS=: read tag flags src <<< "$v"
  ( $NV $flags -I. -c ${src:-fastsum.cu} -o $O/fastsum_$tag.o && \
    nvcc -shared -o $V/lib$tag.so $O/fastsum_$tag.o $(ls $O/*.o | grep -v fastsum_) -lcudart_static -lpthread -ldl -lrt -lz ) &
done
wait
ls -la $V
