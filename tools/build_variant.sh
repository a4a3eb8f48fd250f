#!/bin/bash
# usage: build_variant.sh <tag> [extra nvcc -D flags...]   -> geogram_b200/lib/variants/libb200vor_<tag>.so
set -e
TAG=$1; shift
HERE=$(cd $(dirname $0)/.. && pwd)
OUT=$HERE/geogram_b200/lib/variants; mkdir -p $OUT/obj_$TAG
for f in grid knn knn_team cell tpc otm microbench slab capi; do cp $HERE/geogram_b200/lib/obj/$f.o $OUT/obj_$TAG/; done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" \
  -c $HERE/geogram_b200/csrc/fused.cu -o $OUT/obj_$TAG/fused.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libb200vor_$TAG.so $OUT/obj_$TAG/*.o -lcudart -ldl
rm -rf $OUT/obj_$TAG
echo built $OUT/libb200vor_$TAG.so
