#!/bin/bash
# Developer tool: build A/B variants of libpmnet_b200.so that differ only in scoring.cu compile-time options.
# usage: tools/build_variants.sh tag1:"-DFOO=1 -DBAR" tag2:"..."   -> pharmaconet_b200/_variants/libpmnet_b200_<tag>.so
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/pharmaconet_b200/csrc
OUT=$ROOT/pharmaconet_b200/_variants
OBJ=/tmp/pmnet_variant_obj
mkdir -p $OUT $OBJ
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
for f in conv3d pointwise swin_ops gemm; do
  if [ ! -f $OBJ/$f.o ] || [ $SRC/$f.cu -nt $OBJ/$f.o ]; then nvcc $FLAGS -c $SRC/$f.cu -o $OBJ/$f.o & fi
done
wait
for spec in "$@"; do
  tag=${spec%%:*}; defs=${spec#*:}
  ( nvcc $FLAGS $defs -Xptxas=-v -c $SRC/scoring.cu -o $OBJ/scoring_$tag.o 2> $OBJ/scoring_$tag.log
    nvcc -shared -o $OUT/libpmnet_b200_$tag.so $OBJ/scoring_$tag.o $OBJ/conv3d.o $OBJ/pointwise.o $OBJ/swin_ops.o $OBJ/gemm.o
    echo "$tag: $(grep -A2 'pmnet_score_kernelILi1' $OBJ/scoring_$tag.log | grep -E 'Used|spill' | tr '\n' ' ' | sed 's/ptxas info *: //g')" ) &
done
wait
