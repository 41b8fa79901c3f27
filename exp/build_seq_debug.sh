#!/bin/bash
# sequence.cu compiled with -G (device debug: no optimisation), everything else as usual
set -e
root=/root/repo; pkg=$root/02-visualodometry_b200; tmp=$(mktemp -d)
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-gencode arch=compute_100a,code=sm_100a -std=c++17 --fmad=false -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math"
for f in $pkg/csrc/*.cu; do
  b=$(basename $f .cu)
  if [ "$b" = "sequence" ]; then $NVCC $FLAGS -G -c $f -o $tmp/$b.o & else $NVCC $FLAGS -O3 -lineinfo -c $f -o $tmp/$b.o & fi
done
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $root/exp/variants/libvo_b200_seqG.so $tmp/*.o -ldl
rm -rf $tmp; echo built seqG
