#!/bin/bash
# exp/build_variant.sh <name> <extra nvcc flags...>: builds exp/variants/libvo_b200_<name>.so with EXTRA flags
# (kernel A/B experiments; select it with VO_B200_LIB=...)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
pkg=$root/02-visualodometry_b200
tmp=$(mktemp -d)
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --fmad=false -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math"
for f in $pkg/csrc/*.cu; do
  $NVCC $FLAGS "$@" -c $f -o $tmp/$(basename $f .cu).o &
done
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $root/exp/variants/libvo_b200_$name.so $tmp/*.o -ldl
rm -rf $tmp
echo built exp/variants/libvo_b200_$name.so
