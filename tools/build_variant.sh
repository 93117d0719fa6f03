#!/bin/bash
# Builds a variant of the product library into build_variants/lib_<name>.so (A/B runs: tools/ab_variants.py).
#   tools/build_variant.sh <name> "<extra nvcc flags>" [git-rev]      (git-rev: build the csrc of that revision)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; EXTRA=$2; REV=$3
SRC=$ROOT/lidar_camera_calibration_b200/csrc
INC=$ROOT/include
if [ -n "$REV" ]; then
  TMP=$(mktemp -d); mkdir -p $TMP/csrc $TMP/include
  for f in ilcc2_kernels.cu ilcc2_api.cu ilcc2_pnp.cu ilcc2_internal.h; do git -C $ROOT show $REV:lidar_camera_calibration_b200/csrc/$f > $TMP/csrc/$f; done
  git -C $ROOT show $REV:include/ilcc2_b200.h > $TMP/include/ilcc2_b200.h
  SRC=$TMP/csrc; INC=$TMP/include
fi
mkdir -p $ROOT/build_variants
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC \
  -I$INC -I$SRC $EXTRA -shared -o $ROOT/build_variants/lib_$NAME.so $SRC/ilcc2_kernels.cu $SRC/ilcc2_api.cu $SRC/ilcc2_pnp.cu
echo built build_variants/lib_$NAME.so
