#!/bin/bash
# Builds libfidget_cuda.so (sm_100a only) in-tree and the CPU oracle.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
SRC=fidget_b200/csrc
OUT=${OUT:-fidget_b200/libfidget_cuda.so}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  -Xcompiler -fPIC,-O2,-ffp-contract=off -shared"
if [ "$1" = "-v" ]; then FLAGS="$FLAGS -Xptxas -v"; fi
# EXTRA="-DSOME_SWITCH" OUT=fidget_b200/libfidget_cuda_variant.so ./build.sh builds a variant for FIDGET_B200_LIB
FLAGS="$FLAGS $EXTRA"
$NVCC $FLAGS -o $OUT $SRC/cuda/kernels.cu $SRC/cuda/coop.cu $SRC/cuda/bulk.cu $SRC/cuda/tail2d.cu $SRC/cuda/octree.cu $SRC/cuda/effects.cu $SRC/cuda/capi.cu $SRC/cuda/schedule.cu $SRC/cuda/render.cu $SRC/cuda/octree_capi.cu $SRC/cuda/mesh.cu $SRC/cuda/effects_capi.cu $SRC/host/tape.cc $SRC/host/host_capi.cc
make -s -C oracle liboracle.so
echo "built $OUT and oracle/liboracle.so"
