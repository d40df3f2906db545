#!/bin/bash
# Experimental builds of the fused decoder kernels (timing / tracing only; some variants drop work
# and give wrong gradients).  Each lands in pyroved_b200/csrc/variants/libpvb_<name>.so and is
# selected with PVB_LIB=<path>.
# usage: tools/build_variants.sh name:file.cu:"-Dflag ..." ...      (file = pvb_sdec_tc.cu | pvb_sdec_tc2.cu)
set -e
cd "$(dirname "$0")/../pyroved_b200/csrc"
./build.sh > /dev/null
mkdir -p variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
for spec in "$@"; do
  name=${spec%%:*}
  rest=${spec#*:}
  file=${rest%%:*}
  defs=${rest#*:}
  obj=${file%.cu}.o
  OTHERS=$(ls *.o | grep -v "^$obj$")
  $NVCC $FLAGS $defs -c $file -o variants/${name}_$obj 2>/dev/null
  $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o variants/libpvb_$name.so $OTHERS variants/${name}_$obj
  echo "built variants/libpvb_$name.so ($file $defs)"
done
