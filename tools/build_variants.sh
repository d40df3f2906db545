#!/bin/bash
# Experimental builds of the interleaved decoder kernel (timing / tracing only; some variants drop
# work and give wrong gradients).  Each lands in pyroved_b200/csrc/variants/libpvb_<name>.so and is
# selected with PVB_LIB=<path>.  usage: tools/build_variants.sh name:"-Dflag ..." ...
set -e
cd "$(dirname "$0")/../pyroved_b200/csrc"
./build.sh > /dev/null
mkdir -p variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
OTHERS=$(ls *.o | grep -v "^pvb_sdec_tc2.o$")
for spec in "$@"; do
  name=${spec%%:*}
  defs=${spec#*:}
  $NVCC $FLAGS $defs -c pvb_sdec_tc2.cu -o variants/tc2_$name.o
  $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o variants/libpvb_$name.so $OTHERS variants/tc2_$name.o
  echo "built variants/libpvb_$name.so ($defs)"
done
