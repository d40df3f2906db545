#!/bin/bash
# Build libpvb.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall"
SRCS="pvb_core.cu pvb_gemm.cu pvb_latent.cu pvb_sdec_simt.cu pvb_mlp.cu pvb_conv.cu pvb_conv_tc.cu pvb_norm.cu pvb_peer.cu pvb_conv3d.cu"
[ -f pvb_sdec_tc.cu ] && SRCS="$SRCS pvb_sdec_tc.cu pvb_sdec_tc2.cu"
OBJS=""
PIDS=""
FAILED=""
for s in $SRCS; do
  o="${s%.cu}.o"
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ pvb_common.cuh -nt "$o" ] || [ pvb_fold.cuh -nt "$o" ] || [ umma.cuh -nt "$o" ] || [ pvb_sdec_tc.cuh -nt "$o" ] || [ ../../include/pvb.h -nt "$o" ]; then
    # compile to a temporary name: a failed compile must not leave the previous object to be linked
    ( $NVCC $FLAGS ${PVB_EXTRA_FLAGS} -c "$s" -o "$o.tmp" && mv "$o.tmp" "$o" || { rm -f "$o" "$o.tmp"; exit 1; } ) &
    PIDS="$PIDS $!"
  fi
  OBJS="$OBJS $o"
done
for p in $PIDS; do
  wait $p || FAILED=1
done
if [ -n "$FAILED" ]; then
  echo "build failed" >&2
  exit 1
fi
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libpvb.so $OBJS
echo "built $(pwd)/libpvb.so"
