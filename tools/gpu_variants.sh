#!/bin/bash
# time every experimental build of the decoder kernel (tools/build_variants.sh) + the default one
O=gpurun_out
mkdir -p $O
python tools/time_sdec.py > $O/variants.log 2>&1
for f in pyroved_b200/csrc/variants/libpvb_*.so; do
  PVB_LIB=$PWD/$f timeout 300 python tools/time_sdec.py >> $O/variants.log 2>&1
done
cat $O/variants.log
