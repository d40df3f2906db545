#!/bin/bash
# Profiling visit (1 GPU): launch lists of one cfg2 step and one VED step, one `ncu --set full`
# capture of the fused decoder kernel.  Outputs under gpurun_out/<tag>_*; summarise here with
# tools/launch_list.py / tools/ncu_summary.py and copy into profiles/.
TAG=${1:-prof}
O=gpurun_out
mkdir -p $O
export PVB_BENCH_SKIP_CPU=1 PVB_BENCH_CONFIGS=0
# every launch of bench.py --steps 2 (warm-up 3: eager, capture, replay; then the timed blocks)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $O/${TAG}_launches_ivae.csv python bench.py --steps 2 --warmup 3 > $O/${TAG}_ivae_ncu.log 2>&1
VED_CPU=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file $O/${TAG}_launches_ved.csv python tools/bench_ved.py > $O/${TAG}_ved_ncu.log 2>&1
# the dominant kernel, once, full set (skip the eager + capture launches)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sdec_tc_kernel -s 3 -c 1 \
  -f -o $O/${TAG}_sdec python bench.py --steps 2 --warmup 3 > $O/${TAG}_sdec_ncu.log 2>&1
ls -la $O/${TAG}_*
