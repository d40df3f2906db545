#!/bin/bash
# compute-sanitizer passes (memcheck, racecheck, synccheck) over eager SVI steps of the hot paths.
# usage: tools/sanitize.sh [outdir] [cases...]   (run on a GPU box; logs are text, summarised at the end)
OUT=${1:-gpurun_out/sanitizer_r02}
shift
CASES=${@:-ivae jivae ved}
mkdir -p "$OUT"
SAN=${SAN:-/usr/local/cuda/bin/compute-sanitizer}
for c in $CASES; do
  for tool in memcheck racecheck synccheck; do
    log="$OUT/${tool}_${c}.log"
    if [ "$c" = "peer" ]; then
      # one sanitizer per rank (never a multi-rank command under one sanitizer)
      timeout ${SAN_TIMEOUT:-900} python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port 29533 --no-python \
        $SAN --tool $tool --log-file "$OUT/${tool}_${c}_rank%q{LOCAL_RANK}.log" \
        python tools/sanitize_step.py peer > "$OUT/${tool}_${c}.stdout" 2>&1
    else
      timeout ${SAN_TIMEOUT:-900} $SAN --tool $tool --log-file "$log" \
        python tools/sanitize_step.py $c > "$OUT/${tool}_${c}.stdout" 2>&1
    fi
    echo "$tool $c rc=$?" >> "$OUT/summary.txt"
  done
done
grep -H "ERROR SUMMARY\|RACECHECK SUMMARY\|hazard" "$OUT"/*.log | sort | uniq -c | sort -rn | head -40 >> "$OUT/summary.txt"
cat "$OUT/summary.txt"
