#!/bin/bash
# Multi-GPU visit: bench.py at N ranks (fused NVLink exchange, dp_check, configs block).
# usage: tools/gpu_scale.sh <tag> "<N list>" [steps]      PVB_PEER_TWO_SHOT is passed through
TAG=${1:-scale}
NS=${2:-"8"}
STEPS=${3:-50}
O=gpurun_out
mkdir -p $O
for N in $NS; do
  for form in ${FORMS:-default}; do
    if [ "$form" = "default" ]; then unset PVB_PEER_TWO_SHOT; else export PVB_PEER_TWO_SHOT=$form; fi
    out=$O/${TAG}_n${N}_${form}
    if [ "$N" = "1" ]; then
      timeout 600 python bench.py --gpus 1 --steps $STEPS --warmup 5 > $out.json 2> $out.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
        --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N --steps $STEPS --warmup 5 \
        > $out.json 2> $out.err
    fi
    tail -c 200 $out.err
    python - <<PY
import json
try:
    d = json.loads(open("$out.json").read().strip().splitlines()[-1])
    print("N=$N form=$form: value %.0f e2e %.0f ms/step %.4f exch=%s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"]["exchange"][:30]))
    print("   dp_check", {k: v for k, v in (d.get("dp_check") or {}).items() if k != "note"})
    for k, c in d.get("configs", {}).items():
        print("   ", k, "value %.0f e2e %.0f ms/batch %.3f" % (c.get("value", 0), c.get("e2e", {}).get("value", 0), c.get("ms_per_batch", 0)) if c and "value" in c else c)
except Exception as e:
    print("parse failed $out", e)
PY
  done
done
