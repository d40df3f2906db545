#!/bin/bash
# 2-GPU visit: the N = 2 bench (fused NVLink exchange, dp_check) in both exchange forms, and the
# sanitizer over the exchange kernel.  usage: tools/gpu_check2.sh <tag> [steps]
TAG=${1:-chk2}
STEPS=${2:-50}
O=gpurun_out
mkdir -p $O
for form in 0 1; do
  PVB_PEER_TWO_SHOT=$form timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
    --master-addr 127.0.0.1 --master-port 2951$form bench.py --gpus 2 --steps $STEPS --warmup 5 \
    > $O/${TAG}_bench_n2_twoshot$form.json 2> $O/${TAG}_bench_n2_twoshot$form.err
  tail -c 300 $O/${TAG}_bench_n2_twoshot$form.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_bench_n2_twoshot$form.json").read().strip().splitlines()[-1])
    print("N=2 two_shot=$form: value %.0f e2e %.0f ms/step %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]),
          d.get("dp_check"), {k: (round(c.get("value", 0)), round(c.get("ms_per_batch", 0), 3)) for k, c in d.get("configs", {}).items() if c})
except Exception as e:
    print("parse failed", e)
PY
done
SAN_TIMEOUT=300 timeout 700 tools/sanitize.sh $O/sanitizer_r02 peer > $O/${TAG}_san_peer.log 2>&1
tail -6 $O/${TAG}_san_peer.log
