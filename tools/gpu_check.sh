#!/bin/bash
# One GPU-box visit: kernel-level check of both decoder-kernel variants, the whole GPU suite, a short
# bench.  Everything lands in gpurun_out/<tag>_*.  usage: tools/gpu_check.sh <tag> [bench steps]
TAG=${1:-chk}
STEPS=${2:-20}
O=gpurun_out
mkdir -p $O
PT="python -m pytest -q -p no:cacheprovider --timeout 300"
# 1. the two variants of the fused decoder kernel against PyTorch (a protocol error traps, see
#    mbar_wait in csrc/pvb_sdec_tc2.cu; the timeout is the second line of defence)
timeout 600 $PT tests/test_gpu_kernels.py -k sdec_tc 2>&1 | tail -15 > $O/${TAG}_sdec.log
tail -3 $O/${TAG}_sdec.log
# 2. whole GPU suite
timeout 1200 $PT tests -m gpu 2>&1 | tail -40 > $O/${TAG}_tests.log
tail -4 $O/${TAG}_tests.log
# 3. bench (headline + configs block)
timeout 900 python bench.py --steps $STEPS --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
tail -c 400 $O/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("bench: value %.0f  e2e %.0f  ms/step %.4f  sdec us %.1f frac %.3f" % (
        d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["us"], d["roofline"]["frac"]))
    for k, c in d.get("configs", {}).items():
        print(" ", k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in c.items()
                      if kk in ("value", "ms_per_batch", "error")}, c.get("e2e", {}).get("value"),
              (c.get("roofline") or {}).get("frac"))
except Exception as e:
    print("bench parse failed:", e)
PY
