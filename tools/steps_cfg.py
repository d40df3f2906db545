"""Run a few device-resident steps of one bench workload (cfg3 / cfg4 / cfg5 / cfg2), for launch lists:
ncu --metrics gpu__time_duration.sum ... python tools/steps_cfg.py cfg3"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import benchlib as bl  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
m, tr = bl.build(name, "cuda:0")
w = bl.WORKLOADS[name]
B = w["batch"]
kw = {}
if name == "cfg3":
    kw = dict(scale_factor=[3.0, 3.0])
if name == "cfg5":
    kw = dict(scale_factor=4.0)
if w["kind"] == "ssivae":
    xu = tuple(t.cuda() for t in bl.synth(name, B, seed=1))
    xs = tuple(t.cuda() for t in bl.synth(name, B, seed=2, labelled=True))
    for i in range(6):
        tr.svi.step(*xu)
        if i % 2:
            tr.svi.step(*xs)
            tr.svi._step(xs, {"aux_loss_multiplier": 50.0}, train=True, update=True, mode="aux") if hasattr(tr.svi, "_step") else None
else:
    data = tuple(t.cuda() for t in bl.synth(name, B, seed=1))
    for i in range(6):
        tr.svi.step(*data, **kw)
torch.cuda.synchronize()
print("done", name)
