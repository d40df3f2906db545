"""Time latent_side_bwd at the cfg3 (jiVAE, 10240 instances) and cfg2 shapes inside their step programmes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchlib as bl
for name in ("cfg2", "cfg3", "cfg4"):
    m, tr = bl.build(name, "cuda:0")
    B = bl.WORKLOADS[name]["batch"]
    data = tuple(t.cuda() for t in bl.synth(name, B, seed=1))
    kw = dict(scale_factor=[3.0, 3.0]) if name == "cfg3" else {}
    for _ in range(4):
        tr.svi.step(*data, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        tr.svi.step(*data, **kw)
    e1.record(); torch.cuda.synchronize()
    print(os.environ.get("PVB_LIB", "default")[-14:], name, "step %.1f us" % (e0.elapsed_time(e1) / 20 * 1e3))
