"""cfg5 side measurement (not the headline bench): VED 64x64 image -> 128-point spectrum,
default filters, batch 512 on one GPU; prints samples/s of the SVI step and the CPU port's."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import pyroved_b200 as pv  # noqa: E402

B = int(os.environ.get("VED_B", "512"))
m = pv.models.VED((64, 64), (128,), latent_dim=2, seed=1, device="cuda:0")
tr = pv.trainers.SVItrainer(m, device="cuda:0")
g = torch.Generator().manual_seed(0)
x = torch.rand(B, 1, 64, 64, generator=g).cuda()
y = torch.rand(B, 1, 128, generator=g).cuda()
for _ in range(3):
    tr.svi.step(x, y, scale_factor=4.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    tr.svi.step(x, y, scale_factor=4.0, _sync=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
out = {"workload": "VED 64x64 -> 128, default filters, batch {}".format(B), "ms_per_step": ms,
       "samples_per_s": B / ms * 1e3, "gflop_per_step": 0.71 * B,
       "tflops": 0.71 * B / ms}
if os.environ.get("VED_CPU", "1") == "1":
    from oracle import svi_port as sp
    torch.set_num_threads(os.cpu_count())
    cfg = sp.VedCfg((64, 64), (128,), 2)
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    xb, yb = x[:64].cpu(), y[:64].cpu()
    eps = torch.randn(64, 2)
    sp.loss_and_grads(sp.ved_loss, sd, cfg, xb, yb, eps, 4.0)
    t0 = time.perf_counter()
    for _ in range(3):
        sp.loss_and_grads(sp.ved_loss, sd, cfg, xb, yb, eps, 4.0)
    out["cpu_port_samples_per_s"] = 3 * 64 / (time.perf_counter() - t0)
    out["cpu_cores"] = os.cpu_count()
print(json.dumps(out))
