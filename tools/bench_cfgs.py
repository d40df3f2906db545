"""Side measurements for BASELINE configs 3 and 4 (parity-test configs, not the headline bench):
jiVAE 28x28, 10 classes, batch 1024; ssiVAE 64x64, 4 classes, batch 256 per GPU (unsupervised step
+ auxiliary step).  Prints samples/s of the SVI step on one GPU."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import pyroved_b200 as pv  # noqa: E402


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


g = torch.Generator().manual_seed(0)
out = {}
B = 1024
m = pv.models.jiVAE((28, 28), latent_dim=2, discrete_dim=10, invariances=['r'], seed=1, device="cuda:0")
tr = pv.trainers.SVItrainer(m, enumerate_parallel=True, device="cuda:0")
x = (torch.rand(B, 28, 28, generator=g) < 0.3).float().cuda()
ms = timed(lambda: tr.svi.step(x, scale_factor=[3, 3], _sync=False))
out["cfg3_jiVAE_28x28_K10_B1024"] = {"ms_per_step": ms, "samples_per_s": B / ms * 1e3,
                                      "decoder_rows_per_step": B * 10 * 784}
B = 256
m = pv.models.ssiVAE((64, 64), latent_dim=2, num_classes=4, invariances=['r'], seed=1, device="cuda:0")
tr = pv.trainers.auxSVItrainer(m, device="cuda:0")
x = (torch.rand(B, 4096, generator=g) < 0.3).float().cuda()
ms = timed(lambda: (tr.svi.step(x, _sync=False), tr.svi.step_aux(x, _sync=False)))
out["cfg4_ssiVAE_64x64_K4_B256_unsup+aux"] = {"ms_per_step": ms, "samples_per_s": B / ms * 1e3,
                                               "decoder_rows_per_step": B * 4 * 4096}
print(json.dumps(out))
