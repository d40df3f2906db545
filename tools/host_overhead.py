"""How much of the end-to-end epoch loop (SVItrainer.train) is host time?  The loop only blocks in
Event.synchronize() (loss read-back three steps behind): blocked time ~ 0 means the launch thread,
not the GPU, sets the pace."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import pyroved_b200 as pv  # noqa: E402

B, NB = 512, 400
m = pv.models.iVAE((28, 28), latent_dim=2, invariances=["r", "t"], seed=1, device="cuda:0")
tr = pv.trainers.SVItrainer(m, seed=1, device="cuda:0")
x = (torch.rand(B * 48, 28, 28) < 0.3).float()
loader = pv.utils.TensorBatchLoader(x, batch_size=B) if hasattr(pv.utils, "TensorBatchLoader") else \
    pv.utils.init_dataloader(x, batch_size=B, shuffle=False)
tr.train(loader)
tr.train(loader)
blocked = [0.0]
orig = torch.cuda.Event.synchronize


def timed(self):
    t0 = time.perf_counter()
    orig(self)
    blocked[0] += time.perf_counter() - t0


torch.cuda.Event.synchronize = timed
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 0
for _ in range(8):
    tr.train(loader)
    n += 48
torch.cuda.synchronize()
wall = time.perf_counter() - t0
print("steps {}  wall {:.1f} us/step  blocked in Event.synchronize {:.1f} us/step  -> host work {:.1f} us/step".format(
    n, wall / n * 1e6, blocked[0] / n * 1e6, (wall - blocked[0]) / n * 1e6))
