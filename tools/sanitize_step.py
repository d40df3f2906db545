"""One or two eager SVI steps of a hot path, for compute-sanitizer (tools/sanitize.sh).
usage: python tools/sanitize_step.py {ivae|jivae|ved|peer|ops}
  ivae : cfg2 model, batch 128 (784 tiles of the fused decoder kernel over 148 CTAs), 2 steps
  jivae: enumerated jiVAE 28x28, 3 classes, batch 32, 2 steps
  ved  : VED 64x64 -> 128 default filters, batch 8 (tcgen05 convolutions), 2 steps
  ops  : the kernels the small cases above do not reach: encoder batch 512 (tensor-core small-batch linear +
         grouped weight gradients), the 32768 -> 4 skinny layer at batch 512 (cluster kernel, dx, dW)
  peer : under torchrun --nproc-per-node 2: cfg2 model, batch 64 per rank, 3 steps through the fused
         NVLink all-reduce + Adam kernel
CUDA graphs are off so that every launch is checked individually."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PVB_CUDA_GRAPHS"] = "0"
import torch  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "ivae"
g = torch.Generator().manual_seed(0)
if what == "peer":
    import torch.distributed as dist
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda:{}".format(lr)))
    dev = "cuda:{}".format(lr)
else:
    dev = "cuda:0"
import pyroved_b200 as pv  # noqa: E402

if what in ("ivae", "peer"):
    B = 128 if what == "ivae" else 64
    m = pv.models.iVAE((28, 28), 2, ['r', 't'], seed=1, device=dev)
    tr = pv.trainers.SVItrainer(m, device=dev)
    x = (torch.rand(B, 28, 28, generator=g) < 0.3).float().to(dev)
    ls = [tr.svi.step(x) for _ in range(3 if what == "peer" else 2)]
elif what == "jivae":
    m = pv.models.jiVAE((28, 28), 2, 3, ['r'], seed=1, device=dev)
    tr = pv.trainers.SVItrainer(m, enumerate_parallel=True, device=dev)
    x = (torch.rand(32, 28, 28, generator=g) < 0.3).float().to(dev)
    ls = [tr.svi.step(x, scale_factor=[3., 3.]) for _ in range(2)]
elif what == "ops":
    from pyroved_b200 import ops
    m = pv.models.iVAE((28, 28), 2, ['r', 't'], seed=1, device=dev)
    tr = pv.trainers.SVItrainer(m, device=dev)
    x = (torch.rand(512, 28, 28, generator=g) < 0.3).float().to(dev)
    tr.svi.step(x)                         # allocates the batch-512 programme
    prog = next(iter(tr.svi.programs.values()))
    enc = prog.enc
    l0 = enc.layers[0]
    for M in (512, 1024):                  # 16-wide and 32-wide tiles
        xi = torch.rand(M, 784, generator=g).to(dev)
        ops.linear_fwd(xi, l0.weight.data, l0.bias.data, "tanh")
    ops.mlp_wgrad(enc._wgrad, enc.M)
    xs = torch.randn(512, 32768, generator=g).to(dev)
    Ws, bs = torch.randn(4, 32768, generator=g).to(dev) * 0.01, torch.zeros(4, device=dev)
    gs = torch.randn(512, 4, generator=g).to(dev)
    ys = ops.linear_fwd(xs, Ws, bs, None)
    dxs, dWs, dbs = torch.empty_like(xs), torch.zeros_like(Ws), torch.zeros_like(bs)
    ops.linear_bwd(xs, Ws, None, None, gs, gs, dxs, False, dWs, dbs, None)
    ls = [float(ys.sum()), float(dWs.sum())]
else:
    m = pv.models.VED((64, 64), (128,), latent_dim=2, seed=1, device=dev)
    tr = pv.trainers.SVItrainer(m, device=dev)
    x = torch.rand(8, 1, 64, 64, generator=g).to(dev)
    y = torch.rand(8, 1, 128, generator=g).to(dev)
    ls = [tr.svi.step(x, y, scale_factor=4.0) for _ in range(2)]
torch.cuda.synchronize()
print(what, "losses", ls, "peer" if getattr(tr.svi, "peer", None) is not None else "")
assert all(v == v for v in ls)
if what == "peer":
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()
