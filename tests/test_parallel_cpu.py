"""CPU tests (gloo, world_size 2) of the data-parallel host logic: batch
sharding, the SUM all-reduce of the flat [gradients | loss] buffer, and that
shard gradients of the SVI loss add up to the full-batch gradient (checked
with the oracle port, which is the CPU restatement of the reference step)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyroved_b200 import parallel
from oracle import svi_port as sp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    r, w = parallel.init_process_group("gloo")
    assert (r, w) == (rank, world)
    cfg = sp.Cfg((12, 12), 2, ['r', 't'])
    sd = sp.init_ivae_state(cfg, seed=1)
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(8, 12, 12, generator=g) < 0.3).float()
    eps = torch.randn(8, cfg.z_dim, generator=g)
    xs, es = parallel.shard(x), parallel.shard(eps)
    assert xs.shape[0] == 4
    assert parallel.noise_first_index(rank, 4 * cfg.z_dim) == rank * 4 * cfg.z_dim
    res, grads = sp.loss_and_grads(sp.ivae_loss, sd, cfg, xs, es)
    names = list(grads)
    flat = torch.cat([grads[n].reshape(-1) for n in names] + [res["loss"].reshape(1)])
    parallel.allreduce_sum_(flat)
    if rank == 0:
        full, gfull = sp.loss_and_grads(sp.ivae_loss, sd, cfg, x, eps)
        ref = torch.cat([gfull[n].reshape(-1) for n in names] + [full["loss"].reshape(1)])
        out.put(((flat - ref).abs().max().item(), ref.abs().max().item()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_gradients_sum_to_full_batch_gradient():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, scale = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err <= 1e-4 * scale, (err, scale)


def test_shard_bounds_and_loader():
    assert parallel.shard_bounds(512, 3, 8) == (192, 256)
    try:
        parallel.shard_bounds(10, 0, 4)
        assert False
    except ValueError:
        pass
    import pyroved_b200 as pv
    x = torch.arange(40.).reshape(20, 2)
    y = torch.arange(20.).reshape(20, 1)
    loader = pv.utils.init_dataloader(x, y, batch_size=10, shuffle=False)
    parts = [list(parallel.ShardedLoader(loader, rank=r, world=2)) for r in range(2)]
    for b in range(2):
        xb = torch.cat([parts[0][b][0], parts[1][b][0]])
        assert torch.equal(xb, x[b * 10:(b + 1) * 10])
        assert parts[0][b][1].shape == (5, 1)
    assert len(parallel.ShardedLoader(loader, rank=0, world=2).dataset) == 10
