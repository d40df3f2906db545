"""Multi-GPU test (needs >= 2 GPUs, run under `gpurun --gpus 2`): two NCCL
ranks on half batches must reproduce the single-GPU full-batch step."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data():
    g = torch.Generator().manual_seed(11)
    x = (torch.rand(64, 28, 28, generator=g) < 0.3).float()
    eps = torch.randn(64, 5, generator=g)
    return x, eps


def _worker(rank, world, port, q, two_shot):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank), PVB_PEER_TWO_SHOT=two_shot)
    import pyroved_b200 as pv
    from pyroved_b200 import parallel
    parallel.init_process_group("nccl")
    dev = "cuda:{}".format(rank)
    x, eps = _data()
    m = pv.models.iVAE((28, 28), 2, ['r', 't'], seed=1, device=dev)
    tr = pv.trainers.SVItrainer(m, seed=1, device=dev)
    losses = []
    for _ in range(3):
        losses.append(tr.svi.step(parallel.shard(x).to(dev), _eps=parallel.shard(eps).to(dev)))
    # auto-generated noise is keyed by the global sample index
    tr.svi.step(parallel.shard(x).to(dev))
    prog = next(iter(tr.svi.programs.values()))
    assert tr.svi.peer is not None and tr.svi.peer.two_shot == (two_shot == "1")
    if rank == 0:
        q.put((losses, {k: v.cpu() for k, v in m.state_dict().items()}, prog.eps.cpu()))
    else:
        q.put(("eps1", prog.eps.cpu()))
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("two_shot", ["0", "1"], ids=["one-shot", "two-shot"])
def test_two_rank_step_equals_single_gpu_step(two_shot):
    """Fused NVLink exchange (csrc/pvb_peer.cu), both forms: the one-shot sum every rank does for
    small worlds, and the reduce-scatter + all-gather form used from 4 ranks on (forced here at 2
    ranks): three data-parallel steps on half batches == three single-GPU steps on the full batch."""
    import pyroved_b200 as pv
    x, eps = _data()
    m = pv.models.iVAE((28, 28), 2, ['r', 't'], seed=1, device="cuda:0")
    tr = pv.trainers.SVItrainer(m, seed=1, device="cuda:0")
    ref_losses = [tr.svi.step(x.cuda(), _eps=eps.cuda()) for _ in range(3)]
    tr.svi.step(x.cuda())
    ref_eps = next(iter(tr.svi.programs.values())).eps.cpu()
    ref_sd = {k: v.cpu() for k, v in m.state_dict().items()}

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, two_shot)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    main = [g for g in got if g[0] != "eps1"][0]
    other = [g for g in got if g[0] == "eps1"][0]
    losses, sd, eps0 = main
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 2e-4 * abs(b), (losses, ref_losses)
    # weights after 3 DP steps (before the 4th, which used fresh noise)
    assert torch.equal(torch.cat([eps0, other[1]]), ref_eps)
    for k in ref_sd:
        assert torch.allclose(sd[k], ref_sd[k], atol=2e-4), k
