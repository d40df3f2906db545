"""GPU tests of the data-staging layer (reference utils/data.py:6-52,
trainers/svi.py:95-115, trainers/auxsvi.py:102-128): the GPU-resident loader with on-device
shuffle, the host loaders, and the pipelined epoch loops of both trainers against plain
step-by-step training."""
import pytest
import torch

import pyroved_b200 as pv
from pyroved_b200 import ops

pytestmark = pytest.mark.gpu


def test_gather_rows_kernel():
    g = torch.Generator().manual_seed(0)
    for row in (784, 7, 4096, 130):
        src = torch.randn(1000, row, generator=g).cuda()
        idx = torch.randperm(1000, generator=g)[:333].cuda()
        dst = torch.full((333, row), -7.0, device="cuda")
        ops.gather_rows(src, idx, dst)
        assert torch.equal(dst, src[idx]), row
    # multi-dimensional samples and an unaligned view
    src = torch.randn(64, 1, 5, 9, generator=g).cuda()
    idx = torch.tensor([3, 3, 0, 63], device="cuda")
    dst = torch.empty(4, 1, 5, 9, device="cuda")
    ops.gather_rows(src, idx, dst)
    assert torch.equal(dst, src[idx])


def test_device_batch_loader_protocol_and_shuffle():
    """Same protocol as the DataLoader the reference builds: tuples (x,) / (x, y), len() = number
    of batches, .dataset has a length; every epoch is a permutation of the dataset (drawn on the
    device), different from epoch to epoch, reproducible with a seed; ragged last batch."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1000, 28, 28, generator=g)
    y = torch.arange(1000, dtype=torch.float32)[:, None]
    ld = pv.utils.DeviceBatchLoader(x, y, batch_size=128, shuffle=True, device="cuda:0", seed=5)
    assert len(ld) == 8 and len(ld.dataset) == 1000
    seen = []
    for xb, yb in ld:
        assert xb.is_cuda and xb.shape[1:] == (28, 28) and yb.shape[1:] == (1,)
        ids = yb[:, 0].long().cpu()
        assert torch.equal(xb.cpu(), x[ids])            # rows travel together
        seen.append(ids.clone())
    assert [len(s) for s in seen] == [128] * 7 + [104]
    first = torch.cat(seen)
    assert torch.equal(first.sort().values, torch.arange(1000))
    second = torch.cat([yb[:, 0].long().cpu().clone() for _, yb in ld])
    assert torch.equal(second.sort().values, torch.arange(1000)) and not torch.equal(first, second)
    ld2 = pv.utils.DeviceBatchLoader(x, y, batch_size=128, shuffle=True, device="cuda:0", seed=5)
    assert torch.equal(torch.cat([yb[:, 0].long().cpu().clone() for _, yb in ld2]), first)
    # shuffle=False: views of the resident tensors, in order
    ld3 = pv.utils.DeviceBatchLoader(x, batch_size=300, shuffle=False, device="cuda:0")
    got = torch.cat([xb.cpu() for (xb,) in ld3])
    assert torch.equal(got, x) and len(ld3) == 4


def _ivae(seed=1):
    m = pv.models.iVAE((28, 28), 2, ['r', 't'], seed=seed, device="cuda:0")
    return m, pv.trainers.SVItrainer(m, seed=1, device="cuda:0")


def test_training_from_the_device_loader_equals_training_from_the_same_batches():
    """SVItrainer.train over a DeviceBatchLoader (on-device shuffle, no H2D per step) == svi.step on
    the same permuted batches one by one: same epoch loss, same weights."""
    g = torch.Generator().manual_seed(2)
    x = (torch.rand(4 * 512, 28, 28, generator=g) < 0.25).float()
    ma, ta = _ivae()
    mb, tb = _ivae()
    ld = pv.utils.DeviceBatchLoader(x, batch_size=512, shuffle=True, device="cuda:0", seed=11)
    for epoch in range(2):
        got = ta.train(ld) * len(ld.dataset)
        perm = ld.last_perm.cpu()
        ref = sum(tb.svi.step(x[perm[i * 512:(i + 1) * 512]].cuda()) for i in range(4))
        assert abs(got - ref) <= 1e-6 * abs(ref), (epoch, got, ref)
    for (k, a), (_, b) in zip(ma.state_dict().items(), mb.state_dict().items()):
        assert torch.allclose(a, b, atol=1e-6, rtol=0), (k, (a - b).abs().max().item())


def test_host_loader_shuffle_path():
    """TensorBatchLoader(shuffle=True): every epoch a permutation of the pinned dataset."""
    x = torch.arange(100, dtype=torch.float32)[:, None].repeat(1, 8)
    ld = pv.utils.TensorBatchLoader(x, batch_size=32, shuffle=True)
    a = torch.cat([xb[:, 0].clone() for (xb,) in ld])
    b = torch.cat([xb[:, 0].clone() for (xb,) in ld])
    assert torch.equal(a.sort().values, torch.arange(100.)) and torch.equal(b.sort().values, torch.arange(100.))
    assert not torch.equal(a, b)


def test_aux_trainer_pipelined_epoch_equals_compute_loss_loop():
    """auxSVItrainer.train (pipelined: staged uploads, two optimisation steps per batch without a
    host synchronisation, losses through the pinned ring) == the reference's loop written with
    compute_loss (trainers/auxsvi.py:102-128): same epoch loss, same weights, same schedule
    (a labelled batch after every unlabelled batch with i % p == 1)."""
    # batches of 512: every kernel of the step then sums in a fixed order (smaller batches take a
    # split-K GEMM whose atomics make the last bit of a gradient run-dependent, which Adam's
    # g / sqrt(v) turns into lr-sized differences on weights whose gradient is rounding noise)
    g = torch.Generator().manual_seed(3)
    xu = (torch.rand(6 * 512, 256, generator=g) < 0.3).float()
    xs = (torch.rand(2 * 512, 256, generator=g) < 0.3).float()
    ys = pv.utils.to_onehot(torch.randint(0, 3, (2 * 512,), generator=g), 3)
    lu = pv.utils.init_dataloader(xu, batch_size=512, shuffle=False)
    ls = pv.utils.init_dataloader(xs, ys, batch_size=512, shuffle=False)

    def make():
        m = pv.models.ssiVAE((16, 16), 2, 3, ['r'], seed=1, device="cuda:0")
        return m, pv.trainers.auxSVItrainer(m, seed=1, device="cuda:0")
    ma, ta = make()
    mb, tb = make()
    for epoch in range(2):
        got = ta.train(lu, ls, aux_loss_multiplier=20.0)
        # the reference loop
        p = (len(ls) + len(lu)) // len(ls)
        it = iter(ls)
        tot, cnt = 0.0, 0
        for i, (xb,) in enumerate(lu):
            tot += tb.compute_loss(xb, aux_loss_multiplier=20.0)
            cnt += xb.shape[0]
            if i % p == 1:
                xl, yl = next(it)
                tb.compute_loss(xl, yl, aux_loss_multiplier=20.0)
        ref = tot / cnt
        assert abs(got - ref) <= 1e-5 * abs(ref), (epoch, got, ref)
    for (k, a), (_, b) in zip(ma.state_dict().items(), mb.state_dict().items()):
        assert torch.allclose(a, b, atol=5e-6, rtol=0), (k, (a - b).abs().max().item())


def test_scale_factor_annealing_does_not_grow_the_graph_cache():
    """KL annealing: a new scale_factor every epoch must not pile up captured graphs."""
    m, tr = _ivae()
    x = (torch.rand(64, 28, 28) < 0.3).float().cuda()
    for e in range(tr.svi.MAX_GRAPHS + 40):
        tr.svi.step(x, scale_factor=1.0 + 0.01 * e)
        tr.svi.step(x, scale_factor=1.0 + 0.01 * e)
    assert len(tr.svi.graphs) <= tr.svi.MAX_GRAPHS


def test_graph_capture_with_discarded_engines_pending_collection(monkeypatch):
    """A trainer that went out of scope leaves its CUDA graphs in reference cycles until the cyclic collector
    runs; destroying a graph while ANOTHER capture is under way invalidates that capture
    (cudaErrorStreamCaptureInvalidated -- seen once in bench.py's configs block).  The engine collects before a
    capture and keeps the collector off during it.  Here a captured graph sits in a garbage cycle and the
    collector is made to run in the middle of the next engine's capture (a hook on the decoder launch)."""
    import gc
    x = (torch.rand(64, 12, 12, generator=torch.Generator().manual_seed(0)) < 0.3).float().cuda()
    gc.disable()
    try:
        m = pv.models.iVAE((12, 12), 2, ['r', 't'], seed=1, device="cuda:0")
        tr = pv.trainers.SVItrainer(m, seed=1, device="cuda:0")
        for _ in range(3):
            tr.svi.step(x)                       # eager, capture, replay
        assert any(not isinstance(g, str) for g in tr.svi.graphs.values())
        tr._cycle = tr                           # whatever the object graph looks like: certainly a cycle now
        del m, tr                                # garbage, not collected yet (the collector is off)
        m2 = pv.models.iVAE((12, 12), 2, ['r', 't'], seed=2, device="cuda:0")
        tr2 = pv.trainers.SVItrainer(m2, seed=2, device="cuda:0")
        tr2.svi.step(x)                          # eager
        orig = ops.sdec_tc_step
        calls = []

        def hooked(*a, **k):
            if gc.isenabled():                   # the engine's guard turns the collector off during a capture
                calls.append(gc.collect())
            return orig(*a, **k)
        monkeypatch.setattr(ops, "sdec_tc_step", hooked)
        gc.enable()
        l1 = tr2.svi.step(x)                     # capture
        l2 = tr2.svi.step(x)                     # replay
        assert l1 == l1 and l2 == l2
        assert not calls                         # the collector never ran inside the capture
    finally:
        gc.enable()
