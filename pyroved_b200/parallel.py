"""Data parallelism for the SVI step: one process per GPU, batches sharded
across ranks, ONE all-reduce(SUM) of the flat gradient buffer (+ loss slot)
per optimizer step over NCCL (NVLink / NVSwitch), identical fused Adam on
every rank.

The reference has no distributed code at all; the loss of
its SVI step is a SUM over the batch (Trace_ELBO), so gradients of shards add
up exactly -- the collective is SUM, not AVG, and no learning-rate rescaling
is involved.  Noise is drawn from a counter-based generator keyed by the
GLOBAL sample index, so the ELBO does not depend on the number of GPUs.
"""
import os

import torch
import torch.distributed as dist


def init_process_group(backend=None, device=None):
    """Initialise torch.distributed from torchrun's environment (RANK,
    LOCAL_RANK, WORLD_SIZE, MASTER_ADDR, MASTER_PORT).  Returns (rank, world)."""
    if dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return 0, 1
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        kw["device_id"] = torch.device("cuda:{}".format(local) if device is None else device)
    dist.init_process_group(backend, **kw)
    return dist.get_rank(), dist.get_world_size()


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n, rank, world):
    """Contiguous slice [lo, hi) of a global batch of n samples owned by `rank`
    (n must divide evenly: every rank runs the same static kernel shapes)."""
    if n % world != 0:
        raise ValueError("global batch {} is not divisible by world size {}".format(n, world))
    per = n // world
    return rank * per, (rank + 1) * per


def shard(t, rank=None, world=None):
    """This rank's contiguous slice of a global batch tensor."""
    if rank is None or world is None:
        rank, world = rank_world()
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def noise_first_index(rank, n_local):
    """Global index of the first noise element of this rank's shard."""
    return rank * n_local


def allreduce_sum_(flat, group=None):
    """In-place SUM all-reduce of the flat [gradients | loss] buffer."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


class PeerExchange:
    """Symmetric-memory plumbing for the fused all-reduce + Adam kernel (csrc/pvb_peer.cu): every
    rank owns two staging buffers (used by epoch parity) and a block of epoch flags in CUDA
    symmetric memory (torch.distributed._symmetric_memory: VMM allocations whose handles the
    per-GPU processes exchange) mapped into all ranks; the kernel gets the peer pointers as two
    small device arrays.  PyTorch only provides the allocation / handle exchange here -- the data
    path is the kernel's own loads over NVLink.  The gradient buffer itself stays local."""

    def __init__(self, n_floats, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import ops
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.n_floats = int(n_floats)
        self.stage = [symm_mem.empty(n_floats, dtype=torch.float32, device=device) for _ in range(2)]
        self.flags = symm_mem.empty(max(64, ops.peer_flag_words()), dtype=torch.int32, device=device)
        for t in self.stage:
            t.zero_()
        self.flags.zero_()
        hs = [symm_mem.rendezvous(t, group.group_name) for t in self.stage]
        hf = symm_mem.rendezvous(self.flags, group.group_name)
        self._handles = (hs, hf)
        ptrs = [int(p) for h in hs for p in h.buffer_ptrs]        # [parity][rank]
        self.stage_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=device)
        self.peer_flags = torch.tensor([int(p) for p in hf.buffer_ptrs], dtype=torch.int64,
                                       device=device)
        self.state = torch.zeros(ops.peer_state_words(), dtype=torch.int32, device=device)
        # reduce-scatter + all-gather (inbound 2 (W-1)/W instead of W-1 buffer sizes, one more flag
        # round) pays off for large buffers on >= 4 ranks: measured at 8 GPUs, 4.5 MB (ssiVAE 64x64):
        # 1.613 vs 1.668 ms per batch; 0.6 MB (iVAE 28x28): 0.251 vs 0.242 ms per step, so small
        # buffers keep the one-shot form.  PVB_PEER_TWO_SHOT=0/1 forces either (every rank alike)
        env = os.environ.get("PVB_PEER_TWO_SHOT")
        self.two_shot = ((self.world >= 4 and self.n_floats >= (1 << 19)) if env is None
                         else (env == "1"))
        torch.cuda.synchronize(device)
        dist.barrier(group)       # nobody signals before every rank's flags are zeroed


def peer_exchange_enabled():
    """The fused NVLink exchange is used for NCCL (CUDA) process groups unless PVB_PEER_REDUCE=0."""
    return (os.environ.get("PVB_PEER_REDUCE", "1") != "0" and dist.is_available()
            and dist.is_initialized() and dist.get_world_size() > 1
            and dist.get_backend() == "nccl")


class ShardedLoader:
    """Wraps a loader of GLOBAL batches; yields this rank's shard of each
    tensor (keeps the (x,) / (x, y) tuple protocol of the trainers)."""

    def __init__(self, loader, rank=None, world=None):
        self.loader = loader
        r, w = rank_world()
        self.rank = r if rank is None else rank
        self.world = w if world is None else world
        self.dataset = _ShardedLen(loader.dataset, self.world)

    def __iter__(self):
        for batch in self.loader:
            n = batch[0].shape[0]
            if n % self.world != 0:
                # a ragged last batch: trim it evenly (every rank must run the same static
                # shapes); the few dropped samples are seen in other epochs when shuffling
                n = n // self.world * self.world
                if n == 0:
                    continue
                batch = [t[:n] for t in batch]
            yield [shard(t, self.rank, self.world) for t in batch]

    def __len__(self):
        return len(self.loader)


class _ShardedLen:
    """len() = samples this rank sees per epoch (the trainer divides by it)."""

    def __init__(self, dataset, world):
        self.n = len(dataset) // world

    def __len__(self):
        return self.n
