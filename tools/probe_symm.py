"""Probe: does torch symmetric memory (CUDA VMM handles exchanged between the per-GPU processes)
work on this box?  torchrun --nproc-per-node 2 tools/probe_symm.py"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyroved_b200 import parallel  # noqa: E402

rank, world = parallel.init_process_group()
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
import torch.distributed._symmetric_memory as symm_mem  # noqa: E402

t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
h = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "ptrs", [hex(p) for p in h.buffer_ptrs], "multicast", h.has_multicast_support,
      hex(h.multicast_ptr), "signal pad", h.signal_pad_size, flush=True)
t.fill_(float(rank + 1))
h.barrier(0)
peer = h.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
print(rank, "peer value", float(peer[12345]), flush=True)
h.barrier(0)
x = torch.randn(152076, device=dev)
for _ in range(5):
    dist.all_reduce(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    dist.all_reduce(x)
e1.record()
torch.cuda.synchronize()
print(rank, "nccl all_reduce 0.61 MB: {:.1f} us".format(e0.elapsed_time(e1) * 5), flush=True)
dist.destroy_process_group()
