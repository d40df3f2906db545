"""Data-loader helpers with the reference's call signatures
(reference utils/data.py:6-52)."""
import torch


def init_dataloader(*args, random_sampler=False, shuffle=True, **kwargs):
    """TensorDataset + DataLoader(batch_size=100 by default, num_workers=0).
    Extra kwarg `pin_memory=True` pins host batches so the trainer's H2D copy
    is asynchronous."""
    device_ = kwargs.get("device")
    generator_ = torch.Generator(device_) if device_ else None
    batch_size = kwargs.get("batch_size", 100)
    pin = bool(kwargs.get("pin_memory", False))
    dataset = torch.utils.data.TensorDataset(*args)
    if random_sampler:
        sampler = torch.utils.data.RandomSampler(dataset)
        return torch.utils.data.DataLoader(dataset, batch_size=batch_size, sampler=sampler,
                                           generator=generator_, pin_memory=pin)
    return torch.utils.data.DataLoader(dataset, batch_size=batch_size, shuffle=shuffle,
                                       generator=generator_, pin_memory=pin)


def init_ssvae_dataloaders(data_unsup, data_sup, data_val, **kwargs):
    """Three loaders for semi-supervised training (reference utils/data.py:41-52;
    its `sampler=True` kwarg is ignored there, so it is here too)."""
    kwargs.pop("sampler", None)
    loader_unsup = init_dataloader(data_unsup, **kwargs)
    loader_sup = init_dataloader(*data_sup, **kwargs)
    loader_val = init_dataloader(*data_val, **kwargs)
    return loader_unsup, loader_sup, loader_val


class TensorBatchLoader:
    """Contiguous mini-batches sliced from (pinned) host tensors: no per-sample collation, so a
    batch is one zero-copy view and one H2D copy.  Same protocol as the DataLoader that
    `init_dataloader` returns, as far as `SVItrainer.train` uses it: iteration yields tuples
    `(x,)` or `(x, y)` (reference trainers/svi.py:105-111) and `.dataset` has a length.
    shuffle=True permutes the samples once per epoch on the host (index_select into a pinned
    buffer)."""

    def __init__(self, *tensors, batch_size=100, shuffle=False, pin_memory=True, drop_last=False):
        n = tensors[0].shape[0]
        assert all(t.shape[0] == n for t in tensors)
        self.tensors = [t.contiguous() for t in tensors]
        if pin_memory and torch.cuda.is_available():
            self.tensors = [t if t.is_pinned() else t.pin_memory() for t in self.tensors]
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), shuffle, drop_last
        self.dataset = torch.utils.data.TensorDataset(*self.tensors)
        self._perm_buf = None

    def __len__(self):
        n = self.tensors[0].shape[0]
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        src = self.tensors
        n = src[0].shape[0]
        if self.shuffle:
            perm = torch.randperm(n)
            if self._perm_buf is None:
                self._perm_buf = [torch.empty_like(t).pin_memory() if t.is_pinned()
                                  else torch.empty_like(t) for t in src]
            for t, b in zip(src, self._perm_buf):
                torch.index_select(t, 0, perm, out=b)
            src = self._perm_buf
        for i in range(len(self)):
            lo = i * self.batch_size
            yield tuple(t[lo:lo + self.batch_size] for t in src)


class DeviceBatchLoader:
    """GPU-resident replacement for `init_dataloader` (reference utils/data.py:6-52): the whole
    dataset lives in HBM (fp32; 180 GB per B200 holds ~57 M 28x28 images), every epoch draws a
    permutation ON THE DEVICE, and each mini-batch is gathered by a hand-written row-gather kernel
    (`pvb_gather_rows`, HBM-bound) into one of two device buffers whose addresses recur -- so the
    trainers' step graphs take their input from it without any host -> device traffic.  Same
    protocol as the DataLoader the reference builds, as far as the trainers use it: iteration
    yields tuples `(x,)` / `(x, y)`, `len()` is the number of batches, `.dataset` has a length.

    shuffle=False yields zero-copy views of the resident tensors.  `seed` makes the epoch
    permutations reproducible (one generator, advanced every epoch)."""

    def __init__(self, *tensors, batch_size=100, shuffle=True, device=None, seed=None,
                 drop_last=False):
        if device is None:
            device = tensors[0].device if tensors[0].is_cuda else "cuda"
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceBatchLoader keeps the dataset in GPU memory; got device={!r}"
                               .format(str(device)))
        n = tensors[0].shape[0]
        assert all(t.shape[0] == n for t in tensors)
        self.tensors = [t.to(self.device, torch.float32).contiguous() for t in tensors]
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), bool(shuffle), drop_last
        self.dataset = _Len(n)
        self.gen = torch.Generator(device=self.device)
        if seed is not None:
            self.gen.manual_seed(int(seed))
        self._bufs = None
        self.last_perm = None

    def __len__(self):
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        from .. import ops
        n, bs = len(self.dataset), self.batch_size
        if not self.shuffle:
            for i in range(len(self)):
                yield tuple(t[i * bs:(i + 1) * bs] for t in self.tensors)
            return
        with torch.cuda.device(self.device):
            perm = torch.randperm(n, device=self.device, generator=self.gen)
            self.last_perm = perm
            if self._bufs is None:
                self._bufs = [[torch.empty((bs, *t.shape[1:]), device=self.device) for t in self.tensors]
                              for _ in range(2)]
            for i in range(len(self)):
                idx = perm[i * bs:(i + 1) * bs]
                out = []
                for t, b in zip(self.tensors, self._bufs[i % 2]):
                    ops.gather_rows(t, idx, b)
                    out.append(b[:idx.numel()])
                yield tuple(out)


class _Len:
    def __init__(self, n):
        self.n = int(n)

    def __len__(self):
        return self.n
