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
