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
