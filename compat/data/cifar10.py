"""build_cifar10_data (upstream data/cifar10.py:9-53); synthetic tensors when the dataset is not on disk"""
import os
import warnings

from ._synthetic import synthetic_loaders


def build_cifar10_data(data_path: str = '', input_size: int = 224, batch_size: int = 64, workers: int = 4,
                       dist_sample: bool = False):
    root = os.path.expanduser(data_path)
    if os.path.isdir(os.path.join(root, 'cifar-10-batches-py')):
        import torch
        import torchvision.transforms as T
        from torchvision.datasets import CIFAR10
        norm = T.Normalize((0.4914, 0.4822, 0.4465), (0.2023, 0.1994, 0.2010))
        train = CIFAR10(root=root, train=True, download=False,
                        transform=T.Compose([T.RandomCrop(32, padding=4), T.RandomHorizontalFlip(), T.ToTensor(), norm]))
        val = CIFAR10(root=root, train=False, download=False, transform=T.Compose([T.ToTensor(), norm]))
        samp = (lambda d: torch.utils.data.distributed.DistributedSampler(d)) if dist_sample else (lambda d: None)
        ts, vs = samp(train), samp(val)
        return (torch.utils.data.DataLoader(train, batch_size=batch_size, shuffle=ts is None, num_workers=workers, pin_memory=True, sampler=ts),
                torch.utils.data.DataLoader(val, batch_size=batch_size, shuffle=False, num_workers=workers, pin_memory=True, sampler=vs))
    warnings.warn(f'CIFAR-10 not found under {root!r}: serving synthetic 32x32 tensors')
    return synthetic_loaders(1024, 256, (3, 32, 32), 10, batch_size)
