"""build_imagenet_data (upstream data/imagenet.py); synthetic tensors when the dataset is not on disk"""
import os
import warnings

from ._synthetic import synthetic_loaders


def build_imagenet_data(data_path: str = '', input_size: int = 224, batch_size: int = 64, workers: int = 4,
                        dist_sample: bool = False):
    root = os.path.expanduser(data_path)
    if os.path.isdir(os.path.join(root, 'train')) and os.path.isdir(os.path.join(root, 'val')):
        import torch
        import torchvision.datasets as datasets
        import torchvision.transforms as T
        norm = T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
        train = datasets.ImageFolder(os.path.join(root, 'train'),
                                     T.Compose([T.RandomResizedCrop(input_size), T.RandomHorizontalFlip(), T.ToTensor(), norm]))
        val = datasets.ImageFolder(os.path.join(root, 'val'),
                                   T.Compose([T.Resize(256), T.CenterCrop(input_size), T.ToTensor(), norm]))
        samp = (lambda d: torch.utils.data.distributed.DistributedSampler(d)) if dist_sample else (lambda d: None)
        ts, vs = samp(train), samp(val)
        return (torch.utils.data.DataLoader(train, batch_size=batch_size, shuffle=ts is None, num_workers=workers, pin_memory=True, sampler=ts),
                torch.utils.data.DataLoader(val, batch_size=batch_size, shuffle=False, num_workers=workers, pin_memory=True, sampler=vs))
    warnings.warn(f'ImageNet not found under {root!r}: serving synthetic {input_size}x{input_size} tensors')
    return synthetic_loaders(1024, 128, (3, input_size, input_size), 1000, batch_size)
