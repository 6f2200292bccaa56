import torch
from torch.utils.data import DataLoader, TensorDataset


def synthetic_loaders(n_train, n_val, shape, n_classes, batch_size, seed=0):
    g = torch.Generator().manual_seed(seed)
    mk = lambda n: TensorDataset(torch.randn((n,) + shape, generator=g), torch.randint(0, n_classes, (n,), generator=g))
    return (DataLoader(mk(n_train), batch_size=batch_size, shuffle=False),
            DataLoader(mk(n_val), batch_size=batch_size, shuffle=False))
