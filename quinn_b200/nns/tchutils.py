"""numpy <-> torch helpers with the reference's dtype contract (quinn/nns/tchutils.py:9-41):
the default dtype is double and float arrays are cast to it."""
import numpy as np
import torch

torch.set_default_dtype(torch.double)


def tch(arr, device='cpu', rgrad=False):
    """numpy array (or list of arrays) -> tensor of the default dtype (tchutils.py:11-28)."""
    if isinstance(arr, list):
        arr = np.array(arr)
    t = torch.tensor(arr, requires_grad=rgrad, device=device)
    return t.to(torch.get_default_dtype()) if t.is_floating_point() else t


def npy(arr):
    """tensor -> numpy (tchutils.py:31-41)."""
    return arr.detach().cpu().numpy()


def print_nnparams(nnmodel, names_only=False):
    assert isinstance(nnmodel, torch.nn.Module)
    for name, param in nnmodel.named_parameters():
        if names_only:
            print(f"{name}, shape {tuple(param.shape)}")
        else:
            print(name, param.data)
