"""Operand normalisation shared by the distance functions.

Mirrors prograph/distance/utils.py:7-39 (same name, arguments and error behaviour).
"""
import torch
import torch.nn.functional as F


def clean_input(X, Y, verbose=False):
    """Return X, Y as 2-D tensors with a common second dimension.

    * an empty operand raises ``ValueError`` (utils.py:29-30);
    * 1-D operands become a single row (utils.py:31);
    * the narrower operand is right-padded with zeros (utils.py:32-38), so a pad token
      matches another pad token and differs from every residue.
    """
    if X.shape[0] == 0 or Y.shape[0] == 0:
        raise ValueError("You cannot pass an empty tensor: it would be padded with zeros and the distance "
                         "from every sequence to the origin would be returned. Pass a tensor of zeros "
                         "explicitly if that is what you want.")
    X = torch.atleast_2d(torch.as_tensor(X))
    Y = torch.atleast_2d(torch.as_tensor(Y))
    dx, dy = X.shape[1], Y.shape[1]
    if dx != dy:
        if verbose:
            print("X and Y have different sequence lengths (dimension 1)")
        if dy > dx:
            X = F.pad(X, (0, dy - dx))
        else:
            Y = F.pad(Y, (0, dx - dy))
    return X, Y


def result_device(X, Y):
    """Results live where the inputs live (the reference returns a tensor on its inputs' device)."""
    for t in (X, Y):
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    return torch.device("cpu")


_INT_DTYPES = (torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64, torch.bool)


def is_integer_dtype(dt):
    return dt in _INT_DTYPES
