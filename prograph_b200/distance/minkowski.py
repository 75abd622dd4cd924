"""Minkowski distance, drop-in for prograph/distance/minkowski.py:8-41."""
import torch

from ..engine import get_engine
from .utils import clean_input, result_device, is_integer_dtype


def staged_dtype(X, Y, p):
    """Compute dtype as torch's type promotion picks it for
    ``pow(sum(pow(X - Y, p)), 1/p)``: fp16 stays fp16, float32/float64 stay, integers stay
    exact (int64) for positive integer p and promote to float32 otherwise."""
    dt = torch.result_type(X, Y)
    if is_integer_dtype(dt):
        return torch.int64 if (float(p) == int(p) and p > 0) else torch.float32
    if dt in (torch.float16, torch.float32, torch.float64):
        return dt
    return torch.float32


def minkowski_matrix(eng, X, Y, p=2, similarity=False):
    dt = staged_dtype(X, Y, p)
    Xd, Yd = eng.to_device(X, dt), eng.to_device(Y, dt)
    return eng.minkowski_tile(Xd, Yd, 0, Yd.shape[0], p=p, similarity=similarity)


def minkowski(X, Y, p=2, similarity=False):
    """Pairwise Minkowski "distances" between the rows of X (N, D) and Y (M, D):
    ``pow(sum(pow(X - Y[:, None, :], p), axis=2), 1/p)`` -- without an absolute value,
    exactly as the reference computes it (minkowski.py:36), so p=1 gives signed sums and
    odd p can give NaN.  Shape (M, N).  fp16 in -> fp16 out with every step rounded to
    fp16 as torch does (the dtype ``build_graph`` feeds, prograph.py:726); float32 ->
    float32; integers -> exact integer power sum, float32 root.  ``similarity`` returns
    ``1/(1+d)`` (minkowski.py:39-40).
    """
    X, Y = clean_input(X, Y)
    dev = result_device(X, Y)
    out = minkowski_matrix(get_engine(), X, Y, p, similarity)
    return out if out.device == dev else out.to(dev)
