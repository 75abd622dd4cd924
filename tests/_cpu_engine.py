"""A checker-backed stand-in for prograph_b200.engine.CudaEngine, used ONLY by the CPU tests
to exercise the host logic above the C ABI (dispatch, sharding, gathering, list assembly)
on machines without a GPU.  It answers the engine calls with the oracle; the product never
imports this module."""
import numpy as np
import torch

from oracle import prograph_oracle as O


class _Packed:
    def __init__(self, tokens):
        self.tokens = np.asarray(tokens)
        self.rows, self.L = self.tokens.shape
        self.planes, self.words = 5, max(1, -(-self.L // 32))


class CheckerEngine:
    device = torch.device("cpu")

    def __init__(self):
        self.rows_seen = (0, 0)

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype)

    def pack(self, tokens, planes=None, words=None):
        t = tokens.numpy() if isinstance(tokens, torch.Tensor) else np.asarray(tokens)
        if t.dtype.kind == "f" and not np.all(t == np.round(t)) or t.min() < 0 or t.max() > 255:
            raise OverflowError("not tokens")
        return _Packed(t.astype(np.int64))

    def hamming_knn(self, own, row0, rows, stream, k, drop=1, similarity=False):
        self.rows_seen = (row0, rows)
        D = O.hamming(stream.tokens, own.tokens[row0:row0 + rows], similarity=similarity)
        key = D.astype(np.float32) if similarity else D
        order = np.argsort(-key if similarity else key, axis=1, kind="stable")[:, drop:drop + k]
        return torch.from_numpy(order.astype(np.int64)), torch.from_numpy(np.take_along_axis(D, order, axis=1))

    def hamming_eps(self, own, row0, rows, stream, lut, similarity=False):
        self.rows_seen = (row0, rows)
        D = O.hamming(stream.tokens, own.tokens[row0:row0 + rows])
        lut = np.asarray(lut, dtype=np.uint32)
        keep = ((lut[D >> 5] >> (D & 31).astype(np.uint32)) & 1).astype(bool)
        r, c = np.nonzero(keep)
        W = O.hamming(stream.tokens, own.tokens[row0:row0 + rows], similarity=True) if similarity else D
        indptr = np.concatenate([[0], np.cumsum(keep.sum(1))]).astype(np.int64)
        return torch.from_numpy(indptr), torch.from_numpy(c.astype(np.int64)), torch.from_numpy(W[r, c])
