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

    # ---- symmetric kNN build: the same contract as CudaEngine, answered with the oracle ----
    SYM_MAX_LIST = 32
    EMPTY = np.int64(-1)

    @staticmethod
    def _keys(D, cols):
        return (D.astype(np.int64) << 32) | cols.astype(np.int64)

    @classmethod
    def _smallest(cls, keys, k1):
        """k1 smallest keys per row (rows may hold EMPTY = all ones, which sorts last as uint64)."""
        u = np.sort(keys.view(np.uint64), axis=1)[:, :k1]
        if u.shape[1] < k1:
            u = np.concatenate([u, np.full((u.shape[0], k1 - u.shape[1]), np.uint64(2**64 - 1))], axis=1)
        return u.view(np.int64)

    def hamming_knn_boot(self, table, row0, rows, boot_rows, k1):
        D = O.hamming(table.tokens[:boot_rows], table.tokens[row0:row0 + rows])
        keys = self._keys(D, np.broadcast_to(np.arange(boot_rows), D.shape))
        return torch.from_numpy(np.ascontiguousarray(self._smallest(keys, k1)))

    def hamming_knn_sym(self, table, k1, rank=0, world=1, lists=None, boot_rows=0, mode=0):
        """Candidates this rank sees.  Every 256-row block sweeps the stream rows from its own
        block on (row side) and feeds the rows of later blocks (column side); blocks of the
        bootstrap rows start behind them and have no column side.  mode 0: this rank takes blocks
        rank, rank+world, ...; mode 1: all blocks, restricted to the rank's band of stream rows
        (the band limits come from the library's host-side planner, pg_knn_sym_band)."""
        n = table.rows
        self.sym_calls = getattr(self, "sym_calls", 0) + 1
        self.sym_modes = getattr(self, "sym_modes", []) + [mode]
        cand = [[] for _ in range(n)]
        if boot_rows:
            for r in range(n):
                cand[r].extend(int(v) for v in lists[r].numpy() if v != -1)
        band = (0, n)
        if mode == 1:
            import ctypes as C
            from prograph_b200 import _lib
            a, b = C.c_int64(0), C.c_int64(0)
            _lib.check(_lib.load().pg_knn_sym_band(n, table.words, int(boot_rows), rank, world, C.byref(a), C.byref(b)))
            band = (a.value, b.value)
        T = table.tokens
        blocks = range(-(-n // 256)) if mode == 1 else range(rank, -(-n // 256), world)
        for rb in blocks:
            a, b = rb * 256, min(n, rb * 256 + 256)
            start = max(boot_rows if a < boot_rows else a, band[0])
            end = band[1]
            if start >= end:
                continue
            D = O.hamming(T[start:end], T[a:b])                      # (b-a, end-start)
            for i in range(a, b):
                for j in range(start, end):
                    d = int(D[i - a, j - start])
                    cand[i].append((d << 32) | j)
                    if a >= boot_rows and j >= b:
                        cand[j].append((d << 32) | i)
        out = np.full((n, k1), -1, dtype=np.int64)
        for r in range(n):
            u = sorted(set(cand[r]))[:k1]
            out[r, :len(u)] = u
        return torch.from_numpy(out)

    def knn_lists_finalize(self, lists, row0, rows, k, drop=1, similarity=False):
        lists = lists.numpy()
        if lists.ndim == 2:
            lists = lists[None]
        idx = np.full((rows, k), -1, dtype=np.int64)
        w = np.zeros((rows, k), dtype=np.float32 if similarity else np.int64)
        for r in range(rows):
            keys = sorted(set(int(v) for v in lists[:, row0 + r].reshape(-1) if v != -1))[drop:drop + k]
            for j, key in enumerate(keys):
                idx[r, j] = key & 0xffffffff
                d = key >> 32
                w[r, j] = np.float32(1.0) / np.float32(1 + d) if similarity else d
        return torch.from_numpy(idx), torch.from_numpy(w)

    def hamming_eps(self, own, row0, rows, stream, lut, similarity=False):
        self.rows_seen = (row0, rows)
        D = O.hamming(stream.tokens, own.tokens[row0:row0 + rows])
        lut = np.asarray(lut, dtype=np.uint32)
        keep = ((lut[D >> 5] >> (D & 31).astype(np.uint32)) & 1).astype(bool)
        r, c = np.nonzero(keep)
        W = O.hamming(stream.tokens, own.tokens[row0:row0 + rows], similarity=True) if similarity else D
        indptr = np.concatenate([[0], np.cumsum(keep.sum(1))]).astype(np.int64)
        return torch.from_numpy(indptr), torch.from_numpy(c.astype(np.int64)), torch.from_numpy(W[r, c])
