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
        self.informative = None


class CheckerEngine:
    """The full engine interface graph.py uses on the fused Hamming path, answered by the oracle."""
    device = torch.device("cpu")
    sharded_pack = False          # tables are packed whole on every rank
    GEMM_MAX_WIDTH = 0            # no tensor-core path

    def __init__(self):
        self.rows_seen = (0, 0)
        self.sym_checks = 0

    def sym_check(self):
        self.sym_checks += 1

    def check_edge_budget(self, nnz):
        pass

    def knn_lists_merge(self, lists, k, drop=1):
        """(n_lists, rows, k1) -> (rows, k) merged unique keys after `drop`, -1 = missing."""
        lists = lists.numpy()
        rows = lists.shape[1]
        out = np.full((rows, k), -1, dtype=np.int64)
        for r in range(rows):
            keys = sorted(set(int(v) for v in lists[:, r].reshape(-1) if v != -1))[drop:drop + k]
            out[r, :len(keys)] = keys
        return torch.from_numpy(out)

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype)

    def pack(self, tokens, planes=None, words=None):
        t = tokens.numpy() if isinstance(tokens, torch.Tensor) else np.asarray(tokens)
        if t.dtype.kind == "f" and not np.all(t == np.round(t)) or t.min() < 0 or t.max() > 255:
            raise OverflowError("not tokens")
        return _Packed(t.astype(np.int64))

    def packed_words(self, L_):
        from prograph_b200 import _lib
        return int(_lib.load().pg_packed_words(int(L_)))

    def varying_columns(self, table):
        return np.nonzero((table.tokens != table.tokens[0]).any(axis=0))[0].astype(np.int32)

    def compact_columns(self, table, cols):
        self.compacted = getattr(self, "compacted", 0) + 1
        return _Packed(table.tokens[:, cols] if len(cols) else np.zeros((table.rows, 1), dtype=np.int64))

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

    def _band(self, table, boot_rows, rank, world, mode):
        """Stream-row band of a rank (mode 1) from the library's host-side planner."""
        if mode != 1:
            return 0, table.rows
        import ctypes as C
        from prograph_b200 import _lib
        a, b = C.c_int64(0), C.c_int64(0)
        _lib.check(_lib.load().pg_knn_sym_band(table.rows, table.words, int(boot_rows), rank, world, C.byref(a),
                                               C.byref(b)))
        return a.value, b.value

    def _triangle(self, table, rank, world, mode, boot_rows=0):
        """The (own rows a..b, stream rows start..end, distances) pieces of the triangle of unordered
        pairs this rank sweeps.  Every 256-row block sweeps the stream rows from its own block on;
        blocks of the bootstrap rows start behind them.  mode 0: blocks rank, rank+world, ...;
        mode 1: all blocks, restricted to the rank's band of stream rows."""
        n, T = table.rows, table.tokens
        band = self._band(table, boot_rows, rank, world, mode)
        blocks = range(-(-n // 256)) if mode == 1 else range(rank, -(-n // 256), world)
        for rb in blocks:
            a, b = rb * 256, min(n, rb * 256 + 256)
            start, end = max(boot_rows if a < boot_rows else a, band[0]), band[1]
            if start < end:
                yield a, b, start, end, O.hamming(T[start:end], T[a:b])        # (b-a, end-start)

    def hamming_knn_sym(self, table, k1, rank=0, world=1, lists=None, boot_rows=0, mode=0):
        """Candidates this rank sees: row side for the own row, column side for the rows of later
        blocks; blocks of the bootstrap rows have no column side."""
        n = table.rows
        self.sym_calls = getattr(self, "sym_calls", 0) + 1
        self.sym_modes = getattr(self, "sym_modes", []) + [mode]
        cand = [[] for _ in range(n)]
        if boot_rows:
            for r in range(n):
                cand[r].extend(int(v) for v in lists[r].numpy() if v != -1)
        for a, b, start, end, D in self._triangle(table, rank, world, mode, boot_rows):
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

    def hamming_eps_degrees(self, own, row0, rows, stream, lut):
        indptr, _, _ = self.hamming_eps(own, row0, rows, stream, lut)
        return indptr[1:] - indptr[:-1]

    def hamming_eps_mean_degree(self, own, row0, rows, stream, lut):
        deg = self.hamming_eps_degrees(own, row0, rows, stream, lut)
        return float(deg.sum()) / rows, int(deg.max())

    # symmetric epsilon sweep: edge keys row << 40 | column << 12 | distance (the checker's own packing)
    def hamming_eps_sym(self, table, lut, rank=0, world=1, mode=0, capacity=None):
        self.eps_sym_calls = getattr(self, "eps_sym_calls", 0) + 1
        lut = np.asarray(lut, dtype=np.uint32)
        hits = np.nonzero([(lut[d >> 5] >> (d & 31)) & 1 for d in range(len(lut) * 32)])[0]
        if len(hits) and hits[-1] - hits[0] + 1 != len(hits):
            from prograph_b200._lib import Unsupported
            raise Unsupported("not a contiguous range")
        keys = []
        for a, b, start, end, D in self._triangle(table, rank, world, mode):
            keep = ((lut[D >> 5] >> (D & 31).astype(np.uint32)) & 1).astype(bool)
            for i, j in zip(*np.nonzero(keep)):
                gi, gj, d = a + int(i), start + int(j), int(D[i, j])
                keys.append((gi << 40) | (gj << 12) | d)
                if gj >= b:
                    keys.append((gj << 40) | (gi << 12) | d)
        keys = keys + [-1] * (7 + rank)                         # unused slots of the last chunk
        return torch.tensor(keys, dtype=torch.int64), len(keys) - 7 - rank

    def edge_keys_to_csr(self, keys, rows, words, nnz, similarity=False):
        k = np.sort(keys.numpy().view(np.uint64))[:nnz].astype(np.int64)
        row, col, d = k >> 40, (k >> 12) & ((1 << 28) - 1), k & 4095
        indptr = np.concatenate([[0], np.cumsum(np.bincount(row, minlength=rows))]).astype(np.int64)
        w = (np.float32(1) / (1 + d).astype(np.float32)).astype(np.float32) if similarity else d
        return torch.from_numpy(indptr), torch.from_numpy(col), torch.from_numpy(w)

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

    def hamming_eps(self, own, row0, rows, stream, lut, similarity=False, capture=None):
        self.rows_seen = (row0, rows)
        self.captures = getattr(self, "captures", []) + [capture]
        D = O.hamming(stream.tokens, own.tokens[row0:row0 + rows])
        lut = np.asarray(lut, dtype=np.uint32)
        keep = ((lut[D >> 5] >> (D & 31).astype(np.uint32)) & 1).astype(bool)
        r, c = np.nonzero(keep)
        W = O.hamming(stream.tokens, own.tokens[row0:row0 + rows], similarity=True) if similarity else D
        indptr = np.concatenate([[0], np.cumsum(keep.sum(1))]).astype(np.int64)
        return torch.from_numpy(indptr), torch.from_numpy(c.astype(np.int64)), torch.from_numpy(W[r, c])
