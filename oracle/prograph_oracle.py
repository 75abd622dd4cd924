"""CPU oracle for prograph's graph-construction hot path.  TEST INFRASTRUCTURE ONLY.

A plain numpy restatement of the reference algorithm (acmater/prograph, read-only at
/root/reference in the build container).  Every function cites the reference
file:line it follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module;
the product (``prograph_b200``) never does and fails loudly without its CUDA library.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function here
against outputs of the unmodified reference captured by
``tests/golden/make_golden.py`` (the reference's own unit vectors from
tests/tests.py:41-208, data/synthetic_data.csv, data/knntest.csv,
data/knntest_pgraph.pkl and a seeded ragged mutational library).

The arithmetic of the reference lives in torch (``!=``, ``sub``, ``pow``, ``sum``,
``sort``, ``where``; reference pin torch>=1.8.1 in setup.py:136, 2.11.0 here).  What is
restated below is torch's published semantics for those ops:
  * fp16 elementwise ops compute in fp32 and round once to fp16;
  * ``sum`` over fp16 accumulates in fp32 and rounds the total to fp16;
  * ``pow(x, 2)`` is ``x*x`` and ``pow(x, 0.5)`` is ``sqrt(x)`` (the optimised scalar
    exponents), other exponents go through ``powf``;
  * an int64 tensor raised to a float exponent, or divided, promotes to float32;
  * ``sort`` ascending puts NaN last; the tie order used by the parity contract is
    the *stable* one (value, then index) -- SURVEY.md §8(c).
"""
import operator
from functools import reduce

import numpy as np

# --------------------------------------------------------------------------
# distance/utils.py:7-39
# --------------------------------------------------------------------------


def clean_input(X, Y):
    """distance/utils.py:29-39: empty operand -> ValueError; promote to 2-D; the
    narrower operand is right-padded with zeros."""
    X, Y = np.asarray(X), np.asarray(Y)
    if X.shape[0] == 0 or Y.shape[0] == 0:
        raise ValueError("You cannot pass an empty tensor.")
    X, Y = np.atleast_2d(X), np.atleast_2d(Y)
    if X.shape[1] != Y.shape[1]:
        if Y.shape[1] > X.shape[1]:
            X = np.pad(X, ((0, 0), (0, Y.shape[1] - X.shape[1])))
        else:
            Y = np.pad(Y, ((0, 0), (0, X.shape[1] - Y.shape[1])))
    return X, Y


def _similarity(d):
    """``1/(1+d)`` (hamming.py:37-38, minkowski.py:39-40).  Integer d promotes to
    float32; fp16 stays fp16 with both the add and the divide rounded to fp16."""
    if d.dtype.kind in "iu":
        return (np.float32(1.0) / (1 + d).astype(np.float32)).astype(np.float32)
    if d.dtype == np.float16:
        one_plus = (np.float32(1.0) + d.astype(np.float32)).astype(np.float16)
        with np.errstate(divide="ignore"):
            return (np.float32(1.0) / one_plus.astype(np.float32)).astype(np.float16)
    with np.errstate(divide="ignore"):
        return (d.dtype.type(1.0) / (d.dtype.type(1.0) + d)).astype(d.dtype)


# --------------------------------------------------------------------------
# distance/hamming.py:8-39
# --------------------------------------------------------------------------


def hamming(X, Y, similarity=False, chunk=64):
    """hamming.py:34: ``sum(X != Y[:,None,:], axis=2)`` -> (M, N) int64; rows follow
    Y (queries), columns follow X (dataset)."""
    X, Y = clean_input(X, Y)
    M, N = Y.shape[0], X.shape[0]
    out = np.empty((M, N), dtype=np.int64)
    for m0 in range(0, M, chunk):
        out[m0:m0 + chunk] = (X[None, :, :] != Y[m0:m0 + chunk, None, :]).sum(axis=2)
    if similarity:
        return _similarity(out)
    return out


# --------------------------------------------------------------------------
# distance/minkowski.py:8-41
# --------------------------------------------------------------------------


def _pow_f32(x, p):
    """torch.pow(float32 tensor, python scalar): the exponent is cast to float32;
    2 -> x*x, 3 -> x*x*x, 0.5 -> sqrt, 1 -> identity (the optimised exponents of
    ATen's pow_tensor_scalar kernels), powf otherwise."""
    x = x.astype(np.float32)
    p = float(np.float32(p))
    if p == 2:
        return x * x
    if p == 1:
        return x
    if p == 0.5:
        return np.sqrt(x)
    if p == 3:
        return (x * x) * x
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.power(x, np.float32(p)).astype(np.float32)


def _pow_f16(x, p):
    """torch.pow(fp16 tensor, python scalar): the exponent is cast to fp16 first
    (``exp_scalar.to<scalar_t>()``), every multiply of the optimised exponents rounds
    to fp16, the general case is powf on the fp32-widened operands."""
    x = x.astype(np.float16)
    p = float(np.float16(p))
    xf = x.astype(np.float32)
    if p == 2:
        return (xf * xf).astype(np.float16)
    if p == 1:
        return x
    if p == 0.5:
        return np.sqrt(xf).astype(np.float16)
    if p == 3:
        return ((xf * xf).astype(np.float16).astype(np.float32) * xf).astype(np.float16)
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        return np.power(xf, np.float32(p)).astype(np.float16)


def minkowski(X, Y, p=2, similarity=False):
    """minkowski.py:36: ``pow(sum(pow(X - Y[:,None,:], p), axis=2), 1/p)`` -- no abs.
    dtype chain per SURVEY.md Appendix A.3/A.7: fp16 in -> fp16 out with every step
    rounded to fp16 (sum accumulated in fp32); float32 in -> float32; integer in ->
    exact integer power sum, then a float32 root."""
    X, Y = clean_input(X, Y)
    if X.dtype.kind in "iu" or Y.dtype.kind in "iu":
        if X.dtype.kind == "f" or Y.dtype.kind == "f":
            ft = np.result_type(X.dtype if X.dtype.kind == "f" else np.float32,
                                Y.dtype if Y.dtype.kind == "f" else np.float32)
            X, Y = X.astype(ft), Y.astype(ft)
    if X.dtype.kind in "iu":
        diff = X[None, :, :].astype(np.int64) - Y[:, None, :].astype(np.int64)
        if float(p) == int(p) and p >= 0:
            s = np.sum(diff ** int(p), axis=2)
            root = _pow_f32(s.astype(np.float32), 1.0 / p)
        else:  # int tensor ** float scalar promotes to float32 first
            s = np.sum(_pow_f32(diff.astype(np.float32), p), axis=2, dtype=np.float32)
            root = _pow_f32(s, 1.0 / p)
        d = root.astype(np.float32)
    elif X.dtype == np.float16 or Y.dtype == np.float16:
        X, Y = X.astype(np.float16), Y.astype(np.float16)
        diff = (X[None, :, :].astype(np.float32) - Y[:, None, :].astype(np.float32)).astype(np.float16)
        pw = _pow_f16(diff, p)
        with np.errstate(over="ignore"):
            s = np.sum(pw.astype(np.float32), axis=2, dtype=np.float32).astype(np.float16)
        d = _pow_f16(s, 1.0 / p)
    else:
        ft = np.result_type(X.dtype, Y.dtype)
        diff = X[None, :, :].astype(ft) - Y[:, None, :].astype(ft)
        if ft == np.float32:
            pw = _pow_f32(diff, p)
            s = np.sum(pw, axis=2, dtype=np.float32)
            d = _pow_f32(s, 1.0 / p)
        else:
            with np.errstate(invalid="ignore"):
                d = np.power(np.sum(np.power(diff, p), axis=2), 1.0 / p)
    if similarity:
        return _similarity(d)
    return d


# --------------------------------------------------------------------------
# prograph.py:454-474, 488-505  (tokeniser and mutation masks)
# --------------------------------------------------------------------------

AMINO_ACIDS = "ACDEFGHIKLMNPQRSTVWY"


def tokenize(sequences, amino_acids=AMINO_ACIDS):
    """prograph.py:454-474: letters -> 1..len(alphabet), right-pad with 0."""
    arr = np.array(sequences, dtype="bytes").reshape(-1, 1).view("S1")
    tok = np.zeros(arr.shape, dtype=np.int64)
    for i, ch in enumerate(amino_acids):
        tok[arr == ch.encode("utf-8")] = i + 1
    return tok


def boolean_mutant_array(tokenized, ref_row):
    """prograph.py:488-492."""
    return tokenized != tokenized[ref_row]


def calc_mutated_positions(tokenized, seed_tokens):
    """prograph.py:494-505: positions where any row differs from the seed."""
    return np.where(~np.all(tokenized == np.asarray(seed_tokens).reshape(1, -1), axis=0))[0]


def get_mutated_positions(mutant_array, mutated_positions, positions):
    """prograph.py:349-368: rows whose mutations avoid every *other* dataset-mutated
    position."""
    constants = np.setdiff1d(mutated_positions, positions)
    return np.all(~mutant_array[:, constants], axis=1)


def indexing(tokenized, ref_row, distances=None, positions=None, Bool="or", complement=False):
    """prograph.py:254-343 without the unseeded ``percentage`` sub-sampling."""
    assert Bool in ("or", "and"), "Not a valid boolean value."
    N, L = tokenized.shape
    d_data = hamming(tokenized, tokenized[ref_row].reshape(1, -1))
    idxs = []
    if distances is not None:
        if type(distances) == int:
            distances = [distances]
        assert type(distances) == list, "Distances must be provided as integer or list"
        for d in distances:
            assert d in np.unique(d_data), f"{d} is not a valid distance"
        idxs.append(reduce(np.union1d, [np.where(d_data == d)[1] for d in distances]))
    if positions is not None:
        mut = boolean_mutant_array(tokenized, ref_row)
        op = np.logical_or if Bool == "or" else np.logical_and
        working = reduce(op, [mut[:, pos] for pos in positions])
        for pos in range(L):
            if pos not in positions:
                working = working & ~mut[:, pos]      # prograph.py:322-324 simplifies to this
        idxs.append(np.where(working)[0])
    idxs = reduce(np.intersect1d, idxs) if idxs else np.arange(N)
    assert len(idxs) != 0, "No possible valid indices have been provided."
    if complement:
        return idxs, np.setdiff1d(np.arange(N), idxs)
    return idxs


# --------------------------------------------------------------------------
# prograph.py:526-588  (distance-to-dataset queries)
# --------------------------------------------------------------------------


def calc_neighbours(tokenized, row, eps=1, distance=hamming, comp=operator.eq):
    """prograph.py:544: column indices where comp(d, eps); no d>0 filter here."""
    d = distance(tokenized, tokenized[row].reshape(1, -1))
    return np.where(comp(d, eps))[1]


def neighbourhood_mask(tokenized, row, eps):
    """prograph.py:587: always Hamming, ``<= eps``."""
    return (hamming(tokenized, np.atleast_2d(tokenized[row])) <= eps).flatten()


def nearest_neighbour(tokenized, query_tokens, distance=hamming):
    """Intended semantics of prograph.py:567-569 (the shipped body raises NameError at
    :565): per query row argmin over the dataset and the overall minimum."""
    d = distance(tokenized, np.atleast_2d(query_tokens))
    return np.argmin(d, axis=1), np.min(d)


# --------------------------------------------------------------------------
# prograph.py:656-765  (build_graph)
# --------------------------------------------------------------------------


def stage_fp16(rep):
    """prograph.py:726: every representation is rounded to fp16 before any metric."""
    return np.asarray(rep).astype(np.float16)


def _cmp_in_dtype(comp, a, b, dtype):
    """torch compares a tensor with a python scalar in the tensor's dtype: a float
    scalar against an fp16 tensor is rounded to fp16 first;
    an integer tensor against a float scalar is compared in float32."""
    if dtype.kind in "iu":
        if isinstance(a, np.ndarray):
            a = a.astype(np.float32) if isinstance(b, float) else a
        else:
            b = b.astype(np.float32) if isinstance(a, float) else b
        return comp(a, b)
    if isinstance(a, np.ndarray):
        return comp(a, dtype.type(b))
    return comp(dtype.type(a), b)


def knn_from_distances(D, k, descending=False):
    """prograph.py:757-762 with the stable tie order: sort each row by (value, index),
    drop sorted position 0 whatever it is, keep the next k."""
    D = np.asarray(D)
    key = D.astype(np.float32) if D.dtype == np.float16 else D
    if descending:
        if key.dtype.kind == "f":
            # NaN first for descending, then large -> small; stable within ties
            order = np.argsort(np.where(np.isnan(key), -np.inf, -key), axis=1, kind="stable")
        else:
            order = np.argsort(-key, axis=1, kind="stable")
    else:
        order = np.argsort(key, axis=1, kind="stable")      # numpy puts NaN last, like torch
    sel = order[:, 1:k + 1]
    return sel.astype(np.int64), np.take_along_axis(D, sel, axis=1)


def build_graph(rep, eps=None, k=None, similarity=False, distance=hamming,
                comp=operator.le, idxs=None, batch_size=8):
    """prograph.py:656-765.  ``rep`` is the (N, D) representation before fp16 staging.
    Returns the reference's list of (indices int64, weights) tuples."""
    if operator.xor(bool(eps), bool(k)) is False:
        raise ValueError("Epsilon or K must be provided, but both cannot be.")
    if k is not None and not isinstance(k, int):
        raise TypeError("K must be provided as an integer.")
    if similarity and eps:
        eps = 1 / (1 + eps)
    X = stage_fp16(rep)
    if idxs is not None:
        X = X[idxs, :]
    N = X.shape[0]
    out = []
    for b0 in range(0, N, batch_size):            # get_every_n, prograph.py:617-624
        D = distance(X, X[b0:b0 + batch_size], similarity=similarity)
        if eps:
            if similarity:
                keep = _cmp_in_dtype(comp, eps, D, D.dtype) & (D < 1)      # :734
            else:
                keep = _cmp_in_dtype(comp, D, eps, D.dtype) & (D > 0)      # :736
            for r in range(D.shape[0]):
                cols = np.where(keep[r])[0].astype(np.int64)
                if len(cols):
                    out.append((cols, D[r, cols]))
                else:                                                     # :753
                    out.append((np.array([], dtype=int), np.array([], dtype=int)))
        else:
            I, W = knn_from_distances(D, k, descending=similarity)         # :757-762
            out.extend((I[r], W[r]) for r in range(D.shape[0]))
    return out


def to_csr(lst):
    indptr = np.zeros(len(lst) + 1, dtype=np.int64)
    for i, (a, _) in enumerate(lst):
        indptr[i + 1] = indptr[i] + len(a)
    idx = np.concatenate([np.asarray(a) for a, _ in lst]) if lst else np.zeros(0, np.int64)
    wl = [np.asarray(w) for _, w in lst if len(w)]
    w = np.concatenate(wl) if wl else np.zeros(0, np.int64)
    return indptr, idx, w


# --------------------------------------------------------------------------
# CPU baseline leg of bench.py: the reference's own loop body with torch CPU ops
# --------------------------------------------------------------------------


def reference_knn_batch_torch(X, batch, k, similarity=False):
    """One iteration of the reference kNN loop (prograph.py:756-762) exactly as it executes
    on the host: the broadcasting compare of hamming.py:34 on fp16-staged tokens, then a full
    sort of every row (stable, the parity contract) and the [:, 1:k+1] slice."""
    import torch
    d = torch.sum(X != batch[:, None, :], axis=2)
    if similarity:
        d = 1 / (1 + d)
    s = torch.sort(d, dim=1, descending=similarity, stable=True)
    return s.indices[:, 1:k + 1].numpy(), s.values[:, 1:k + 1].numpy()


def knn_batch_uint8(tokens_u8, batch_u8, k, threads=1, chunk=32768):
    """Best-effort CPU restatement of one kNN batch (hamming.py:34 + prograph.py:757-762) for the
    bench's second baseline: uint8 compare instead of the fp16 staging, column chunks spread
    over `threads` threads (numpy releases the GIL in the compare / sum), and a partial selection
    of the k+1 smallest (distance, index) keys instead of the reference's full sort.  Same
    result as knn_from_distances(hamming(...), k)."""
    from concurrent.futures import ThreadPoolExecutor
    n = tokens_u8.shape[0]
    bounds = list(range(0, n, chunk)) + [n]

    def part(i):
        a, b = bounds[i], bounds[i + 1]
        return (tokens_u8[None, a:b, :] != batch_u8[:, None, :]).sum(axis=2, dtype=np.int32)

    if threads > 1:
        with ThreadPoolExecutor(threads) as pool:
            parts = list(pool.map(part, range(len(bounds) - 1)))
    else:
        parts = [part(i) for i in range(len(bounds) - 1)]
    D = np.concatenate(parts, axis=1)                                   # (batch, n) int32
    keys = (D.astype(np.int64) << 32) | np.arange(n, dtype=np.int64)[None, :]
    kk = min(k + 1, n)
    sel = np.sort(np.partition(keys, kk - 1, axis=1)[:, :kk], axis=1)[:, 1:]
    return (sel & 0xffffffff).astype(np.int64), (sel >> 32).astype(np.int64)
