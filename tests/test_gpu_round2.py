"""Round-2 additions on the GPU: vectorised pack / mask kernels against the ballot kernels and the
oracle, batched neighbourhood queries and clustering, the query module at BASELINE.json configs[4]
scale (100 000 queries x 1 000 000 library rows, L=256, every metric of prograph/distance), and the
error word of the symmetric sweep.  Reference lines: hamming.py:34-38, minkowski.py:36-40,
prograph.py:488-492, 526-615."""
import functools
import operator

import numpy as np
import pytest
import torch

from oracle import prograph_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from prograph_b200.engine import get_engine
    return get_engine()


def np_(t):
    return t.cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


# ------------------------------------------------------------------ informative columns
@pytest.mark.parametrize("L,alphabet,n", [(56, 20, 3000), (256, 20, 5000), (300, 200, 1500), (33, 20, 7), (1000, 20, 900),
                                          (2000, 20, 600)])
def test_varying_and_compacted_columns(eng, L, alphabet, n):
    """pg_varying_columns against numpy and pg_compact_columns against a pack of the selected token
    columns, for 5 and 8 planes, vector and scalar row lengths, short last words."""
    rng = np.random.default_rng(L + n)
    wt = rng.integers(1, alphabet + 1, size=L)
    X = np.tile(wt, (n, 1))
    sites = np.sort(rng.choice(L, size=max(1, L // 7), replace=False))
    X[:, sites] = rng.integers(1, alphabet + 1, size=(n, len(sites)))
    X[0] = wt
    X[n - 1, sites[-1]] = wt[sites[-1]] % alphabet + 1             # the last row alone makes a column vary
    tab = eng.pack(X.astype(np.uint8))
    cols = eng.varying_columns(tab)
    want = np.nonzero((X != X[0]).any(axis=0))[0]
    np.testing.assert_array_equal(cols, want)
    small = eng.compact_columns(tab, cols)
    ref = eng.pack(X[:, want].astype(np.uint8), planes=tab.planes)
    assert (small.planes, small.words, small.rows) == (ref.planes, ref.words, n)
    assert torch.equal(small.data, ref.data)
    # identical rows: nothing varies
    same = eng.pack(np.tile(wt, (50, 1)).astype(np.uint8))
    assert len(eng.varying_columns(same)) == 0


@pytest.mark.parametrize("n,L,sites", [(6000, 56, 4), (70000, 200, 30), (5000, 1900, 100)])
def test_graph_builds_on_informative_columns_match_the_oracle(eng, n, L, sites):
    """build_neighbours on libraries with few varying positions (graph.informative_table drops the
    constant ones: W=2 -> 1, W=8 -> 1, and a 1900-residue table that no fused sweep covers comes back
    into range): kNN and epsilon CSR against the C oracle on the FULL rows."""
    from oracle import c_oracle as CO
    from prograph_b200 import build_neighbours, graph
    rng = np.random.default_rng(n + L)
    wt = rng.integers(1, 21, size=L)
    X = np.tile(wt, (n, 1)).astype(np.uint8)
    pos = np.sort(rng.choice(L, size=sites, replace=False))
    muts = rng.integers(0, 4, size=(n, sites)) == 0                # each site mutated with probability 1/4
    X[:, pos] = np.where(muts, rng.integers(1, 21, size=(n, sites)), wt[pos]).astype(np.uint8)
    tab = eng.pack(X)
    small = graph.informative_table(eng, tab)
    assert small.L <= sites and small.words == eng.packed_words(small.L) < tab.words
    srows = np.unique(np.concatenate([[0, n - 1], rng.choice(n, size=40, replace=False)]))
    D = CO.hamming_rows(CO.pack(X), L, srows)
    knn = build_neighbours(X, k=8)
    for i, r in enumerate(srows):
        order = np.lexsort((np.arange(n), D[i]))[1:9]
        np.testing.assert_array_equal(knn.idx[r], order)
        np.testing.assert_array_equal(knn.w[r], D[i, order])
    g = build_neighbours(X, eps=1)
    for i, r in enumerate(srows):
        cols = np.nonzero((D[i] <= 1) & (D[i] > 0))[0]
        a, b = g.indptr[r], g.indptr[r + 1]
        np.testing.assert_array_equal(g.idx[a:b], cols)
        np.testing.assert_array_equal(g.w[a:b], D[i, cols])


# ------------------------------------------------------------------ any k, like the reference's whole-row sort
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.float64, torch.int32, torch.int64])
@pytest.mark.parametrize("descending", [False, True])
def test_tile_topk_beyond_the_shared_memory_sort(eng, dtype, descending):
    """k + drop > 4096: pg_tile_topk sorts whole rows (two stable radix sorts); same (value, index)
    order as torch.sort(stable=True) (prograph.py:757-762), heavy ties included."""
    rng = np.random.default_rng(3)
    rows, n, k, drop = 5, 9000, 5000, 1
    vals = rng.integers(0, 40, size=(rows, n))                     # 40 distinct values: long runs of ties
    tile = torch.from_numpy(vals).to(eng.device).to(dtype)
    idx, val = eng.tile_topk(tile, k, drop=drop, descending=descending)
    sv, si = torch.sort(tile.cpu().to(torch.float64), dim=1, stable=True, descending=descending)
    assert torch.equal(idx.cpu(), si[:, drop:drop + k])
    assert torch.equal(val.cpu().to(torch.float64), sv[:, drop:drop + k])
    # k beyond the row: the reference's slice simply ends; missing slots are marked -1
    idx, _ = eng.tile_topk(tile[:2, :4500].contiguous(), 4499, drop=1, descending=descending)
    assert idx.shape == (2, 4499) and int((idx < 0).sum()) == 0


def test_build_graph_accepts_any_k(eng):
    """build_neighbours(k=4200) on 4 300 rows: Hamming on tokens (k beyond the fused lists -> int64 tiles
    -> whole-row sort) and Minkowski, against the oracle's stable sort."""
    from prograph_b200 import build_neighbours, minkowski
    rng = np.random.default_rng(8)
    n, L, k = 4300, 40, 4200
    X = rng.integers(1, 21, size=(n, L))
    D = O.hamming(X, X)
    got = build_neighbours(X, k=k)
    ri, rw = O.knn_from_distances(D, k)
    assert np.array_equal(got.idx, ri) and np.array_equal(got.w, rw)
    Dm = O.minkowski(X.astype(np.float16), X.astype(np.float16))
    gm = build_neighbours(X, k=k, distance=minkowski)
    mi, mw = O.knn_from_distances(Dm, k)
    assert np.array_equal(gm.idx, mi) and np.array_equal(gm.w, mw)


# ------------------------------------------------------------------ dense epsilon graphs: captures sized from the degree sample
@pytest.mark.parametrize("L,alphabet,eps", [(8, 3, 4), (40, 2, 18), (300, 2, 145), (600, 2, 290)])
def test_eps_count_pass_keeps_every_hit_when_the_workspace_allows(eng, L, alphabet, eps):
    """pg_eps_workspace_bytes_capture: with room for the largest degree the count pass keeps every hit
    and the fill pass is a copy; with too little room the overflowing rows get the fill sweep.  Same
    CSR as the default (128-slot) workspace and as the oracle (prograph.py:731-753)."""
    from prograph_b200 import graph
    rng = np.random.default_rng(L)
    n = 3000
    X = rng.integers(1, alphabet + 1, size=(n, L))
    tab = eng.pack(X)
    lut = graph.distance_lut(tab.words * 32, operator.le, eps, False)
    ref = eng.hamming_eps(tab, 0, n, tab, lut)
    D = O.hamming(X, X)
    keep = (D <= eps) & (D > 0)
    deg = keep.sum(1)
    assert deg.max() > 400                                    # denser than the default captures
    np.testing.assert_array_equal(np_(ref[0]), np.concatenate([[0], np.cumsum(deg)]))
    np.testing.assert_array_equal(np_(ref[1]), np.nonzero(keep)[1])
    for capture in (int(deg.max()), int(np.median(deg)), 130, 1 << 20):
        got = eng.hamming_eps(tab, 0, n, tab, lut, capture=capture)
        for a, b in zip(got, ref):
            assert torch.equal(a, b), capture
    sim = eng.hamming_eps(tab, 5, 1000, tab, lut, similarity=True, capture=int(deg.max()))
    ref_sim = eng.hamming_eps(tab, 5, 1000, tab, lut, similarity=True)
    for a, b in zip(sim, ref_sim):
        assert torch.equal(a, b)


def test_dense_epsilon_build_uses_the_sampled_degree(eng):
    """build_neighbours on a table whose eps graph is dense (sampled mean degree > 384): the one-sided
    count / fill path with captures sized from the sample, against the oracle on sampled rows."""
    from oracle import c_oracle as CO
    from prograph_b200 import build_neighbours
    rng = np.random.default_rng(12)
    n, L = 40000, 12
    X = rng.integers(1, 4, size=(n, L)).astype(np.uint8)
    g = build_neighbours(X, eps=4)
    srows = np.unique(np.concatenate([[0, n - 1], rng.choice(n, size=30, replace=False)]))
    D = CO.hamming_rows(CO.pack(X), L, srows)
    assert np.diff(g.indptr).mean() > 384
    for i, r in enumerate(srows):
        cols = np.nonzero((D[i] <= 4) & (D[i] > 0))[0]
        a, b = g.indptr[r], g.indptr[r + 1]
        np.testing.assert_array_equal(g.idx[a:b], cols)
        np.testing.assert_array_equal(g.w[a:b], D[i, cols])


# ------------------------------------------------------------------ pack / masks
@pytest.mark.parametrize("L", [1, 31, 32, 33, 56, 100, 255, 256, 257, 600])
@pytest.mark.parametrize("alphabet", [20, 200])
def test_vectorised_pack_equals_ballot_pack(eng, L, alphabet):
    """uint8 tokens take pack_bytes_kernel (128-bit / 32-bit / byte loads depending on the row pitch,
    one multiply per 4 tokens and plane); int64 tokens take the ballot kernel.  Same planes, for
    5 planes (alphabet 20) and 8 planes (alphabet 200), ragged widths and short last words."""
    rng = np.random.default_rng(L * 7 + alphabet)
    n = 777
    X = rng.integers(0, alphabet + 1, size=(n, L))
    a = eng.pack(X.astype(np.uint8))
    b = eng.pack(X.astype(np.int64))
    assert (a.planes, a.words) == (b.planes, b.words) and a.planes == (5 if alphabet < 32 else 8)
    assert torch.equal(a.data, b.data)
    # rows of the padding (and words past L) are zero
    assert int(a.data[n:].abs().sum().item()) == 0
    if a.planes == 5 or a.words <= 8:                      # the shapes the fused sweep covers
        D = O.hamming(X, X[:5])
        np.testing.assert_array_equal(np_(eng.hamming_tile(a, a, 0, 512))[:5], D)


def test_vectorised_pack_flags_tokens_that_do_not_fit(eng):
    X = np.full((40, 64), 7, dtype=np.uint8)
    X[17, 33] = 40                                       # does not fit 5 planes
    flag = torch.zeros(1, dtype=torch.int32, device=eng.device)
    from prograph_b200 import _lib as L
    from prograph_b200.engine import _ptr
    t = eng.to_device(X)
    out = eng.empty((512, 5, 2), torch.int32)
    L.check(eng.lib.pg_pack_tokens(_ptr(t), L.U8, 40, 64, 64, _ptr(out), 5, 2, _ptr(flag), eng._stream()))
    assert int(flag.item()) == 1
    tab = eng.pack(X)                                    # the engine retries with 8 planes
    assert tab.planes == 8
    with pytest.raises(OverflowError):
        eng.pack(X, planes=5)


@pytest.mark.parametrize("L", [3, 33, 56, 100, 256])
def test_pack_chars_and_mutant_bool(eng, L):
    """Letters -> planes in one pass equals tokenise -> pack; the (N, L) boolean mutant array equals
    prograph.py:488-492 for row pitches with and without 128-bit stores."""
    rng = np.random.default_rng(L)
    letters = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
    n = 1500
    chars = letters[rng.integers(0, 20, size=(n, L))]
    short = rng.random(n) < 0.2                           # ragged library: zero bytes pad shorter strings
    for r in np.nonzero(short)[0]:
        chars[r, rng.integers(1, L + 1):] = 0
    lut = np.zeros(256, dtype=np.uint8)
    lut[letters] = np.arange(1, 21)
    tok = lut[chars].astype(np.int64)
    a = eng.pack_chars(chars, lut)
    b = eng.pack(tok, planes=5)
    assert torch.equal(a.data, b.data)
    for ref in (0, 7):
        want = O.boolean_mutant_array(tok, ref)
        got = np_(eng.mutant_bool(a, a.row(ref))).view(np.bool_)
        np.testing.assert_array_equal(got, want)


# ------------------------------------------------------------------ batched neighbourhoods
def test_flags_tile_and_clustering_batches(eng, tmp_path):
    from prograph_b200 import Prograph
    rng = np.random.default_rng(4)
    n, L = 3000, 40
    wt = rng.integers(0, 20, size=L)
    X = np.tile(wt, (n, 1))
    for i in range(1, n):
        pos = rng.choice(L, size=rng.integers(1, 6), replace=False)
        X[i, pos] = (X[i, pos] + rng.integers(1, 20, size=len(pos))) % 20
    alphabet = "ACDEFGHIKLMNPQRSTVWY"
    seqs = ["".join(alphabet[t] for t in row) for row in X]
    seqs = list(dict.fromkeys(seqs))                       # the frame's sequence -> row map wants unique rows
    import pandas as pd
    path = tmp_path / "lib.csv"
    pd.DataFrame({"Sequence": seqs, "Fitness": np.arange(len(seqs), dtype=float)}).to_csv(path)
    pg = Prograph(str(path))
    tok = pg.tokenized
    table = pg._device_tokens()
    front = np.array([0, 5, 17, len(seqs) - 1])
    flags = np_(eng.hamming_flags(table, eng.gather_packed(table, front), len(front), 0, 3))
    np.testing.assert_array_equal(flags.astype(bool), O.hamming(tok, tok[front]) <= 3)
    assert eng.hamming_flags(table, eng.gather_packed(table, front), len(front), 2, 1).sum().item() == 0
    ref = None
    for batch in (1, 7, 128):
        clusters = pg.neighbourhood_clustering(3, batch=batch)
        covered = np.zeros(len(pg), dtype=bool)
        for seed, members in clusters.items():             # insertion order = the reference's greedy order
            assert not covered[seed]
            np.testing.assert_array_equal(np.asarray(members.index), np.where(O.neighbourhood_mask(tok, seed, 3))[0])
            covered[np.asarray(members.index)] = True
        assert covered.all()
        keys = list(clusters)
        assert keys == sorted(keys)
        assert ref is None or keys == ref
        ref = keys


# ------------------------------------------------------------------ configs[4]: 100k queries x 1M library
def _mutational(n, L, seed):
    from bench import make_tokens
    return make_tokens(n, L, "mutational", seed=seed)


def test_config5_queries_every_metric(eng):
    """BASELINE.json configs[4] / SURVEY.md §8(d) C5: library = C4-M (1M x 256), 100 000 queries from
    the same wild type (default_rng(1)).  Every consumer of the query module against sampled oracle
    rows: Hamming argmin/min, Hamming similarity tile, `<= eps` counts, materialised (1024, N) tiles,
    Minkowski p=2 (int64 in -> float32 out, +- similarity) argmin and tiles, p in {1, 3} on a
    subsample (the no-abs quirk: signed sums, NaN for negative cubes)."""
    from oracle import c_oracle as CO
    from prograph_b200 import hamming, minkowski, query
    n, m, L = 1_000_000, 100_000, 256
    X = _mutational(n, L, 0)
    Q = _mutational(m, L, 1)
    assert np.array_equal(X[0][X[0] == Q[0]], Q[0][X[0] == Q[0]])
    lib = query.Library(X)
    rng = np.random.default_rng(2)
    sample = np.sort(rng.choice(m, size=24, replace=False))
    sample[0], sample[-1] = 0, m - 1
    planes = CO.pack(np.concatenate([X, Q[sample]]))
    D = CO.hamming_rows(planes, L, n + np.arange(len(sample)), threads=16)[:, :n].astype(np.int64)   # (24, n)
    del planes
    # Hamming: argmin / min for all 100k queries
    idx, d = query.nearest(lib, Q)
    assert idx.shape == (m,) and d.dtype == torch.int64
    np.testing.assert_array_equal(np_(idx)[sample], D.argmin(axis=1))
    np.testing.assert_array_equal(np_(d)[sample], D.min(axis=1))
    # <= eps counts (calc_neighbours semantics: no d > 0 filter), two predicates
    for eps, comp in ((3, operator.le), (4, operator.eq)):
        cnt = query.count_within(lib, Q, eps, comp)
        np.testing.assert_array_equal(np_(cnt)[sample], comp(D, eps).sum(axis=1))
    # materialised (1024, N) tiles: the public distance functions on the first 1024 queries
    first = np.arange(0, 1024, 128)
    D1 = CO.hamming_rows(CO.pack(np.concatenate([X, Q[first]])), L, n + np.arange(len(first)), threads=16)[:, :n]
    T = query.tile(lib, Q, 0, 1024)
    assert T.shape == (1024, n) and T.dtype == torch.int64
    np.testing.assert_array_equal(np_(T[torch.as_tensor(first, device=T.device)]), D1)
    S = query.tile(lib, Q, 0, 1024, similarity=True)
    assert S.dtype == torch.float32
    np.testing.assert_array_equal(np_(S[torch.as_tensor(first, device=S.device)]),
                                  (np.float32(1) / (1 + D1).astype(np.float32)).astype(np.float32))
    del T, S
    # the drop-in distance function itself on a few queries against the whole library
    D8 = CO.hamming_rows(CO.pack(np.concatenate([X, Q[:8]])), L, n + np.arange(8), threads=16)[:, :n]
    np.testing.assert_array_equal(np_(hamming(X, Q[:8])), D8)
    # Minkowski p=2 on int64 tokens: exact integer sum, float32 root (minkowski.py:36-40)
    X64 = X.astype(np.int64)
    few = sample[:6]
    S2 = np.stack([((X.astype(np.int32) - Q[q].astype(np.int32)) ** 2).sum(axis=1, dtype=np.int64) for q in few])
    root = np.sqrt(S2.astype(np.float32)).astype(np.float32)            # IEEE-rounded, as torch's CUDA sqrt
    lib64 = query.Library(X64)
    mi, mv = query.nearest(lib64, Q.astype(np.int64), distance=minkowski)
    assert mv.dtype == torch.float32
    np.testing.assert_array_equal(np_(mi)[few], np.argmin(S2, axis=1))
    np.testing.assert_array_equal(np_(mv)[few], root.min(axis=1))
    fewq = Q[few].astype(np.int64)
    T2 = query.tile(lib64, fewq, distance=minkowski)
    np.testing.assert_array_equal(np_(T2), root)
    T2s = query.tile(lib64, fewq, distance=minkowski, similarity=True)
    np.testing.assert_array_equal(np_(T2s), (np.float32(1) / (np.float32(1) + root)).astype(np.float32))
    # p = 1 and p = 3 (no abs): signed sums; cube roots of negative sums are NaN -- subsample of the library
    sub = X64[:20_000]
    for p in (1, 3):
        want = np_(O.minkowski(sub, fewq, p=p))
        got = np_(query.tile(query.Library(sub), fewq, distance=functools.partial(minkowski, p=p)))
        assert got.dtype == np.float32
        if p == 1:
            np.testing.assert_array_equal(got, (sub[None, :, :] - fewq[:, None, :]).sum(axis=2).astype(np.float32))
        else:
            s3 = ((sub[None, :, :] - fewq[:, None, :]) ** 3).sum(axis=2)
            assert np.array_equal(np.isnan(got), s3 < 0)
            ok = s3 >= 0
            np.testing.assert_allclose(got[ok], np.cbrt(s3[ok].astype(np.float64)), rtol=4e-7)
        np.testing.assert_allclose(got, want, rtol=4e-7, equal_nan=True)


def test_sym_status_reports_ok_and_knows_its_workspace(eng):
    rng = np.random.default_rng(3)
    X = rng.integers(1, 21, size=(5000, 64)).astype(np.uint8)
    tab = eng.pack(X)
    eng.hamming_knn_sym(tab, 17)
    eng.sym_check()                 # no lock was given up
    eng.sym_check()                 # nothing pending: a no-op


def test_integration_stub_as_documented(eng):
    """INTEGRATION.md §2 shows the ctypes stub a maintainer of the reference would add.  Run it exactly
    as printed (only the library path is rewritten) and compare with the oracle."""
    import os
    import re
    from prograph_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = next(b for b in blocks if "# prograph/_b200.py" in b)
    stub = stub.replace('C.CDLL("libprograph_b200.so")', f'C.CDLL({_lib.LIB_PATH!r})')
    ns = {}
    exec(compile(stub, "INTEGRATION.md:_b200.py", "exec"), ns)
    rng = np.random.default_rng(12)
    X = rng.integers(1, 21, size=(1500, 100))
    Y = rng.integers(1, 21, size=(37, 100))
    Xd, Yd = torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda()
    np.testing.assert_array_equal(np_(ns["hamming_cuda"](Xd, Yd)), O.hamming(X, Y))
    np.testing.assert_array_equal(np_(ns["hamming_cuda"](Xd, Yd, similarity=True)), O.hamming(X, Y, similarity=True))
    np.testing.assert_array_equal(np_(ns["hamming_cuda"](Xd.to(torch.float16), Yd.to(torch.float16))), O.hamming(X, Y))
    idx, w = ns["knn_cuda"](Xd, 16)
    ri, rw = O.knn_from_distances(O.hamming(X, X), 16)
    np.testing.assert_array_equal(np_(idx), ri)
    np.testing.assert_array_equal(np_(w), rw)
    with pytest.raises(OverflowError):
        ns["pack"](torch.full((4, 8), 40, dtype=torch.int64, device="cuda"))


@pytest.mark.parametrize("dtype", ["int64", "float16", "float32"])
def test_minkowski_p1_rank_one_tile_and_its_fallback(eng, dtype):
    """p = 1 without abs is sum(x) - sum(y) (minkowski.py:36): integer-valued rows take the rank-1 tile
    (row sums + outer difference), rows with fractional or large values are recognised on the device
    and take the element-wise kernel; both must give the reference's rounding chain bit for bit."""
    from prograph_b200 import minkowski
    rng = np.random.default_rng(5)
    X = rng.integers(0, 21, size=(900, 256))
    Y = rng.integers(0, 21, size=(70, 256))
    cast = {"int64": np.int64, "float16": np.float16, "float32": np.float32}[dtype]
    for sim in (False, True):
        got = np_(minkowski(X.astype(cast), Y.astype(cast), p=1, similarity=sim))
        want = O.minkowski(X.astype(cast), Y.astype(cast), p=1, similarity=sim)
        assert got.dtype == want.dtype
        np.testing.assert_array_equal(got, want)
    if dtype != "int64":
        Xf = (rng.integers(-40, 40, size=(300, 7)) / 8.0).astype(cast)          # fractional: not a rank-1 case
        Yf = (rng.integers(-40, 40, size=(33, 7)) / 8.0).astype(cast)
        np.testing.assert_array_equal(np_(minkowski(Xf, Yf, p=1)), O.minkowski(Xf, Yf, p=1))
        Xb = X.astype(cast) * 20                                                  # integers, but beyond 255
        got = np_(minkowski(Xb, Y.astype(cast), p=1))
        want = O.minkowski(Xb, Y.astype(cast), p=1)
        if dtype == "float32":
            np.testing.assert_allclose(got, want, rtol=3e-7)                      # fp32 sum order
        else:
            np.testing.assert_array_equal(got, want)
    else:
        big = rng.integers(-2**40, 2**40, size=(50, 9))
        np.testing.assert_array_equal(np_(minkowski(big, big[:7], p=1)), O.minkowski(big, big[:7], p=1))


@pytest.mark.parametrize("p", [1, 2, 3])
def test_fp16_chain_in_native_half_arithmetic(eng, p):
    """The element-wise fp16 path runs its per-element chain (x - y, powers) in native half2
    arithmetic.  With two components per row the fp32 sum has no order to disagree on, so the result
    must equal the reference's rounding chain bit for bit -- including overflow to inf, subnormal
    differences, negative bases under the odd exponents (NaN roots), signed zeros."""
    from prograph_b200 import minkowski
    rng = np.random.default_rng(40 + p)
    X = (rng.standard_normal((3000, 2)) * rng.choice([1e-4, 0.1, 1.0, 30.0, 300.0], size=(3000, 1))).astype(np.float16)
    Y = (rng.standard_normal((257, 2)) * rng.choice([1e-4, 0.1, 1.0, 30.0, 300.0], size=(257, 1))).astype(np.float16)
    X[:4] = np.array([[65504, 65504], [-65504, 1], [6e-8, -6e-8], [0.0, -0.0]], dtype=np.float16)
    Y[:2] = np.array([[-65504, -65504], [6e-8, 6e-8]], dtype=np.float16)
    for sim in (False, True):
        with np.errstate(all="ignore"):
            want = O.minkowski(X, Y, p=p, similarity=sim)
        got = np_(minkowski(X, Y, p=p, similarity=sim))
        assert got.dtype == np.float16
        ok = ~np.isnan(want)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        ulp = np.abs(got.view(np.int16)[ok].astype(np.int32) - want.view(np.int16)[ok].astype(np.int32))
        # p = 1, 2: bit for bit.  p = 3: the cube root is powf(s, fp16(1/3)) on both sides, but torch's pow and
        # CUDA's powf are different implementations: 1 fp16 ulp, as stated in DESIGN.md §2 (only the root).
        assert ulp.max() <= (1 if p == 3 else 0), int(ulp.max())
    # wide rows of small integers: every partial sum is exact, so the order cannot matter either
    Xi = rng.integers(-6, 7, size=(700, 300)).astype(np.float16)
    Yi = rng.integers(-6, 7, size=(65, 300)).astype(np.float16)
    with np.errstate(all="ignore"):
        want = O.minkowski(Xi, Yi, p=p)
    got = np_(minkowski(Xi, Yi, p=p))
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    ulp = np.abs(got.view(np.int16)[ok].astype(np.int32) - want.view(np.int16)[ok].astype(np.int32))
    assert ulp.max() <= (1 if p == 3 else 0), int(ulp.max())


def test_edge_budget_refuses_graphs_that_cannot_fit(eng):
    """A graph whose CSR cannot fit in device memory is refused with MemoryError before anything is
    allocated (the C4-M eps=2 graph of the bench: 1.6e10 edges = 257 GB); small requests pass, also
    when the cached free-memory reading is used."""
    eng.check_edge_budget(1000)
    eng.check_edge_budget(1000)
    with pytest.raises(MemoryError, match="does not fit in device memory"):
        eng.check_edge_budget(16_071_968_078)
