"""Parity of the symmetric kNN build (pg_hamming_knn_boot / pg_hamming_knn_sym /
pg_knn_lists_finalize, prograph_b200/csrc/pg_sweep_sym.cuh) with the oracle, the one-sided fused
sweep and, at the headline size, sampled oracle rows.  Reference semantics: prograph.py:755-765
(sort each row of distances, drop sorted position 0, keep k; ties by ascending index)."""
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


def mutational(rng, n, L, alphabet=20, max_mut=8):
    wt = rng.integers(1, alphabet + 1, size=L)
    X = np.tile(wt, (n, 1))
    m = rng.integers(1, max_mut + 1, size=n)
    for j in range(max_mut):
        rows = np.nonzero(m > j)[0]
        pos = rng.integers(0, L, size=len(rows))
        X[rows, pos] = (X[rows, pos] - 1 + rng.integers(1, alphabet, size=len(rows))) % alphabet + 1
    if n > 200:
        X[100] = X[99]            # duplicates: the positional self-drop keeps one of the pair
        X[n - 1] = X[0]
    return X.astype(np.int64)


def sym_knn(eng, tab, k, drop=1, world=1, boot=0, similarity=False, mode=0):
    """All ranks of a `world`-GPU build emulated on one device, merged like graph._hamming_knn_sym
    (mode 0: interleaved row blocks per rank, mode 1: bands of stream rows per rank)."""
    k1 = k + drop
    seed = eng.hamming_knn_boot(tab, 0, tab.rows, boot, k1) if boot else None
    lists = [eng.hamming_knn_sym(tab, k1, r, world, lists=seed.clone() if boot else None, boot_rows=boot, mode=mode)
             for r in range(world)]
    return eng.knn_lists_finalize(torch.stack(lists), 0, tab.rows, k, drop, similarity)


@pytest.mark.parametrize("L,k,alphabet", [(3, 3, 4), (20, 16, 2), (56, 16, 20), (100, 5, 20), (256, 16, 20),
                                          (256, 31, 20), (300, 16, 20), (512, 7, 20), (64, 16, 200)])
@pytest.mark.parametrize("world,boot,mode", [(1, 0, 0), (1, 512, 0), (2, 1024, 0), (3, 0, 1), (4, 512, 1)])
def test_symmetric_knn_uniform_ties(eng, L, k, alphabet, world, boot, mode):
    """iid-uniform tokens: the k-th place of almost every row is a tie, so the (distance, index)
    order of the merged row / column candidates is what is being tested."""
    rng = np.random.default_rng(L * 131 + k)
    n = 2100
    X = rng.integers(1, alphabet + 1, size=(n, L)).astype(np.int64)
    tab = eng.pack(X)
    D = O.hamming(X, X)
    ri, rw = O.knn_from_distances(D, k)
    idx, w = sym_knn(eng, tab, k, world=world, boot=boot, mode=mode)
    np.testing.assert_array_equal(np_(idx), ri)
    np.testing.assert_array_equal(np_(w), rw)
    if world == 1:
        idx, w = sym_knn(eng, tab, k, world=world, boot=boot, similarity=True)
        ri, rw = O.knn_from_distances(O.hamming(X, X, similarity=True), k, descending=True)
        np.testing.assert_array_equal(np_(idx), ri)
        np.testing.assert_array_equal(np_(w), rw)


@pytest.mark.parametrize("n", [5, 256, 257, 513, 3000])
def test_symmetric_knn_ragged_sizes_and_duplicates(eng, n):
    rng = np.random.default_rng(n)
    X = mutational(rng, n, 256)
    tab = eng.pack(X.astype(np.uint8))
    k = min(16, n - 1)
    ri, rw = O.knn_from_distances(O.hamming(X, X), k)
    # 17 ranks: more lists than the merge kernel keeps cursors for (its rescanning variant), and
    # more ranks than row blocks / bands with work
    for world, boot, mode in ((1, 0, 0), (2, 0, 0), (1, 512 if n >= 512 else 0, 0), (2, 0, 1),
                              (8, 512 if n >= 512 else 0, 1), (17, 0, 0), (17, 0, 1)):
        idx, w = sym_knn(eng, tab, k, world=world, boot=boot, mode=mode)
        np.testing.assert_array_equal(np_(idx), ri)
        np.testing.assert_array_equal(np_(w), rw)
    # nearest neighbour including the row itself (drop = 0): position 0 is the smallest (d, index)
    idx, w = sym_knn(eng, tab, 1, drop=0)
    D = O.hamming(X, X)
    np.testing.assert_array_equal(np_(idx)[:, 0], np.argmin(D, axis=1))
    np.testing.assert_array_equal(np_(w)[:, 0], D.min(axis=1))


def test_symmetric_knn_adversarial_descending(eng):
    """Distances to the early rows strictly improve with the column index: every column inserts
    on the row side, and the column side sees its candidates in worst-first order."""
    L, n = 256, 1500
    X = np.ones((n, L), dtype=np.int64)
    for i in range(1, n):
        X[i, : max(0, L - (i * L) // n)] = 2
    tab = eng.pack(X)
    ri, rw = O.knn_from_distances(O.hamming(X, X), 16)
    for boot in (0, 512):
        idx, w = sym_knn(eng, tab, 16, boot=boot)
        np.testing.assert_array_equal(np_(idx), ri)
        np.testing.assert_array_equal(np_(w), rw)


def test_symmetric_knn_degenerate_tables(eng):
    """All rows identical (every distance 0, every comparison a tie decided by the index) and a
    two-point table: the lists are the lowest indices after the positional self-drop."""
    from prograph_b200 import graph
    n, L, k = 5000, 256, 16
    X = np.full((n, L), 7, dtype=np.uint8)
    tab = eng.pack(X)
    want = np.tile(np.arange(1, k + 1), (n, 1))          # sorted (0, index) keys, position 0 dropped
    for world, boot, mode in ((1, 0, 0), (1, 512, 0), (3, 512, 1)):
        idx, w = sym_knn(eng, tab, k, world=world, boot=boot, mode=mode)
        np.testing.assert_array_equal(np_(idx), want)
        assert int(np_(w).max()) == 0
    # two clusters far apart, through the public build above the symmetric threshold
    n = graph.SYM_MIN_ROWS + 700
    X = np.full((n, 64), 3, dtype=np.uint8)
    X[1::2] = 11
    got = graph.build_neighbours(X, k=4)
    even, odd = np.arange(0, n, 2), np.arange(1, n, 2)
    # row r's neighbours: the lowest indices of its own cluster in (0, index) order, sorted position 0 dropped
    np.testing.assert_array_equal(got.idx[0], even[1:5])
    np.testing.assert_array_equal(got.idx[2], even[1:5])         # position 0 is index 0, not the row itself
    np.testing.assert_array_equal(got.idx[n - 1 if (n - 1) % 2 else n - 2], odd[1:5])
    np.testing.assert_array_equal(got.idx[1], odd[1:5])
    assert got.w.max() == 0
    eps = graph.build_neighbours(X[:40000], eps=70)              # every cross-cluster pair: d = 64 <= 70, d > 0
    assert np.all(eps.degrees() == 20000) and np.all(eps.w == 64)
    np.testing.assert_array_equal(eps.idx[:20000], odd[:20000])  # row 0: all odd rows, ascending


@pytest.mark.parametrize("band_mb", ["0", "0.25", "24"])
@pytest.mark.parametrize("L", [40, 100, 256, 300])
def test_symmetric_sweeps_under_every_schedule(eng, monkeypatch, band_mb, L):
    """The work-item schedule of the symmetric sweeps (PG_SYM_BAND_MB: width of the L2 column bands the
    per-CTA item lists are ordered by; 0 = one band, 0.25 = many narrow bands even on these small
    tables) never changes the kNN lists or the epsilon graph: any arrival order gives the oracle's."""
    from prograph_b200.graph import distance_lut
    monkeypatch.setenv("PG_SYM_BAND_MB", band_mb)
    rng = np.random.default_rng(77 + L)
    for n in (513, 3001):
        X = mutational(rng, n, L)
        tab = eng.pack(X.astype(np.uint8))
        D = O.hamming(X, X)
        ri, rw = O.knn_from_distances(D, 16)
        for world, boot, mode in ((1, 0, 0), (1, 512, 0), (2, 0, 1)):
            idx, w = sym_knn(eng, tab, 16, world=world, boot=boot, mode=mode)
            np.testing.assert_array_equal(np_(idx), ri)
            np.testing.assert_array_equal(np_(w), rw)
        keys, edges = eng.hamming_eps_sym(tab, distance_lut(tab.words * 32, operator.le, 3, False))
        ip, ei, ew = eng.edge_keys_to_csr(keys, n, tab.words, edges)
        keep = (D <= 3) & (D > 0)
        r, c = np.nonzero(keep)
        np.testing.assert_array_equal(np_(ip), np.concatenate([[0], np.cumsum(keep.sum(1))]))
        np.testing.assert_array_equal(np_(ei), c)
        np.testing.assert_array_equal(np_(ew), D[r, c])
    U = rng.integers(1, 21, size=(2100, L)).astype(np.int64)          # heavy ties
    tab = eng.pack(U)
    ri, rw = O.knn_from_distances(O.hamming(U, U), 31)
    idx, w = sym_knn(eng, tab, 31, boot=512)
    np.testing.assert_array_equal(np_(idx), ri)
    np.testing.assert_array_equal(np_(w), rw)


def test_exchange_merge_then_widen_equals_direct_finalize(eng):
    """The multi-GPU exchange path: per-rank lists -> pg_knn_lists_merge (8-byte keys of a row block)
    -> pg_knn_lists_finalize(n_lists=1, drop=0), against the direct G-way finalize."""
    rng = np.random.default_rng(9)
    n, k = 3000, 16
    X = mutational(rng, n, 256)
    tab = eng.pack(X.astype(np.uint8))
    seed = eng.hamming_knn_boot(tab, 0, n, 512, k + 1)
    lists = torch.stack([eng.hamming_knn_sym(tab, k + 1, r, 3, lists=seed.clone(), boot_rows=512, mode=1) for r in range(3)])
    want_i, want_w = eng.knn_lists_finalize(lists, 0, n, k, 1)
    for sim in (False, True):
        parts = []
        for r0, r1 in ((0, 1024), (1024, 2048), (2048, n)):
            keys = eng.knn_lists_merge(lists[:, r0:r1].contiguous(), k, drop=1)
            assert keys.shape == (r1 - r0, k)
            parts.append(keys)
        keys = torch.cat(parts)
        gi, gw = eng.knn_lists_finalize(keys, 0, n, k, 0, sim)
        np.testing.assert_array_equal(np_(gi), np_(want_i))
        ri, rw = O.knn_from_distances(O.hamming(X, X, similarity=sim), k, descending=sim)
        np.testing.assert_array_equal(np_(gi), ri)
        np.testing.assert_array_equal(np_(gw), rw)
    eng.sym_check()                                                      # no lock was ever given up


def test_symmetric_unsupported_shapes_fall_back(eng):
    """Lists longer than 32 entries are not covered: the engine says so and build_neighbours
    takes the one-sided sweep (never a CPU path)."""
    from prograph_b200 import _lib, graph
    rng = np.random.default_rng(1)
    X = rng.integers(1, 21, size=(1200, 64)).astype(np.int64)
    tab = eng.pack(X)
    with pytest.raises(_lib.Unsupported):
        eng.hamming_knn_sym(tab, 40)
    old = graph.SYM_MIN_ROWS
    graph.SYM_MIN_ROWS = 0
    try:
        got = graph.build_neighbours(X, k=40)
        got16 = graph.build_neighbours(X, k=16)
    finally:
        graph.SYM_MIN_ROWS = old
    D = O.hamming(X, X)
    for res, k in ((got, 40), (got16, 16)):
        ri, rw = O.knn_from_distances(D, k)
        np.testing.assert_array_equal(res.idx, ri)
        np.testing.assert_array_equal(res.w, rw)


def test_public_build_crosses_into_symmetric_path(eng):
    """build_neighbours on a table just above graph.SYM_MIN_ROWS: the whole result equals the
    one-sided sweep's, sampled rows equal the oracle."""
    from prograph_b200 import graph
    rng = np.random.default_rng(2)
    n, L = graph.SYM_MIN_ROWS + 1500, 56
    X = mutational(rng, n, L, max_mut=4).astype(np.uint8)
    eng.time_sweeps(True)
    eng.sweep_times(reset=True)
    got = graph.build_neighbours(X, k=16)
    assert len(eng.sweep_times(reset=True)) == 2          # bootstrap sweep + symmetric sweep
    eng.time_sweeps(False)
    tab = eng.pack(X)
    idx, w = eng.hamming_knn(tab, 0, n, tab, 16, drop=1)
    np.testing.assert_array_equal(got.idx, np_(idx))
    np.testing.assert_array_equal(got.w, np_(w))
    sample = rng.choice(n, size=48, replace=False)
    ri, rw = O.knn_from_distances(O.hamming(X.astype(np.int64), X[sample].astype(np.int64)), 16)
    np.testing.assert_array_equal(got.idx[sample], ri)
    np.testing.assert_array_equal(got.w[sample], rw)
    sim = graph.build_neighbours(X, k=3, similarity=True)
    np.testing.assert_array_equal(sim.idx, got.idx[:, :3])
    np.testing.assert_array_equal(sim.w, (np.float32(1) / (1 + got.w[:, :3]).astype(np.float32)).astype(np.float32))


# ------------------------------------------------------------------ symmetric epsilon graph
def sym_eps(eng, tab, lut, similarity=False, world=1, mode=0, capacity=None):
    """All ranks of a `world`-GPU symmetric epsilon build emulated on one device (graph.hamming_eps_graph)."""
    keys, edges = [], 0
    for r in range(world):
        k, e = eng.hamming_eps_sym(tab, lut, r, world, mode=mode, capacity=capacity)
        keys.append(k)
        edges += e
    return eng.edge_keys_to_csr(torch.cat(keys), tab.rows, tab.words, edges, similarity)


@pytest.mark.parametrize("L,eps,alphabet", [(3, 1, 4), (20, 12, 4), (56, 2, 20), (100, 3, 20), (256, 4, 20), (300, 5, 20),
                                            (64, 5, 200)])
@pytest.mark.parametrize("world,mode", [(1, 0), (2, 0), (3, 1)])
def test_symmetric_eps_vs_oracle(eng, L, eps, alphabet, world, mode):
    """prograph.py:731-753: neighbours in ascending index order, weights = distances (or float32
    similarities), `d > 0` guard (duplicates are not neighbours)."""
    from prograph_b200.graph import distance_lut
    rng = np.random.default_rng(L + eps)
    n = 2100
    X = mutational(rng, n, L, alphabet=alphabet, max_mut=min(8, L)) if L >= 56 else \
        rng.integers(1, alphabet + 1, size=(n, L)).astype(np.int64)
    tab = eng.pack(X)
    for similarity in (False, True):
        e = 1 / (1 + eps) if similarity else eps
        lut = distance_lut(tab.words * 32, operator.le, e, similarity)
        indptr, idx, w = O.to_csr(O.build_graph(X, eps=eps, similarity=similarity))
        # a 1024-slot buffer is too small for every case here: the sweep reports what it needs and runs again
        for capacity in (None, 1024):
            gi, gx, gw = sym_eps(eng, tab, lut, similarity, world, mode, capacity)
            np.testing.assert_array_equal(np_(gi), indptr)
            np.testing.assert_array_equal(np_(gx), idx)
            np.testing.assert_array_equal(np_(gw), w)
            assert np_(gw).dtype == (np.float32 if similarity else np.int64)


def test_symmetric_eps_other_comparisons_and_empty(eng):
    from prograph_b200 import _lib
    from prograph_b200.graph import distance_lut
    rng = np.random.default_rng(9)
    X = rng.integers(1, 5, size=(1500, 20)).astype(np.int64)
    tab = eng.pack(X)
    D = O.hamming(X, X)
    for comp, eps in ((operator.ge, 18), (operator.gt, 17), (operator.lt, 9), (operator.eq, 10)):
        lut = distance_lut(tab.words * 32, comp, eps, False)
        gi, gx, gw = sym_eps(eng, tab, lut)
        keep = comp(D, eps) & (D > 0)
        rows, cols = np.nonzero(keep)
        np.testing.assert_array_equal(np_(gi), np.concatenate([[0], np.cumsum(keep.sum(1))]))
        np.testing.assert_array_equal(np_(gx), cols)
        np.testing.assert_array_equal(np_(gw), D[rows, cols])
    # nothing passes: all-zero indptr, empty arrays
    gi, gx, gw = sym_eps(eng, tab, distance_lut(tab.words * 32, operator.gt, 25, False))
    assert np_(gi).tolist() == [0] * 1501 and gx.numel() == 0 and gw.numel() == 0
    # `!=` is not one range of distances: the symmetric sweep declines, the caller takes count / fill
    with pytest.raises(_lib.Unsupported):
        eng.hamming_eps_sym(tab, distance_lut(tab.words * 32, operator.ne, 10, False))


def test_gb1_library_symmetric_eps_known_answers(eng):
    """C3: the full 20^4 library, L=56.  eps=1: every degree = 4*19 = 76, nnz =
    12 160 000, through the public build (symmetric sweep above graph.SYM_EPS_MIN_ROWS); the whole
    CSR equals the one-sided count / fill passes'.  eps=2 (degree 2242) is dense: one-sided path."""
    import itertools
    from prograph_b200 import graph
    wt = np.frombuffer(b"MTYKLILNGKTLKGETTTEAVDAATAEKVFKQYANDNGVDGEWTYDDATKTFTVTE", dtype=np.uint8)
    lut256 = np.zeros(256, dtype=np.uint8)
    for i, ch in enumerate("ACDEFGHIKLMNPQRSTVWY"):
        lut256[ord(ch)] = i + 1
    combos = np.array(list(itertools.product(range(1, 21), repeat=4)), dtype=np.uint8)
    X = np.tile(lut256[wt], (len(combos), 1))
    X[:, [38, 39, 40, 53]] = combos
    n = len(X)
    assert n == 160_000 >= graph.SYM_EPS_MIN_ROWS
    eng.time_sweeps(True)
    eng.sweep_times(reset=True)
    got = graph.build_neighbours(X, eps=1)
    assert len(eng.sweep_times(reset=True)) == 2                 # degree sample + symmetric sweep
    eng.time_sweeps(False)
    assert np.all(got.degrees() == 76) and len(got.idx) == 12_160_000
    assert np.all(got.w == 1)
    tab = eng.pack(X)
    lut = graph.distance_lut(tab.words * 32, operator.le, 1, False)
    ip, ix, w = eng.hamming_eps(tab, 0, n, tab, lut)
    np.testing.assert_array_equal(got.indptr, np_(ip))
    np.testing.assert_array_equal(got.idx, np_(ix))
    np.testing.assert_array_equal(got.w, np_(w))
    assert eng.hamming_eps_mean_degree(tab, *graph._eps_sample(n), tab,
                                       graph.distance_lut(tab.words * 32, operator.le, 2, False)) == (2242.0, 2242)
