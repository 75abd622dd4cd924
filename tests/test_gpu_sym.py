"""Parity of the symmetric kNN build (pg_hamming_knn_boot / pg_hamming_knn_sym /
pg_knn_lists_finalize, prograph_b200/csrc/pg_sweep_sym.cuh) with the oracle, the one-sided fused
sweep and, at the headline size, sampled oracle rows.  Reference semantics: prograph.py:755-765
(sort each row of distances, drop sorted position 0, keep k; ties by ascending index)."""
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
    for world, boot, mode in ((1, 0, 0), (2, 0, 0), (1, 512 if n >= 512 else 0, 0), (2, 0, 1),
                              (8, 512 if n >= 512 else 0, 1)):
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
