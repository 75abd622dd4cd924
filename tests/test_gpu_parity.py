"""Parity of the CUDA path with the oracle and the reference's golden vectors.

Every test here runs the sm_100a kernels through the C ABI (prograph_b200 -> ctypes ->
libprograph_b200.so) and compares with oracle/prograph_oracle.py on the same seeded inputs
and with tests/golden/*.npz.  Bar: bit-exact for integer / index work and for fp16 results;
float32 roots within 1 ulp (the golden vectors come from torch's CPU sqrt, which is not
correctly rounded -- see test_oracle_golden.ulps32).
"""
import functools
import operator

import numpy as np
import pandas as pd
import pytest
import torch

from oracle import prograph_oracle as O
from conftest import assert_csr_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pgb():
    import prograph_b200
    return prograph_b200


@pytest.fixture(scope="module")
def eng(pgb):
    from prograph_b200.engine import get_engine
    return get_engine()


def np_(t):
    return t.cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def ulps32(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


def mutational_library(rng, n, L, max_mut=8, alphabet=20, dup_every=97):
    wt = rng.integers(1, alphabet + 1, size=L)
    X = np.tile(wt, (n, 1))
    for i in range(1, n):
        m = rng.integers(1, max_mut + 1)
        pos = rng.choice(L, size=min(m, L), replace=False)
        X[i, pos] = (X[i, pos] - 1 + rng.integers(1, alphabet, size=len(pos))) % alphabet + 1
    for i in range(dup_every, n, dup_every):
        X[i] = X[i - 1]
    return X.astype(np.int64)


# ------------------------------------------------------------------ distance functions
def test_reference_unit_vectors(pgb, g_distance):
    g = g_distance
    X, Y = torch.from_numpy(g["t_X"]), torch.from_numpy(g["t_Y"])
    for got, key in ((pgb.hamming(X, Y), "t_ham_2d2d"), (pgb.hamming(X, Y[0]), "t_ham_2d1d"),
                     (pgb.hamming(X[1], Y[0]), "t_ham_1d1d")):
        assert got.dtype == torch.int64 and not got.is_cuda
        np.testing.assert_array_equal(np_(got), g[key])
    for got, key in ((pgb.minkowski(X, Y), "t_min_2d2d"), (pgb.minkowski(X, Y[0]), "t_min_2d1d"),
                     (pgb.minkowski(X[1], Y[0]), "t_min_1d1d")):
        assert got.dtype == torch.float32
        assert ulps32(np_(got), g[key]).max() <= 1
    # tests/tests.py:195: exact float32 values of the correctly rounded root
    np.testing.assert_array_equal(np_(pgb.minkowski(X, Y)), O.minkowski(g["t_X"], g["t_Y"]))
    np.testing.assert_array_equal(np_(pgb.minkowski(X, Y, p=1)), g["t_min_p1"])
    with pytest.raises(ValueError):
        pgb.hamming(torch.Tensor([4, 5, 6]), torch.Tensor())
    with pytest.raises(ValueError):
        pgb.minkowski(torch.Tensor([4, 5, 6]), torch.Tensor())


def test_hamming_tokens_golden(pgb, g_distance):
    g = g_distance
    X, Y = g["i_X"], g["i_Y"]
    np.testing.assert_array_equal(np_(pgb.hamming(X, Y)), g["i_ham"])
    sim = pgb.hamming(X, Y, similarity=True)
    assert sim.dtype == torch.float32
    np.testing.assert_array_equal(np_(sim), g["i_ham_sim"])
    Xh, Yh = torch.from_numpy(X).half(), torch.from_numpy(Y).half()
    np.testing.assert_array_equal(np_(pgb.hamming(Xh, Yh)), g["h_ham"])
    np.testing.assert_array_equal(np_(pgb.hamming(Xh, Yh, similarity=True)), g["h_ham_sim"])
    np.testing.assert_array_equal(np_(pgb.hamming(g["r_X"], g["r_Y"])), g["r_ham_xy"])
    np.testing.assert_array_equal(np_(pgb.hamming(g["r_Y"], g["r_X"])), g["r_ham_yx"])
    # real values: element-wise != kernel
    np.testing.assert_array_equal(np_(pgb.hamming(g["f_X"], g["f_Y"])), g["f_ham"])
    # results stay on the device of the inputs
    out = pgb.hamming(torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda())
    assert out.is_cuda
    np.testing.assert_array_equal(np_(out), g["i_ham"])


@pytest.mark.parametrize("L", [1, 20, 32, 33, 56, 64, 100, 128, 200, 256, 300, 512, 513, 700, 1500, 2000])
@pytest.mark.parametrize("alphabet", [20, 200])
def test_hamming_matrix_vs_oracle(pgb, L, alphabet):
    rng = np.random.default_rng(L * 1000 + alphabet)
    X = rng.integers(0, alphabet + 1, size=(700, L)).astype(np.int64)
    Y = rng.integers(0, alphabet + 1, size=(37, L)).astype(np.int64)
    Y[3] = X[5]
    Y[4, : L // 2] = X[6, : L // 2]
    np.testing.assert_array_equal(np_(pgb.hamming(X, Y)), O.hamming(X, Y))
    np.testing.assert_array_equal(np_(pgb.hamming(X, Y, similarity=True)), O.hamming(X, Y, similarity=True))


def test_hamming_matrix_many_queries(pgb):
    rng = np.random.default_rng(5)
    X = rng.integers(1, 21, size=(1300, 56)).astype(np.uint8)
    Y = rng.integers(1, 21, size=(1100, 56)).astype(np.uint8)
    np.testing.assert_array_equal(np_(pgb.hamming(X, Y)), O.hamming(X, Y))


def test_minkowski_golden(pgb, g_distance):
    g = g_distance
    X, Y = g["i_X"], g["i_Y"]
    for p in (1, 2, 3):
        got = np_(pgb.minkowski(X, Y, p=p))
        ref = g[f"i_min_p{p}"]
        assert got.dtype == ref.dtype
        np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
        ok = ~np.isnan(ref)
        assert ulps32(got[ok], ref[ok]).max() <= (1 if p < 3 else 4)
        np.testing.assert_array_equal(got, O.minkowski(X, Y, p=p)) if p < 3 else None
    Xh, Yh = torch.from_numpy(X).half(), torch.from_numpy(Y).half()
    for p in (1, 2):
        got = pgb.minkowski(Xh, Yh, p=p)
        assert got.dtype == torch.float16
        np.testing.assert_array_equal(np_(got), g[f"h_min_p{p}"])
        np.testing.assert_array_equal(np_(pgb.minkowski(Xh, Yh, p=p, similarity=True)), g[f"h_min_p{p}_sim"])
    got, ref = np_(pgb.minkowski(Xh, Yh, p=3)), g["h_min_p3"]
    np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.abs(got[ok].view(np.int16).astype(int) - ref[ok].view(np.int16).astype(int)).max() <= 1
    assert ulps32(np_(pgb.minkowski(g["r_X"], g["r_Y"])), g["r_min_xy"]).max() <= 1
    # real-valued float32: the fp32 sum order differs from torch's, allow 2 ulp
    assert ulps32(np_(pgb.minkowski(g["f_X"], g["f_Y"])), g["f_min_p2"]).max() <= 2
    # two-component fp16 embeddings: order independent -> bit exact
    E = torch.from_numpy(g["e_X"]).half()
    np.testing.assert_array_equal(np_(pgb.minkowski(E, E[:16])), g["e_min_p2_h"])


# ------------------------------------------------------------------ fused kNN / eps (engine level)
@pytest.mark.parametrize("L,k", [(3, 3), (20, 16), (56, 16), (100, 5), (256, 16), (256, 40), (400, 16), (512, 7), (700, 16), (1000, 5), (1792, 3)])
def test_fused_knn_uniform_ties(eng, L, k):
    """iid-uniform tokens: the k-th place of almost every row is a tie."""
    rng = np.random.default_rng(L + k)
    n = 2100
    X = rng.integers(1, 5 if L < 10 else 21, size=(n, L)).astype(np.int64)
    tab = eng.pack(X)
    idx, w = eng.hamming_knn(tab, 0, n, tab, k, drop=1)
    D = O.hamming(X, X)
    ri, rw = O.knn_from_distances(D, k)
    np.testing.assert_array_equal(np_(idx), ri)
    np.testing.assert_array_equal(np_(w), rw)
    idx, w = eng.hamming_knn(tab, 0, n, tab, k, drop=1, similarity=True)
    ri, rw = O.knn_from_distances(O.hamming(X, X, similarity=True), k, descending=True)
    np.testing.assert_array_equal(np_(idx), ri)
    np.testing.assert_array_equal(np_(w), rw)


def test_fused_knn_mutational_duplicates_and_row_ranges(eng):
    rng = np.random.default_rng(11)
    X = mutational_library(rng, 3000, 256)
    tab = eng.pack(X.astype(np.uint8))
    D = O.hamming(X, X)
    ri, rw = O.knn_from_distances(D, 16)
    idx, w = eng.hamming_knn(tab, 0, 3000, tab, 16, drop=1)
    np.testing.assert_array_equal(np_(idx), ri)
    np.testing.assert_array_equal(np_(w), rw)
    # a row shard (what one rank of a multi-GPU build computes)
    idx, w = eng.hamming_knn(tab, 1234, 777, tab, 16, drop=1)
    np.testing.assert_array_equal(np_(idx), ri[1234:2011])
    np.testing.assert_array_equal(np_(w), rw[1234:2011])


def test_fused_knn_adversarial_descending(eng):
    """Distances to row 0 strictly improve with the column index: every column inserts."""
    L, n = 256, 600
    base = np.ones(L, dtype=np.int64)
    X = np.tile(base, (n, 1))
    for i in range(1, n):
        X[i, : max(0, L - (i * L) // n)] = 2      # later rows are closer to all-ones
    X[0] = 1
    tab = eng.pack(X)
    idx, w = eng.hamming_knn(tab, 0, n, tab, 16, drop=1)
    ri, rw = O.knn_from_distances(O.hamming(X, X), 16)
    np.testing.assert_array_equal(np_(idx), ri)
    np.testing.assert_array_equal(np_(w), rw)


def test_fused_knn_queries_vs_dataset(eng):
    rng = np.random.default_rng(3)
    X = mutational_library(rng, 5000, 100)
    Q = mutational_library(np.random.default_rng(4), 300, 100)
    tx = eng.pack(X)
    tq = eng.pack(Q, planes=tx.planes, words=tx.words)
    idx, d = eng.hamming_knn(tq, 0, 300, tx, 1, drop=0)
    D = O.hamming(X, Q)
    np.testing.assert_array_equal(np_(idx)[:, 0], np.argmin(D, axis=1))
    np.testing.assert_array_equal(np_(d)[:, 0], D.min(axis=1))


@pytest.mark.parametrize("L", [3, 56, 256, 300, 900])
def test_fused_eps_vs_oracle(eng, L):
    from prograph_b200.graph import distance_lut
    rng = np.random.default_rng(L)
    X = mutational_library(rng, 2500, L, max_mut=min(4, L))
    tab = eng.pack(X)
    D = O.hamming(X, X)
    for comp, eps, sim in ((operator.le, 2, False), (operator.lt, 3, False), (operator.eq, 2, False),
                           (operator.ne, 2, False), (operator.ge, 3, False), (operator.le, 1.5, False),
                           (operator.le, 2, True)):
        e = 1 / (1 + eps) if sim else eps
        lut = distance_lut(tab.words * 32, comp, e, sim)
        indptr, idx, w = eng.hamming_eps(tab, 0, 2500, tab, lut, similarity=sim)
        if sim:
            S = O.hamming(X, X, similarity=True)
            keep = comp(np.float32(e), S) & (S < 1)
            W = S
        else:
            keep = comp(D, eps) & (D > 0)
            W = D
        rows, cols = np.nonzero(keep)
        np.testing.assert_array_equal(np_(indptr), np.concatenate([[0], np.cumsum(keep.sum(1))]))
        np.testing.assert_array_equal(np_(idx), cols)
        np.testing.assert_array_equal(np_(w), W[rows, cols])


# ------------------------------------------------------------------ tile consumers
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.int64, torch.int32, torch.float64])
@pytest.mark.parametrize("descending", [False, True])
def test_tile_topk_stable(eng, dtype, descending):
    rng = np.random.default_rng(17)
    a = rng.integers(-6, 7, size=(9, 5000)).astype(np.float64) / 2.0     # heavy ties, signed, +-0
    if dtype.is_floating_point:
        a[2, 100] = np.nan
        a[3, :40] = np.nan
        a[4, 7] = -0.0
    else:
        a = np.round(a)
    t = torch.from_numpy(a).to(dtype).cuda()
    for k, drop in ((1, 0), (16, 1), (300, 1)):
        idx, val = eng.tile_topk(t, k, drop=drop, descending=descending)
        ref = torch.sort(t.cpu().double() if dtype != torch.float16 else t.cpu().float(), dim=1,
                         descending=descending, stable=True)
        np.testing.assert_array_equal(np_(idx), np_(ref.indices[:, drop:drop + k]))
        got, want = np_(val).astype(np.float64), np_(ref.values[:, drop:drop + k]).astype(np.float64)
        np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
        np.testing.assert_array_equal(got[~np.isnan(got)], want[~np.isnan(want)])


def test_tile_threshold(eng):
    rng = np.random.default_rng(23)
    a = (rng.integers(0, 9, size=(13, 3001)) / 4.0).astype(np.float16)
    t = torch.from_numpy(a).cuda()
    for code, op in ((0, operator.lt), (1, operator.le), (2, operator.eq), (3, operator.ne), (4, operator.ge),
                     (5, operator.gt)):
        for eps, swap, guard in ((1.0, False, 1), (0.75, True, 2), (0.1, False, 0)):
            indptr, idx, val = eng.tile_threshold(t, code, eps, swap=swap, guard=guard)
            e = np.float16(eps)
            keep = op(e, a) if swap else op(a, e)
            if guard == 1:
                keep &= a > 0
            elif guard == 2:
                keep &= a < 1
            rows, cols = np.nonzero(keep)
            np.testing.assert_array_equal(np_(indptr), np.concatenate([[0], np.cumsum(keep.sum(1))]))
            np.testing.assert_array_equal(np_(idx), cols)
            np.testing.assert_array_equal(np_(val), a[rows, cols])


# ------------------------------------------------------------------ build_graph on golden fixtures
def make_csv(tmp_path, g, name):
    path = tmp_path / f"{name}.csv"
    pd.DataFrame({"Sequence": list(g["sequences"]), "Fitness": g["fitness"]}).to_csv(path)
    return str(path)


@pytest.fixture(scope="module")
def pg_synth(pgb, g_synthetic, tmp_path_factory):
    return pgb.Prograph(file=make_csv(tmp_path_factory.mktemp("synth"), g_synthetic, "synthetic_data"))


@pytest.fixture(scope="module")
def pg_lib(pgb, g_library, tmp_path_factory):
    return pgb.Prograph(file=make_csv(tmp_path_factory.mktemp("lib"), g_library, "library"))


@pytest.fixture(scope="module")
def pg_knn(pgb, g_knntest, tmp_path_factory):
    g = g_knntest
    pg = pgb.Prograph(file=make_csv(tmp_path_factory.mktemp("knn"), g, "knntest"))
    pg.graph["Embedded"] = [row for row in g["embedded"]]
    return pg


def test_synthetic_build(pgb, pg_synth, g_synthetic):
    g, pg = g_synthetic, pg_synth
    np.testing.assert_array_equal(pg.tokenized, g["tokenized"])
    np.testing.assert_array_equal(pg.mutated_positions, g["mutated_positions"])
    np.testing.assert_array_equal(pg.sequence_mutation_locations, g["mutant_array_seed"])
    assert_csr_equal(list(pg("Neighbours")), g, "nb_eps1")
    assert np.all(pg.degree() == 27)                                              # tests.py:158
    assert_csr_equal(pg.build_graph(eps=2), g, "nb_eps2")
    assert_csr_equal(pg.build_graph(eps=2, comp=operator.lt), g, "nb_eps2_lt")
    assert_csr_equal(pg.build_graph(eps=2, comp=operator.eq), g, "nb_eps2_eq")
    assert_csr_equal(pg.build_graph(eps=1.5), g, "nb_eps1p5")
    assert_csr_equal(pg.build_graph(eps=2, similarity=True), g, "nb_eps2_sim")
    assert_csr_equal(pg.build_graph(eps=1, batch_size=7), g, "nb_eps1_b7")
    sub = g["sub_idxs"]
    assert_csr_equal(pg.build_graph(eps=1, idxs=sub), g, "nb_eps1_sub")
    assert_csr_equal(pg.build_graph(eps=3, idxs=sub, comp=operator.ge), g, "nb_eps3_ge_sub")
    assert_csr_equal(pg.build_graph(eps=1, idxs=sub, comp=operator.gt), g, "nb_eps1_gt_sub")
    assert_csr_equal(pg.build_graph(eps=2, idxs=sub, comp=operator.ne), g, "nb_eps2_ne_sub")
    assert_csr_equal(pg.build_graph(eps=2, distance=pgb.minkowski), g, "nb_min_eps2")
    for k in (1, 3, 16):
        assert_csr_equal(pg.build_graph(k=k), g, f"knn{k}")
        assert_csr_equal(pg.build_graph(k=k, similarity=True), g, f"knn{k}_sim")
        assert_csr_equal(pg.build_graph(k=k), g, f"knn{k}_unstable", check_idx=False)
    assert_csr_equal(pg.build_graph(k=4, idxs=sub), g, "knn4_sub")
    assert_csr_equal(pg.build_graph(k=200, idxs=sub), g, "knn200_sub")
    assert_csr_equal(pg.build_graph(k=3, distance=pgb.minkowski), g, "knn3_min")
    with pytest.raises(ValueError):
        pg.build_graph(k=0)
    with pytest.raises(TypeError):
        pg.build_graph(k=0.5)
    with pytest.raises(ValueError):
        pg.build_graph(k=2, eps=1)
    with pytest.raises(ValueError):
        pg.build_graph()
    # a user-supplied metric honouring the plug-in protocol (README.md:48)
    def my_metric(X, Y, similarity=False):
        d = torch.sum(X != Y[:, None, :], axis=2)
        return 1 / (1 + d) if similarity else d
    assert_csr_equal(pg.build_graph(eps=2, distance=my_metric), g, "nb_eps2")
    assert_csr_equal(pg.build_graph(k=3, distance=my_metric), g, "knn3")
    assert_csr_equal(pg.build_graph(eps=2, distance=my_metric, comp=lambda a, b: a <= b), g, "nb_eps2")


def test_synthetic_indexing_and_queries(pg_synth, g_synthetic):
    g, pg = g_synthetic, pg_synth
    same = np.testing.assert_array_equal
    same(pg.boolean_mutant_array("LDC"), g["mutant_array_LDC"])
    same(pg.indexing(positions=[1, 2]), g["ix_pos12"])
    same(pg.indexing(positions=[1, 2], Bool="and"), g["ix_pos12_and"])
    same(pg.indexing(positions=[0]), g["ix_pos0"])
    same(pg.indexing(distances=3), g["ix_d3"])
    same(pg.indexing(distances=2), g["ix_d2"])
    same(pg.indexing(distances=[1, 3]), g["ix_d13"])
    same(pg.indexing(positions=[1, 2], distances=2), g["ix_pos12_d2"])
    a, b = pg.indexing(positions=[1, 2], distances=2, complement=True)
    same(a, g["ix_pos12_d2_c0"])
    same(b, g["ix_pos12_d2_c1"])
    same(pg.indexing(reference_seq="LDC", positions=[1]), g["ix_ref_LDC_pos1"])
    same(pg.indexing(reference_seq="LDC", distances=1), g["ix_ref_LDC_d1"])
    same(pg.indexing(reference_seq=500, distances=2, positions=[0, 1]), g["ix_ref_500_d2_pos01"])
    assert len(pg.indexing(percentage=0.7)) == 700                                  # tests.py:49
    assert len(pg.indexing(positions=[1, 2], distances=2, percentage=0.3)) == 24    # tests.py:51
    with pytest.raises(AssertionError):
        pg.indexing(distances=[1, 2, 4])                                            # tests.py:94
    assert pg[pg.indexing(reference_seq="LDC", positions=[1])]["Sequence"][901] == "LAC"   # tests.py:98
    same(pg.get_mutated_positions(np.array([0])), g["gmp_0"])
    same(pg.get_mutated_positions(np.array([1, 2])), g["gmp_12"])
    same(pg.calc_neighbours(seq="ACL"), g["cn_ACL"])
    same(pg.calc_neighbours(seq="ACL"), pg["ACL"]["Neighbours"][0])                # tests.py:64
    same(pg.calc_neighbours(seq="ACL", eps=2), g["cn_ACL_eps2"])
    same(pg.calc_neighbours(seq="ACL", eps=2, comp=operator.le), g["cn_ACL_le2"])
    same(pg.calc_neighbours(seq=77, eps=3, comp=operator.ge), g["cn_77_ge3"])
    from prograph_b200 import minkowski
    same(pg.calc_neighbours(seq=77, eps=2, distance=minkowski, comp=operator.le), g["cn_77_min_le2"])
    same(pg.neighbourhood("ACL", 1).index.to_numpy(), g["nh_ACL_1"])
    same(pg.neighbourhood("ACL", 2).index.to_numpy(), g["nh_ACL_2"])
    text = str(pg)
    assert f"Max Distance        : {int(g['str_max'])}" in text
    assert f"Number of Distances : {int(g['str_nuniq'])}" in text
    A = pg.adjacency()
    assert tuple(A.shape) == tuple(g["adj_shape"])
    same(A.row, g["adj_row"]); same(A.col, g["adj_col"]); same(A.data, g["adj_data"])
    assert A.data.dtype == g["adj_data"].dtype
    same(pg.degree(), g["degree"])
    same(pg.degree(boolean_weights=True), g["degree_bool"])
    assert np.array_equal(A.todense()[:3, :3], [[0, 1, 1], [1, 0, 1], [1, 1, 0]])   # tests.py:137
    # query / access surface
    assert pg["AAC"]["Sequence"] == "AAC" and pg("Sequence")[26] == "ADH"
    assert pg[(1, 2, 2)]["Sequence"] == "ACC" and len(pg) == 1000
    near, dmin = pg.nearest_neighbour(["ACL", "AAA"])
    assert list(near["Sequence"]) == ["ACL", "AAA"] and dmin == 0


def test_knntest_fixture(pgb, pg_knn, g_knntest):
    g, pg = g_knntest, pg_knn
    assert_csr_equal(list(pg("Neighbours")), g, "csv_nb")
    mk = pgb.minkowski
    for k in (1, 2, 3, 5, 9):
        assert_csr_equal(pg.build_graph(representation="Embedded", k=k, distance=mk), g, f"knn{k}")
        assert_csr_equal(pg.build_graph(representation="Embedded", k=k, distance=mk, similarity=True), g, f"knn{k}_sim")
    assert_csr_equal(pg.build_graph(representation="Embedded", eps=2, distance=mk), g, "eps2")
    assert_csr_equal(pg.build_graph(representation="Embedded", eps=2, distance=mk, similarity=True), g, "eps2_sim")
    assert_csr_equal(pg.build_graph(representation="Embedded", eps=1.25, distance=mk, comp=operator.lt), g, "eps1p25_lt")
    assert_csr_equal(pg.build_graph(representation="Embedded", eps=0.1, distance=mk), g, "eps0p1")
    # tests/tests.py:141-154, 159-167
    L1 = [x[0] for x in pg.build_graph(representation="Embedded", k=1, distance=mk)]
    assert np.all(np.array(L1).reshape(-1,) == np.array([1, 0, 3, 2, 5, 4]))
    L2 = [x[0] for x in pg.build_graph(representation="Embedded", k=2, distance=mk)]
    assert np.all(L2 == np.array([[1, 3], [0, 3], [3, 4], [2, 4], [5, 2], [4, 2]]))
    with pytest.raises(ValueError):
        pg.build_graph(representation="Embedded", k=0, distance=mk)
    with pytest.raises(TypeError):
        pg.build_graph(representation="Embedded", k=0.5, distance=mk)
    pg.graph["Weighted"] = pg.build_graph(eps=2, representation="Embedded", distance=mk)
    np.testing.assert_almost_equal(pg.degree(graph="Weighted", boolean_weights=True), g["deg_eps2_bool"])
    pg.graph["Weighted"] = pg.build_graph(k=1, representation="Embedded", distance=mk)
    np.testing.assert_array_equal(pg.degree(graph="Weighted"), g["deg_k1"])
    np.testing.assert_almost_equal(pg.degree(graph="Weighted"), np.array([0.5, 0.5, 1., 1., 0.79052734, 0.79052734]))
    # minkowski with another exponent through functools.partial keeps the fused dispatch
    p1 = pg.build_graph(representation="Embedded", k=2, distance=functools.partial(mk, p=1))
    ref = O.build_graph(g["embedded"], k=2, distance=functools.partial(O.minkowski, p=1))
    for (a, b), (c, d) in zip(p1, ref):
        np.testing.assert_array_equal(a, c)
        np.testing.assert_array_equal(b, d)


def test_library_fixture(pgb, pg_lib, g_library):
    g, pg = g_library, pg_lib
    np.testing.assert_array_equal(pg.tokenized, g["tokenized"])
    np.testing.assert_array_equal(pg.mutated_positions, g["mutated_positions"])
    np.testing.assert_array_equal(pg.sequence_mutation_locations, g["mutant_array_seed"])
    np.testing.assert_array_equal(np_(pgb.hamming(pg.tokenized, pg.tokenized)), g["dmat"])
    assert_csr_equal(list(pg("Neighbours")), g, "nb_eps1")
    assert_csr_equal(pg.build_graph(eps=3), g, "nb_eps3")
    assert_csr_equal(pg.build_graph(eps=3, similarity=True), g, "nb_eps3_sim")
    assert_csr_equal(pg.build_graph(eps=4, comp=operator.eq), g, "nb_eps4_eq")
    assert_csr_equal(pg.build_graph(eps=3.0, distance=pgb.minkowski), g, "nb_min_eps3")
    for k in (1, 16, 40):
        assert_csr_equal(pg.build_graph(k=k), g, f"knn{k}")
        assert_csr_equal(pg.build_graph(k=k, similarity=True), g, f"knn{k}_sim")
        assert_csr_equal(pg.build_graph(k=k), g, f"knn{k}_unstable", check_idx=False)
    assert_csr_equal(pg.build_graph(k=4, distance=pgb.minkowski), g, "knn4_min")
    assert_csr_equal(pg.build_graph(k=4, distance=pgb.minkowski, similarity=True), g, "knn4_min_sim")
    same = np.testing.assert_array_equal
    same(pg.indexing(distances=2), g["ix_d2"])
    same(pg.indexing(distances=[1, 2]), g["ix_d12"])
    same(pg.indexing(positions=[int(x) for x in g["ix_pos_list"]]), g["ix_pos"])
    same(pg.indexing(reference_seq=5, positions=[int(x) for x in g["ix_pos_ref5_list"]]), g["ix_pos_ref5"])
    same(pg.get_mutated_positions(g["ix_pos_list"]), g["gmp"])
    same(pg.calc_neighbours(seq=9, eps=2), g["cn_9_eq2"])
    same(pg.calc_neighbours(seq=5, eps=0), g["cn_5_eq0"])
    same(pg.neighbourhood(9, 2).index.to_numpy(), g["nh_9_2"])
    clusters = pg.neighbourhood_clustering(2)
    covered = np.zeros(len(pg), dtype=bool)
    for seed, members in clusters.items():
        assert not covered[seed]
        same(np.asarray(members.index), np.where(O.neighbourhood_mask(g["tokenized"], seed, 2))[0])
        covered[np.asarray(members.index)] = True
    assert covered.all()


# ------------------------------------------------------------------ size-independent properties at scale
def test_gb1_style_library_known_answers(eng):
    """C3: all 20^4 variants at 4 sites of a 56-mer: every epsilon=1 degree is 4*19 = 76 and
    epsilon=2 gives 76 + 6*19^2 = 2242."""
    from prograph_b200.graph import distance_lut
    import itertools
    wt = np.frombuffer(b"MTYKLILNGKTLKGETTTEAVDAATAEKVFKQYANDNGVDGEWTYDDATKTFTVTE", dtype=np.uint8)
    aa = "ACDEFGHIKLMNPQRSTVWY"
    lut256 = np.zeros(256, dtype=np.uint8)
    for i, ch in enumerate(aa):
        lut256[ord(ch)] = i + 1
    base = lut256[wt]
    combos = np.array(list(itertools.product(range(1, 21), repeat=4)), dtype=np.uint8)
    X = np.tile(base, (len(combos), 1))
    X[:, [38, 39, 40, 53]] = combos
    tab = eng.pack(X)
    n = len(X)
    indptr, idx, w = eng.hamming_eps(tab, 0, n, tab, distance_lut(tab.words * 32, operator.le, 1, False))
    deg = np.diff(np_(indptr))
    assert n == 160000 and np.all(deg == 76) and int(np_(indptr)[-1]) == 12160000
    assert np.all(np_(w) == 1)
    ii = np_(idx).reshape(n, 76)
    assert np.all(np.diff(ii, axis=1) > 0)                       # ascending neighbour order
    rows = np.repeat(np.arange(n), 76)
    assert np.all((X[rows[::997]] != X[np_(idx)[::997]]).sum(1) == 1)
    indptr2, _, w2 = eng.hamming_eps(tab, 0, 4096, tab, distance_lut(tab.words * 32, operator.le, 2, False))
    assert np.all(np.diff(np_(indptr2)) == 2242)
    assert int((np_(w2) == 1).sum()) == 4096 * 76


def test_large_knn_sampled_rows(eng):
    """C4-shaped (scaled to 200k x 256): sampled rows against the oracle, and the symmetric
    checksum sum_i d(i, nn_1(i)) computed two ways."""
    rng = np.random.default_rng(0)
    n, L = 200_000, 256
    X = mutational_library(rng, n, L, dup_every=1009).astype(np.uint8)
    tab = eng.pack(X)
    idx, w = eng.hamming_knn(tab, 0, n, tab, 16, drop=1)
    idx, w = np_(idx), np_(w)
    sample = rng.choice(n, size=64, replace=False)
    D = O.hamming(X, X[sample])
    ri, rw = O.knn_from_distances(D, 16)
    np.testing.assert_array_equal(idx[sample], ri)
    np.testing.assert_array_equal(w[sample], rw)
    assert np.all(np.diff(w, axis=1) >= 0)                       # weights ascending per row
    pick = rng.choice(n, size=5000, replace=False)
    assert np.all((X[pick] != X[idx[pick, 0]]).sum(1) == w[pick, 0])


# ------------------------------------------------------------------ next rows: fused tokeniser, sidecar
def test_pack_chars_matches_tokenise_then_pack(eng, pg_lib, g_library):
    seqs = list(g_library["sequences"])                       # ragged lengths: zero padded
    codes = pg_lib._letter_codes(seqs)
    a = eng.pack_chars(codes, pg_lib._letter_table(), planes=5)
    b = eng.pack(g_library["tokenized"], planes=5)
    assert (a.rows, a.L, a.words) == (b.rows, b.L, b.words)
    assert torch.equal(a.data, b.data)
    # letters outside the alphabet and pad bytes are token 0, like the reference tokeniser
    odd = pg_lib._letter_codes(["AXZ", "C"])
    t = eng.pack_chars(odd, pg_lib._letter_table(), planes=5)
    ref = eng.pack(np.array([[1, 0, 0], [2, 0, 0]]), planes=5)
    assert torch.equal(t.data, ref.data)


def test_save_reload_with_csr_sidecar(pgb, pg_synth, tmp_path):
    assert pgb.save(pg_synth, name="synth", directory=str(tmp_path) + "/")
    assert (tmp_path / "synth.pkl").exists() and (tmp_path / "synth.graph.npz").exists()
    again = pgb.Prograph(file=str(tmp_path / "synth.pkl"))
    assert again[0]["Sequence"] == "AAA"                                          # tests.py:121
    t = again._table_for("Neighbours")
    z = np.load(tmp_path / "synth.graph.npz")
    np.testing.assert_array_equal(t.idx, z["idx"])
    np.testing.assert_array_equal(again.degree(), pg_synth.degree())
    A, B = again.adjacency(), pg_synth.adjacency()
    np.testing.assert_array_equal(A.row, B.row)
    np.testing.assert_array_equal(A.col, B.col)
    np.testing.assert_array_equal(A.data, B.data)
    # a pickle without a sidecar still loads (lists are flattened on demand)
    (tmp_path / "synth.graph.npz").unlink()
    plain = pgb.Prograph(file=str(tmp_path / "synth.pkl"))
    np.testing.assert_array_equal(plain.degree(), pg_synth.degree())


def test_headline_config_one_million_uniform(eng):
    """The bench workload itself (C4-U: 1 M x 256 iid-uniform tokens, k=16, all 10^12 ordered pairs
    through the symmetric build): sampled rows bit for bit against the oracle, plus
    size-independent properties."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import make_tokens
    n, L, k = 1_000_000, 256, 16
    X = make_tokens(n, L, "uniform")
    tab = eng.pack(X)
    from prograph_b200 import graph
    assert n >= graph.SYM_MIN_ROWS                                # i.e. the symmetric build, as in bench.py
    idx, w, row0 = graph.hamming_knn_graph(eng, tab, k, False, 0, 1, None)
    assert row0 == 0
    # every one of the 16 M entries against the one-sided sweep (10^12 pair evaluations, ~3 s): the lock
    # regime of the full-size build (tens of millions of row-lock acquisitions) never occurs in the
    # small-table tests
    oi, ow = eng.hamming_knn(tab, 0, n, tab, k, drop=1)
    assert torch.equal(idx, oi) and torch.equal(w, ow)
    del oi, ow
    eng.sym_check()
    idx, w = np_(idx), np_(w)
    rng = np.random.default_rng(123)
    # rows of the bootstrap block, of the first / last row blocks and random ones
    sample = np.concatenate([[0, 5000, 8191, 8192, n - 1], rng.choice(n, size=19, replace=False)])
    D = O.hamming(X, X[sample], chunk=8)
    ri, rw = O.knn_from_distances(D, k)
    np.testing.assert_array_equal(idx[sample], ri)
    np.testing.assert_array_equal(w[sample], rw)
    assert idx.min() >= 0 and idx.max() < n
    assert np.all(np.diff(w, axis=1) >= 0)
    ties = np.diff(w, axis=1) == 0
    assert np.all(np.diff(idx, axis=1)[ties] > 0)                 # equal distances: ascending index
    pick = rng.choice(n, size=4000, replace=False)
    for j in (0, k - 1):
        assert np.all((X[pick] != X[idx[pick, j]]).sum(1) == w[pick, j])
    # mutuality (what the symmetric build must get right on both sides of every pair): if j is
    # i's nearest neighbour and strictly closer than j's own k-th neighbour, j lists i
    pick = rng.choice(n, size=200_000, replace=False)
    j, d = idx[pick, 0], w[pick, 0]
    must = d < w[j, k - 1]
    assert must.sum() > 1000
    assert np.all((idx[j[must]] == pick[must][:, None]).any(axis=1))
