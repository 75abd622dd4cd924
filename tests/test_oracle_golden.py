"""Pin the CPU oracle against outputs of the unmodified reference (tests/golden)."""
import operator

import numpy as np
import pytest

from oracle import prograph_oracle as O
from conftest import assert_csr_equal


def ulps32(a, b):
    """Distance in float32 ulps.  torch's *CPU* sqrt (MKL VML) is not correctly rounded for
    ~0.6 % of inputs (measured in tests/golden/make_golden.py's container); the oracle and
    the CUDA kernels use the IEEE-rounded sqrt that torch's CUDA kernel also uses, so float32
    roots are compared with the CPU-generated golden vectors to within 1 ulp."""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


def close32(a, b, ulps=1):
    a, b = np.asarray(a), np.asarray(b)
    assert a.dtype == b.dtype == np.float32, (a.dtype, b.dtype)
    assert a.shape == b.shape
    np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(a)
    assert ulps32(a[ok], b[ok]).max(initial=0) <= ulps


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.dtype == b.dtype, (a.dtype, b.dtype)
    np.testing.assert_array_equal(a, b)


# ---- distance functions (tests/tests.py:175-208 and random cases) -------------
def test_reference_unit_vectors(g_distance):
    g = g_distance
    X, Y = g["t_X"], g["t_Y"]
    same(O.hamming(X, Y), g["t_ham_2d2d"])
    same(O.hamming(X, Y[0]), g["t_ham_2d1d"])
    same(O.hamming(X[1], Y[0]), g["t_ham_1d1d"])
    close32(O.minkowski(X, Y), g["t_min_2d2d"])
    close32(O.minkowski(X, Y[0]), g["t_min_2d1d"])
    close32(O.minkowski(X[1], Y[0]), g["t_min_1d1d"])
    same(O.minkowski(X, Y, p=1), g["t_min_p1"])
    assert np.array_equal(O.hamming(X, Y), [[0, 3], [3, 3]])
    assert O.minkowski(X, Y)[1, 0] == np.float32(10.3923048454)


def test_empty_operand_raises():
    with pytest.raises(ValueError):
        O.hamming(np.array([4, 5, 6.0]), np.array([]))
    with pytest.raises(ValueError):
        O.minkowski(np.array([4, 5, 6.0]), np.array([]))


def test_int64_tokens(g_distance):
    g = g_distance
    X, Y = g["i_X"], g["i_Y"]
    same(O.hamming(X, Y), g["i_ham"])
    same(O.hamming(X, Y, similarity=True), g["i_ham_sim"])
    same(O.minkowski(X, Y, p=1), g["i_min_p1"])
    same(O.minkowski(X, Y, p=1, similarity=True), g["i_min_p1_sim"])
    close32(O.minkowski(X, Y, p=2), g["i_min_p2"])
    close32(O.minkowski(X, Y, p=2, similarity=True), g["i_min_p2_sim"], ulps=2)
    # the oracle's root is the IEEE-rounded one
    S = ((X[None].astype(np.int64) - Y[:, None]) ** 2).sum(2)
    same(O.minkowski(X, Y, p=2), np.sqrt(S.astype(np.float64)).astype(np.float32))
    # p=3: cube root of negative sums is NaN in the reference (no abs, minkowski.py:36)
    got, ref = O.minkowski(X, Y, p=3), g["i_min_p3"]
    assert got.dtype == ref.dtype
    np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    np.testing.assert_allclose(got[ok], ref[ok], rtol=3e-7)   # powf(x, 1/3): <= 2 ulp fp32


def test_fp16_staged_tokens(g_distance):
    g = g_distance
    X, Y = g["i_X"].astype(np.float16), g["i_Y"].astype(np.float16)
    same(O.hamming(X, Y), g["h_ham"])
    same(O.hamming(X, Y, similarity=True), g["h_ham_sim"])
    for p in (1, 2):
        same(O.minkowski(X, Y, p=p), g[f"h_min_p{p}"])
        same(O.minkowski(X, Y, p=p, similarity=True), g[f"h_min_p{p}_sim"])
    got, ref = O.minkowski(X, Y, p=3), g["h_min_p3"]
    np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.all(np.abs(got[ok].view(np.int16).astype(int) - ref[ok].view(np.int16).astype(int)) <= 1)  # 1 fp16 ulp


def test_ragged_padding(g_distance):
    g = g_distance
    same(O.hamming(g["r_X"], g["r_Y"]), g["r_ham_xy"])
    same(O.hamming(g["r_Y"], g["r_X"]), g["r_ham_yx"])
    close32(O.minkowski(g["r_X"], g["r_Y"]), g["r_min_xy"])


def test_real_valued(g_distance):
    g = g_distance
    X, Y = g["f_X"], g["f_Y"]
    same(O.hamming(X, Y), g["f_ham"])
    np.testing.assert_allclose(O.minkowski(X, Y), g["f_min_p2"], rtol=2.5e-7)
    np.testing.assert_allclose(O.minkowski(X, Y, similarity=True), g["f_min_p2_sim"], rtol=2.5e-7)
    got = O.minkowski(X.astype(np.float16), Y.astype(np.float16))
    ref = g["fh_min_p2"]
    assert got.dtype == ref.dtype
    assert np.all(np.abs(got.view(np.int16).astype(int) - ref.view(np.int16).astype(int)) <= 1)
    # two components: the fp32 sum is order independent, so fp16 results are bit exact
    E = g["e_X"].astype(np.float16)
    same(O.minkowski(E, E[:16]), g["e_min_p2_h"])


# ---- data/synthetic_data.csv --------------------------------------------------
def test_tokenize_and_masks(g_synthetic, g_library):
    for g in (g_synthetic, g_library):
        tok = O.tokenize(list(g["sequences"]))
        same(tok, g["tokenized"])
        same(O.calc_mutated_positions(tok, tok[0]), g["mutated_positions"])
        same(O.boolean_mutant_array(tok, 0), g["mutant_array_seed"])
    assert np.array_equal(O.tokenize("ACA"), [[1, 2, 1]])                      # tests.py:125
    assert np.array_equal(O.tokenize(["ACCCACAAA", "ACAA"])[1], [1, 2, 1, 1, 0, 0, 0, 0, 0])  # tests.py:129


def test_synthetic_eps_graphs(g_synthetic):
    g = g_synthetic
    tok = g["tokenized"]
    assert_csr_equal(O.build_graph(tok, eps=1), g, "nb_eps1")
    assert_csr_equal(O.build_graph(tok, eps=1, batch_size=7), g, "nb_eps1_b7")
    assert_csr_equal(O.build_graph(tok, eps=2), g, "nb_eps2")
    assert_csr_equal(O.build_graph(tok, eps=2, comp=operator.lt), g, "nb_eps2_lt")
    assert_csr_equal(O.build_graph(tok, eps=2, comp=operator.eq), g, "nb_eps2_eq")
    assert_csr_equal(O.build_graph(tok, eps=1.5), g, "nb_eps1p5")
    assert_csr_equal(O.build_graph(tok, eps=2, similarity=True), g, "nb_eps2_sim")
    sub = g["sub_idxs"]
    assert_csr_equal(O.build_graph(tok, eps=1, idxs=sub), g, "nb_eps1_sub")
    assert_csr_equal(O.build_graph(tok, eps=3, idxs=sub, comp=operator.ge), g, "nb_eps3_ge_sub")
    assert_csr_equal(O.build_graph(tok, eps=1, idxs=sub, comp=operator.gt), g, "nb_eps1_gt_sub")
    assert_csr_equal(O.build_graph(tok, eps=2, idxs=sub, comp=operator.ne), g, "nb_eps2_ne_sub")
    assert_csr_equal(O.build_graph(tok, eps=2, distance=O.minkowski), g, "nb_min_eps2")
    deg = np.diff(g["nb_eps1_indptr"])
    assert np.all(deg == 27)                                                   # tests.py:158


def test_synthetic_knn(g_synthetic):
    g = g_synthetic
    tok = g["tokenized"]
    for k in (1, 3, 16):
        assert_csr_equal(O.build_graph(tok, k=k), g, f"knn{k}")
        assert_csr_equal(O.build_graph(tok, k=k, similarity=True), g, f"knn{k}_sim")
        # the unmodified (unstable-sort) reference must agree on the weights bit for bit
        assert_csr_equal(O.build_graph(tok, k=k), g, f"knn{k}_unstable", check_idx=False)
    sub = g["sub_idxs"]
    assert_csr_equal(O.build_graph(tok, k=4, idxs=sub), g, "knn4_sub")
    assert_csr_equal(O.build_graph(tok, k=200, idxs=sub), g, "knn200_sub")
    assert_csr_equal(O.build_graph(tok, k=3, distance=O.minkowski), g, "knn3_min")
    with pytest.raises(ValueError):
        O.build_graph(tok, k=0)
    with pytest.raises(TypeError):
        O.build_graph(tok, k=0.5)
    with pytest.raises(ValueError):
        O.build_graph(tok, k=2, eps=1)


def test_synthetic_indexing(g_synthetic):
    g = g_synthetic
    tok = g["tokenized"]
    seqs = list(g["sequences"])
    ldc = seqs.index("LDC")
    same(O.boolean_mutant_array(tok, ldc), g["mutant_array_LDC"])
    same(O.indexing(tok, 0, positions=[1, 2]), g["ix_pos12"])
    same(O.indexing(tok, 0, positions=[1, 2], Bool="and"), g["ix_pos12_and"])
    same(O.indexing(tok, 0, positions=[0]), g["ix_pos0"])
    same(O.indexing(tok, 0, distances=3), g["ix_d3"])
    same(O.indexing(tok, 0, distances=2), g["ix_d2"])
    same(O.indexing(tok, 0, distances=[1, 3]), g["ix_d13"])
    same(O.indexing(tok, 0, positions=[1, 2], distances=2), g["ix_pos12_d2"])
    a, b = O.indexing(tok, 0, positions=[1, 2], distances=2, complement=True)
    same(a, g["ix_pos12_d2_c0"])
    same(b, g["ix_pos12_d2_c1"])
    same(O.indexing(tok, ldc, positions=[1]), g["ix_ref_LDC_pos1"])
    same(O.indexing(tok, ldc, distances=1), g["ix_ref_LDC_d1"])
    same(O.indexing(tok, 500, distances=2, positions=[0, 1]), g["ix_ref_500_d2_pos01"])
    with pytest.raises(AssertionError):
        O.indexing(tok, 0, distances=[1, 2, 4])                                # tests.py:94
    assert len(g["ix_pos12"]) == 99 and len(g["ix_d3"]) == 729 and len(g["ix_d13"]) == 756
    mut = O.boolean_mutant_array(tok, 0)
    same(O.get_mutated_positions(mut, g["mutated_positions"], np.array([0])), g["gmp_0"])
    same(O.get_mutated_positions(mut, g["mutated_positions"], np.array([1, 2])), g["gmp_12"])


def test_synthetic_queries(g_synthetic):
    g = g_synthetic
    tok = g["tokenized"]
    acl = list(g["sequences"]).index("ACL")
    same(O.calc_neighbours(tok, acl), g["cn_ACL"])
    same(O.calc_neighbours(tok, acl, eps=2), g["cn_ACL_eps2"])
    same(O.calc_neighbours(tok, acl, eps=2, comp=operator.le), g["cn_ACL_le2"])
    same(O.calc_neighbours(tok, 77, eps=3, comp=operator.ge), g["cn_77_ge3"])
    same(O.calc_neighbours(tok, 77, eps=2, distance=O.minkowski, comp=operator.le), g["cn_77_min_le2"])
    same(np.where(O.neighbourhood_mask(tok, acl, 1))[0], g["nh_ACL_1"])
    same(np.where(O.neighbourhood_mask(tok, acl, 2))[0], g["nh_ACL_2"])
    d = O.hamming(tok, tok[0].reshape(1, -1))
    assert d.max() == g["str_max"] and len(np.unique(d)) == g["str_nuniq"]
    # tests.py:64: calc_neighbours("ACL") equals the stored eps=1 neighbour list
    lo, hi = g["nb_eps1_indptr"][acl], g["nb_eps1_indptr"][acl + 1]
    same(O.calc_neighbours(tok, acl), g["nb_eps1_idx"][lo:hi])


# ---- data/knntest_pgraph.pkl (fp16 minkowski on 2-D embeddings) ----------------
def test_knntest(g_knntest):
    g = g_knntest
    E = g["embedded"]
    same(O.minkowski(E.astype(np.float16), E.astype(np.float16)), g["dmat_h"])
    tok = O.tokenize(list(g["sequences"]))
    assert_csr_equal(O.build_graph(tok, eps=1), g, "csv_nb")
    assert_csr_equal(O.build_graph(tok, eps=1), g, "pkl_nb")
    for k in (1, 2, 3, 5, 9):
        assert_csr_equal(O.build_graph(E, k=k, distance=O.minkowski), g, f"knn{k}")
        assert_csr_equal(O.build_graph(E, k=k, distance=O.minkowski, similarity=True), g, f"knn{k}_sim")
        assert_csr_equal(O.build_graph(E, k=k, distance=O.minkowski), g, f"knn{k}_unstable", check_idx=False)
    assert_csr_equal(O.build_graph(E, eps=2, distance=O.minkowski), g, "eps2")
    assert_csr_equal(O.build_graph(E, eps=2, distance=O.minkowski, similarity=True), g, "eps2_sim")
    assert_csr_equal(O.build_graph(E, eps=1.25, distance=O.minkowski, comp=operator.lt), g, "eps1p25_lt")
    assert_csr_equal(O.build_graph(E, eps=0.1, distance=O.minkowski), g, "eps0p1")
    # the two kNN answers pinned in tests/tests.py:141-148 (row 2 is a tie -> lower index)
    k1 = O.build_graph(E, k=1, distance=O.minkowski)
    assert [int(x[0][0]) for x in k1] == [1, 0, 3, 2, 5, 4]
    k2 = O.build_graph(E, k=2, distance=O.minkowski)
    assert [list(map(int, x[0])) for x in k2] == [[1, 3], [0, 3], [3, 4], [2, 4], [5, 2], [4, 2]]
    assert float(k1[4][1][0]) == 0.79052734375                                  # tests.py:167


# ---- seeded ragged mutational library --------------------------------------------
def test_library(g_library):
    g = g_library
    tok = g["tokenized"]
    same(O.hamming(tok, tok), g["dmat"])
    assert_csr_equal(O.build_graph(tok, eps=1), g, "nb_eps1")
    assert_csr_equal(O.build_graph(tok, eps=3), g, "nb_eps3")
    assert_csr_equal(O.build_graph(tok, eps=3, similarity=True), g, "nb_eps3_sim")
    assert_csr_equal(O.build_graph(tok, eps=4, comp=operator.eq), g, "nb_eps4_eq")
    assert_csr_equal(O.build_graph(tok, eps=3.0, distance=O.minkowski), g, "nb_min_eps3")
    for k in (1, 16, 40):
        assert_csr_equal(O.build_graph(tok, k=k), g, f"knn{k}")
        assert_csr_equal(O.build_graph(tok, k=k, similarity=True), g, f"knn{k}_sim")
        assert_csr_equal(O.build_graph(tok, k=k), g, f"knn{k}_unstable", check_idx=False)
    assert_csr_equal(O.build_graph(tok, k=4, distance=O.minkowski), g, "knn4_min")
    assert_csr_equal(O.build_graph(tok, k=4, distance=O.minkowski, similarity=True), g, "knn4_min_sim")
    same(O.indexing(tok, 0, distances=2), g["ix_d2"])
    same(O.indexing(tok, 0, distances=[1, 2]), g["ix_d12"])
    same(O.indexing(tok, 0, positions=[int(x) for x in g["ix_pos_list"]]), g["ix_pos"])
    same(O.indexing(tok, 5, positions=[int(x) for x in g["ix_pos_ref5_list"]]), g["ix_pos_ref5"])
    mut = O.boolean_mutant_array(tok, 0)
    same(O.get_mutated_positions(mut, g["mutated_positions"], g["ix_pos_list"]), g["gmp"])
    same(O.calc_neighbours(tok, 9, eps=2), g["cn_9_eq2"])
    same(O.calc_neighbours(tok, 5, eps=0), g["cn_5_eq0"])
    same(np.where(O.neighbourhood_mask(tok, 9, 2))[0], g["nh_9_2"])


def test_best_effort_uint8_batch_matches_the_restatement():
    """bench.py's second CPU baseline (oracle.knn_batch_uint8) returns what the pinned
    restatement returns, ties included."""
    rng = np.random.default_rng(4)
    X = rng.integers(1, 4, size=(3000, 40)).astype(np.uint8)
    for threads, chunk in ((1, 32768), (3, 700)):
        idx, w = O.knn_batch_uint8(X, X[100:108], 16, threads=threads, chunk=chunk)
        D = O.hamming(X.astype(np.int64), X[100:108].astype(np.int64))
        ri, rw = O.knn_from_distances(D, 16)
        np.testing.assert_array_equal(idx, ri)
        np.testing.assert_array_equal(w, rw)


def test_c_restatement_matches_the_numpy_oracle():
    """oracle/hamming_knn_cpu.c (built by `make -C oracle`, i.e. by __graft_entry__.build()): packed
    bit planes + popcount + sorted k+1 lists return exactly what the pinned numpy restatement
    returns -- duplicates, heavy ties, ragged widths, fewer rows than k+1, row sub-ranges, threads."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "oracle")], check=True, capture_output=True)
    from oracle import c_oracle as CO
    rng = np.random.default_rng(1)
    for n, L, k, alphabet in ((3000, 256, 16, 20), (500, 40, 5, 3), (10, 3, 16, 3), (2000, 100, 31, 20), (700, 64, 1, 2)):
        X = rng.integers(1, alphabet + 1, size=(n, L)).astype(np.uint8)
        if n > 200:
            X[100] = X[99]
        planes = CO.pack(X)
        np.testing.assert_array_equal(planes, CO.pack(X, use_c=True))
        ri, rw = O.knn_from_distances(O.hamming(X.astype(np.int64), X.astype(np.int64)), k)
        kk = ri.shape[1]
        for threads in (1, 3):
            idx, d = CO.hamming_knn(planes, L, 0, n, k, threads=threads)
            np.testing.assert_array_equal(idx[:, :kk], ri)
            np.testing.assert_array_equal(d[:, :kk], rw)
            assert np.all(idx[:, kk:] == -1)
        lo, cnt = 7, min(13, n - 7)
        idx, d = CO.hamming_knn(planes, L, lo, cnt, k, threads=2)
        np.testing.assert_array_equal(idx[:, :kk], ri[lo:lo + cnt])
        # arbitrary query rows (what bench.py's parity block samples) and plain distance rows
        rows = np.array([n - 1, 0, n // 2, 3, 3], dtype=np.int64)
        idx, d = CO.hamming_knn_rows(planes, L, rows, k, threads=2)
        np.testing.assert_array_equal(idx[:, :kk], ri[rows])
        np.testing.assert_array_equal(d[:, :kk], rw[rows])
        D = CO.hamming_rows(planes, L, rows, threads=3)
        np.testing.assert_array_equal(D, O.hamming(X.astype(np.int64), X[rows].astype(np.int64)))
    with pytest.raises(ValueError):
        CO.hamming_knn_rows(planes, L, np.array([n]), k)
    with pytest.raises(OverflowError):
        CO.pack(np.full((4, 8), 40, dtype=np.uint8))
