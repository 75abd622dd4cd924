"""Parity of the tcgen05 int8 Minkowski (p=2) path with the element-wise kernels, the oracle and
torch's stable sort.  Bit-exact: the contraction is exact in int32 and the rounding chain of
minkowski.py:36-40 is applied to the exact integer sum."""
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
    return t.cpu().numpy()


@pytest.mark.parametrize("L,hi", [(256, 21), (56, 21), (20, 21), (100, 32), (256, 256), (33, 200)])
def test_gemm_tile_matches_elementwise_and_oracle(eng, L, hi):
    rng = np.random.default_rng(L + hi)
    X = rng.integers(0, hi, size=(1000, L)).astype(np.int64)
    Y = rng.integers(0, hi, size=(300, L)).astype(np.int64)
    Y[7] = X[11]
    tx = eng.gemm_pack(X)
    ty = eng.gemm_pack(Y, K=tx.K)
    Xd, Yd = torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda()
    for sim in (False, True):
        got = eng.minkowski2_gemm_tile(tx, ty, 1, similarity=sim)
        ref = eng.minkowski_tile(Xd, Yd, 0, 300, p=2, similarity=sim)
        assert got.dtype == torch.float32
        np.testing.assert_array_equal(np_(got), np_(ref))
        np.testing.assert_array_equal(np_(got)[:40], O.minkowski(X, Y[:40], similarity=sim))
    if hi <= 32:     # fp16 chain: exact only while every squared difference is an fp16 integer
        txh = eng.gemm_pack(X, max_token=31)
        tyh = eng.gemm_pack(Y, max_token=31, K=txh.K)
        for sim in (False, True):
            got = eng.minkowski2_gemm_tile(txh, tyh, 0, similarity=sim)
            ref = eng.minkowski_tile(Xd.half(), Yd.half(), 0, 300, p=2, similarity=sim)
            assert got.dtype == torch.float16
            np.testing.assert_array_equal(np_(got).view(np.uint16), np_(ref).view(np.uint16))
    else:
        with pytest.raises(OverflowError):
            eng.gemm_pack(X, max_token=31)


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("sim", [False, True])
def test_gemm_knn_matches_stable_sort(eng, kind, sim):
    rng = np.random.default_rng(9 + kind)
    n, L = 3000, 256
    wt = rng.integers(1, 21, size=L)
    X = np.tile(wt, (n, 1))
    for i in range(1, n):
        pos = rng.choice(L, size=rng.integers(1, 9), replace=False)
        X[i, pos] = rng.integers(1, 21, size=len(pos))
    X[100] = X[99]
    tab = eng.gemm_pack(X, max_token=31 if kind == 0 else 255)
    tile = eng.minkowski2_gemm_tile(tab, tab, kind, similarity=sim)
    for k, drop in ((16, 1), (1, 0), (24, 1)):
        idx, val = eng.minkowski2_gemm_knn(tab, tab, k, drop, kind, similarity=sim)
        ref = torch.sort(tile.float().cpu(), dim=1, descending=sim, stable=True)
        np.testing.assert_array_equal(np_(idx), ref.indices[:, drop:drop + k].numpy())
        np.testing.assert_array_equal(np_(val).astype(np.float32), ref.values[:, drop:drop + k].numpy())


def test_gemm_knn_uniform_and_queries(eng):
    rng = np.random.default_rng(2)
    X = rng.integers(1, 21, size=(5000, 56)).astype(np.uint8)
    Q = rng.integers(1, 21, size=(700, 56)).astype(np.uint8)
    tx = eng.gemm_pack(X)
    tq = eng.gemm_pack(Q, K=tx.K)
    idx, val = eng.minkowski2_gemm_knn(tx, tq, 1, 0, 1)
    D = O.minkowski(X.astype(np.int64), Q.astype(np.int64))
    np.testing.assert_array_equal(np_(idx)[:, 0], np.argmin(D, axis=1))
    np.testing.assert_array_equal(np_(val)[:, 0], D.min(axis=1))


@pytest.mark.parametrize("sim", [False, True])
def test_gemm_eps_graph_matches_tile_threshold(eng, sim):
    """Fused epsilon epilogue (S-range test) against the threshold consumer on the materialised
    fp16 tile, for every ordering comparison."""
    import operator
    from prograph_b200 import _lib as L
    from prograph_b200.graph import minkowski_s_range
    rng = np.random.default_rng(21)
    n, Lw = 2500, 100
    wt = rng.integers(1, 21, size=Lw)
    X = np.tile(wt, (n, 1))
    for i in range(1, n):
        pos = rng.choice(Lw, size=rng.integers(1, 5), replace=False)
        X[i, pos] = rng.integers(1, 21, size=len(pos))
    X[40] = X[39]
    tab = eng.gemm_pack(X, max_token=31)
    tile = eng.minkowski2_gemm_tile(tab, tab, 0, similarity=sim)
    for comp, code, eps in ((operator.le, L.LE, 12.0), (operator.lt, L.LT, 9.5), (operator.ge, L.GE, 30.0),
                            (operator.gt, L.GT, 25.0), (operator.eq, L.EQ, 10.0), (operator.le, L.LE, 0.01)):
        e = 1 / (1 + eps) if sim else eps
        r = minkowski_s_range(tab.K * 31 * 31, comp, e, sim)
        assert r is not None
        ref = eng.tile_threshold(tile, code, e, swap=sim, guard=2 if sim else 1)
        if r[0] > r[1]:
            assert int(ref[0][-1]) == 0
            continue
        got = eng.minkowski2_gemm_eps(tab, tab, r[0], r[1], 0, similarity=sim)
        for a, b in zip(got, ref):
            np.testing.assert_array_equal(np_(a).view(np.uint16) if a.dtype == torch.float16 else np_(a),
                                          np_(b).view(np.uint16) if b.dtype == torch.float16 else np_(b))
    assert minkowski_s_range(1000, operator.ne, 3.0, False) is None
