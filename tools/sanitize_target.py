#!/usr/bin/env python3
"""CI-size pass over the kernels with hand-rolled synchronisation, for compute-sanitizer:

    compute-sanitizer --tool racecheck python tools/sanitize_target.py     # shared-memory hazards
    compute-sanitizer --tool memcheck  python tools/sanitize_target.py     # out-of-bounds / misaligned

symmetric kNN sweep (bootstrap + row locks + ring of bulk copies), symmetric epsilon sweep, the
one-sided sweeps, the tcgen05 Minkowski kernel (mbarrier pipeline, TMEM), the vectorised helpers.
Every result is also compared with the oracle, so a sanitizer-clean run is a correct run."""
import operator
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from oracle import prograph_oracle as O
    from prograph_b200 import minkowski, query
    from prograph_b200.engine import get_engine
    from prograph_b200.graph import distance_lut
    eng = get_engine()
    rng = np.random.default_rng(0)
    n, L, k = 2600, 256, 16
    wt = rng.integers(1, 21, size=L)
    X = np.tile(wt, (n, 1))
    for i in range(1, n):
        pos = rng.choice(L, size=rng.integers(1, 9), replace=False)
        X[i, pos] = (X[i, pos] - 1 + rng.integers(1, 20, size=len(pos))) % 20 + 1
    X = X.astype(np.uint8)
    D = O.hamming(X.astype(np.int64), X.astype(np.int64))
    ri, rw = O.knn_from_distances(D, k)
    tab = eng.pack(X)
    # symmetric kNN: bootstrap + sweep (two emulated ranks, bands) + merge + widen
    seed = eng.hamming_knn_boot(tab, 0, n, 512, k + 1)
    lists = torch.stack([eng.hamming_knn_sym(tab, k + 1, r, 2, lists=seed.clone(), boot_rows=512, mode=1) for r in range(2)])
    eng.sym_check()
    keys = eng.knn_lists_merge(lists, k, drop=1)
    idx, w = eng.knn_lists_finalize(keys, 0, n, k, 0)
    assert np.array_equal(idx.cpu().numpy(), ri) and np.array_equal(w.cpu().numpy(), rw), "symmetric kNN"
    # symmetric epsilon sweep -> CSR
    lut = distance_lut(tab.words * 32, operator.le, 3, False)
    ekeys, edges = eng.hamming_eps_sym(tab, lut)
    ip, ei, ew = eng.edge_keys_to_csr(ekeys, n, tab.words, edges)
    keep = (D <= 3) & (D > 0)
    assert np.array_equal(ei.cpu().numpy(), np.nonzero(keep)[1]), "symmetric eps"
    # one-sided sweeps: kNN, count / fill, tiles, flags
    oi, ow = eng.hamming_knn(tab, 0, n, tab, k, drop=1)
    assert np.array_equal(oi.cpu().numpy(), ri), "one-sided kNN"
    cp, ci, cw = eng.hamming_eps(tab, 0, n, tab, lut)
    assert np.array_equal(ci.cpu().numpy(), np.nonzero(keep)[1]), "count / fill"
    assert np.array_equal(eng.hamming_tile(tab, tab, 0, 512).cpu().numpy(), D[:512]), "tile"
    fl = eng.hamming_flags(tab, eng.gather_packed(tab, np.array([0, 9, 77])), 3, 0, 3).cpu().numpy().astype(bool)
    assert np.array_equal(fl, D[[0, 9, 77]] <= 3), "flags"
    # tcgen05 Minkowski kernel: kNN, tile, epsilon graph
    g = eng.gemm_pack(X, max_token=31)
    mi, mv = eng.minkowski2_gemm_knn(g, g, 4, 1, 0)
    Md = O.minkowski(X.astype(np.float16), X.astype(np.float16))
    qi, qv = O.knn_from_distances(Md, 4)
    assert np.array_equal(mi.cpu().numpy(), qi), "gemm kNN"
    assert np.array_equal(eng.minkowski2_gemm_tile(g, g, 0).cpu().numpy(), Md), "gemm tile"
    ni, nv = query.nearest(query.Library(X.astype(np.int64)), X[:300].astype(np.int64), distance=minkowski)
    assert int(nv.max().item()) == 0, "query.nearest"
    # helpers
    letters = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
    lut256 = np.zeros(256, dtype=np.uint8)
    lut256[letters] = np.arange(1, 21)
    assert torch.equal(eng.pack_chars(letters[X - 1], lut256).data, tab.data), "pack_chars"
    assert np.array_equal(eng.mutant_bool(tab, tab.row(0)).cpu().numpy().astype(bool), X != X[0]), "mutant_bool"
    torch.cuda.synchronize()
    print("sanitize target ok")


if __name__ == "__main__":
    main()
