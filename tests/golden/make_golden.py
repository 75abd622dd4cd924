#!/usr/bin/env python3
"""Generate golden vectors by running the UNMODIFIED reference (acmater/prograph)
on CPU in the build container.

This script is test infrastructure. It is the only place in the repo that imports
/root/reference, and it only runs in the build container (the GPU box has no
/root/reference).  Outputs are small .npz fixtures committed under tests/golden/.

Shims applied (none of them touch the arithmetic of the path):
  * a two-class ``colorama`` stub (imported at prograph/prograph.py:15, used only
    for terminal colours at :516);
  * ``torch.as_tensor(..., device="cuda:0")`` redirected to CPU, because
    prograph/prograph.py:726 hard-codes cuda:0 and this container has no GPU.
    The pandas Series of row arrays is stacked first (pandas 3 cannot hand an
    object Series to torch directly).
  * for the kNN *index* vectors only, a second run wraps ``torch.sort`` so that it
    is called with ``stable=True``.  The reference calls torch.sort with the
    default stable=False (prograph.py:758-760); on CPU that kernel is visibly
    unstable, on CUDA the N>4096 path is the stable kernel.  The parity contract
     is the stable order; the unmodified run's *weights* are also
    stored and must match bit for bit.

Usage:  python tests/golden/make_golden.py     (writes tests/golden/*.npz)
"""
import os
import sys
import io
import types
import operator
import contextlib
import tempfile

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

# ---- shims -----------------------------------------------------------------
colorama = types.ModuleType("colorama")
colorama.Fore = type("Fore", (), {"GREEN": ""})
colorama.Style = type("Style", (), {"RESET_ALL": ""})
sys.modules["colorama"] = colorama

import torch  # noqa: E402
import pandas as pd  # noqa: E402

_orig_as_tensor = torch.as_tensor


def _cpu_as_tensor(data, dtype=None, device=None):
    if isinstance(data, pd.Series):
        data = np.stack(list(data))
    return _orig_as_tensor(data, dtype=dtype)


torch.as_tensor = _cpu_as_tensor
sys.path.insert(0, REF)
import tqdm as _tqdm  # noqa: E402

_tqdm.tqdm = lambda it, *a, **k: it  # silence progress bars

from prograph import Prograph  # noqa: E402
from prograph.distance import hamming, minkowski  # noqa: E402

_orig_sort = torch.sort


@contextlib.contextmanager
def stable_sort():
    def _sort(x, dim=-1, descending=False, stable=False):
        return _orig_sort(x, dim=dim, descending=descending, stable=True)
    torch.sort = _sort
    try:
        yield
    finally:
        torch.sort = _orig_sort


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def csr(lst, prefix, out):
    """Flatten the reference's list of (idx, weight) tuples into CSR arrays."""
    indptr = np.zeros(len(lst) + 1, dtype=np.int64)
    for i, (a, _) in enumerate(lst):
        indptr[i + 1] = indptr[i] + len(a)
    idx = np.concatenate([np.asarray(a) for a, _ in lst]) if len(lst) else np.zeros(0, np.int64)
    # rows without neighbours carry int arrays even for float metrics (prograph.py:753)
    wl = [np.asarray(w) for _, w in lst if len(w)]
    w = np.concatenate(wl) if wl else np.zeros(0, np.int64)
    out[prefix + "_indptr"] = indptr
    out[prefix + "_idx"] = idx
    out[prefix + "_w"] = w
    out[prefix + "_idx_dtype"] = np.array(str(idx.dtype))
    out[prefix + "_w_dtype"] = np.array(str(w.dtype))


def t2n(t):
    return t.detach().cpu().numpy()


# ---- A. distance functions -------------------------------------------------
def gen_distance():
    out = {}
    rng = np.random.default_rng(1234)
    # the reference's own unit vectors (tests/tests.py:175-208)
    X = torch.Tensor([[1, 2, 3], [4, 5, 6]])
    Y = torch.Tensor([[1, 2, 3], [7, 8, 9]])
    out["t_X"], out["t_Y"] = t2n(X), t2n(Y)
    out["t_ham_2d2d"] = t2n(hamming(X, Y))
    out["t_ham_2d1d"] = t2n(hamming(X, Y[0]))
    out["t_ham_1d1d"] = t2n(hamming(X[1], Y[0]))
    out["t_min_2d2d"] = t2n(minkowski(X, Y))
    out["t_min_2d1d"] = t2n(minkowski(X, Y[0]))
    out["t_min_1d1d"] = t2n(minkowski(X[1], Y[0]))
    out["t_min_p1"] = t2n(minkowski(X, Y, p=1))
    out["t_min_p3"] = t2n(minkowski(X, Y, p=3))

    # random integer tokens, int64 (the dtype calc_neighbours/indexing use)
    Xi = rng.integers(0, 21, size=(37, 19)).astype(np.int64)
    Yi = rng.integers(0, 21, size=(11, 19)).astype(np.int64)
    Yi[3] = Xi[5]
    Yi[4, :10] = Xi[6, :10]
    out["i_X"], out["i_Y"] = Xi, Yi
    out["i_ham"] = t2n(hamming(Xi, Yi))
    out["i_ham_sim"] = t2n(hamming(Xi, Yi, similarity=True))
    for p in (1, 2, 3):
        out[f"i_min_p{p}"] = t2n(minkowski(Xi, Yi, p=p))
        out[f"i_min_p{p}_sim"] = t2n(minkowski(Xi, Yi, p=p, similarity=True))
    # fp16-staged tokens (what build_graph feeds, prograph.py:726)
    Xh, Yh = torch.as_tensor(Xi, dtype=torch.float16), torch.as_tensor(Yi, dtype=torch.float16)
    out["h_ham"] = t2n(hamming(Xh, Yh))
    out["h_ham_sim"] = t2n(hamming(Xh, Yh, similarity=True))
    for p in (1, 2, 3):
        out[f"h_min_p{p}"] = t2n(minkowski(Xh, Yh, p=p))
        out[f"h_min_p{p}_sim"] = t2n(minkowski(Xh, Yh, p=p, similarity=True))
    # ragged widths: zero right-pad of the narrower operand (distance/utils.py:32-38)
    Xr = rng.integers(1, 21, size=(5, 12)).astype(np.int64)
    Yr = rng.integers(1, 21, size=(3, 7)).astype(np.int64)
    Yr[1] = Xr[2, :7]
    out["r_X"], out["r_Y"] = Xr, Yr
    out["r_ham_xy"] = t2n(hamming(Xr, Yr))
    out["r_ham_yx"] = t2n(hamming(Yr, Xr))
    out["r_min_xy"] = t2n(minkowski(Xr, Yr))
    # real-valued embeddings, float32 and fp16
    Xf = rng.normal(size=(9, 5)).astype(np.float32)
    Yf = rng.normal(size=(4, 5)).astype(np.float32)
    out["f_X"], out["f_Y"] = Xf, Yf
    out["f_min_p2"] = t2n(minkowski(Xf, Yf))
    out["f_min_p2_sim"] = t2n(minkowski(Xf, Yf, similarity=True))
    out["f_ham"] = t2n(hamming(Xf, Yf))
    Xfh, Yfh = torch.as_tensor(Xf, dtype=torch.float16), torch.as_tensor(Yf, dtype=torch.float16)
    out["fh_min_p2"] = t2n(minkowski(Xfh, Yfh))
    out["fh_min_p2_sim"] = t2n(minkowski(Xfh, Yfh, similarity=True))
    # two-component embeddings (order-independent fp32 sum): larger fp16 case
    Xe = (rng.integers(0, 64, size=(64, 2)) / 8.0).astype(np.float32)
    out["e_X"] = Xe
    out["e_min_p2_h"] = t2n(minkowski(torch.as_tensor(Xe, dtype=torch.float16),
                                      torch.as_tensor(Xe[:16], dtype=torch.float16)))
    np.savez_compressed(os.path.join(OUT, "distance.npz"), **out)
    print("distance.npz", len(out))


# ---- B. synthetic_data.csv -------------------------------------------------
def gen_synthetic():
    out = {}
    os.chdir(REF)
    pg = quiet(Prograph, file="data/synthetic_data.csv")
    out["sequences"] = np.array(list(pg("Sequence")), dtype="U")
    out["fitness"] = pg("Fitness").to_numpy()
    out["tokenized"] = pg.tokenized
    out["mutated_positions"] = pg.mutated_positions
    out["mutant_array_seed"] = pg.sequence_mutation_locations
    out["mutant_array_LDC"] = pg.boolean_mutant_array("LDC")
    csr(list(pg("Neighbours")), "nb_eps1", out)
    csr(pg.build_graph(eps=2), "nb_eps2", out)
    csr(pg.build_graph(eps=2, comp=operator.lt), "nb_eps2_lt", out)
    csr(pg.build_graph(eps=2, comp=operator.eq), "nb_eps2_eq", out)
    csr(pg.build_graph(eps=1.5), "nb_eps1p5", out)
    csr(pg.build_graph(eps=2, similarity=True), "nb_eps2_sim", out)
    csr(pg.build_graph(eps=1, batch_size=7), "nb_eps1_b7", out)
    sub = np.arange(0, 1000, 7)
    out["sub_idxs"] = sub
    csr(pg.build_graph(eps=1, idxs=sub), "nb_eps1_sub", out)
    csr(pg.build_graph(eps=3, idxs=sub, comp=operator.ge), "nb_eps3_ge_sub", out)
    csr(pg.build_graph(eps=1, idxs=sub, comp=operator.gt), "nb_eps1_gt_sub", out)
    csr(pg.build_graph(eps=2, idxs=sub, comp=operator.ne), "nb_eps2_ne_sub", out)
    csr(pg.build_graph(eps=2, distance=minkowski), "nb_min_eps2", out)
    # kNN: unmodified run gives the weights; stable run gives the index contract
    for k in (1, 3, 16):
        csr(pg.build_graph(k=k), f"knn{k}_unstable", out)
        with stable_sort():
            csr(pg.build_graph(k=k), f"knn{k}", out)
            csr(pg.build_graph(k=k, similarity=True), f"knn{k}_sim", out)
    with stable_sort():
        csr(pg.build_graph(k=4, idxs=sub), "knn4_sub", out)
        csr(pg.build_graph(k=3, distance=minkowski), "knn3_min", out)
        csr(pg.build_graph(k=200, idxs=sub), "knn200_sub", out)   # k+1 > N(=143): slice shortens
    # indexing (prograph.py:254-343) and the calls pinned by tests/tests.py:41-53,92-98
    out["ix_pos12"] = pg.indexing(positions=[1, 2])
    out["ix_pos12_and"] = pg.indexing(positions=[1, 2], Bool="and")
    out["ix_pos0"] = pg.indexing(positions=[0])
    out["ix_d3"] = pg.indexing(distances=3)
    out["ix_d2"] = pg.indexing(distances=2)
    out["ix_d13"] = pg.indexing(distances=[1, 3])
    out["ix_pos12_d2"] = pg.indexing(positions=[1, 2], distances=2)
    a, b = pg.indexing(positions=[1, 2], distances=2, complement=True)
    out["ix_pos12_d2_c0"], out["ix_pos12_d2_c1"] = a, b
    out["ix_ref_LDC_pos1"] = pg.indexing(reference_seq="LDC", positions=[1])
    out["ix_ref_LDC_d1"] = pg.indexing(reference_seq="LDC", distances=1)
    out["ix_ref_500_d2_pos01"] = pg.indexing(reference_seq=500, distances=2, positions=[0, 1])
    out["gmp_0"] = pg.get_mutated_positions(np.array([0]))
    out["gmp_12"] = pg.get_mutated_positions(np.array([1, 2]))
    # distance-to-dataset queries
    out["cn_ACL"] = pg.calc_neighbours(seq="ACL")
    out["cn_ACL_eps2"] = pg.calc_neighbours(seq="ACL", eps=2)
    out["cn_ACL_le2"] = pg.calc_neighbours(seq="ACL", eps=2, comp=operator.le)
    out["cn_77_ge3"] = pg.calc_neighbours(seq=77, eps=3, comp=operator.ge)
    out["cn_77_min_le2"] = pg.calc_neighbours(seq=77, eps=2, distance=minkowski, comp=operator.le)
    out["nh_ACL_1"] = pg.neighbourhood("ACL", 1).index.to_numpy()
    out["nh_ACL_2"] = pg.neighbourhood("ACL", 2).index.to_numpy()
    d = hamming(pg.tokenized, pg.tokenized[0].reshape(1, -1))
    out["str_max"] = np.array(int(torch.max(d)))
    out["str_nuniq"] = np.array(len(np.unique(d)))
    out["adj_shape"] = np.array(pg.adjacency().shape)
    A = pg.adjacency()
    out["adj_row"], out["adj_col"], out["adj_data"] = A.row, A.col, A.data
    out["degree"] = pg.degree()
    out["degree_bool"] = pg.degree(boolean_weights=True)
    np.savez_compressed(os.path.join(OUT, "synthetic.npz"), **out)
    print("synthetic.npz", len(out))


# ---- C. knntest ------------------------------------------------------------
def gen_knntest():
    out = {}
    os.chdir(REF)
    kn = quiet(Prograph, "data/knntest_pgraph.pkl")
    out["sequences"] = np.array(list(kn("Sequence")), dtype="U")
    out["fitness"] = kn("Fitness").to_numpy()
    out["embedded"] = np.stack(list(kn("Embedded")))
    csr(list(kn("Neighbours")), "pkl_nb", out)
    fresh = quiet(Prograph, "data/knntest.csv")
    csr(list(fresh("Neighbours")), "csv_nb", out)
    E = torch.as_tensor(out["embedded"], dtype=torch.float16)
    out["dmat_h"] = t2n(minkowski(E, E))
    for k in (1, 2, 3, 5, 9):
        csr(kn.build_graph(representation="Embedded", k=k, distance=minkowski), f"knn{k}_unstable", out)
        with stable_sort():
            csr(kn.build_graph(representation="Embedded", k=k, distance=minkowski), f"knn{k}", out)
            csr(kn.build_graph(representation="Embedded", k=k, distance=minkowski, similarity=True),
                f"knn{k}_sim", out)
    csr(kn.build_graph(representation="Embedded", eps=2, distance=minkowski), "eps2", out)
    csr(kn.build_graph(representation="Embedded", eps=2, distance=minkowski, similarity=True), "eps2_sim", out)
    csr(kn.build_graph(representation="Embedded", eps=1.25, distance=minkowski, comp=operator.lt), "eps1p25_lt", out)
    csr(kn.build_graph(representation="Embedded", eps=0.1, distance=minkowski), "eps0p1", out)
    kn.graph["Weighted"] = kn.build_graph(eps=2, representation="Embedded", distance=minkowski)
    out["deg_eps2_bool"] = kn.degree(graph="Weighted", boolean_weights=True)
    kn.graph["Weighted"] = kn.build_graph(k=1, representation="Embedded", distance=minkowski)
    out["deg_k1"] = kn.degree(graph="Weighted")
    np.savez_compressed(os.path.join(OUT, "knntest.npz"), **out)
    print("knntest.npz", len(out))


# ---- D. random mutational library (ragged, duplicates) ----------------------
def gen_library():
    out = {}
    rng = np.random.default_rng(7)
    AA = np.array(list("ACDEFGHIKLMNPQRSTVWY"))
    N, L = 300, 23
    wt = rng.integers(0, 20, size=L)
    toks = np.tile(wt, (N, 1))
    for i in range(1, N):
        m = rng.integers(1, 5)
        pos = rng.choice(L - 3, size=m, replace=False)       # last 3 positions never mutate
        for q in pos:
            toks[i, q] = (toks[i, q] + rng.integers(1, 20)) % 20
    toks[17] = toks[5]                                        # duplicates exercise d>0
    toks[18] = toks[5]
    toks[250] = toks[3]
    seqs = ["".join(AA[r]) for r in toks]
    lens = np.full(N, L)
    for i in (40, 41, 42, 120):                               # ragged: shorter sequences get pad 0
        lens[i] = L - (i % 3) - 1
        seqs[i] = seqs[i][: lens[i]]
    fit = rng.normal(size=N)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "lib.csv")
        pd.DataFrame({"Sequence": seqs, "Fitness": fit}).to_csv(path)
        pg = quiet(Prograph, file=path)
    out["sequences"] = np.array(seqs, dtype="U")
    out["fitness"] = fit
    out["tokenized"] = pg.tokenized
    out["mutated_positions"] = pg.mutated_positions
    out["mutant_array_seed"] = pg.sequence_mutation_locations
    csr(list(pg("Neighbours")), "nb_eps1", out)
    csr(pg.build_graph(eps=3), "nb_eps3", out)
    csr(pg.build_graph(eps=3, similarity=True), "nb_eps3_sim", out)
    csr(pg.build_graph(eps=4, comp=operator.eq), "nb_eps4_eq", out)
    csr(pg.build_graph(eps=3.0, distance=minkowski), "nb_min_eps3", out)
    for k in (1, 16, 40):
        csr(pg.build_graph(k=k), f"knn{k}_unstable", out)
        with stable_sort():
            csr(pg.build_graph(k=k), f"knn{k}", out)
            csr(pg.build_graph(k=k, similarity=True), f"knn{k}_sim", out)
    with stable_sort():
        csr(pg.build_graph(k=4, distance=minkowski), "knn4_min", out)
        csr(pg.build_graph(k=4, distance=minkowski, similarity=True), "knn4_min_sim", out)
    out["ix_d2"] = pg.indexing(distances=2)
    out["ix_d12"] = pg.indexing(distances=[1, 2])
    mp = [int(x) for x in pg.mutated_positions[:3]]
    out["ix_pos_list"] = np.array(mp)
    out["ix_pos"] = pg.indexing(positions=mp)
    rp = sorted(set(mp) | set(int(x) for x in np.where(pg.tokenized[5] != pg.tokenized[0])[0]))
    out["ix_pos_ref5_list"] = np.array(rp)
    out["ix_pos_ref5"] = pg.indexing(reference_seq=5, positions=rp)
    out["gmp"] = pg.get_mutated_positions(np.array(mp))
    out["cn_9_eq2"] = pg.calc_neighbours(seq=9, eps=2)
    out["cn_5_eq0"] = pg.calc_neighbours(seq=5, eps=0)        # duplicates incl. self: no d>0 filter here
    out["nh_9_2"] = pg.neighbourhood(9, 2).index.to_numpy()
    D = t2n(hamming(pg.tokenized, pg.tokenized))
    out["dmat"] = D
    np.savez_compressed(os.path.join(OUT, "library.npz"), **out)
    print("library.npz", len(out))


if __name__ == "__main__":
    gen_distance()
    gen_synthetic()
    gen_knntest()
    gen_library()
