"""CPU-side checks: the C-ABI library loads and exports every declared symbol, and the host
logic (truth tables, CSR views, row sharding, result gathering over gloo) is right.  No
kernel is launched here."""
import operator
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from prograph_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "prograph_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pg_version() >= 100
    # geometry helpers are pure host functions
    assert [lib.pg_packed_words(L) for L in (1, 32, 33, 56, 65, 128, 129, 256, 257, 600)] == \
        [1, 1, 2, 2, 4, 4, 8, 8, 16, 24]
    assert lib.pg_packed_rows(1) == 512 and lib.pg_packed_rows(1_000_000) == 1_000_448
    assert lib.pg_packed_bytes(1000, 256, 5) == 1024 * 5 * 8 * 4


def test_ctypes_signatures_match_the_header():
    """Every declaration of include/prograph_b200.h against the ctypes table of prograph_b200/_lib.py:
    same number of parameters, and the same class of type in every position (pointer / 32-bit int /
    64-bit int / size_t / double) -- a drifted binding would corrupt arguments silently."""
    import ctypes as C
    from prograph_b200 import _lib
    header = open(os.path.join(ROOT, "include", "prograph_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    decls = re.findall(r"\b(?:int|size_t|int64_t|const char\*)\s+(pg_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", header)
    assert len(decls) == len(_lib.SIGNATURES)

    def kind_of_c(param):
        param = param.strip()
        if "*" in param:
            return "ptr"
        base = param.rsplit(" ", 1)[0].replace("const", "").strip() if " " in param else param
        return {"int": "i32", "int64_t": "i64", "size_t": "size", "double": "f64", "uint64_t": "i64"}[base]

    def kind_of_ctypes(t):
        if t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or issubclass(t, C._Pointer):
            return "ptr"
        return {C.c_int: "i32", C.c_int64: "i64", C.c_size_t: "size", C.c_double: "f64"}[t]

    for name, params in decls:
        params = [q for q in params.split(",") if q.strip() and q.strip() != "void"]
        _, argtypes = _lib.SIGNATURES[name]
        assert len(params) == len(argtypes), (name, len(params), len(argtypes))
        for i, (q, t) in enumerate(zip(params, argtypes)):
            assert kind_of_c(q) == kind_of_ctypes(t), (name, i, q.strip(), t)


def test_argument_errors_map_to_reference_exceptions():
    """Bad arguments are rejected on the host side of the ABI, before any launch."""
    from prograph_b200 import _lib
    lib = _lib.load()
    rc = lib.pg_pack_tokens(None, 0, 10, 10, 10, None, 5, 1, None, None)
    assert rc == _lib.ERR_INVALID
    with pytest.raises(ValueError):
        _lib.check(rc)
    assert "null" in _lib.last_error()
    rc = lib.pg_hamming_knn(None, 10, 0, 10, None, 10, 5, 8, 16, 1, 0, None, None, None, 0, None)
    assert rc == _lib.ERR_INVALID


def test_engine_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from prograph_b200.engine import CudaEngine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CudaEngine()
    import prograph_b200
    with pytest.raises(RuntimeError):
        prograph_b200.hamming(np.ones((2, 3)), np.ones((1, 3)))


def test_clean_input_contract():
    from prograph_b200 import clean_input
    X, Y = clean_input(torch.tensor([4, 5, 6]), torch.tensor([[1, 2, 3, 4, 5]]))
    assert X.shape == (1, 5) and Y.shape == (1, 5) and X[0, 3:].tolist() == [0, 0]
    with pytest.raises(ValueError):
        clean_input(torch.Tensor([4, 5, 6]), torch.Tensor())
    with pytest.raises(ValueError):
        clean_input(np.zeros((0, 3)), np.zeros((2, 3)))


def test_validate_and_lut():
    from prograph_b200.graph import validate, distance_lut
    for bad in (dict(eps=None, k=None), dict(eps=1, k=2), dict(eps=0, k=None), dict(eps=None, k=0)):
        with pytest.raises(ValueError):
            validate(**bad)
    with pytest.raises(TypeError):
        validate(None, 0.5)
    validate(1, None)
    validate(None, 3)

    def bits(words):
        return [d for d in range(len(words) * 32) if (words[d >> 5] >> (d & 31)) & 1]
    assert bits(distance_lut(64, operator.le, 2, False)) == [1, 2]
    assert bits(distance_lut(64, operator.lt, 2, False)) == [1]
    assert bits(distance_lut(64, operator.eq, 3, False)) == [3]
    assert bits(distance_lut(64, operator.le, 1.5, False)) == [1]
    assert bits(distance_lut(64, operator.ge, 63, False)) == [63, 64]
    assert bits(distance_lut(64, operator.ne, 2, False)) == [d for d in range(1, 65) if d != 2]
    assert bits(distance_lut(64, operator.eq, 0, False, guard=False)) == [0]
    # similarity: eps already transformed, operands swapped, s < 1 guard (prograph.py:721,734)
    assert bits(distance_lut(64, operator.le, 1 / (1 + 2), True)) == [1, 2]
    assert bits(distance_lut(64, lambda a, b: a <= b, 2, False)) == [1, 2]
    assert bits(distance_lut(64, lambda a, b: True, 2, False)) == list(range(1, 65))   # broadcasts like torch


def test_tables_as_reference_lists():
    from prograph_b200.graph import NeighbourTable, KnnTable
    t = NeighbourTable(np.array([0, 2, 2, 3]), np.array([5, 7, 1]), np.array([1.5, 2.5, 0.5], dtype=np.float16))
    lst = t.as_list()
    assert len(lst) == 3 and lst[0][0].tolist() == [5, 7] and lst[0][1].dtype == np.float16
    assert len(lst[1][0]) == 0 and lst[1][0].dtype == np.dtype(int) and lst[1][1].dtype == np.dtype(int)
    assert np.shares_memory(lst[2][0], t.idx)
    kt = KnnTable(np.arange(6).reshape(3, 2), np.ones((3, 2)))
    assert [a.tolist() for a, _ in kt.as_list()] == [[0, 1], [2, 3], [4, 5]]
    assert kt.to_csr().indptr.tolist() == [0, 2, 4, 6]


def test_row_ranges_cover_everything():
    from prograph_b200 import shard
    for n in (1, 5, 8, 1000, 1_000_000, 999_999):
        for world in (1, 2, 4, 8):
            covered = []
            for r in range(world):
                r0, rows = shard.row_range(n, r, world)
                if n < world or world == 1:
                    assert (r0, rows) == (0, n)
                else:
                    covered.extend([(r0, rows)])
            if covered:
                assert covered[0][0] == 0 and sum(c[1] for c in covered) == n
                for (a, an), (b, _) in zip(covered, covered[1:]):
                    assert a + an == b
    r0, rows = shard.row_range(1_000_000, 7, 8)
    assert r0 % 512 == 0 and rows > 0
    for n, world in ((5, 4), (9, 8), (2100, 4), (513, 2)):
        assert all(shard.row_range(n, r, world)[1] > 0 for r in range(world))


def _gloo_worker(rank, world, port, n, k, tmpdir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from _cpu_engine import CheckerEngine
    from prograph_b200.graph import build_neighbours
    rng = np.random.default_rng(3)
    X = rng.integers(1, 4, size=(n, 12)).astype(np.int64)
    eng = CheckerEngine()
    knn = build_neighbours(X, k=k, engine=eng)
    eps = build_neighbours(X, eps=2, engine=eng)
    sim = build_neighbours(X, eps=2, similarity=True, engine=eng)
    np.savez(os.path.join(tmpdir, f"r{rank}.npz"), idx=knn.idx, w=knn.w, indptr=eps.indptr, eidx=eps.idx, ew=eps.w,
             sindptr=sim.indptr, sidx=sim.idx, sw=sim.w, rows=np.array(eng.rows_seen))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [37, 1301])
def test_sharded_build_over_gloo(tmp_path, n):
    """world_size 2 on CPU: each rank computes only its row block (through a checker-backed
    engine standing in for the CUDA one), the all-gathers rebuild the whole graph on both."""
    import torch.multiprocessing as mp
    from oracle import prograph_oracle as O
    k, world = 5, 2
    port = 29600 + (os.getpid() + n) % 300
    mp.spawn(_gloo_worker, args=(world, port, n, k, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(3)
    X = rng.integers(1, 4, size=(n, 12)).astype(np.int64)
    D = O.hamming(X, X)
    ri, rw = O.knn_from_distances(D, k)
    indptr, eidx, ew = O.to_csr(O.build_graph(X, eps=2))
    sindptr, sidx, sw = O.to_csr(O.build_graph(X, eps=2, similarity=True))
    seen = []
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        np.testing.assert_array_equal(z["idx"], ri)
        np.testing.assert_array_equal(z["w"], rw)
        np.testing.assert_array_equal(z["indptr"], indptr)
        np.testing.assert_array_equal(z["eidx"], eidx)
        np.testing.assert_array_equal(z["ew"], ew)
        np.testing.assert_array_equal(z["sindptr"], sindptr)
        np.testing.assert_array_equal(z["sidx"], sidx)
        np.testing.assert_array_equal(z["sw"], sw)
        seen.append(tuple(z["rows"]))
    # the two ranks worked on disjoint row blocks that tile [0, n)
    assert seen[0][0] == 0 and seen[0][0] + seen[0][1] == seen[1][0] and seen[1][0] + seen[1][1] == n


def _sym_table(n):
    """12 varying sites (3 letters each) inside 70-residue copies of one wild type: the sharded builds
    below also go through graph.informative_table (70 residues = 4 words -> 12 positions = 1 word)."""
    rng = np.random.default_rng(5)
    sites = rng.integers(1, 4, size=(n, 12)).astype(np.int64)
    X = np.tile(np.arange(70, dtype=np.int64) % 20 + 1, (n, 1))
    X[:, 3:63:5] = sites
    return X


def _gloo_sym_worker(rank, world, port, n, k, boot_div, tmpdir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from _cpu_engine import CheckerEngine
    from prograph_b200 import graph
    graph.SYM_MIN_ROWS, graph.SYM_BOOT_DIV = 0, boot_div      # route this small table through the symmetric build
    graph.COMPACT_MIN_ROWS = 1                                # ... on its informative columns
    X = _sym_table(n)
    eng = CheckerEngine()
    graph.SYM_EPS_MIN_ROWS = 0
    knn = graph.build_neighbours(X, k=k, engine=eng)
    sim = graph.build_neighbours(X, k=k, similarity=True, engine=eng)
    eps = graph.build_neighbours(X, eps=2, engine=eng)
    esim = graph.build_neighbours(X, eps=2, similarity=True, engine=eng)
    odd = graph.build_neighbours(X, eps=2, comp=operator.ne, engine=eng)     # not a range: one-sided path
    np.savez(os.path.join(tmpdir, f"r{rank}.npz"), idx=knn.idx, w=knn.w, sidx=sim.idx, sw=sim.w,
             eindptr=eps.indptr, eidx=eps.idx, ew=eps.w, sindptr=esim.indptr, seidx=esim.idx, sew=esim.w,
             oindptr=odd.indptr, oidx=odd.idx, ow=odd.w,
             calls=np.array([eng.sym_calls, graph.sym_boot_rows(n), min(eng.sym_modes), eng.eps_sym_calls, eng.compacted]))
    dist.destroy_process_group()


@pytest.mark.parametrize("n,boot_div", [(700, 8), (1301, 2)])
def test_symmetric_build_over_gloo(tmp_path, n, boot_div):
    """world_size 2 on CPU, symmetric kNN build: ranks take column bands of the triangle, exchange
    their candidate lists (all-to-all: every rank receives all ranks' lists of its rows), merge and
    all-gather the merged keys; with and without the bootstrap pass."""
    import torch.multiprocessing as mp
    from oracle import prograph_oracle as O
    k, world = 5, 2
    port = 29950 + (os.getpid() + n) % 40
    mp.spawn(_gloo_sym_worker, args=(world, port, n, k, boot_div, str(tmp_path)), nprocs=world, join=True)
    X = _sym_table(n)
    ri, rw = O.knn_from_distances(O.hamming(X, X), k)
    si, sw = O.knn_from_distances(O.hamming(X, X, similarity=True), k, descending=True)
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        np.testing.assert_array_equal(z["idx"], ri)
        np.testing.assert_array_equal(z["w"], rw)
        np.testing.assert_array_equal(z["sidx"], si)
        np.testing.assert_array_equal(z["sw"], sw)
        assert z["calls"][0] == 2                       # both builds went through the symmetric sweep
        assert z["calls"][1] == (512 if boot_div == 2 else 0)
        assert z["calls"][2] == 1                       # ranks took column bands of the triangle
        for got, want in (("e", O.build_graph(X, eps=2)), ("s", O.build_graph(X, eps=2, similarity=True)),
                          ("o", O.build_graph(X, eps=2, comp=operator.ne))):
            indptr, eidx, ew = O.to_csr(want)
            np.testing.assert_array_equal(z[got + "indptr"], indptr)
            np.testing.assert_array_equal(z[{"e": "eidx", "s": "seidx", "o": "oidx"}[got]], eidx)
            np.testing.assert_array_equal(z[{"e": "ew", "s": "sew", "o": "ow"}[got]], ew)
        # eps=2 and its similarity form took the symmetric sweep; `d != 2` keeps nearly every pair
        # (dense: above graph.SYM_EPS_MAX_DEGREE for n=1301, not a range for n=700) -> one-sided passes
        assert z["calls"][3] == 2
        assert z["calls"][4] == 5                       # every build swept the 12 informative columns only


def test_symmetric_band_planner_partitions_and_balances():
    """pg_knn_sym_band (host-only planner of the multi-GPU symmetric builds): the bands of stream
    rows tile [0, n) in order, start on stream-tile boundaries, and hold equal shares of the
    triangle's pair evaluations (to within one tile column of the triangle)."""
    import ctypes as C
    from prograph_b200 import _lib
    lib = _lib.load()

    def band(n, words, boot, part, parts):
        a, b = C.c_int64(0), C.c_int64(0)
        _lib.check(lib.pg_knn_sym_band(n, words, boot, part, parts, C.byref(a), C.byref(b)))
        return a.value, b.value

    def evaluations(n, boot, lo, hi):
        """pairs swept by all 256-row blocks inside stream rows [lo, hi)"""
        total = 0
        for a in range(0, n, 256):
            b = min(n, a + 256)
            start = max(boot if a < boot else a, lo)
            total += (b - a) * max(0, hi - start)
        return total

    for n, words, boot in ((1_000_000, 8, 8192), (160_000, 2, 8192), (70_000, 8, 8192), (5000, 1, 0), (513, 16, 512)):
        tile = 32 if words > 16 else 512 // words
        for parts in (1, 2, 3, 8):
            bands = [band(n, words, boot, g, parts) for g in range(parts)]
            assert bands[0][0] == 0 and bands[-1][1] == n
            for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(a % tile == 0 or a == n for a, _ in bands)
            work = [evaluations(n, boot, a, b) for a, b in bands]
            assert sum(work) == evaluations(n, boot, 0, n)
            if n >= 70_000:
                ideal = sum(work) / parts
                assert max(work) - ideal <= n * tile + 1, (n, parts, work)      # one tile column of slack
    with pytest.raises(ValueError):
        band(1000, 8, 0, 3, 3)


def test_symmetric_build_policies():
    """Host-side choices around the symmetric sweeps: bootstrap size, degree-sample rows, and the
    switches that keep small / unusual builds on the one-sided kernels."""
    from prograph_b200 import graph

    class Eng:                      # the attribute _sym_enabled looks at
        SYM_MAX_LIST = 32

    class Tab:
        def __init__(self, rows):
            self.rows = rows

    assert graph.sym_boot_rows(1_000_000) == 8192 and graph.sym_boot_rows(70_000) == 8192
    assert graph.sym_boot_rows(65_536) == 8192 and graph.sym_boot_rows(40_000) == 4608
    assert graph.sym_boot_rows(4000) == 0 and graph.sym_boot_rows(4096) == 512
    assert all(graph.sym_boot_rows(n) % 512 == 0 and graph.sym_boot_rows(n) <= max(0, n // 8) for n in range(1, 200_000, 997))
    for n in (5, 600, 32_768, 160_000, 1_000_000):
        r0, rows = graph._eps_sample(n)
        assert 0 <= r0 and r0 + rows <= n and rows >= min(n, 512) and rows <= 2048 and r0 % 512 == 0
    old = os.environ.pop("PG_KNN_SYM", None)
    try:
        assert graph._sym_enabled(Eng(), Tab(graph.SYM_MIN_ROWS), 17, 1)
        assert not graph._sym_enabled(Eng(), Tab(graph.SYM_MIN_ROWS - 1), 17, 1)
        assert not graph._sym_enabled(Eng(), Tab(1_000_000), 33, 1)          # list too long for the symmetric sweep
        os.environ["PG_KNN_SYM"] = "0"
        assert not graph._sym_enabled(Eng(), Tab(1_000_000), 17, 1)
        os.environ["PG_KNN_SYM"] = "1"
        assert graph._sym_enabled(Eng(), Tab(100), 17, 1)
    finally:
        os.environ.pop("PG_KNN_SYM", None)
        if old is not None:
            os.environ["PG_KNN_SYM"] = old


def test_constant_columns_are_dropped_without_changing_the_graph(monkeypatch):
    """graph.informative_table: a library built around one wild type (here 9 of 70 positions vary) is
    swept on its varying positions only -- same kNN lists and the same epsilon CSR as the oracle
    computes on the full rows (hamming.py:34, prograph.py:731-765); a table whose varying positions
    do not fit a narrower kernel instantiation is left alone."""
    from _cpu_engine import CheckerEngine
    from oracle import prograph_oracle as O
    from prograph_b200 import build_neighbours, graph
    rng = np.random.default_rng(5)
    n, L = 300, 70
    wt = rng.integers(1, 21, size=L)
    X = np.tile(wt, (n, 1))
    sites = np.sort(rng.choice(L, size=9, replace=False))
    X[:, sites] = rng.integers(1, 21, size=(n, 9))
    X[0] = wt
    monkeypatch.setattr(graph, "COMPACT_MIN_ROWS", 1)
    eng = CheckerEngine()
    tab = eng.pack(X)
    small = graph.informative_table(eng, tab)
    assert small.words == 1 and small.L == 9 and np.array_equal(small.tokens, X[:, sites])
    D = O.hamming(X, X)
    knn = build_neighbours(X, k=5, engine=eng)
    ri, rw = O.knn_from_distances(D, 5)
    assert np.array_equal(knn.idx, ri) and np.array_equal(knn.w, rw)
    g = build_neighbours(X, eps=2, engine=eng)
    keep = (D <= 2) & (D > 0)
    rows, cols = np.nonzero(keep)
    assert np.array_equal(g.indptr, np.concatenate([[0], np.cumsum(keep.sum(1))]))
    assert np.array_equal(g.idx, cols) and np.array_equal(g.w, D[rows, cols])
    assert eng.compacted == 3
    # 40 varying positions still need two words: nothing to gain, the table stays as it is
    Y = rng.integers(1, 21, size=(n, 40))
    taby = eng.pack(Y)
    assert graph.informative_table(eng, taby) is taby
    # identical rows: one (empty) column is left and every distance is 0
    Z = np.tile(wt, (n, 1))
    z = graph.informative_table(eng, eng.pack(Z))
    assert z.words == 1 and not z.tokens.any()
    monkeypatch.setattr(graph, "COMPACT_MIN_ROWS", 10**9)
    fresh = eng.pack(X)
    assert graph.informative_table(eng, fresh) is fresh


def test_dense_graph_capture_workspace_plan():
    """pg_eps_workspace_bytes_capture (host-only arithmetic): no request or a small one = the default
    128 slots per (split,row); a dense graph's request cuts the column range into fewer splits so that
    the captures stay within 8 GiB; a request that one split cannot hold falls back to the default."""
    from prograph_b200 import _lib
    lib = _lib.load()
    rows, words = 160_000, 1
    default = lib.pg_eps_workspace_bytes(rows, rows, words)
    counts_only = lib.pg_eps_count_workspace_bytes(rows, rows, words)
    n_splits = (counts_only - 256) // (rows * 8)
    assert n_splits >= 1 and (counts_only - 256) % (rows * 8) == 0
    assert lib.pg_eps_workspace_bytes_capture(rows, rows, words, 0) == default
    assert lib.pg_eps_workspace_bytes_capture(rows, rows, words, 100) == default
    cap = 2866                                             # C3 eps=2: 1.25 x 2242 + 64
    dense = lib.pg_eps_workspace_bytes_capture(rows, rows, words, cap)
    capture_bytes = 8 << 30
    splits = min(n_splits, capture_bytes // (rows * cap * 8))
    assert 1 <= splits < n_splits
    aux = 256 + -(-rows // 256) * 256 + rows * 8
    sizes = {s * rows * 8 + aux + s * rows * cap * 8 for s in range(1, splits + 1)}
    assert dense in sizes and dense - aux <= capture_bytes + splits * rows * 8
    assert lib.pg_eps_workspace_bytes_capture(rows, rows, words, 1 << 20) == default      # 1.3 TB: not even one split
    small = lib.pg_eps_workspace_bytes_capture(3000, 3000, 8, 1000)                       # small tables keep their splits
    assert small > lib.pg_eps_workspace_bytes(3000, 3000, 8)


def test_symmetric_planner_covers_the_triangle_once_in_l2_bands():
    """pg_knn_sym_plan (host-only): over all ranks the work items cover every (row block, stream tile)
    of the triangle exactly once; every CTA walks its items band by band (so that co-resident CTAs
    stream the same L2-sized part of the table); the CTAs' loads are balanced to a fraction of a percent."""
    from prograph_b200.engine import sym_plan
    for n, planes, words, boot, world in ((300_000, 5, 8, 8192, 1), (300_000, 5, 8, 8192, 3), (160_000, 5, 2, 8192, 2),
                                          (70_000, 8, 4, 0, 1), (5_000, 5, 1, 512, 2)):
        tile = 512 // words
        n_tiles, n_blocks = -(-n // tile), -(-n // 256)
        seen = np.zeros((n_blocks, n_tiles), dtype=np.int32)
        for rank in range(world):
            it = sym_plan(n, planes, words, boot, rank, world, 1 if world > 1 else 0, grid=296)
            assert np.all(it[:, 2] > it[:, 1])
            for rb, t0, t1, bt, cta in it:
                seen[rb, t0:t1] += 1
                assert bt == (1 if rb * 256 < boot else 0)
            # per-CTA order: first tiles never go back by more than one band
            load = np.bincount(it[:, 4], weights=(it[:, 2] - it[:, 1]).astype(np.float64))
            if n >= 160_000:
                # narrow rows trade balance for fewer list merges (chunks of at least 32 * 8/W tiles)
                limit = 1.02 if words >= 8 else 1.06
                assert load.max() / load.mean() < limit, (n, world, rank, load.max() / load.mean())
            band_tiles = max(64, int(24 * 2**20 / (tile * planes * words * 4)))
            for cta in np.unique(it[:, 4])[:8]:
                mine = it[it[:, 4] == cta]
                starts = mine[:, 1] // band_tiles
                assert np.all(np.diff(starts) >= -1), (n, world, cta)
        # the triangle: block rb sweeps the tiles from its own diagonal tile on (boot blocks: from `boot` on)
        want = np.zeros_like(seen)
        for rb in range(n_blocks):
            first = (boot if rb * 256 < boot else rb * 256) // tile
            want[rb, first:] = 1
        np.testing.assert_array_equal(seen, want)


def test_row_sums_match_the_reference_per_row_sum_bit_for_bit():
    """graph.row_sums_f32 (degree of a weighted graph, prograph.py:819-821) equals a Python loop of
    np.sum over the rows' float32 weights, for ragged rows, fp16 / float32 / integer weights."""
    from prograph_b200.graph import row_sums_f32
    rng = np.random.default_rng(0)
    deg = rng.integers(0, 300, size=1500)
    deg[5], deg[7] = 0, 1000
    indptr = np.concatenate([[0], np.cumsum(deg)])
    for w in (rng.random(indptr[-1]).astype(np.float16), (rng.random(indptr[-1]) * 1000).astype(np.float32),
              rng.integers(0, 50, indptr[-1])):
        ref = np.array([np.sum(w[indptr[r]:indptr[r + 1]].astype(np.float32)) for r in range(len(deg))], dtype=np.float32)
        got = row_sums_f32(indptr, w)
        assert got.dtype == np.float32
        np.testing.assert_array_equal(got, ref)


def test_sidecar_fingerprint_rejects_a_rewritten_pickle(tmp_path):
    """io.attach_sidecar only registers <name>.graph.npz when its fingerprint matches the pickled
    Neighbours column (a pickle rewritten by another writer must not pick up stale CSR arrays)."""
    import pandas as pd
    from prograph_b200 import io as pio
    from prograph_b200.graph import NeighbourTable

    class Holder:
        def __init__(self, frame):
            self.graph, self.file, self.remembered = frame, str(tmp_path / "lib.csv"), None

        def _table_for(self, name):
            col = list(self.graph[name])
            indptr = np.concatenate([[0], np.cumsum([len(x[0]) for x in col])])
            return NeighbourTable(indptr, np.concatenate([x[0] for x in col]), np.concatenate([x[1] for x in col]))

        def _remember(self, table, lists):
            self.remembered = table

    nb = [(np.array([1, 2]), np.array([1, 1])), (np.array([0]), np.array([1])), (np.array([0]), np.array([1]))]
    h = Holder(pd.DataFrame({"Sequence": ["AA", "AC", "CA"], "Neighbours": nb, "Tokenized": [0, 1, 2]}))
    pio.save(h, name="lib_pgraph", directory=str(tmp_path) + "/")
    pkl = str(tmp_path / "lib_pgraph.pkl")
    loaded = Holder(pd.read_pickle(pkl))
    assert pio.attach_sidecar(loaded, pkl) and loaded.remembered.n_rows == 3
    # same row count and same first row, but another graph: the round-1 check accepted this
    frame = pd.read_pickle(pkl)
    frame.at[2, "Neighbours"] = (np.array([0, 1]), np.array([1, 2]))
    stale = Holder(frame)
    assert not pio.attach_sidecar(stale, pkl) and stale.remembered is None


def _gloo_output_worker(rank, world, port, n, tmpdir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from _cpu_engine import CheckerEngine
    from prograph_b200 import graph, shard
    graph.SYM_MIN_ROWS = graph.SYM_EPS_MIN_ROWS = 0
    rng = np.random.default_rng(8)
    X = rng.integers(1, 4, size=(n, 12)).astype(np.int64)
    eng = CheckerEngine()
    knn = graph.build_neighbours(X, k=4, engine=eng, output="sharded")
    eps = graph.build_neighbours(X, eps=2, engine=eng, output="sharded")
    code = shard.agree(shard.NO_MEMORY if rank == 1 else shard.OK, world, None, torch.device("cpu"))
    np.savez(os.path.join(tmpdir, f"o{rank}.npz"), idx=knn.idx, w=knn.w, row0=knn.row0, indptr=eps.indptr, eidx=eps.idx,
             ew=eps.w, erow0=eps.row0, code=code, checks=eng.sym_checks)
    dist.destroy_process_group()


def test_sharded_output_and_collective_agreement_over_gloo(tmp_path):
    """output="sharded": every rank returns its own row block (symmetric path: all-to-all of the lists,
    merge, no final all-gather) and the blocks tile the oracle's graph; shard.agree hands every rank
    the worst status any rank reported, so that all of them leave a failing build together."""
    import torch.multiprocessing as mp
    from oracle import prograph_oracle as O
    from prograph_b200 import shard
    n, world = 1301, 2
    port = 29870 + os.getpid() % 50
    mp.spawn(_gloo_output_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(8)
    X = rng.integers(1, 4, size=(n, 12)).astype(np.int64)
    ri, rw = O.knn_from_distances(O.hamming(X, X), 4)
    indptr, eidx, ew = O.to_csr(O.build_graph(X, eps=2))
    rows_seen = 0
    for r in range(world):
        z = np.load(tmp_path / f"o{r}.npz")
        r0, rows = shard.row_range(n, r, world)
        assert int(z["row0"]) == r0 == int(z["erow0"]) and z["idx"].shape[0] == rows
        np.testing.assert_array_equal(z["idx"], ri[r0:r0 + rows])
        np.testing.assert_array_equal(z["w"], rw[r0:r0 + rows])
        a, b = indptr[r0], indptr[r0 + rows]
        np.testing.assert_array_equal(z["indptr"], indptr[r0:r0 + rows + 1] - a)
        np.testing.assert_array_equal(z["eidx"], eidx[a:b])
        np.testing.assert_array_equal(z["ew"], ew[a:b])
        assert int(z["code"]) == shard.NO_MEMORY and int(z["checks"]) >= 1
        rows_seen += rows
    assert rows_seen == n


def test_methods_outside_the_path_come_from_the_reference_when_it_is_installed(monkeypatch):
    """Prograph.__call__("sklearn") / fit / graph_to_networkx are not part of the path: they are looked
    up on the reference's class (bound to our object) when `prograph` is importable, and fail with a
    clear message when it is not.  Checked with a stand-in module -- no GPU, no construction."""
    import types
    from prograph_b200.prograph import Prograph
    pg = object.__new__(Prograph)                       # no __init__: nothing here touches the device
    pg.__dict__["graph"] = {"Sequence": "frame"}
    monkeypatch.setitem(sys.modules, "prograph", None)  # import prograph -> ImportError
    with pytest.raises(AttributeError, match="outside the graph-construction path"):
        pg.sklearn_data()
    with pytest.raises(AttributeError):
        pg("sklearn")

    class RefPrograph:
        def sklearn_data(self, split=0.8):
            return ("reference adapter", self.graph, split)

    monkeypatch.setitem(sys.modules, "prograph", types.SimpleNamespace(Prograph=RefPrograph))
    assert pg("sklearn", split=0.5) == ("reference adapter", {"Sequence": "frame"}, 0.5)
    assert pg.sklearn_data() == ("reference adapter", {"Sequence": "frame"}, 0.8)


def test_bench_workload_helpers():
    """bench.py's synthetic inputs and its accounting of evaluated pairs (host-only): the GB1-style
    library has the known epsilon-graph answers of SURVEY.md 8c, the query set shares the library's
    wild type, and the pair count of the symmetric sweep equals what the library's planner hands out."""
    import bench
    from prograph_b200.engine import sym_plan
    G = bench.make_gb1_library()
    assert G.shape == (160_000, 56) and G.dtype == np.uint8 and G.min() >= 1 and G.max() <= 20
    assert len({row.tobytes() for row in G[::997]}) == len(G[::997])
    d = (G[:4000] != G[0]).sum(1)
    # rows 0..3999: site 0 fixed, site 1 over its first 10 residues, sites 2 and 3 over all 20
    assert np.array_equal(np.bincount(d, minlength=4), [1, 9 + 19 + 19, 9 * 19 * 2 + 19 * 19, 9 * 19 * 19])
    M = bench.make_tokens(5000, 64, "mutational")
    Q = bench.make_tokens(300, 64, "mutational", seed=1)
    wt = np.array([np.bincount(M[:, c]).argmax() for c in range(64)])
    assert ((Q != wt).sum(1) <= 8).all() and ((M != wt).sum(1) <= 8).all() and not np.array_equal(M[:300], Q)
    assert np.array_equal(bench.make_tokens(100, 8, "uniform"), np.random.default_rng(0).integers(1, 21, size=(100, 8), dtype=np.uint8))
    n, boot, words = 300_000, 8192, 8
    it = sym_plan(n, 5, words, boot, 0, 1, 0, grid=296)
    tile = 512 // words
    planned = 0
    for rb, t0, t1, _, _ in it.astype(np.int64):
        rows = min(n, rb * 256 + 256) - rb * 256
        planned += int(rows * (min(n, t1 * tile) - t0 * tile))
    exact = bench.triangle_pairs(n, boot, (0, n))
    # the planner deals whole tiles: a block's first tile may start before its diagonal / the bootstrap edge
    assert exact <= planned <= exact + (-(-n // 256)) * 256 * tile
