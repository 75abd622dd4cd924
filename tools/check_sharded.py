#!/usr/bin/env python3
"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/check_sharded.py

Every rank builds the kNN and epsilon graphs of the same synthetic library through the
public API (one-sided builds: row blocks sharded over the ranks; symmetric builds: bands of the
triangle, all-to-all of the candidate lists; packed table and results all-gathered over NCCL) and
compares them, bit for bit, with the unsharded build of the same library on its own GPU; the
output="sharded" tables must be the rank's rows of the replicated ones.  `--big N` adds one table of
N rows (e.g. 1000000) whose symmetric multi-rank kNN / eps=1 graphs are compared with the one-sided
sweeps of every rank's own row block (bench.py runs the same check after its timed region)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    big = int(sys.argv[sys.argv.index("--big") + 1]) if "--big" in sys.argv else 0
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    from prograph_b200 import build_neighbours, minkowski
    from prograph_b200.engine import get_engine
    from prograph_b200.graph import distance_lut
    from bench import make_tokens
    import operator
    eng = get_engine()
    ok = True
    from prograph_b200 import graph
    for n, L, sym_min in ((20000, 256, None), (3001, 56, None), (700, 20, None), (70000, 256, None), (20000, 256, 0),
                          (3001, 56, 0)):
        # sym_min = 0 routes small tables through the symmetric build too (interleaved row blocks,
        # all-gather + merge of the candidate lists); 70000 rows take it by default
        default_min, default_eps_min = graph.SYM_MIN_ROWS, graph.SYM_EPS_MIN_ROWS
        if sym_min is not None:
            graph.SYM_MIN_ROWS = graph.SYM_EPS_MIN_ROWS = sym_min
        X = make_tokens(n, L, "mutational")
        knn = build_neighbours(X, k=16)
        eps = build_neighbours(X, eps=2 if n < 50000 else 1)
        graph.SYM_MIN_ROWS, graph.SYM_EPS_MIN_ROWS = default_min, default_eps_min
        tab = eng.pack(X)
        ri, rw = eng.hamming_knn(tab, 0, n, tab, 16, drop=1)
        ip, ei, ew = eng.hamming_eps(tab, 0, n, tab, distance_lut(tab.words * 32, operator.le, 2 if n < 50000 else 1, False))
        good = (np.array_equal(knn.idx, ri.cpu().numpy()) and np.array_equal(knn.w, rw.cpu().numpy())
                and np.array_equal(eps.indptr, ip.cpu().numpy()) and np.array_equal(eps.idx, ei.cpu().numpy())
                and np.array_equal(eps.w, ew.cpu().numpy()))
        # sharded output: this rank's rows of the same graphs, no final all-gather
        graph.SYM_MIN_ROWS = graph.SYM_EPS_MIN_ROWS = sym_min if sym_min is not None else default_min
        if sym_min is None:
            graph.SYM_EPS_MIN_ROWS = default_eps_min
        sk = build_neighbours(X, k=16, output="sharded")
        se = build_neighbours(X, eps=2 if n < 50000 else 1, output="sharded")
        graph.SYM_MIN_ROWS, graph.SYM_EPS_MIN_ROWS = default_min, default_eps_min
        r0, nr = sk.row0, sk.n_rows
        a, b = eps.indptr[se.row0], eps.indptr[se.row0 + se.n_rows]
        good_s = (np.array_equal(sk.idx, knn.idx[r0:r0 + nr]) and np.array_equal(sk.w, knn.w[r0:r0 + nr])
                  and np.array_equal(se.indptr, eps.indptr[se.row0:se.row0 + se.n_rows + 1] - a)
                  and np.array_equal(se.idx, eps.idx[a:b]) and np.array_equal(se.w, eps.w[a:b]))
        counts = torch.tensor([nr, se.n_rows], device="cuda")
        dist.all_reduce(counts)
        good_s &= counts.tolist() == [n, n]                  # the shards tile the graph
        print(f"rank {rank}/{world}: n={n} L={L} sharded == unsharded: {good}, output='sharded' rows: {good_s}", flush=True)
        ok &= good and good_s
    # tile path (minkowski on a small embedding) through the same sharding
    rng = np.random.default_rng(5)
    E = (rng.integers(0, 64, size=(1500, 2)) / 8.0)
    a = build_neighbours(E, k=3, distance=minkowski)
    Eh = torch.from_numpy(E).cuda().to(torch.float16)
    tile = eng.minkowski_tile(Eh, Eh, 0, 1500)
    ri, rw = eng.tile_topk(tile, 3, drop=1)
    good = np.array_equal(a.idx, ri.cpu().numpy()) and np.array_equal(a.w, rw.cpu().numpy())
    print(f"rank {rank}/{world}: minkowski tile path sharded == unsharded: {good}", flush=True)
    ok &= good
    if big:
        from prograph_b200 import shard
        X = make_tokens(big, 256, "mutational")
        r0, nr = shard.row_range(big, rank, world)
        tab = eng.pack(X)
        knn = build_neighbours(X, k=16)
        ri, rw = eng.hamming_knn(tab, r0, nr, tab, 16, drop=1)
        good = np.array_equal(knn.idx[r0:r0 + nr], ri.cpu().numpy()) and np.array_equal(knn.w[r0:r0 + nr], rw.cpu().numpy())
        del knn, ri, rw
        eps = build_neighbours(X, eps=1)
        cnt = eng.hamming_eps_degrees(tab, r0, nr, tab, distance_lut(tab.words * 32, operator.le, 1, False))
        good &= np.array_equal(np.diff(eps.indptr)[r0:r0 + nr], cnt.cpu().numpy())
        print(f"rank {rank}/{world}: n={big} symmetric sharded kNN == one-sided rows, eps degrees == count sweep: {good}",
              flush=True)
        ok &= good
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print("sharded parity ok")


if __name__ == "__main__":
    main()
