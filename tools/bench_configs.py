#!/usr/bin/env python3
"""Throughput of the other BASELINE.json configurations on one GPU (not bench lines: the
headline bench is bench.py; these numbers go to profiles/ for the record).

  C3  GB1-style 4-site library, 20^4 = 160 000 sequences, L=56: epsilon graphs + kNN
  C4  1 M x 256: kNN on the mutational distribution, epsilon=2 graph
  C5  100 k queries vs the 1 M library: every metric in prograph/distance

    python tools/bench_configs.py [--quick]
"""
import argparse
import itertools
import json
import operator
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, out


def gb1_library():
    wt = np.frombuffer(b"MTYKLILNGKTLKGETTTEAVDAATAEKVFKQYANDNGVDGEWTYDDATKTFTVTE", dtype=np.uint8)
    lut = np.zeros(256, dtype=np.uint8)
    for i, ch in enumerate("ACDEFGHIKLMNPQRSTVWY"):
        lut[ord(ch)] = i + 1
    combos = np.array(list(itertools.product(range(1, 21), repeat=4)), dtype=np.uint8)
    X = np.tile(lut[wt], (len(combos), 1))
    X[:, [38, 39, 40, 53]] = combos
    return X


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default="gpurun_out/configs.jsonl")
    args = ap.parse_args()
    from prograph_b200.engine import get_engine
    from prograph_b200.graph import distance_lut, hamming_knn_graph, hamming_eps_graph
    from prograph_b200 import _lib as L
    from bench import make_tokens
    eng = get_engine()
    rows = []

    def rec(cfg, what, pairs, ms, **extra):
        r = {"config": cfg, "what": what, "pairs": pairs, "ms": round(ms, 3), "gpairs_per_s": round(pairs / ms / 1e6, 2)}
        r.update(extra)
        rows.append(r)
        print(json.dumps(r), flush=True)

    # ---- C3 ---------------------------------------------------------------------------------
    X = gb1_library()
    n = len(X)
    tab = eng.pack(X)
    for eps in (1, 2):
        lut = distance_lut(tab.words * 32, operator.le, eps, False)
        ms, (ip, ix, w) = timed(lambda: eng.hamming_eps(tab, 0, n, tab, lut))
        deg = (ip[1:] - ip[:-1])
        rec("C3", f"hamming eps={eps} graph (count+scan+fill), CSR int64", n * n * 2.0, ms, nnz=int(ip[-1]),
            degree=int(deg[0]), degree_uniform=bool((deg == deg[0]).all()), note="pairs counts both sweeps")
    ms, _ = timed(lambda: eng.hamming_knn(tab, 0, n, tab, 16, drop=1))
    rec("C3", "hamming kNN k=16, one-sided sweep", float(n) * n, ms)
    # the shipped build_graph path: symmetric sweeps where they apply (N^2 ordered pairs counted)
    for eps in (1, 2):
        lut = distance_lut(tab.words * 32, operator.le, eps, False)
        ms, (ip, ix, w) = timed(lambda: hamming_eps_graph(eng, tab, lut, False, 0, 1, None))
        rec("C3", f"hamming eps={eps} graph, shipped path (degree sample + symmetric sweep + sort, or count/fill when dense)",
            float(n) * n, ms, nnz=int(ip[-1]))
    ms, _ = timed(lambda: hamming_knn_graph(eng, tab, 16, False, 0, 1, None))
    rec("C3", "hamming kNN k=16, shipped path (bootstrap + symmetric sweep)", float(n) * n, ms)

    # ---- C4 ---------------------------------------------------------------------------------
    n4 = 262144 if args.quick else 1_000_000
    M = make_tokens(n4, 256, "mutational")
    tabm = eng.pack(M)
    ms, _ = timed(lambda: eng.hamming_knn(tabm, 0, n4, tabm, 16, drop=1), reps=1)
    rec("C4-M", f"hamming kNN k=16, N={n4}, one-sided sweep", float(n4) * n4, ms)
    ms, _ = timed(lambda: hamming_knn_graph(eng, tabm, 16, False, 0, 1, None), reps=1)
    rec("C4-M", f"hamming kNN k=16, N={n4}, shipped path (bootstrap + symmetric sweep)", float(n4) * n4, ms)
    lut = distance_lut(256, operator.le, 1, False)
    ms, (ip, _, _) = timed(lambda: eng.hamming_eps(tabm, 0, n4, tabm, lut), reps=1)
    rec("C4-M", f"hamming eps=1 graph, N={n4}, one-sided count+fill", float(n4) * n4 * 2, ms, nnz=int(ip[-1]),
        note="pairs counts both sweeps")
    ms, (ip, _, _) = timed(lambda: hamming_eps_graph(eng, tabm, lut, False, 0, 1, None), reps=1)
    rec("C4-M", f"hamming eps=1 graph, N={n4}, shipped path (degree sample + symmetric sweep + sort)", float(n4) * n4, ms,
        nnz=int(ip[-1]))
    try:
        eng.hamming_eps(tabm, 0, n4, tabm, distance_lut(256, operator.le, 2, False))
    except MemoryError as e:
        print("C4-M eps=2:", e, flush=True)

    # ---- C5 ---------------------------------------------------------------------------------
    nq = 16384 if args.quick else 100_000
    Q = make_tokens(nq, 256, "mutational")[::-1].copy()
    rng = np.random.default_rng(1)
    pos = rng.integers(0, 256, size=nq)
    Q[np.arange(nq), pos] = (Q[np.arange(nq), pos] % 20) + 1
    tq = eng.pack(Q, planes=tabm.planes, words=tabm.words)
    ms, _ = timed(lambda: eng.hamming_knn(tq, 0, nq, tabm, 1, drop=0), reps=1)
    rec("C5", f"hamming nearest neighbour (argmin,min) {nq} queries x {n4}", float(nq) * n4, ms)
    ms, _ = timed(lambda: eng.hamming_knn(tq, 0, nq, tabm, 1, drop=0, similarity=True), reps=1)
    rec("C5", f"hamming similarity nearest {nq} x {n4}", float(nq) * n4, ms)
    ms, _ = timed(lambda: eng.hamming_eps(tq, 0, nq, tabm, distance_lut(256, operator.le, 2, False, guard=False)), reps=1)
    rec("C5", f"hamming <=2 neighbourhoods {nq} x {n4} (count+fill)", float(nq) * n4 * 2, ms)
    ms, _ = timed(lambda: eng.hamming_tile(tabm, tq, 0, 1024, weight=L.W_I64))
    rec("C5", f"hamming materialised tile 1024 x {n4} int64", 1024.0 * n4, ms, out_gbs=round(1024.0 * n4 * 8 / ms / 1e6, 1))
    ms, _ = timed(lambda: eng.hamming_tile(tabm, tq, 0, 1024, weight=L.W_SIM_F32))
    rec("C5", f"hamming similarity tile 1024 x {n4} float32", 1024.0 * n4, ms, out_gbs=round(1024.0 * n4 * 4 / ms / 1e6, 1))
    # Minkowski p=2 on the integer tokens: tcgen05 int8 contraction
    gm = eng.gemm_pack(M, max_token=31)
    gq = eng.gemm_pack(Q, max_token=31, K=gm.K)
    for kind, name in ((1, "int64 tokens -> float32"), (0, "fp16 chain")):
        ms, _ = timed(lambda: eng.minkowski2_gemm_knn(gm, gq, 1, 0, kind), reps=1)
        rec("C5", f"minkowski p=2 nearest neighbour {nq} x {n4}, {name}, tcgen05", float(nq) * n4, ms)
    ms, _ = timed(lambda: eng.minkowski2_gemm_knn(gm, gq, 16, 0, 1, similarity=True), reps=1)
    rec("C5", f"minkowski p=2 similarity top-16 {nq} x {n4}, tcgen05", float(nq) * n4, ms)
    ms, _ = timed(lambda: eng.minkowski2_gemm_knn(gm, gm, 16, 1, 0), reps=1)
    rec("C4-M", f"minkowski p=2 kNN k=16 graph N={n4}, fp16 chain, tcgen05", float(n4) * n4, ms)
    g1k = eng.gemm_pack(Q[:8192], max_token=31, K=gm.K)
    ms, _ = timed(lambda: eng.minkowski2_gemm_tile(gm, g1k, 1), reps=1)
    rec("C5", f"minkowski p=2 tile 8192 x {n4} float32, tcgen05", 8192.0 * n4, ms, out_gbs=round(8192.0 * n4 * 4 / ms / 1e6, 1))
    del g1k
    Md = torch.from_numpy(M.astype(np.int64)).cuda()
    Qd = torch.from_numpy(Q[:1024].astype(np.int64)).cuda()
    for p, sim in ((2, False), (2, True), (1, False), (3, False)):
        ms, _ = timed(lambda: eng.minkowski_tile(Md, Qd, 0, 256 if p == 3 else 1024, p=p, similarity=sim), reps=1)
        qn = 256 if p == 3 else 1024
        rec("C5", f"minkowski p={p}{' similarity' if sim else ''} tile {qn} x {n4}, int64 tokens -> float32",
            float(qn) * n4, ms)
    Mh, Qh = Md.to(torch.float16), Qd.to(torch.float16)
    ms, tile = timed(lambda: eng.minkowski_tile(Mh, Qh, 0, 1024, p=2), reps=1)
    rec("C5", f"minkowski p=2 tile 1024 x {n4}, fp16 chain", 1024.0 * n4, ms)
    ms, _ = timed(lambda: eng.tile_topk(tile, 16, drop=0), reps=1)
    rec("C5", f"tile top-16 of 1024 x {n4} fp16", 1024.0 * n4, ms)
    ms, _ = timed(lambda: eng.tile_threshold(tile, L.LE, 20.0, guard=1), reps=1)
    rec("C5", f"tile threshold<=20 of 1024 x {n4} fp16 (count+fill)", 1024.0 * n4, ms)
    # masks on the 1M table
    ms, mut = timed(lambda: eng.mutant_bits(tabm, tabm.row(0)))
    rec("masks", f"mutant_bits N={n4}", float(n4), ms, gbs=round(n4 * (160 + 32) / ms / 1e6, 1))
    ms, _ = timed(lambda: eng.flag_indices(eng.select_rows(mut, dist_lut=distance_lut(256, operator.le, 3, False, guard=False))))
    rec("masks", f"select_rows + flag_indices N={n4}", float(n4), ms)
    ms, _ = timed(lambda: eng.mutant_bool(tabm, tabm.row(0)))
    rec("masks", f"mutant_bool (N, L) N={n4}", float(n4), ms, gbs=round(n4 * (160 + 256) / ms / 1e6, 1))
    ms, _ = timed(lambda: eng.pack(M))
    rec("pack", f"pack uint8 tokens N={n4} (incl. H2D of 256 MB pageable)", float(n4), ms)
    Mdev = torch.from_numpy(M).cuda()
    ms, _ = timed(lambda: eng.pack(Mdev))
    rec("pack", f"pack uint8 tokens resident N={n4}", float(n4), ms, gbs=round(n4 * (256 + 160) / ms / 1e6, 1))
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
