#!/usr/bin/env python3
"""C3 (GB1-style 20^4 x 56 library) through the shipped graph builders with and without the
informative-column compaction (graph.informative_table), plus the cost of looking at the columns of
the 1 M x 256 uniform table, where nothing can be dropped.

    python tools/c3_time.py [--n 1000000]
"""
import argparse
import operator
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--quick", action="store_true", help="informative columns only, no 1 M-row column census")
    args = ap.parse_args()
    from bench import make_gb1_library, make_tokens
    from prograph_b200 import graph
    from prograph_b200.engine import get_engine
    from tools.bench_configs import timed
    eng = get_engine()
    G = make_gb1_library()
    n = len(G)
    dev = torch.from_numpy(G).to(eng.device)
    from prograph_b200 import trace

    def phases_of(fn):
        """Per-phase and per-sweep-launch milliseconds of one more call."""
        eng.time_sweeps(True)
        eng.sweep_times(reset=True)
        trace.enable_timing(True)
        trace.phases()
        fn()
        ph = trace.phases()
        trace.enable_timing(False)
        sw = eng.sweep_times(reset=True)
        eng.time_sweeps(False)
        return {k: round(v, 3) for k, v in ph.items()}, [round(v, 3) for v in sw]

    for min_rows, label in ((10**9, "full rows"), (4096, "informative columns")):
        if args.quick and min_rows != 4096:
            continue
        graph.COMPACT_MIN_ROWS = min_rows
        for eps in (1, 2):
            lut = graph.distance_lut(64, operator.le, eps, False)
            fn = lambda: graph.hamming_eps_graph(eng, eng.pack(dev), lut, False, 0, 1, None)
            ms, (ip, _, _) = timed(fn, reps=3)
            print(f"C3 eps={eps} [{label}]: {ms:.3f} ms  nnz={int(ip[-1])}  {n * n / ms / 1e6:.1f} Gpairs/s", flush=True)
            print("   phases", *phases_of(fn), flush=True)
        fn = lambda: graph.hamming_knn_graph(eng, eng.pack(dev), 16, False, 0, 1, None)
        ms, _ = timed(fn, reps=3)
        print(f"C3 kNN k=16 [{label}]: {ms:.3f} ms  {n * n / ms / 1e6:.1f} Gpairs/s", flush=True)
        print("   phases", *phases_of(fn), flush=True)
    if args.quick:
        return
    U = torch.from_numpy(make_tokens(args.n, 256, "uniform")).to(eng.device)
    tab = eng.pack(U)
    ms, cols = timed(lambda: eng.varying_columns(tab), reps=5)
    print(f"varying_columns on {args.n} x 256 uniform: {ms:.3f} ms ({tab.data.numel() * 4 / ms / 1e6:.0f} GB/s), "
          f"{len(cols)} columns vary", flush=True)
    M = torch.from_numpy(make_tokens(args.n, 256, "mutational")).to(eng.device)
    tabm = eng.pack(M)
    print("mutational:", len(eng.varying_columns(tabm)), "columns vary")


if __name__ == "__main__":
    main()
