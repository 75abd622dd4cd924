#!/usr/bin/env python3
"""One symmetric kNN build (bootstrap sweep + symmetric sweep + finalise) of the bench workload,
twice; run under ncu for the captures in profiles/.

    ncu --set full --clock-control none --import-source on -k regex:sweep_sym -s 1 -c 1 \
        -o gpurun_out/sym python tools/profile_sym.py --n 1000000
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--length", type=int, default=256)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--dist", default="uniform")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--eps", type=int, default=0, help="epsilon graph instead of kNN (0 = kNN)")
    args = ap.parse_args()
    from bench import make_tokens
    from prograph_b200 import graph
    from prograph_b200.engine import get_engine
    eng = get_engine()
    import operator
    tab = eng.pack(make_tokens(args.n, args.length, args.dist))
    lut = graph.distance_lut(tab.words * 32, operator.le, args.eps, False) if args.eps else None
    for _ in range(args.reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if args.eps:
            graph.hamming_eps_graph(eng, tab, lut, False, 0, 1, None)
        else:
            graph.hamming_knn_graph(eng, tab, args.k, False, 0, 1, None)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print(f"n={args.n} build {ms:.2f} ms  {args.n * args.n / ms / 1e6:.1f} Gpairs/s (N^2 ordered pairs)", flush=True)


if __name__ == "__main__":
    main()
