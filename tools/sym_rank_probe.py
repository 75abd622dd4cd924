#!/usr/bin/env python3
"""What one rank of a G-GPU symmetric build does, timed on a single GPU: its share of the
bootstrap rectangle plus its interleaved row blocks of the triangle (PG_SYM_STATS=1 prints the
slow-path counters).   python tools/sym_rank_probe.py --n 1000000 --world 8 --boot 8192,32768"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--boot", default="8192")
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--dist", default="uniform")
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--ranks", default="0")
    args = ap.parse_args()
    from bench import make_tokens
    from prograph_b200 import shard
    from prograph_b200.engine import get_engine
    eng = get_engine()
    n, k1 = args.n, args.k + 1
    tab = eng.pack(make_tokens(n, 256, args.dist))

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    for boot in [int(b) for b in args.boot.split(",")]:
      for rank in [int(r) for r in args.ranks.split(",")]:
        for rep in range(2):
            # the full bootstrap lists are needed as the seed; time only this rank's share of them
            row0, rows = shard.row_range(n, rank, args.world)
            a = ev()
            eng.hamming_knn_boot(tab, row0, rows, boot, k1)
            b = ev()
            seed = eng.hamming_knn_boot(tab, 0, n, boot, k1)
            c = ev()
            eng.hamming_knn_sym(tab, k1, rank, args.world, lists=seed, boot_rows=boot, mode=args.mode)
            d = ev()
            torch.cuda.synchronize()
        print(f"n={n} world={args.world} rank={rank} mode={args.mode} boot={boot}: boot share {a.elapsed_time(b):.1f} ms, symmetric sweep "
              f"{c.elapsed_time(d):.1f} ms, sum {a.elapsed_time(b) + c.elapsed_time(d):.1f} ms", flush=True)


if __name__ == "__main__":
    main()
