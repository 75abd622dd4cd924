#!/usr/bin/env python3
"""kNN (k=16) graph build of a small L=256 table through the shipped path, for PG_SYM_CHUNK sweeps
(minimum chunk of the symmetric sweep's work items, in tiles).

    PG_SYM_CHUNK=64 python tools/chunk_time.py 70000 100000
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from bench import make_tokens
    from prograph_b200 import graph
    from prograph_b200.engine import get_engine
    from tools.bench_configs import timed
    eng = get_engine()
    for n in [int(a) for a in sys.argv[1:]] or [70000]:
        for kind in ("uniform", "mutational"):
            dev = torch.from_numpy(make_tokens(n, 256, kind)).to(eng.device)
            ms, _ = timed(lambda: graph.hamming_knn_graph(eng, eng.pack(dev), 16, False, 0, 1, None), reps=3)
            print(f"n={n} {kind}: {ms:.3f} ms  {n * n / ms / 1e6:.1f} Gpairs/s", flush=True)


if __name__ == "__main__":
    main()
