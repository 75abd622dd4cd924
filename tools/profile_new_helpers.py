#!/usr/bin/env python3
"""Launches of the late round-2 helper kernels for ncu: column census / compaction, mutant_bool,
and the capture copy of a dense epsilon graph (GB1 eps=2)."""
import operator
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from bench import make_gb1_library, make_tokens
    from prograph_b200 import graph
    from prograph_b200.engine import get_engine
    eng = get_engine()
    tab = eng.pack(torch.from_numpy(make_tokens(1_000_000, 256, "mutational")).to(eng.device))
    for _ in range(2):
        cols = eng.varying_columns(tab)
        eng.compact_columns(tab, cols[:128])              # 256 -> 128 positions: W 8 -> 4
        eng.mutant_bool(tab, tab.row(0).contiguous())
    g = eng.pack(torch.from_numpy(make_gb1_library()).to(eng.device))
    lut = graph.distance_lut(64, operator.le, 2, False)
    for _ in range(2):
        ip, _, _ = graph.hamming_eps_graph(eng, g, lut, False, 0, 1, None)
    torch.cuda.synchronize()
    print(len(cols), int(ip[-1]))


if __name__ == "__main__":
    main()
