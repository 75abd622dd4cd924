#!/usr/bin/env python3
"""One epsilon=1 build and one kNN (k=16) build of the GB1-style 20^4 x 56 library (C3) through the
shipped graph builders, for ncu (-k regex:sweep_sym picks the two symmetric sweeps)."""
import operator
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from bench import make_gb1_library
    from prograph_b200 import graph
    from prograph_b200.engine import get_engine
    eng = get_engine()
    dev = torch.from_numpy(make_gb1_library()).to(eng.device)
    lut = graph.distance_lut(64, operator.le, 1, False)
    ip, _, _ = graph.hamming_eps_graph(eng, eng.pack(dev), lut, False, 0, 1, None)
    idx, _, _ = graph.hamming_knn_graph(eng, eng.pack(dev), 16, False, 0, 1, None)
    torch.cuda.synchronize()
    print("nnz", int(ip[-1]), "knn", tuple(idx.shape))


if __name__ == "__main__":
    main()
