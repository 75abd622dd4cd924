#!/usr/bin/env python3
"""Quick timing of the tcgen05 Minkowski path (after tests/test_gpu_gemm.py is green)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from prograph_b200.engine import get_engine
    from bench import make_tokens
    eng = get_engine()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    drop = 1 if k > 1 else 0
    X = make_tokens(n, 256, "uniform")
    tab = eng.gemm_pack(X, max_token=31)
    for kind in (0, 1):
        eng.minkowski2_gemm_knn(tab, tab, k, drop, kind)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.minkowski2_gemm_knn(tab, tab, k, drop, kind)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print(f"minkowski2 gemm kNN k={k} kind={kind} N={n}: {ms:.2f} ms  {n * n / ms / 1e6:.1f} Gpairs/s  "
              f"{2.0 * 256 * n * n / ms / 1e9:.1f} int8 TOPS", flush=True)
    q = eng.gemm_pack(X[:4096], max_token=31)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.minkowski2_gemm_tile(tab, q, 1)
    a.record()
    eng.minkowski2_gemm_tile(tab, q, 1)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f"minkowski2 gemm tile 4096 x {n} float32: {ms:.2f} ms  {4096.0 * n / ms / 1e6:.1f} Gpairs/s "
          f"({4096.0 * n * 4 / ms / 1e6:.0f} GB/s written)")


if __name__ == "__main__":
    main()
