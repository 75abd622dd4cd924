#!/usr/bin/env python3
"""Timing of the element-wise tile kernel (the fallback of every metric that is not Hamming on tokens or
Minkowski p=2 on integer tokens): 4096 queries x N rows, D = 256, per dtype and exponent."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from prograph_b200.engine import get_engine
    eng = get_engine()
    n, m, D = 262144, 4096, 256
    rng = np.random.default_rng(0)
    base = rng.random((n, D)) * 4 - 2
    for name, dt in (("float16", torch.float16), ("float32", torch.float32), ("float64", torch.float64), ("int64", torch.int64)):
        X = torch.from_numpy(base if dt != torch.int64 else (base * 10).astype(np.int64)).to(eng.device).to(dt)
        Y = X[:m].contiguous()
        for p in ((2, 3) if dt != torch.int64 else (2, 3)):
            eng.minkowski_tile(X, Y, 0, m, p=p)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.minkowski_tile(X, Y, 0, m, p=p)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            print(json.dumps({"kernel": "elem_tile_kernel minkowski", "dtype": name, "p": p, "D": D, "pairs": n * m,
                              "ms": round(ms, 2), "gpairs_per_s": round(n * m / ms / 1e6, 1)}), flush=True)
        if dt in (torch.float32, torch.int64):
            eng.hamming_values_tile(X, Y, 0, m)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.hamming_values_tile(X, Y, 0, m)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            print(json.dumps({"kernel": "elem_tile_kernel hamming on values", "dtype": name, "D": D, "pairs": n * m,
                              "ms": round(ms, 2), "gpairs_per_s": round(n * m / ms / 1e6, 1)}), flush=True)
        del X, Y


if __name__ == "__main__":
    main()
