#!/usr/bin/env python3
"""The HBM-bound helper kernels of the path at the bench size (1M rows x 256 residues), each launched
three times (ncu: `-k regex:"pack_bytes|mutant_bool|knn_keys_widen|knn_lists_finalize|edge_decode"`),
with CUDA-event timings and the algorithmic GB/s printed for the record.

    python tools/profile_helpers.py [--n 1000000]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--length", type=int, default=256)
    args = ap.parse_args()
    from bench import make_tokens
    from prograph_b200.engine import get_engine
    eng = get_engine()
    n, L, k = args.n, args.length, 16
    X = make_tokens(n, L, "mutational")
    dev = torch.from_numpy(X).to(eng.device)
    tab = eng.pack(dev)
    words, planes = tab.words, tab.planes
    packed_bytes = tab.data.shape[0] * planes * words * 4
    letters = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
    lut = np.zeros(256, dtype=np.uint8)
    lut[letters] = np.arange(1, 21)
    chars = torch.from_numpy(letters[X - 1]).to(eng.device)
    lists = torch.randint(0, 1 << 40, (n, k + 1), dtype=torch.int64, device=eng.device).sort(dim=1).values
    stack = torch.stack([lists, lists + 1, lists + 2, lists + 3])
    out = []

    def rec(name, ms, algo_bytes):
        out.append({"kernel": name, "ms": round(ms, 4), "algorithmic_GB": round(algo_bytes / 1e9, 4),
                    "GBps": round(algo_bytes / ms / 1e6, 1), "frac_of_6559": round(algo_bytes / ms / 1e6 / 6559.4, 3)})

    rec("pack_bytes_kernel<5> (uint8 tokens -> planes)", timed(lambda: eng.pack(dev)), n * L + packed_bytes)
    rec("pack_bytes_kernel<5,chars> (letters -> planes)", timed(lambda: eng.pack_chars(chars, lut)), n * L + packed_bytes)
    rec("mutant_bool_kernel (planes -> (N, L) bool)", timed(lambda: eng.mutant_bool(tab, tab.row(0))), n * L + n * planes * words * 4)
    rec("mutant_bits_kernel (planes -> (N, words) masks)", timed(lambda: eng.mutant_bits(tab, tab.row(0))),
        n * planes * words * 4 + n * words * 4)
    rec("knn_keys_widen_kernel (lists -> idx, w)", timed(lambda: eng.knn_lists_finalize(lists, 0, n, k, 1)),
        n * (k + 1) * 8 + n * k * 16)
    rec("knn_lists_finalize_kernel (4-way merge -> idx, w)", timed(lambda: eng.knn_lists_finalize(stack, 0, n, k, 1)),
        4 * n * (k + 1) * 8 + n * k * 16)
    rec("knn_lists_finalize_kernel (4-way merge -> keys)", timed(lambda: eng.knn_lists_merge(stack, k, 1)),
        4 * n * (k + 1) * 8 + n * k * 8)
    for r in out:
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
