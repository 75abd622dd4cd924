#!/usr/bin/env python3
"""Symmetric kNN sweep (pg_hamming_knn_sym) against the one-sided fused sweep (pg_hamming_knn):
bit-exact comparison over shapes / alphabets / tie patterns, emulated multi-rank merge, timing.

    python tools/check_sym.py [--big] [--n 262144]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prograph_b200.engine import get_engine  # noqa: E402


def mutational(rng, n, L, alphabet=20, max_mut=8, dup=True):
    wt = rng.integers(1, alphabet + 1, size=L)
    X = np.tile(wt, (n, 1))
    m = rng.integers(1, max_mut + 1, size=n)
    for j in range(max_mut):
        rows = np.nonzero(m > j)[0]
        pos = rng.integers(0, L, size=len(rows))
        X[rows, pos] = (X[rows, pos] - 1 + rng.integers(1, alphabet, size=len(rows))) % alphabet + 1
    if dup and n > 200:
        X[100] = X[99]
        X[n - 1] = X[0]
    return X.astype(np.uint8)


def run_sym(eng, tab, k, drop, world=1, similarity=False, boot=None, mode=0):
    k1 = k + drop
    if boot is None:
        boot = int(os.environ.get("PG_BOOT", "-1"))
    if boot < 0:
        boot = min(8192, tab.rows // 32 // 512 * 512)
    boot = min(boot, tab.rows // 512 * 512)
    seed = eng.hamming_knn_boot(tab, 0, tab.rows, boot, k1) if boot else None
    lists = [eng.hamming_knn_sym(tab, k1, r, world, lists=seed.clone() if boot else None, boot_rows=boot, mode=mode)
             for r in range(world)]
    return eng.knn_lists_finalize(torch.stack(lists), 0, tab.rows, k, drop, similarity)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--n", type=int, default=262144)
    ap.add_argument("--time-only", action="store_true")
    args = ap.parse_args()
    eng = get_engine()
    rng = np.random.default_rng(7)
    ok = True
    cases = [
        ("mut L=256", mutational(rng, 3000, 256), 16, 1, 1),
        ("mut L=256 world=2", mutational(rng, 5000, 256), 16, 1, 2),
        ("mut L=256 world=3", mutational(rng, 2049, 256), 16, 1, 3),
        ("uniform L=256", rng.integers(1, 21, size=(20000, 256), dtype=np.uint8), 16, 1, 1),
        ("uniform L=256 k=31", rng.integers(1, 21, size=(9000, 256), dtype=np.uint8), 31, 1, 1),
        ("uniform L=256 k=1 drop=0", rng.integers(1, 21, size=(9000, 256), dtype=np.uint8), 1, 0, 1),
        ("binary L=20 (ties)", rng.integers(1, 3, size=(30000, 20), dtype=np.uint8), 16, 1, 1),
        ("mut L=56", mutational(rng, 40000, 56, max_mut=4), 16, 1, 1),
        ("mut L=56 world=2", mutational(rng, 40000, 56, max_mut=4), 5, 1, 2),
        ("mut L=100", mutational(rng, 10000, 100), 16, 1, 1),
        ("mut L=300", mutational(rng, 6000, 300), 16, 1, 1),
        ("alphabet 200 L=64", mutational(rng, 6000, 64, alphabet=200), 16, 1, 1),
        ("tiny n=5", mutational(rng, 5, 40, dup=False), 3, 1, 1),
        ("n=257", mutational(rng, 257, 256), 16, 1, 1),
        ("n=513 world=2", mutational(rng, 513, 33), 16, 1, 2),
    ]
    if not args.time_only:
        for name, X, k, drop, world in cases:
            tab = eng.pack(torch.from_numpy(X))
            kk = min(k, X.shape[0] - drop)
            ri, rw = eng.hamming_knn(tab, 0, tab.rows, tab, kk, drop=drop)
            good = True
            for boot in (0, 512, 2048):
                if boot > X.shape[0]:
                    continue
                for mode in (0, 1):
                    si, sw = run_sym(eng, tab, kk, drop, world, boot=boot, mode=mode)
                    good &= bool(torch.equal(ri, si)) and bool(torch.equal(rw, sw))
            ok &= good
            print(f"{name:28s} n={X.shape[0]:6d} planes={tab.planes} words={tab.words:2d} k={kk} world={world}: "
                  f"{'ok' if good else 'MISMATCH'}", flush=True)
            if not good:
                bad = (ri != si).any(dim=1).nonzero().flatten()
                print("   first bad rows", bad[:8].tolist(), "of", bad.numel())
                r = int(bad[0])
                print("   ref", ri[r].tolist(), rw[r].tolist())
                print("   sym", si[r].tolist(), sw[r].tolist())
    sizes = [65536, args.n] + ([1_000_000] if args.big else [])
    for n in sizes:
        for kind in ("uniform", "mutational"):
            X = rng.integers(1, 21, size=(n, 256), dtype=np.uint8) if kind == "uniform" else mutational(rng, n, 256)
            tab = eng.pack(torch.from_numpy(X))
            reps = 1 if n >= 500000 else 3
            t_ref = timed(lambda: eng.hamming_knn(tab, 0, tab.rows, tab, 16, drop=1), reps)
            t_sym = timed(lambda: run_sym(eng, tab, 16, 1), reps)
            ri, rw = eng.hamming_knn(tab, 0, tab.rows, tab, 16, drop=1)
            si, sw = run_sym(eng, tab, 16, 1)
            good = bool(torch.equal(ri, si)) and bool(torch.equal(rw, sw))
            ok &= good
            print(f"time n={n} {kind}: one-sided {t_ref:.2f} ms ({n * n / t_ref / 1e6:.1f} Gpairs/s)  symmetric "
                  f"{t_sym:.2f} ms ({n * n / t_sym / 1e6:.1f} Gpairs/s of N^2)  {'ok' if good else 'MISMATCH'}", flush=True)
    print("ALL OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
