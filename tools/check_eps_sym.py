#!/usr/bin/env python3
"""Symmetric epsilon sweep (pg_hamming_eps_sym + pg_edge_keys_to_csr) against the one-sided
count / fill passes: bit-exact CSR comparison and timing on the C3 / C4 shapes.

    python tools/check_eps_sym.py [--big]
"""
import argparse
import operator
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prograph_b200.engine import get_engine  # noqa: E402
from prograph_b200.graph import distance_lut  # noqa: E402
from tools.bench_configs import gb1_library, timed  # noqa: E402
from tools.check_sym import mutational  # noqa: E402


def sym_csr(eng, tab, lut, similarity=False, world=1, mode=0, capacity=None):
    keys, edges = [], 0
    for r in range(world):
        k, e = eng.hamming_eps_sym(tab, lut, r, world, mode=mode, capacity=capacity)
        keys.append(k)
        edges += e
    return eng.edge_keys_to_csr(torch.cat(keys), tab.rows, tab.words, edges, similarity)


def same(a, b):
    return all(bool(torch.equal(x, y)) for x, y in zip(a, b))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    args = ap.parse_args()
    eng = get_engine()
    rng = np.random.default_rng(3)
    ok = True
    cases = [
        ("mut L=256 eps=3", mutational(rng, 5000, 256), 3, operator.le, 1, 0),
        ("mut L=256 eps=3 world=3 bands", mutational(rng, 5000, 256), 3, operator.le, 3, 1),
        ("mut L=56 eps=2 world=2", mutational(rng, 40000, 56, max_mut=4), 2, operator.le, 2, 0),
        ("uniform L=20 eps=12 (dense)", rng.integers(1, 5, size=(3000, 20), dtype=np.uint8), 12, operator.le, 1, 0),
        ("uniform L=20 d>=18 ", rng.integers(1, 5, size=(3000, 20), dtype=np.uint8), 18, operator.ge, 2, 1),
        ("mut L=300 eps=4", mutational(rng, 3000, 300), 4, operator.lt, 1, 0),
        ("alphabet 200 L=64 eps=5", mutational(rng, 3000, 64, alphabet=200), 5, operator.le, 1, 0),
        ("nothing passes", rng.integers(1, 21, size=(2000, 256), dtype=np.uint8), 1, operator.le, 1, 0),
        ("n=5", mutational(rng, 5, 40, dup=False), 30, operator.le, 1, 0),
    ]
    for name, X, eps, comp, world, mode in cases:
        tab = eng.pack(torch.from_numpy(X))
        n = X.shape[0]
        for similarity in (False, True):
            e = 1 / (1 + eps) if similarity else eps
            lut = distance_lut(tab.words * 32, comp, e, similarity)
            ref = eng.hamming_eps(tab, 0, n, tab, lut, similarity=similarity)
            good = same(ref, sym_csr(eng, tab, lut, similarity, world, mode))
            good &= same(ref, sym_csr(eng, tab, lut, similarity, world, mode, capacity=1024))     # overflow -> resize
            ok &= good
            print(f"{name:32s} n={n:6d} sim={int(similarity)} nnz={int(ref[0][-1]):9d}: {'ok' if good else 'MISMATCH'}",
                  flush=True)
    # timing: C3 (GB1-style 160 000 x 56) and the mutational library
    X = gb1_library()
    tab = eng.pack(X)
    n = len(X)
    for eps in (1, 2):
        lut = distance_lut(tab.words * 32, operator.le, eps, False)
        t_ref, ref = timed(lambda: eng.hamming_eps(tab, 0, n, tab, lut))
        t_sym, got = timed(lambda: sym_csr(eng, tab, lut))
        good = same(ref, got)
        ok &= good
        print(f"C3 eps={eps}: nnz={int(ref[0][-1])} one-sided {t_ref:.2f} ms, symmetric {t_sym:.2f} ms "
              f"({n * n / t_sym / 1e6:.0f} Gpairs/s of N^2)  {'ok' if good else 'MISMATCH'}", flush=True)
    for n in [262144] + ([1_000_000] if args.big else []):
        from bench import make_tokens
        X = make_tokens(n, 256, "mutational")
        tab = eng.pack(X)
        for eps in (1, 2):
            lut = distance_lut(tab.words * 32, operator.le, eps, False)
            t_ref, ref = timed(lambda: eng.hamming_eps(tab, 0, n, tab, lut), reps=1)
            t_sym, got = timed(lambda: sym_csr(eng, tab, lut), reps=1)
            good = same(ref, got)
            ok &= good
            print(f"C4-M n={n} eps={eps}: nnz={int(ref[0][-1])} one-sided {t_ref:.2f} ms, symmetric {t_sym:.2f} ms "
                  f"({n * n / t_sym / 1e6:.0f} Gpairs/s of N^2)  {'ok' if good else 'MISMATCH'}", flush=True)
    print("ALL OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
