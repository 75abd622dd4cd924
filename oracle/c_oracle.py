"""ctypes access to the C restatement of the Hamming kNN hot path (oracle/hamming_knn_cpu.c).
TEST / BASELINE INFRASTRUCTURE ONLY, like the rest of oracle/: tests/ pins it against the numpy
oracle, bench.py times it (in a child process) as the best-effort CPU baseline.

    python -m oracle.c_oracle --n 1000000 --rows 256 --threads 16      # prints one JSON line
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libpg_oracle_c.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: run `make -C oracle` (done by __graft_entry__.build())")
        lib = C.CDLL(LIB_PATH)
        lib.pgo_pack.restype = C.c_int
        lib.pgo_pack.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        lib.pgo_hamming_knn.restype = C.c_int
        lib.pgo_hamming_knn.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p]
        lib.pgo_hamming_knn_rows.restype = C.c_int
        lib.pgo_hamming_knn_rows.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                             C.c_void_p, C.c_void_p]
        lib.pgo_hamming_rows.restype = C.c_int
        lib.pgo_hamming_rows.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        _lib = lib
    return _lib


def pack(tokens_u8, use_c=False):
    """(n, L) uint8 tokens 0..31 -> (n, 5, ceil(L/64)) uint64 bit planes: bit l%64 of word l//64 of
    plane p = bit p of residue l.  numpy by default (fast), `use_c` runs the C routine pgo_pack."""
    t = np.ascontiguousarray(tokens_u8, dtype=np.uint8)
    n, L = t.shape
    W = (L + 63) // 64
    if use_c:
        planes = np.empty((n, 5, W), dtype=np.uint64)
        if load().pgo_pack(t.ctypes.data, n, L, planes.ctypes.data) != 0:
            raise OverflowError("tokens do not fit in 5 bit planes")
        return planes
    if t.size and int(t.max()) > 31:
        raise OverflowError("tokens do not fit in 5 bit planes")
    if L % 64:
        t = np.concatenate([t, np.zeros((n, W * 64 - L), dtype=np.uint8)], axis=1)
    planes = np.empty((n, 5, W), dtype=np.uint64)
    for p in range(5):
        bits = np.packbits((t >> p) & 1, axis=1, bitorder="little")           # (n, W*8) bytes, little endian words
        planes[:, p, :] = bits.view("<u8")
    return planes


def hamming_knn(planes, L, q0, nq, k, threads=1):
    """Sorted positions 1..k of query rows [q0, q0+nq) against all rows: (idx, d) int64 (nq, k)."""
    n = planes.shape[0]
    idx = np.empty((nq, k), dtype=np.int64)
    d = np.empty((nq, k), dtype=np.int64)
    rc = load().pgo_hamming_knn(planes.ctypes.data, n, L, q0, nq, k, threads, idx.ctypes.data, d.ctypes.data)
    if rc != 0:
        raise ValueError(f"pgo_hamming_knn failed ({rc})")
    return idx, d


def hamming_knn_rows(planes, L, rows, k, threads=1):
    """The same for an arbitrary list of query rows: (idx, d) int64 (len(rows), k)."""
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    idx = np.empty((len(rows), k), dtype=np.int64)
    d = np.empty((len(rows), k), dtype=np.int64)
    rc = load().pgo_hamming_knn_rows(planes.ctypes.data, planes.shape[0], L, rows.ctypes.data, len(rows), k, threads,
                                     idx.ctypes.data, d.ctypes.data)
    if rc != 0:
        raise ValueError(f"pgo_hamming_knn_rows failed ({rc})")
    return idx, d


def hamming_rows(planes, L, rows, threads=1):
    """hamming.py:34 for the given query rows against every row: (len(rows), n) int32."""
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    out = np.empty((len(rows), planes.shape[0]), dtype=np.int32)
    rc = load().pgo_hamming_rows(planes.ctypes.data, planes.shape[0], L, rows.ctypes.data, len(rows), threads,
                                 out.ctypes.data)
    if rc != 0:
        raise ValueError(f"pgo_hamming_rows failed ({rc})")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--length", type=int, default=256)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--rows", type=int, default=256, help="query rows to time")
    ap.add_argument("--threads", type=int, default=len(os.sched_getaffinity(0)))
    ap.add_argument("--dist", default="uniform")
    args = ap.parse_args()
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import make_tokens
    X = make_tokens(args.n, args.length, args.dist)
    t0 = time.perf_counter()
    planes = pack(X)
    t_pack = time.perf_counter() - t0
    hamming_knn(planes, args.length, 0, min(args.threads, args.n), args.k, args.threads)      # warm-up
    t0 = time.perf_counter()
    hamming_knn(planes, args.length, 0, min(args.rows, args.n), args.k, args.threads)
    dt = time.perf_counter() - t0
    pairs = float(min(args.rows, args.n)) * args.n
    print(json.dumps({"value": pairs / dt / 1e9, "unit": "Gpairs/s", "cores": args.threads, "kind": "port",
                      "pack_s": t_pack,
                      "sample": f"{min(args.rows, args.n)} query rows x all {args.n} columns, C restatement "
                                f"(5 bit planes of 64-bit words, popcount, insertion into a sorted k+1 list), "
                                f"{args.threads} POSIX threads"}))


if __name__ == "__main__":
    main()
