#!/usr/bin/env python3
"""The measured denominators of the rooflines: integer-pipe probes (mix / LOP3 only / POPC only,
pg_measure_int_peak) and the int8 tensor-pipe probe (pg_measure_i8_mma_peak), one JSON line.
Under ncu: `-k regex:"int_peak|i8_mma_peak"` captures the probe kernels themselves."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from prograph_b200.engine import get_engine
    eng = get_engine()
    mix, _ = eng.int_peak(mix=0, iters=2048)
    lop, _ = eng.int_peak(mix=1, iters=2048)
    popc, _ = eng.int_peak(mix=2, iters=2048)
    i8, ms = eng.i8_mma_peak(batches=int(sys.argv[1]) if len(sys.argv) > 1 else 128)
    sms, clk = 148, 1.965e9
    print(json.dumps({
        "mix_tlaneops": mix / 1e12, "lop3_tlops": lop / 1e12, "popc_tlops": popc / 1e12,
        "lop3_lanes_per_clk_sm": lop / sms / clk, "popc_lanes_per_clk_sm": popc / sms / clk,
        "mix_clk_per_word_and_warp": 7 * 32 / (mix / sms / 4 / clk),
        "i8_mma_pops": i8 / 1e15, "i8_probe_ms": ms,
        "onehot_hamming_bound_gpairs_L256": i8 / (2 * 21 * 256) / 1e9,
        "onehot_hamming_bound_gpairs_L56": i8 / (2 * 21 * 56) / 1e9,
        "popcount_mix_bound_gpairs_L256": mix / 56 / 1e9,
    }))


if __name__ == "__main__":
    main()
