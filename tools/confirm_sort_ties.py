#!/usr/bin/env python3
"""SURVEY.md §8c [confirm on box]: where does torch's CUDA sort -- the one arithmetic step of the
reference's kNN branch whose implementation leaks into results (prograph.py:757-762,
`torch.sort(d)[1][:, 1:k+1]`, unstable by default) -- agree with the stable restatement
(sort by (distance, index)) that defines kNN parity here?  Heavy-tie Hamming rows of N columns.

    python tools/confirm_sort_ties.py            # prints one line per N
"""
import sys

import numpy as np
import torch


def main():
    k = 16
    rng = np.random.default_rng(0)
    print(f"torch {torch.__version__}, device {torch.cuda.get_device_name(0)}")
    for n in (6, 17, 33, 129, 1000, 2048, 4096, 4097, 10000, 100000, 1000000):
        rows = 64 if n <= 100000 else 8
        # distances in a narrow integer band: almost every sorted position is a tie
        d = torch.from_numpy(rng.integers(230, 250, size=(rows, n)).astype(np.int64)).cuda()
        d16 = d.to(torch.float16)                       # the reference sorts the fp16-staged distances' int64 result
        kk = min(k, n - 1)
        stable = torch.sort(d, dim=1, stable=True).indices[:, 1:kk + 1]
        out = []
        for name, t in (("int64", d), ("fp16", d16)):
            default = torch.sort(t, dim=1).indices[:, 1:kk + 1]
            same_idx = bool(torch.equal(default, stable))
            same_val = bool(torch.equal(torch.gather(d, 1, default), torch.gather(d, 1, stable)))
            out.append(f"{name}: indices {'==' if same_idx else '!='} stable, values {'==' if same_val else '!='}")
        print(f"N={n:8d}  " + "; ".join(out), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
