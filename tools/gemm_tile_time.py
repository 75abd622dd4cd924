import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from prograph_b200.engine import get_engine
from bench import make_tokens
eng = get_engine()
M = make_tokens(1_000_000, 256, "mutational")
Q = make_tokens(8192, 256, "mutational")[::-1].copy()
gm = eng.gemm_pack(M, max_token=31)
gq = eng.gemm_pack(Q, max_token=31, K=gm.K)
for kind, name, b in ((1, "float32", 4), (0, "fp16", 2)):
    for rows in (8192, 1024):
        q = gq if rows == 8192 else eng.gemm_pack(Q[:rows], max_token=31, K=gm.K)
        eng.minkowski2_gemm_tile(gm, q, kind); torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = eng.minkowski2_gemm_tile(gm, q, kind); e.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(e)
        print(f"minkowski p=2 tile {rows} x 1000000 {name}: {ms:.2f} ms, {rows*1e6/ms/1e6:.1f} Gpairs/s, out {rows*1e6*b/ms/1e6:.0f} GB/s", flush=True)
        del out
