#!/usr/bin/env python3
"""Headline benchmark: Gpairs/s of the Hamming all-pairs + kNN (k=16) graph build on a
synthetic library of N=1M sequences, L=256 (BASELINE.json configs[3]), at 1/2/4/8 B200.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29500 bench.py --gpus 8 --steps 3 --warmup 3
    python bench.py --impl reference --steps 3 --warmup 1      # CPU arm (oracle port)

A step = one full graph build: pack the token table into bit planes, build the kNN lists of
all N rows with the fused Hamming+top-k sweeps, finalise, and all-gather.  The build is the
symmetric one (prograph_b200/csrc/pg_sweep_sym.cuh): d(i,j) == d(j,i), so after a one-sided
bootstrap pass over the first 8192 columns every unordered pair is evaluated ONCE and offered
to both rows' lists.  The metric still counts all N^2 ordered pairs -- what the reference
evaluates and what `--one-sided` (the previous kernel) evaluates -- while the roofline is
computed on the pair evaluations actually issued (`pairs_evaluated`).  `value` times the
build with the uint8 token table resident in HBM; `e2e` times the public API call
(prograph_b200.build_neighbours) on a pinned HOST token table, i.e. H2D + the same work +
D2H of the neighbour lists.  Ranks own interleaved row blocks of the triangle (total work
fixed -> "strong" scaling) and all-gather their candidate lists; times are CUDA-event times,
max over ranks.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Gpairs/s Hamming all-pairs + kNN graph build, N=1M L=256"
UNIT = "Gpairs/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000, help="sequences (default: the headline 1M)")
    ap.add_argument("--length", type=int, default=256)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--dist", default="uniform", choices=["uniform", "mutational"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--one-sided", action="store_true",
                    help="evaluate all N^2 ordered pairs with the one-sided sweep (no symmetry)")
    return ap.parse_args()


def make_tokens(n, L, kind):
    """U = iid uniform residues (adversarial ties), M = mutational library."""
    rng = np.random.default_rng(0)
    if kind == "uniform":
        return rng.integers(1, 21, size=(n, L), dtype=np.uint8)
    wt = rng.integers(1, 21, size=L, dtype=np.uint8)
    X = np.tile(wt, (n, 1))
    m = rng.integers(1, 9, size=n)
    for j in range(8):
        rows = np.nonzero(m > j)[0]
        pos = rng.integers(0, L, size=len(rows))
        X[rows, pos] = (X[rows, pos] - 1 + rng.integers(1, 20, size=len(rows))) % 20 + 1
    return X


def config_of(args, world):
    return {
        "workload": f"C4-{'U' if args.dist == 'uniform' else 'M'}: synthetic {args.n} sequences, L={args.length}, "
                    f"Hamming kNN k={args.k}, N^2 ordered pairs counted, "
                    + ("one-sided sweep, rows" if args.one_sided else "symmetric sweep (each unordered pair evaluated "
                       "once), bands of the triangle") + f" sharded over {world} GPU(s)",
        "n_sequences": args.n, "seq_len": args.length, "k": args.k, "distribution": args.dist,
        "parallelism": (f"row-block x{world}" if args.one_sided else f"triangle bands x{world}"),
        "l2": "inputs exceed L2 (126 MB): 160 MB packed table + 136 MB of per-row lists + 8 MB of filter words are "
              "re-read every step",
    }


emit = None      # set by main(): writes the one JSON line to the real stdout


# ----------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm restated with torch CPU ops (oracle port)
# ----------------------------------------------------------------------------------------
def cpu_reference_batches(tokens_u8, k, n_batches, budget_s=None):
    """hamming.py:34 + prograph.py:726,756-762 on the host: fp16 staging, batches of 8 query
    rows against all N, full sort per row.  Returns (pairs, seconds, batches done)."""
    import torch
    from oracle import prograph_oracle as O
    X = torch.from_numpy(tokens_u8).to(torch.float16)
    n = X.shape[0]
    t0 = time.perf_counter()
    done = 0
    for b in range(n_batches):
        batch = X[(8 * b) % n:(8 * b) % n + 8]
        O.reference_knn_batch_torch(X, batch, k)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done * 8 * n, dt, done


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    tokens = make_tokens(args.n, args.length, args.dist)
    per_step = 2            # batches of 8 query rows per step: a bounded sample of the build
    for _ in range(args.warmup):
        cpu_reference_batches(tokens, args.k, 1)
    t0 = time.perf_counter()
    pairs = 0
    for _ in range(args.steps):
        p, _, _ = cpu_reference_batches(tokens, args.k, per_step)
        pairs += p
    dt = time.perf_counter() - t0
    val = pairs / dt / 1e9
    sample = f"{per_step * 8} query rows x all {args.n} columns per step (reference batches of 8), torch CPU fp16"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / max(1, args.steps) * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "fp16 compare -> int64", "data": "synthetic",
        "config": config_of(args, args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def ncu_traffic_bytes(n, world, capture):
    """dram__bytes_read.sum + dram__bytes_write.sum of the sweep kernel, per launch, from the
    committed `ncu --set full` capture of this exact configuration (profiles/); None otherwise."""
    if n != 1_000_000 or world != 1:
        return None
    path = os.path.join(ROOT, "profiles", capture)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total = 0.0
    try:
        for line in open(path):
            parts = [p.strip().strip('"') for p in line.split(",")]
            if len(parts) == 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                total += float(parts[2]) * scale.get(parts[1], 1.0)
    except OSError:
        return None
    return total or None


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.one_sided:
        os.environ["PG_KNN_SYM"] = "0"
    from prograph_b200 import build_neighbours
    from prograph_b200 import graph, shard
    from prograph_b200.engine import get_engine
    eng = get_engine()
    n, L, k = args.n, args.length, args.k

    tokens = make_tokens(n, L, args.dist)
    host = torch.from_numpy(tokens).pin_memory()
    dev = host.to(eng.device)
    row0, rows = shard.row_range(n, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        tab = eng.pack(dev)
        return graph.hamming_knn_graph(eng, tab, k, False, rank, world, None)

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=eng.device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.launch_count(reset=True)
    eng.time_sweeps(True)
    eng.sweep_time(reset=True)
    ms_total = timed(step_resident, args.steps)
    sweep_list = eng.sweep_times(reset=True)
    eng.time_sweeps(False)
    launches = eng.launch_count(reset=True)
    clocks = sampler.stop() if rank == 0 else None
    pairs = float(n) * float(n)
    value = pairs * args.steps / (ms_total * 1e-3) / 1e9

    # ---- end to end through the public API, host buffers in, host arrays out -------------
    e2e = None
    if not args.no_e2e:
        def step_e2e():
            return build_neighbours(host, k=k)
        step_e2e()
        e2e_steps = max(1, min(args.steps, 3))
        e2e_ms = timed(step_e2e, e2e_steps)
        kk = min(k, n - 1)
        e2e = {"value": pairs * e2e_steps / (e2e_ms * 1e-3) / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(host.numel()) * world, "d2h_bytes_per_step": int(n * kk * 16) * world,
               "ms_per_step": e2e_ms / e2e_steps,
               "api": "prograph_b200.build_neighbours(pinned uint8 tokens, k=16) -> numpy idx/weights on every rank"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel, on the pair evaluations it actually issued ---------
    words = eng.packed_words(L)
    lane_ops_per_pair = 7 * words                       # 5 LOP3 + 1 POPC + 1 IADD per 32 residues
    symmetric = (not args.one_sided) and len(sweep_list) == 2 * args.steps
    boot = graph.sym_boot_rows(n) if symmetric else 0
    if symmetric:
        # rank 0's share: bootstrap rectangle + its piece of the triangle (one GPU: everything;
        # several: the band of stream rows the library's planner gives rank 0)
        boot_pairs = float(shard.row_range(n, 0, world)[1]) * boot
        band = (0, n) if world == 1 else eng.sym_band(n, words, boot, 0, world)
        tri_pairs = 0.0
        for rb in range(0, -(-n // 256)):
            a, b = rb * 256, min(n, rb * 256 + 256)
            lo = max(boot if a < boot else a, band[0])
            tri_pairs += float(b - a) * max(0, band[1] - lo)
        kernel_ms = sum(sweep_list[1::2]) / args.steps
        boot_ms = sum(sweep_list[0::2]) / args.steps
        pairs_per_launch, kernel_name = tri_pairs, "pg::sweep_sym_kernel<5,8>"
        capture = "r1_ncu_sym.csv"
    else:
        boot_pairs, boot_ms = 0.0, 0.0
        kernel_ms = sum(sweep_list) / max(1, len(sweep_list))
        pairs_per_launch, kernel_name = float(rows) * float(n), "pg::sweep_kernel<5,8,KNN>"
        capture = "r1_ncu_sweep_r1d.csv"
    peak_ops, _ = eng.int_peak(mix=0, iters=2048)
    lop_ops, _ = eng.int_peak(mix=1, iters=2048)
    popc_ops, _ = eng.int_peak(mix=2, iters=2048)
    roofline = None
    if sweep_list:
        achieved = lane_ops_per_pair * pairs_per_launch / (kernel_ms * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        packed_bytes = float(n) * 5 * words * 4
        hbm_algo = packed_bytes + float(n) * (k + 1) * 8     # table read once + lists written once
        roofline = {
            "bound": "int-alu", "kernel": kernel_name, "achieved": achieved / 1e12,
            "peak": peak_ops / 1e12, "unit": "Tlane-op/s", "frac": achieved / peak_ops,
            "peak_source": "measured live: pg_measure_int_peak (register-only kernel, 5 LOP3 : 1 POPC : 1 IMAD like the sweep)",
            "lane_ops_per_pair": lane_ops_per_pair, "pairs_evaluated_per_launch": pairs_per_launch,
            "ordered_pairs_counted_per_step": float(n) * float(n),
            "kernel_ms": kernel_ms, "kernel_launches": args.steps,
            "kernel_share_of_step": kernel_ms * args.steps / ms_total,
            "gpairs_evaluated_per_s_kernel": pairs_per_launch / (kernel_ms * 1e-3) / 1e9,
            "bootstrap": ({"rows": boot, "pairs_evaluated": boot_pairs, "kernel": "pg::sweep_kernel<5,8,KNN>",
                           "kernel_ms": boot_ms} if symmetric else None),
            "lop3_peak_tlops": lop_ops / 1e12, "popc_peak_tlops": popc_ops / 1e12,
            "traffic": ncu_traffic_bytes(n, world, capture),
            "hbm": {"algorithmic_bytes_per_launch": hbm_algo, "achieved_gbs": hbm_algo / (kernel_ms * 1e-3) / 1e9,
                    "peak_gbs": peaks.get("hbm_gbs", 6650.0),
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s",
                    "note": "compute-bound kernel: the HBM roofline is not the binding one"},
        }

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = len(os.sched_getaffinity(0))
        torch.set_num_threads(cores)
        cpu_reference_batches(tokens, k, 1)
        p, dt, done = cpu_reference_batches(tokens, k, 16, budget_s=20.0)
        # second, best-effort CPU number: uint8 compare + partial selection on all
        # cores, so that the reported baseline is not handicapped by fp16-on-CPU and the full sort
        best = None
        try:
            from oracle import prograph_oracle as O
            O.knn_batch_uint8(tokens, tokens[:8], k, threads=cores)
            t0, nb = time.perf_counter(), 0
            while nb < 16 and time.perf_counter() - t0 < 8.0:
                O.knn_batch_uint8(tokens, tokens[8 * nb:8 * nb + 8], k, threads=cores)
                nb += 1
            bt = time.perf_counter() - t0
            best = {"value": nb * 8 * n / bt / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{nb * 8} query rows x all {n} columns (numpy uint8 compare in column chunks on "
                              f"{cores} threads, partial selection of the k+1 smallest (distance, index) keys)"}
        except Exception as exc:            # a reported extra, never a reason to lose the bench line
            best = {"error": repr(exc)}
        # third: the C restatement (oracle/hamming_knn_cpu.c: bit planes + popcount + sorted k+1 lists on
        # all cores), timed in a child process so that nothing it does can cost the bench line
        best_c = None
        try:
            out = subprocess.run([sys.executable, "-m", "oracle.c_oracle", "--n", str(n), "--length", str(L), "--k", str(k),
                                  "--rows", "512", "--threads", str(cores), "--dist", args.dist], cwd=ROOT,
                                 capture_output=True, text=True, timeout=180)
            lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
            best_c = json.loads(lines[-1]) if lines else {"error": (out.stderr or "no output")[-300:]}
        except Exception as exc:
            best_c = {"error": repr(exc)}
        cpu = {"value": p / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port", "best_effort": best,
               "best_effort_c": best_c,
               "sample": f"{done * 8} query rows x all {n} columns (reference batches of 8, fp16 compare + full "
                         f"sort per row, torch CPU with {cores} threads); extrapolated full build "
                         f"{pairs / (p / dt) / 3600:.1f} h"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8 tokens -> 5 bit planes, int32 popcount distances", "data": "synthetic",
        "config": config_of(args, world), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        "roofline": roofline, "cpu_baseline": cpu,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    # stdout must carry exactly one JSON line: park the real stdout, let everything libraries
    # write to fd 1 (NCCL prints its version banner there) go to stderr, and hand the JSON line
    # to the parked descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global emit
    def emit(line):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
