#!/bin/bash
# Round-2 ncu captures (one gpurun call; every command is run once without ncu by the wrapper first).
# Outputs land in gpurun_out/; tools/ncu_extract.py condenses the .ncu-rep files into profiles/*.csv.
set -x
OUT=gpurun_out
NCU="ncu --clock-control none"
# 1. the probe kernels behind the roofline denominators
$NCU --set full -k regex:"int_peak|i8_mma_peak" -c 16 -f -o $OUT/r2_probe python tools/probe_peaks.py 32 > $OUT/r2_ncu_probe.log 2>&1
# 2. the symmetric kNN sweep at a size whose ~40 replays stay short
$NCU --set full --import-source on -k regex:sweep_sym -s 1 -c 1 -f -o $OUT/r2_sym_262k python tools/profile_sym.py --n 262144 > $OUT/r2_ncu_sym_262k.log 2>&1
# 3. DRAM traffic + L2 hit rate of the symmetric sweep at the bench size (few metrics: two replays)
$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
  -k regex:sweep_sym -s 1 -c 1 -f -o $OUT/r2_sym_1m_dram python tools/profile_sym.py --n 1000000 > $OUT/r2_ncu_sym_1m.log 2>&1
# 4. the HBM-bound helpers at 1M rows
$NCU --set full -k regex:"pack_bytes|mutant_bool|mutant_bits|knn_keys_widen|knn_lists_finalize" -c 28 -f -o $OUT/r2_helpers python tools/profile_helpers.py > $OUT/r2_ncu_helpers.log 2>&1
# 5. the tcgen05 Minkowski kernel, k=1 (streaming part) at 262144 rows
$NCU --set full --import-source on -k regex:mink_gemm -s 1 -c 1 -f -o $OUT/r2_gemm_k1 python tools/check_gemm.py 262144 1 > $OUT/r2_ncu_gemm.log 2>&1
# 6. the symmetric epsilon sweep on the GB1 library (W=2)
$NCU --set full -k regex:sweep_sym -s 1 -c 1 -f -o $OUT/r2_sym_eps_gb1 python tools/sym_variants.py --dist gb1 --eps 1 --band 24 --pair 0 > $OUT/r2_ncu_eps_gb1.log 2>&1
# 7. launch list of the bench command (shares of the step)
$NCU --metrics gpu__time_duration.sum -c 2000 --csv --log-file $OUT/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-queries > $OUT/r2_ncu_bench.log 2>&1
ls -la $OUT/*.ncu-rep
