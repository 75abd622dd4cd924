#!/bin/bash
# Round-2 ncu captures (one gpurun call; every command is run once without ncu by the wrapper first).
# gpurun_out/ may carry 64 MiB back: the multi-launch reports are condensed to CSV on the box
# (tools/ncu_extract.py) and deleted; only the two single-launch reports with source are kept.
set -x
OUT=gpurun_out
NCU="ncu --clock-control none"
csv() { ncu -i $OUT/$1.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_extract.py --all > $OUT/$1.csv; rm -f $OUT/$1.ncu-rep; }
# 1. the probe kernels behind the roofline denominators (4 launches each: take the second of every kind)
$NCU --set full -k regex:"int_peak|i8_mma_peak" -c 16 -f -o $OUT/r2_ncu_probe python tools/probe_peaks.py 32 > $OUT/r2_ncu_probe.log 2>&1
csv r2_ncu_probe
# 2. the symmetric kNN sweep at a size whose ~40 replays stay short
$NCU --set full --import-source on -k regex:sweep_sym -s 1 -c 1 -f -o $OUT/r2_sym_262k python tools/profile_sym.py --n 262144 > $OUT/r2_ncu_sym_262k.log 2>&1
ncu -i $OUT/r2_sym_262k.ncu-rep --page raw --csv | python tools/ncu_extract.py > $OUT/r2_ncu_sym_262k.csv
# 3. DRAM traffic + L2 hit rate of the symmetric sweep at the bench size (few metrics: two replays)
$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
  -k regex:sweep_sym -s 1 -c 1 -f -o $OUT/r2_ncu_sym_1m_dram python tools/profile_sym.py --n 1000000 > $OUT/r2_ncu_sym_1m.log 2>&1
csv r2_ncu_sym_1m_dram
# 4. the HBM-bound helpers at 1M rows
$NCU --set full -k regex:"pack_bytes|mutant_bool|mutant_bits|knn_keys_widen|knn_lists_finalize" -c 28 -f -o $OUT/r2_ncu_helpers python tools/profile_helpers.py > $OUT/r2_ncu_helpers.log 2>&1
csv r2_ncu_helpers
# 5. the tcgen05 Minkowski kernel, k=1 (streaming part) at 262144 rows
$NCU --set full --import-source on -k regex:mink_gemm -s 1 -c 1 -f -o $OUT/r2_gemm_k1 python tools/check_gemm.py 262144 1 > $OUT/r2_ncu_gemm.log 2>&1
ncu -i $OUT/r2_gemm_k1.ncu-rep --page raw --csv | python tools/ncu_extract.py > $OUT/r2_ncu_gemm_k1.csv
# 6. launch list of the bench command (shares of the step)
$NCU --metrics gpu__time_duration.sum -c 2000 --csv --log-file $OUT/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-queries --no-eps > $OUT/r2_ncu_bench.log 2>&1
ls -la $OUT/
du -sh $OUT
