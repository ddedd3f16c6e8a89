#!/bin/bash
# SATD 8x8 batch (the HBM-streaming roofline kernel): plain 128-bit-load kernel vs the persistent TMA/mbarrier variant.
# CUDA-event timings first (never under a profiler), then one ncu pass per variant for DRAM / issue utilisation.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/satd_tma.log
: > $OUT
timeout 120 python tools/time_satd.py >> $OUT 2>&1
KVZ_CUDA_SATD_TMA=1 timeout 120 python tools/time_satd.py >> $OUT 2>&1
M=gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 150 ncu --metrics $M --clock-control none -k regex:satd_nxn_kernel -s 3 -c 1 --csv --log-file gpurun_out/satd_plain_ncu.csv python tools/time_satd.py > /dev/null 2>&1
KVZ_CUDA_SATD_TMA=1 timeout 150 ncu --metrics $M --clock-control none -k regex:satd8_tma_kernel -s 3 -c 1 --csv --log-file gpurun_out/satd_tma_ncu.csv python tools/time_satd.py > /dev/null 2>&1
cat $OUT
