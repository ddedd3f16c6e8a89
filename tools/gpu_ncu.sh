#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# build without the phase profile for the capture? (PROF adds clock reads only; keep what is built)
timeout 900 ncu --set full --import-source on --clock-control none -k regex:ctu_frame_kernel -c 1 -f -o gpurun_out/r02_ctu_frame_1080p_medium \
  python tools/ctu_devbench.py --res 1920x1080 --preset medium --frames 1 --slots 1 > gpurun_out/ncu_run.log 2>&1
tail -5 gpurun_out/ncu_run.log
ls -la gpurun_out/*.ncu-rep
