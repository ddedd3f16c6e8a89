#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/quick.log
: > $OUT
one() { echo "## $*" >> $OUT; env KVZ_CUDA_CTU_GRID=$1 timeout 300 python tools/ctu_devbench.py "${@:2}" 2>&1 | grep -v "^  \|phase profile" >> $OUT; }
one 18 --res 3840x2160 --preset veryslow --frames 40 --slots 32
one 12 --res 3840x2160 --preset veryslow --frames 48 --slots 40
one 8 --res 3840x2160 --preset veryslow --frames 64 --slots 56
one 9 --res 1920x1080 --preset medium --frames 128 --slots 32
one 9 --res 1920x1080 --preset medium --frames 160 --slots 52
one 6 --res 1920x1080 --preset medium --frames 200 --slots 76
one 4 --res 1920x1080 --preset medium --frames 240 --slots 110
cat $OUT
