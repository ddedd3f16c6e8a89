#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/quick.log
: > $OUT
one() { echo "## $*" >> $OUT; env $1 timeout 300 python tools/ctu_devbench.py "${@:2}" 2>&1 | grep -v "^  \|phase profile" >> $OUT; }
one KVZ_CUDA_CTU_LEADER0=1 --res 1920x1080 --preset medium --frames 300 --slots 76
one X=1 --res 1920x1080 --preset medium --frames 300 --slots 76
one KVZ_CUDA_CTU_LEADER0=1 --res 3840x2160 --preset veryslow --frames 100 --slots 40
one X=1 --res 3840x2160 --preset veryslow --frames 100 --slots 40
cat $OUT
