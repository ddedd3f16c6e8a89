#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/quick.log
: > $OUT
run() { echo "## $*" >> $OUT; timeout 300 env "${@:1:1}" python tools/ctu_devbench.py "${@:2}" 2>&1 | grep -v "^  \|phase profile" >> $OUT; }
run X=1 --res 1920x1080 --preset medium --frames 96 --slots 48
run CUDA_DEVICE_MAX_CONNECTIONS=32 --res 1920x1080 --preset medium --frames 96 --slots 48
run CUDA_DEVICE_MAX_CONNECTIONS=32 --res 1920x1080 --preset medium --frames 96 --slots 32
echo "## grid15 conn32 slots32" >> $OUT; KVZ_CUDA_CTU_GRID=15 CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python tools/ctu_devbench.py --res 1920x1080 --preset medium --frames 96 --slots 32 2>&1 | grep -v "^  \|phase profile" >> $OUT
run CUDA_DEVICE_MAX_CONNECTIONS=32 --res 3840x2160 --preset veryslow --frames 32 --slots 32
cat $OUT
