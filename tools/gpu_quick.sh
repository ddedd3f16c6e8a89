#!/bin/bash
# quick GPU check of the CTU driver: parity tests + device-only throughput at both workloads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/quick.log
: > $OUT
timeout 900 python -m pytest tests/test_ctu_driver.py -x -q -m gpu 2>&1 | tail -3 >> $OUT
one() { echo "## $*" >> $OUT; env $1 timeout 300 python tools/ctu_devbench.py "${@:2}" 2>&1 | grep -v "^  \|phase profile" >> $OUT; }
one X=1 --res 1920x1080 --preset medium --frames 300 --slots 76
one X=1 --res 3840x2160 --preset veryslow --frames 100 --slots 40
cat $OUT
