#!/bin/bash
# GPU check of the integer motion search / candidate kernels: parity tests, per-test lines kept even if the run is cut short
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout ${1:-110} python -m pytest tests/test_me_search.py -m gpu -x -v -p no:cacheprovider 2>&1 | grep -E "PASSED|FAILED|ERROR|passed|failed|error|assert|Error" | tee gpurun_out/me_gpu.log | tail -40
