#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/test_gpu_full.log 2>&1
tail -4 gpurun_out/test_gpu_full.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -2 gpurun_out/bench_default.err
timeout 600 python bench.py --workload 1080p > gpurun_out/bench_1080p.json 2> gpurun_out/bench_1080p.err
tail -2 gpurun_out/bench_1080p.err
cat gpurun_out/bench_default.json gpurun_out/bench_1080p.json | cut -c1-1200
