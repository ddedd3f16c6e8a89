#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_ctu_driver.py tests/test_abi_and_dropin.py tests/test_framepass.py -x -q -m gpu > gpurun_out/test_gpu_new.log 2>&1
tail -4 gpurun_out/test_gpu_new.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -2 gpurun_out/bench_default.err
timeout 600 python bench.py --workload 1080p > gpurun_out/bench_1080p.json 2> gpurun_out/bench_1080p.err
tail -2 gpurun_out/bench_1080p.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_default_ref.json 2> gpurun_out/bench_default_ref.err
timeout 300 python bench.py --impl reference --workload 1080p > gpurun_out/bench_1080p_ref.json 2> gpurun_out/bench_1080p_ref.err
cat gpurun_out/bench_default.json gpurun_out/bench_1080p.json gpurun_out/bench_default_ref.json gpurun_out/bench_1080p_ref.json | cut -c1-700
