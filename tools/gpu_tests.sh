#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_ctu_driver.py -x -q -m gpu > gpurun_out/test_ctu_gpu.log 2>&1
tail -5 gpurun_out/test_ctu_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --workload 1080p --steps 2 --warmup 1 > gpurun_out/bench_1080p.json 2> gpurun_out/bench_1080p.err
tail -2 gpurun_out/bench_1080p.err; cat gpurun_out/bench_1080p.json | cut -c1-900
