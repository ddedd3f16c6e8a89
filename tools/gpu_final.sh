#!/bin/bash
# Round-end style validation on a GPU box: whole GPU suite, smoke, both bench workloads, reference arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/test_gpu_full.log 2>&1
tail -3 gpurun_out/test_gpu_full.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -2 gpurun_out/bench_default.err
timeout 600 python bench.py --workload 1080p > gpurun_out/bench_1080p.json 2> gpurun_out/bench_1080p.err
tail -2 gpurun_out/bench_1080p.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_default_ref.json 2> gpurun_out/bench_default_ref.err
cat gpurun_out/bench_default.json gpurun_out/bench_1080p.json gpurun_out/bench_default_ref.json | cut -c1-400
