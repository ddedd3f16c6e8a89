#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --workload 2160p --steps 2 --warmup 1 > gpurun_out/bench_2160p.json 2> gpurun_out/bench_2160p.err
timeout 300 python bench.py --impl reference --workload 2160p --steps 2 --warmup 1 > gpurun_out/bench_2160p_ref.json 2> gpurun_out/bench_2160p_ref.err
timeout 300 python bench.py --impl reference --workload 1080p --steps 2 --warmup 1 > gpurun_out/bench_1080p_ref.json 2> gpurun_out/bench_1080p_ref.err
tail -3 gpurun_out/bench_2160p.err; cat gpurun_out/bench_2160p.json gpurun_out/bench_2160p_ref.json gpurun_out/bench_1080p_ref.json | cut -c1-1500
