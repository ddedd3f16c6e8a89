#!/bin/bash
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
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ctu_frame_kernel -c 1 -f -o gpurun_out/r02b_ctu_frame_1080p_medium \
  python tools/ctu_devbench.py --res 1920x1080 --preset medium --frames 1 --slots 1 > gpurun_out/ncu_run.log 2>&1
tail -2 gpurun_out/ncu_run.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_devbench.csv python tools/ctu_devbench.py --res 1920x1080 --preset medium --frames 8 --slots 8 > /dev/null 2>&1
cat gpurun_out/bench_default.json gpurun_out/bench_1080p.json | cut -c1-400
