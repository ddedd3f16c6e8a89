#!/bin/bash
# One gpurun call: GPU test suite, bench lines (RDOQ on / off), ncu launch list and one --set full capture.
set -u
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != "1" ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt; fi
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_rdoq1.log 2>&1; tail -1 gpurun_out/bench_rdoq1.log > gpurun_out/bench_rdoq1.json
python bench.py --steps 10 --warmup 3 --rdoq 0 --no-cpu-baseline > gpurun_out/bench_rdoq0.log 2>&1; tail -1 gpurun_out/bench_rdoq0.log > gpurun_out/bench_rdoq0.json
python bench.py --workload 2160p --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2160p.log 2>&1; tail -1 gpurun_out/bench_2160p.log > gpurun_out/bench_2160p.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log > gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_rdoq.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
PASSES=1 ncu --set full --clock-control none -c 80 -o /tmp/prof_rdoq -f python tools/profile_kernels.py > gpurun_out/ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/prof_rdoq.ncu-rep gpurun_out/r01b > gpurun_out/ncu_summary.log 2>&1
ls -la gpurun_out | tail -12
python - <<'PY'
import json
for f in ("bench_rdoq1", "bench_rdoq0", "bench_2160p", "bench_ref"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d.get("value"), d.get("e2e", {}).get("value"), d.get("cpu_baseline", {}).get("value"), d.get("roofline", {}).get("kernel"))
    except Exception as e:
        print(f, "unreadable", e)
PY
