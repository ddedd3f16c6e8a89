cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_10bit.py -x -q -m gpu -k "full_size" > gpurun_out/test_4320p.log 2>&1
tail -3 gpurun_out/test_4320p.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/test_gpu_full.log 2>&1
tail -3 gpurun_out/test_gpu_full.log
