"""Frames/s of the frame-level INTER pass (BASELINE config 4 kernels) at 1920x1080, R = 8, QP 27: CUDA (frames resident
in HBM, 4 in flight) and the same pass through the reference's AVX2 strategy pointers on all host threads."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import kvazaar_b200 as kb  # noqa: E402
from test_interpass import moving_pair  # noqa: E402

W, H, QP, R = 1920, 1088 - 8, 27, 8
kb.init(0)
pairs = [moving_pair(W, H, seed=s) for s in range(4)]
dev = [(kb.to_dev(c), kb.to_dev(r)) for c, r in pairs]
passes = [kb.InterPass(W, H, QP, R) for _ in range(4)]
streams = [torch.cuda.Stream() for _ in range(4)]


def step():
    for i in range(8):
        with torch.cuda.stream(streams[i % 4]):
            passes[i % 4].run_dev(*dev[i % 4])


for _ in range(3):
    step()
torch.cuda.synchronize()
n0 = kb.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
main = torch.cuda.current_stream()
e0.record(main)
for s in streams:
    s.wait_stream(main)
K = 20
for _ in range(K):
    step()
for s in streams:
    main.wait_stream(s)
e1.record(main)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
out = {"workload": f"{W}x{H} inter pass, 16x16 PUs, full search R={R}, search_frac FME, MC, inter residual, QP{QP}",
       "cuda_frames_per_s": 8 * K / (ms / 1000), "ms_per_frame": ms / (8 * K), "gpu_launches": kb.launch_count() - n0}
try:
    from _oracle import Ref, ref_inter_pass
    ref = Ref()
    cores = os.cpu_count()
    ref_inter_pass(ref, pairs[0][0], pairs[0][1], W, H, QP, R, passes[0].layout, nthreads=cores)
    t0, n = time.perf_counter(), 0
    while time.perf_counter() - t0 < 10:
        ref_inter_pass(ref, pairs[n % 4][0], pairs[n % 4][1], W, H, QP, R, passes[0].layout, nthreads=cores)
        n += 1
    out["reference_avx2_frames_per_s"] = n / (time.perf_counter() - t0)
    out["cores"] = cores
except Exception as e:  # pragma: no cover
    out["reference"] = f"unavailable: {e}"
print(json.dumps(out))
