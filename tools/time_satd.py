"""Times kvz_cuda_satd_nxn_batch(8) on 4 Mi pairs (512 MiB > L2) with CUDA events; KVZ_CUDA_SATD_TMA=1 picks the TMA variant."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import kvazaar_b200 as kb  # noqa: E402
kb.init(0)
n = 4 * 1024 * 1024
g = torch.Generator(device="cuda").manual_seed(7)
a = torch.randint(0, 256, (n * 64,), dtype=torch.uint8, device="cuda", generator=g)
b = torch.randint(0, 256, (n * 64,), dtype=torch.uint8, device="cuda", generator=g)
for _ in range(3):
    kb.satd_nxn_batch(8, a, b, n)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    o = kb.satd_nxn_batch(8, a, b, n)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("tma" if os.environ.get("KVZ_CUDA_SATD_TMA") else "plain", f"{ms:.4f} ms  {n * 132 / ms / 1e6:.0f} GB/s  frac {n * 132 / ms / 1e6 / 6567.7:.3f}  sum {int(o.to(torch.int64).sum())}")
