"""Small driver for ncu captures: a few launches of the frame pass and of the batched SATD kernel (no timing here)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402
import kvazaar_b200 as kb  # noqa: E402
from test_framepass import synth_frame  # noqa: E402

W, H = 1920, 1080
kb.init(0)
fp = kb.FramePass(W, H, 27, 0, int(os.environ.get("RDOQ", "1")))
frames = [kb.to_dev(synth_frame(W, H, frame_idx=i)) for i in range(2)]
for i in range(int(os.environ.get("PASSES", "3"))):
    fp.run_dev(frames[i % 2])
torch.cuda.synchronize()
n_pairs = 4 * 1024 * 1024
g = torch.Generator(device="cuda").manual_seed(7)
a = torch.randint(0, 256, (n_pairs * 64,), dtype=torch.uint8, device="cuda", generator=g)
b = torch.randint(0, 256, (n_pairs * 64,), dtype=torch.uint8, device="cuda", generator=g)
for _ in range(3):
    kb.satd_nxn_batch(8, a, b, n_pairs)
torch.cuda.synchronize()
print("done")
