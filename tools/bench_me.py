#!/usr/bin/env python
"""Device timing of the integer motion search / candidate kernels, next to the reference's own functions.

    python tools/bench_me.py [--res 1920x1080] [--pu 16] [--algo hexbs|dia|tz|full8] [--bitdepth 8] [--iters 20]

Every pu x pu block of a synthetic picture pair is searched (the settings of --preset slow: early termination on, WPP + SAO
limits).  CUDA events on the launching stream around `iters` launches after 3 warm-up launches; the pictures (> L2 for
2160p) stay resident, as they do between the PUs of a picture in the encoder.  The reference arm is oracle/ref_me.c (the
reference's search_inter.c compiled in place) on ONE host thread -- a per-core baseline, not the target.  One JSON line.
NOT YET RUN ON B200 (written after the round's GPU budget was spent).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", default="1920x1080")
    ap.add_argument("--pu", type=int, default=16)
    ap.add_argument("--algo", default="hexbs", choices=["hexbs", "tz", "full8", "full16", "dia"])
    ap.add_argument("--bitdepth", type=int, default=8, choices=[8, 10])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--no-reference", action="store_true")
    a = ap.parse_args()
    w, h = map(int, a.res.split("x"))
    import torch
    import kvazaar_b200 as kb
    from _me_cases import RESULT, grid_case, run_reference, same
    kb.init(0)
    p, cur, rf, pus = grid_case(w, h, a.bitdepth, a.pu)
    p.ime_algorithm = {"hexbs": 0, "tz": 1, "full8": 3, "full16": 4, "dia": 7}[a.algo]
    d_cur, d_ref, d_pus = kb.to_dev(cur), kb.to_dev(rf), kb.to_dev(pus)
    out = None
    for _ in range(3):
        out = kb.me_search_batch(p, d_cur, d_ref, d_pus, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        kb.me_search_batch(p, d_cur, d_ref, d_pus, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    got = out.cpu().numpy().view(RESULT)
    px = 1 if a.bitdepth == 8 else 2
    line = {"kernel": "me_search_kernel", "res": a.res, "pu": a.pu, "algo": a.algo, "bitdepth": a.bitdepth, "pus": int(len(pus)),
            "ms_per_launch": ms, "pus_per_s": len(pus) / ms * 1e3, "points_per_pu": float(got["points"].mean()),
            "sad_pixels_per_s": float(got["points"].sum()) * a.pu * a.pu / ms * 1e3,
            "algorithmic_bytes_per_launch": 2 * w * h * px, "algorithmic_GBps": 2 * w * h * px / ms / 1e6}
    if not a.no_reference:
        from _oracle import Ref
        ref = Ref(a.bitdepth)
        t = time.perf_counter()
        want = run_reference(ref, p, cur, rf, pus)
        dt = time.perf_counter() - t
        line["reference_one_thread_ms"] = dt * 1e3
        line["identical"] = bool(same(got, want))
    print(json.dumps(line))


if __name__ == "__main__":
    main()
