#!/usr/bin/env python
"""Writes the deterministic synthetic I420 clip of SURVEY.md 8(d): tools/synth_yuv.py W H FRAMES out.yuv [seed]"""
import sys
import numpy as np


def synth_frame(width, height, seed=1234, frame_idx=0):
    r = np.random.default_rng(seed + frame_idx)
    y, x = np.mgrid[0:height, 0:width]
    luma = (x + y) * 0.11 + 60 * np.sin((x + 3 * frame_idx) / 37.0) * np.cos(y / 29.0) + 128 + r.integers(-4, 5, (height, width))
    cy, cx = np.mgrid[0:height // 2, 0:width // 2]
    u = 128 + 40 * np.sin(cx / 23.0 + frame_idx * 0.1) + r.integers(-2, 3, cx.shape)
    v = 128 + 40 * np.cos(cy / 19.0) + r.integers(-2, 3, cx.shape)
    return np.concatenate([np.clip(p, 0, 255).astype(np.uint8).ravel() for p in (luma, u, v)])


if __name__ == "__main__":
    w, h, n, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    seed = int(sys.argv[5]) if len(sys.argv) > 5 else 1234
    with open(out, "wb") as f:
        for i in range(n):
            f.write(synth_frame(w, h, seed, i).tobytes())
