#!/usr/bin/env python
"""Deterministic synthetic I420 clips.
   tools/synth_yuv.py W H FRAMES out.yuv [seed] [--noisy]
default: the clip of SURVEY.md 8(d) (ramp + drifting sinusoid + small noise);
--noisy: strong noise, flat patches and sharp edges (exercises high coefficient levels, band SAO, transform skip)."""
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


def noisy_frame(width, height, seed=5, frame_idx=0):
    r = np.random.default_rng(seed * 1000 + frame_idx)
    y, x = np.mgrid[0:height, 0:width]
    base = 128 + 70 * np.sin(x / 9.0 + frame_idx) * np.cos(y / 7.0) + r.integers(-40, 41, (height, width))
    base[(x // 16 + y // 16) % 3 == 0] = r.integers(0, 256)
    base[((x // 4) % 2 == 0) & ((y // 32) % 2 == 1)] += 60
    cy, cx = np.mgrid[0:height // 2, 0:width // 2]
    u = 128 + 50 * np.sin(cx / 5.0) + r.integers(-20, 21, cx.shape)
    v = 128 + 50 * np.cos(cy / 3.0) + r.integers(-30, 31, cx.shape)
    return np.concatenate([np.clip(p, 0, 255).astype(np.uint8).ravel() for p in (base, u, v)])


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    noisy = "--noisy" in sys.argv
    w, h, n, out = int(args[0]), int(args[1]), int(args[2]), args[3]
    seed = int(args[4]) if len(args) > 4 else (5 if noisy else 1234)
    with open(out, "wb") as f:
        for i in range(n):
            f.write((noisy_frame if noisy else synth_frame)(w, h, seed, i).tobytes())
