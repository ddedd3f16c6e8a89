"""Seeded input generators shared by the oracle tests (CPU) and the CUDA parity tests (GPU)."""
import math

import numpy as np


def rng(seed):
    return np.random.default_rng(seed)


from _oracle import aligned, al  # noqa: E402,F401


def rand_pix(r, n, dtype=np.uint8, kind="uniform"):
    mx = 255 if dtype == np.uint8 else 1023
    if kind == "uniform":
        return al(r.integers(0, mx + 1, n).astype(dtype))
    if kind == "extreme":
        return al((r.integers(0, 2, n) * mx).astype(dtype))
    if kind == "smooth":
        base = r.integers(0, mx + 1)
        return al(np.clip(base + r.integers(-6, 7, n).cumsum() // 4, 0, mx).astype(dtype))
    raise ValueError(kind)


KINDS = ("uniform", "extreme", "smooth")


# ---- the fixtures of the reference's own unit tests (tests/satd_tests.c:60-105) ----
def satd_test_bufs(test, log_w):
    w = 1 << log_w
    size = w * w
    i = np.arange(size)
    if test == 0:   # black / white
        return np.zeros(size, np.uint8), np.full(size, 255, np.uint8)
    if test == 1:   # checker; buffer 2 is (buf1 + 1) % 2
        a = (255 * ((((i >> log_w) % 2) + (i % 2)) % 2)).astype(np.uint8)
        b = ((a.astype(np.int32) + 1) % 2).astype(np.uint8)
        return a, b
    col, row = i % w, i // w
    r = np.sqrt(row * row + col * col).astype(np.int64)
    a = (255 // (r + 1)).astype(np.uint8)
    return a, (255 - 255 // (r + 1)).astype(np.uint8)


SATD_GOLDEN = {0: [2040, 4080, 16320, 65280, 261120],      # tests/satd_tests.c:122
               1: [2040, 4080, 16320, 65280, 261120],      # tests/satd_tests.c:140
               2: [3140, 9004, 20481, 67262, 258672]}      # tests/satd_tests.c:159


# tests/intra_sad_tests.c:60-107
def intra_sad_bufs(test, log_w):
    w = 1 << log_w
    if test == 0:
        return np.zeros(w * w, np.uint8), np.full(w * w, 255, np.uint8)
    y, x = np.mgrid[0:w, 0:w]
    val = (np.sqrt((3 - x) ** 2 + (1 - y) ** 2) + 0.5 + 1).astype(np.int64)
    return np.clip(val, 0, 255).astype(np.uint8).ravel(), np.full(w * w, 128, np.uint8)


# tests/dct_tests.c:68-90: radial gradient, slope 255/64 = 3 (integer division), centre (64,64)
def dct_test_buf():
    y, x = np.mgrid[0:64, 0:64]
    val = (3 * np.sqrt((64 - x) ** 2 + (64 - y) ** 2) + 0.5).astype(np.int64)
    return np.clip(val, 0, 255).astype(np.int16).ravel()


# tests/coeff_sum_tests.c:40-54
def coeff_sum_case():
    data = (np.arange(4096, dtype=np.int64) * 16 - 32768).astype(np.int16)
    expected = 2048 * (16 + 32768) // 2 + 2048 * 2047 * 16 // 2
    return data, expected


def rand_refs(r, log2w, dtype=np.uint8, kind="uniform"):
    n = 2 * (1 << log2w) + 1
    top = rand_pix(r, n, dtype, kind)
    left = rand_pix(r, n, dtype, kind)
    left[0] = top[0]
    return top, left


def rand_coeffs(r, n, kind):
    return al(_rand_coeffs(r, n, kind))


def _rand_coeffs(r, n, kind):
    if kind == "residual":     # what a DCT sees: 9-bit residuals
        return r.integers(-255, 256, n).astype(np.int16)
    if kind == "full":         # full int16 range incl. extremes
        a = r.integers(-32768, 32768, n).astype(np.int16)
        a[:: max(1, n // 7)] = 32767
        a[1:: max(1, n // 5)] = -32768
        return a
    if kind == "sparse":
        a = np.zeros(n, np.int16)
        idx = r.integers(0, n, max(1, n // 8))
        a[idx] = r.integers(-2000, 2000, idx.size)
        return a
    if kind == "small":
        return r.integers(-40, 41, n).astype(np.int16)
    raise ValueError(kind)
