"""Inputs of the integer-motion-search parity tests (tests/test_me_search.py, tools/make_golden_me.py): record layouts of
include/kvz_cuda.h (kvz_cuda_me_*), deterministic pictures / PU lists, and the ctypes plumbing shared by the three
implementations (reference shim, host build of the device code, the CUDA library)."""
import ctypes as C

import numpy as np

from kvazaar_b200.api import ME_MERGE as MERGE, ME_PU as PU, ME_RESULT as RESULT, MeParams as Params

assert MERGE.itemsize == 12 and PU.itemsize == 84 and RESULT.itemsize == 24 and C.sizeof(Params) == 56      # include/kvz_cuda.h

# PU shapes the reference searches: 2Nx2N / 2NxN / Nx2N of CUs 8..64 and the asymmetric (AMP) splits
SHAPES = [(8, 8), (16, 16), (32, 32), (64, 64), (16, 8), (8, 16), (32, 16), (16, 32), (64, 32), (32, 64), (8, 4), (4, 8),
          (16, 4), (16, 12), (4, 16), (12, 16), (32, 8), (32, 24), (8, 32), (24, 32), (64, 16), (64, 48), (16, 64), (48, 64)]

# name -> (width, height, bitdepth, params, seed, pus)
CASES = {
    "hexbs_et_sensitive":   dict(w=208, h=136, bd=8, algo=0, steps=-1, et=2, mvc=0, wpp=0, delay=0, qp=27, seed=1, n=400),
    "hexbs_et_on_wpp_sao":  dict(w=208, h=136, bd=8, algo=0, steps=-1, et=1, mvc=0, wpp=1, delay=10, qp=32, seed=2, n=400),
    "hexbs_et_off_margin":  dict(w=136, h=72, bd=8, algo=0, steps=-1, et=0, mvc=4, wpp=1, delay=8, qp=22, seed=3, n=400),
    "hexbs_steps2_frame":   dict(w=320, h=192, bd=8, algo=0, steps=2, et=0, mvc=1, wpp=0, delay=0, qp=37, seed=4, n=400),
    "hexbs_steps0":         dict(w=128, h=128, bd=8, algo=0, steps=0, et=2, mvc=0, wpp=1, delay=0, qp=27, seed=5, n=300),
    "dia_et_sensitive":     dict(w=208, h=136, bd=8, algo=7, steps=-1, et=2, mvc=0, wpp=0, delay=0, qp=27, seed=6, n=400),
    "dia_et_off_steps3":    dict(w=136, h=72, bd=8, algo=7, steps=3, et=0, mvc=4, wpp=1, delay=10, qp=30, seed=7, n=400),
    "hexbs_noisy":          dict(w=208, h=136, bd=8, algo=0, steps=-1, et=0, mvc=0, wpp=0, delay=0, qp=17, seed=8, n=400, noisy=True),
    "hexbs_10bit":          dict(w=208, h=136, bd=10, algo=0, steps=-1, et=1, mvc=0, wpp=1, delay=10, qp=27, seed=9, n=400),
    "dia_10bit_margin":     dict(w=136, h=72, bd=10, algo=7, steps=-1, et=2, mvc=4, wpp=0, delay=0, qp=32, seed=10, n=300),
}


def lambda_sqrt(qp):
    return float(np.sqrt(0.57 * 2.0 ** ((qp - 12) / 3.0)))


def pictures(w, h, bd, seed, noisy=False):
    """current and reference luma planes: smooth structure + texture, the current picture = the reference moved by a
    spatially varying amount (so the searches travel several steps) + noise"""
    r = np.random.default_rng(1000 + seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = 110 + 60 * np.sin(xx / 9.0 + seed) * np.cos(yy / 7.0) + 35 * np.sin((xx + 2 * yy) / 3.3)
    tex = r.integers(-25, 26, (h, w))
    ref = np.clip(base + tex, 0, 255)
    sx = (3 + 4 * np.sin(yy / 40.0)).astype(int)
    sy = (-2 + 3 * np.cos(xx / 50.0)).astype(int)
    cx = np.clip(xx + sx, 0, w - 1)
    cy = np.clip(yy + sy, 0, h - 1)
    amp = 40 if noisy else 4
    cur = np.clip(ref[cy, cx] + r.integers(-amp, amp + 1, (h, w)), 0, 255)
    if bd == 10:
        ref = ref * 4 + r.integers(0, 4, (h, w))
        cur = cur * 4 + r.integers(0, 4, (h, w))
        return np.ascontiguousarray(cur.astype(np.uint16)), np.ascontiguousarray(ref.astype(np.uint16))
    return np.ascontiguousarray(cur.astype(np.uint8)), np.ascontiguousarray(ref.astype(np.uint8))


def pu_list(w, h, seed, n):
    r = np.random.default_rng(2000 + seed)
    pus = np.zeros(n, PU)
    for i in range(n):
        pw, ph = SHAPES[int(r.integers(0, len(SHAPES)))]
        while pw > w or ph > h:
            pw, ph = SHAPES[int(r.integers(0, len(SHAPES)))]
        # positions on the 4-sample grid; a good share touching the picture borders
        edge = r.integers(0, 6)
        x = int(r.integers(0, (w - pw) // 4 + 1)) * 4
        y = int(r.integers(0, (h - ph) // 4 + 1)) * 4
        if edge == 0:
            x = 0
        elif edge == 1:
            x = (w - pw) // 4 * 4
        elif edge == 2:
            y = 0
        elif edge == 3:
            y = (h - ph) // 4 * 4
        pus[i]["x"], pus[i]["y"], pus[i]["w"], pus[i]["h"] = x, y, pw, ph
        pus[i]["mv_cand"] = r.integers(-40, 41, (2, 2))
        if r.integers(0, 4) == 0:
            pus[i]["mv_cand"][1] = pus[i]["mv_cand"][0]
        kind = r.integers(0, 5)
        if kind == 0:
            pus[i]["start_mv"] = 0
        elif kind == 1:
            pus[i]["start_mv"] = r.integers(-600, 601, 2)          # far away, often outside the picture / not allowed
        else:
            pus[i]["start_mv"] = r.integers(-48, 49, 2)
        nm = int(r.integers(0, 6))
        pus[i]["num_merge"] = nm
        for m in range(nm):
            pus[i]["merge"][m]["dir"] = int(r.integers(1, 4))
            pus[i]["merge"][m]["mv"] = r.integers(-64, 65, (2, 2))
            if r.integers(0, 5) == 0:
                pus[i]["merge"][m]["mv"] = 0
            if r.integers(0, 6) == 0 and kind != 0:
                pus[i]["merge"][m]["mv"][:] = pus[i]["start_mv"]       # the start MV is one of the merge candidates
    return pus


def make_case(name):
    c = CASES[name]
    p = Params()
    p.width, p.height, p.bitdepth = c["w"], c["h"], c["bd"]
    p.ime_algorithm, p.me_max_steps, p.me_early_termination = c["algo"], c["steps"], c["et"]
    p.mv_constraint, p.wpp_owf, p.delay_px = c["mvc"], c["wpp"], c["delay"]
    p.max_ref_lcu_right, p.max_ref_lcu_down = 1, 1                  # encoder.c:193-194
    p.lambda_sqrt = lambda_sqrt(c["qp"])
    cur, ref = pictures(c["w"], c["h"], c["bd"], c["seed"], c.get("noisy", False))
    return p, cur, ref, pu_list(c["w"], c["h"], c["seed"], c["n"])


def grid_case(w, h, bd, size=16, seed=77, qp=27):
    """every size x size PU of a picture (the shape of a frame-level call)"""
    p = Params()
    p.width, p.height, p.bitdepth = w, h, bd
    p.ime_algorithm, p.me_max_steps, p.me_early_termination = 0, -1, 1       # --preset slow: hexbs, early termination on
    p.mv_constraint, p.wpp_owf, p.delay_px = 0, 1, 10
    p.max_ref_lcu_right, p.max_ref_lcu_down = 1, 1
    p.lambda_sqrt = lambda_sqrt(qp)
    cur, ref = pictures(w, h, bd, seed)
    nx, ny = w // size, h // size
    pus = np.zeros(nx * ny, PU)
    r = np.random.default_rng(seed)
    for j in range(ny):
        for i in range(nx):
            u = pus[j * nx + i]
            u["x"], u["y"], u["w"], u["h"] = i * size, j * size, size, size
            u["mv_cand"] = r.integers(-24, 25, (2, 2))
            u["start_mv"] = r.integers(-32, 33, 2)
            u["num_merge"] = 2
            u["merge"][0]["dir"], u["merge"][1]["dir"] = 1, 2
            u["merge"][0]["mv"] = r.integers(-32, 33, (2, 2))
            u["merge"][1]["mv"] = r.integers(-32, 33, (2, 2))
    return p, cur, ref, pus


def run_host_api(lib, p, cur, ref, pus):
    """kvz_cuda_call_me_search of `lib` (host buffers): the CUDA library or the host build of the device code"""
    out = np.zeros(len(pus), RESULT)
    lib.kvz_cuda_call_me_search.argtypes = [C.POINTER(Params), C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    rc = lib.kvz_cuda_call_me_search(C.byref(p), cur.ctypes.data, cur.shape[1], ref.ctypes.data, ref.shape[1], pus.ctypes.data, len(pus), out.ctypes.data)
    assert rc == 0, rc
    return out


def run_reference(ref_shim, p, cur, ref, pus):
    """the unmodified reference's own functions (oracle/ref_me.c)"""
    out = np.zeros(len(pus), RESULT)
    ctx = ref_shim.ctx(27, 0, 0, p.width, p.height)
    f = ref_shim.lib.kvzref_me_search
    f.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    rc = f(ctx, C.byref(p), cur.ctypes.data, cur.shape[1], ref.ctypes.data, ref.shape[1], pus.ctypes.data, len(pus), out.ctypes.data)
    assert rc == 0, rc
    return out


def same(a, b):
    """decisions and costs identical (the diagnostic point count is not part of the reference's result)"""
    return np.array_equal(a["mv"], b["mv"]) and np.array_equal(a["bits"], b["bits"]) and np.array_equal(a["cost"], b["cost"])
