"""Deblocking filter, frame level (ref: src/filter.c).

CPU: the oracle restatement (oracle/kvz_oracle.c orc_deblock_frame, two-pass frame form) against the compiled
reference's kvz_filter_deblock_lcu run LCU by LCU; the 20-byte CU record layout against the reference's own
bitfields.  GPU: kvz_cuda_deblock_frame against the oracle and the reference, bit-exact.
"""
import numpy as np
import pytest

from _oracle import make_cu_records, random_cu_grid


def rough_frame(rng, w, h, bitdepth=8):
    """Blocky content: per-8x8 DC levels + small noise, so weak, strong and no-filter decisions all occur."""
    mx = (1 << bitdepth) - 1
    sc = 1 << (bitdepth - 8)

    def plane(pw, ph):
        yy, xx = np.mgrid[0:ph, 0:pw]
        base = 110 + 40 * np.sin(xx / 50.0) + 30 * np.cos(yy / 40.0)
        step = rng.integers(0, 13, ((ph + 7) // 8, (pw + 7) // 8))
        step = np.kron(step, np.ones((8, 8), np.int64))[:ph, :pw]
        noise = rng.integers(-2, 3, (ph, pw)) * (xx > pw // 2)          # left half noise-free: strong filter fires
        out = (base.astype(np.int64) + step + noise) * sc + rng.integers(0, sc, (ph, pw))
        return np.clip(out, 0, mx)
    dt = np.uint8 if bitdepth == 8 else np.uint16
    return plane(w, h).astype(dt).ravel(), plane(w // 2, h // 2).astype(dt).ravel(), plane(w // 2, h // 2).astype(dt).ravel()


def test_cu_record_layout(ref):
    """The 20-byte records equal the reference's cu_info_t memory image (src/cu.h:126-165)."""
    assert ref.lib.kvzref_sizeof_cu_info() == 20
    rng = np.random.default_rng(5)
    for _ in range(200):
        type_ = int(rng.integers(1, 3))
        depth, part, trd = int(rng.integers(0, 4)), int(rng.integers(0, 8)), int(rng.integers(0, 5))
        cbf, qp, mv_dir = int(rng.integers(0, 1 << 15)), int(rng.integers(0, 52)), int(rng.integers(1, 4))
        mv = rng.integers(-3000, 3000, 4).astype(np.int16)
        mref = rng.integers(0, 16, 2).astype(np.uint8)
        want = ref.make_cu_info(type_, depth, part, trd, cbf, qp, mv_dir, mv, mref)
        got = make_cu_records(np.array(type_), np.array(depth), np.array(part), np.array(trd), np.array(cbf), np.array(qp),
                              np.array(mv_dir), mv.reshape(2, 2), mref)
        # bytes 2/3/7/19 and (for intra) the mode bytes are padding or fields the filter never reads
        keep = [0, 1, 4, 5, 6] + ([8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18] if type_ != 1 else [])
        assert np.array_equal(want[keep] & ([0xff] * 4 + [0xff] + ([0xff] * 10 + [0xc0] if type_ != 1 else [])),
                              got[keep] & ([0xff] * 4 + [0xff] + ([0xff] * 10 + [0xc0] if type_ != 1 else [])))


CASES = [
    # w, h, qp, beta, tc, slice_type (0 B, 1 P, 2 I), per_cu_qp, intra_only
    (128, 64, 27, 0, 0, 2, 0, True),
    (192, 136, 32, 0, 0, 2, 0, True),
    (200, 120, 22, 2, -2, 2, 1, True),
    (192, 136, 30, 0, 0, 1, 0, False),
    (256, 72, 37, -3, 3, 0, 0, False),
    (136, 200, 45, 6, 6, 0, 1, False),
    (64, 64, 20, 0, 0, 1, 0, False),
]


@pytest.mark.parametrize("case", CASES)
def test_oracle_deblock_vs_reference(orc, ref, case):
    w, h, qp, beta, tc, slice_type, per_cu_qp, intra_only = case
    rng = np.random.default_rng(hash(case) & 0xffff)
    y, u, v = rough_frame(rng, w, h)
    cus = random_cu_grid(rng, w, h, intra_only=intra_only)
    ref_lx = rng.integers(0, 3, (2, 16)).astype(np.uint8)
    want = ref.deblock_frame(y, u, v, cus, w, h, qp, beta, tc, slice_type, per_cu_qp, ref_lx)
    got = orc.deblock_frame(y, u, v, cus, w, h, qp, beta, tc, int(slice_type == 0), per_cu_qp, ref_lx)
    changed = sum(int(np.count_nonzero(a != b)) for a, b in zip((y, u, v), want))
    assert changed > 0
    if w >= 128:
        d = (np.asarray(want[0]).reshape(h, w) != y.reshape(h, w))
        assert d[:, 2::8].any() or d[:, 5::8].any(), "strong luma filter never fired"      # p2/q2 only change when strong
    for name, a, b in zip("yuv", got, want):
        assert np.array_equal(a, np.asarray(b)), f"plane {name}: {np.count_nonzero(a != np.asarray(b))} samples differ"


def test_oracle_deblock_10bit(orc10, ref10):
    rng = np.random.default_rng(77)
    w, h = 192, 72
    y, u, v = rough_frame(rng, w, h, 10)
    cus = random_cu_grid(rng, w, h)
    want = ref10.deblock_frame(y, u, v, cus, w, h, 30, 0, 0, 1, 0, np.zeros((2, 16), np.uint8))
    got = orc10.deblock_frame(y, u, v, cus, w, h, 30, 0, 0, 0, 0, None)
    for a, b in zip(got, want):
        assert np.array_equal(a, np.asarray(b))


# ---------------------------------------------------------------------------------------------- GPU parity
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES + [(1920, 1080, 27, 0, 0, 2, 0, True), (832, 480, 32, 1, -1, 0, 1, False)])
def test_cuda_deblock_vs_oracle_and_reference(cuda_lib, orc, ref, case):
    from kvazaar_b200 import api
    w, h, qp, beta, tc, slice_type, per_cu_qp, intra_only = case
    rng = np.random.default_rng(hash(case) & 0xffff)
    y, u, v = rough_frame(rng, w, h)
    cus = random_cu_grid(rng, w, h, intra_only=intra_only)
    ref_lx = rng.integers(0, 3, (2, 16)).astype(np.uint8)
    want = orc.deblock_frame(y, u, v, cus, w, h, qp, beta, tc, int(slice_type == 0), per_cu_qp, ref_lx)
    dy, du, dv = api.to_dev(y), api.to_dev(u), api.to_dev(v)
    api.deblock_frame(dy, du, dv, api.to_dev(cus), w, h, qp, beta, tc, int(slice_type == 0), per_cu_qp, ref_lx)
    for name, a, b in zip("yuv", (dy, du, dv), want):
        assert np.array_equal(a.cpu().numpy(), b), f"plane {name} differs from the oracle"
    if w * h <= 832 * 480:
        rw = ref.deblock_frame(y, u, v, cus, w, h, qp, beta, tc, slice_type, per_cu_qp, ref_lx)
        for a, b in zip((dy, du, dv), rw):
            assert np.array_equal(a.cpu().numpy(), np.asarray(b))
    # host-buffer form with a padded stride
    stride = w + 16
    hy = np.zeros((h, stride), np.uint8); hy[:, :w] = y.reshape(h, w)
    hu = np.zeros((h // 2, stride // 2), np.uint8); hu[:, :w // 2] = u.reshape(h // 2, w // 2)
    hv = np.zeros((h // 2, stride // 2), np.uint8); hv[:, :w // 2] = v.reshape(h // 2, w // 2)
    api.call_deblock_frame(hy, hu, hv, stride, np.ascontiguousarray(cus), w, h, qp, beta, tc, int(slice_type == 0), per_cu_qp, ref_lx)
    assert np.array_equal(hy[:, :w].ravel(), want[0]) and np.array_equal(hu[:, :w // 2].ravel(), want[1])
    assert np.array_equal(hv[:, :w // 2].ravel(), want[2])


@pytest.mark.gpu
def test_cuda_deblock_10bit(cuda_lib, orc10):
    from kvazaar_b200 import api
    rng = np.random.default_rng(78)
    w, h = 320, 136
    y, u, v = rough_frame(rng, w, h, 10)
    cus = random_cu_grid(rng, w, h)
    want = orc10.deblock_frame(y, u, v, cus, w, h, 33, 0, 0, 0, 0, None)
    dy, du, dv = api.to_dev(y), api.to_dev(u), api.to_dev(v)
    api.deblock_frame(dy, du, dv, api.to_dev(cus), w, h, 33)
    for a, b in zip((dy, du, dv), want):
        assert np.array_equal(a.cpu().numpy().view(np.uint16), b)
