"""Pins oracle/kvz_oracle.c (our CPU restatement) before anything trusts it.

Three anchors:
  1. the golden constants of the reference's own greatest suites (satd/sad/coeff_sum);
  2. the UNMODIFIED reference compiled into oracle/_ref (generic AND the selected AVX2
     implementations) on seeded random / extreme inputs;
  3. regenerated tables (DCT matrices, scan orders, chroma QP map) against the
     reference's exported tables.
All CPU; no GPU needed.
"""
import numpy as np
import pytest

import _cases as cs

SIZES = (4, 8, 16, 32, 64)


# --------------------------------------------------------------------- goldens
@pytest.mark.parametrize("test", [0, 1, 2])
@pytest.mark.parametrize("log_w", [2, 3, 4, 5, 6])
def test_satd_goldens(orc, test, log_w):
    a, b = cs.satd_test_bufs(test, log_w)
    n = 1 << log_w
    assert orc.satd_nxn(n, a, b) == cs.SATD_GOLDEN[test][log_w - 2]
    assert orc.satd_nxn(n, b, a) == cs.SATD_GOLDEN[test][log_w - 2]


@pytest.mark.parametrize("log_w", [2, 3, 4, 5, 6])
def test_intra_sad_goldens(orc, log_w):
    n = 1 << log_w
    a, b = cs.intra_sad_bufs(0, log_w)
    assert orc.sad_nxn(n, a, b) == 255 * n * n            # tests/intra_sad_tests.c:136-150
    a, b = cs.intra_sad_bufs(1, log_w)
    assert orc.sad_nxn(n, a, b) == int(np.abs(a.astype(int) - b.astype(int)).sum())


def test_coeff_abs_sum_golden(orc):
    data, expected = cs.coeff_sum_case()
    assert orc.coeff_abs_sum(data) == expected


def test_reg_sad_overflow_golden(orc):
    # tests/sad_tests.c:286-333: 64x64 of 0 vs PIXEL_MAX
    a = np.zeros(64 * 64, np.uint8)
    b = np.full(64 * 64, 255, np.uint8)
    assert orc.reg_sad(a, b, 64, 64, 64, 64) == 64 * 64 * 255


# ---------------------------------------------------------------------- tables
def test_dct_matrices_match_reference_tables(orc, ref):
    for n in (4, 8, 16, 32):
        # probe the oracle's matrix through its transform: DCT of a unit impulse column
        for i in range(n):
            blk = np.zeros((n, n), np.int16)
            blk[:, i] = 1      # every row has a 1 in column i -> first pass gives M[k][i] (before shift)
            # easier: compare full transforms on random data below; here pin the table via ref export
        mat = np.array([[ref.dct_coef(n, k, i) for i in range(n)] for k in range(n)])
        # reconstruct the oracle matrix from impulse responses with a large amplitude
        amp = 256
        for i in range(n):
            row = np.zeros(n * n, np.int16)
            row[i] = amp       # only first row, column i
            out = orc.dct(n, 8, row).reshape(n, n).astype(np.int64)
            # first pass: tmp[k][0] = (M[k][i]*amp + add) >> s1 ; second pass mixes with M[:,0] = 64
            s1 = int(np.log2(n)) - 1
            s2 = int(np.log2(n)) + 6
            tmp = (mat[:, i] * amp + (1 << (s1 - 1) if s1 else 0)) >> s1 if s1 else mat[:, i] * amp
            exp = np.zeros((n, n), np.int64)
            for k2 in range(n):
                exp[k2, :] = (mat[k2, 0] * tmp + (1 << (s2 - 1))) >> s2
            assert np.array_equal(out, exp), (n, i)


def test_scan_tables_match_reference(orc, ref):
    for s in range(3):
        for log2 in range(1, 6):
            assert np.array_equal(orc.scan_table(s, log2), ref.scan_table(s, log2)), (s, log2)


def test_scaled_qp_matches_reference(orc, ref):
    for type_ in (0, 2, 3):
        for qp in range(0, 52):
            for off in (0, 12):
                assert orc.lib.orc_get_scaled_qp(type_, qp, off) == ref.lib.kvzref_get_scaled_qp(type_, qp, off)


def test_reference_registry_has_expected_entries(ref):
    ents = ref.entries()
    types = {t for t, _, _ in ents}
    for t in ("satd_8x8", "dct_32x32", "quant", "angular_pred", "sao_band_ddistortion", "array_checksum",
              "sample_quarterpel_luma", "get_extended_block"):
        assert t in types
    assert any(n == "avx2" for _, n, _ in ents)
    assert ref.selected_name("satd_8x8") in ("avx2", "generic")


# -------------------------------------------------------------- picture group
IMPLS = ("generic", "")   # "" = highest priority registered (AVX2 on this host)


@pytest.mark.parametrize("kind", cs.KINDS)
@pytest.mark.parametrize("n", SIZES)
def test_sad_satd_nxn_vs_reference(orc, ref, n, kind):
    r = cs.rng(100 + n)
    for it in range(6):
        a = cs.rand_pix(r, n * n, kind=kind)
        b = cs.rand_pix(r, n * n, kind=cs.KINDS[(it + 1) % 3])
        for impl in IMPLS:
            assert orc.sad_nxn(n, a, b) == ref.nxn("sad", n, a, b, impl)
            assert orc.satd_nxn(n, a, b) == ref.nxn("satd", n, a, b, impl)
        preds = cs.aligned(2 * 32 * 32 + n * n, np.uint8)    # pred_buffer layout: rows 1024 px apart
        preds[: n * n] = a
        preds[1024: 1024 + n * n] = cs.rand_pix(r, n * n, kind=kind)
        for impl in IMPLS:
            assert np.array_equal(orc.satd_nxn_dual(n, preds, b), ref.nxn_dual("satd", n, preds, b, impl))
            assert np.array_equal(orc.sad_nxn_dual(n, preds, b), ref.nxn_dual("sad", n, preds, b, "generic"))


PU_SHAPES = [(w, h) for w in (4, 8, 12, 16, 24, 32, 48, 64) for h in (4, 8, 12, 16, 24, 32, 48, 64)]


def test_reg_sad_and_any_size_vs_reference(orc, ref):
    r = cs.rng(7)
    for (w, h) in PU_SHAPES:
        s1, s2 = 96, 80
        a = cs.rand_pix(r, s1 * 64)
        b = cs.rand_pix(r, s2 * 64, kind="smooth")
        for impl in IMPLS:
            assert orc.reg_sad(a, b, w, h, s1, s2) == ref.reg_sad(a, b, w, h, s1, s2, impl), (w, h, impl)
            assert orc.satd_any_size(w, h, a, s1, b, s2) == ref.satd_any_size(w, h, a, s1, b, s2, impl), (w, h)


def test_satd_any_size_quad_vs_reference_including_quirk(orc, ref):
    r = cs.rng(8)
    for (w, h) in PU_SHAPES:
        if w < 8 and h < 8:
            continue
        preds = [cs.rand_pix(r, 64 * 64 + 64) for _ in range(4)]
        orig = cs.rand_pix(r, 96 * 64 + 64)
        got = orc.satd_any_size_quad(w, h, preds, 64, orig, 96)
        for impl in IMPLS:
            assert np.array_equal(got, ref.satd_any_size_quad(w, h, preds, 64, orig, 96, impl)), (w, h, impl)
    # the documented quirk: 16x12 differs from the straight any_size result
    preds = [cs.rand_pix(r, 64 * 64) for _ in range(4)]
    orig = cs.rand_pix(r, 64 * 64)
    quad = orc.satd_any_size_quad(16, 12, preds, 64, orig, 64)
    straight = [orc.satd_any_size(16, 12, p, 64, orig, 64) for p in preds]
    assert list(quad) != straight


def test_ssd_ver_hor_sad_var_vs_reference(orc, ref):
    r = cs.rng(9)
    for width in (4, 8, 16, 32, 64):
        a = cs.rand_pix(r, 64 * 64)
        b = cs.rand_pix(r, 64 * 64, kind="extreme")
        for impl in IMPLS:
            assert orc.pixels_calc_ssd(a, b, 64, 64, width) == ref.pixels_calc_ssd(a, b, 64, 64, width, impl)
    for (w, h) in [(8, 8), (16, 4), (12, 16), (64, 64), (24, 32)]:
        pic = cs.rand_pix(r, 100 * 64)
        refp = cs.rand_pix(r, 100 * 64)
        for impl in IMPLS:
            assert orc.ver_sad(pic, refp, w, h, 100) == ref.ver_sad(pic, refp, w, h, 100, impl)
        for (left, right) in [(3, 0), (0, 5), (w - 1, 0), (0, w - 1), (1, 0), (0, 1)]:  # callers pass exactly one non-zero (src/image.c:326-387)
            for impl in IMPLS:
                assert orc.hor_sad(pic, refp, w, h, 100, 100, left, right) == \
                    ref.hor_sad(pic, refp, w, h, 100, 100, left, right, impl), (w, h, left, right, impl)
    buf = cs.rand_pix(r, 4096)
    assert orc.pixel_var(buf) == ref.pixel_var(buf, "generic")


def test_bipred_average_vs_reference(orc, ref):
    r = cs.rng(10)
    for (w, h) in [(8, 8), (16, 8), (32, 32), (64, 64), (8, 16)]:
        px = [dict(y=cs.rand_pix(r, w * h), u=cs.rand_pix(r, w * h // 4), v=cs.rand_pix(r, w * h // 4)) for _ in range(2)]
        im = [dict(y=r.integers(-2000, 18000, w * h).astype(np.int16), u=r.integers(-2000, 18000, w * h // 4).astype(np.int16),
                   v=r.integers(-2000, 18000, w * h // 4).astype(np.int16)) for _ in range(2)]
        for f0 in range(4):
            for f1 in range(4):
                for impl in IMPLS:
                    ry, ru, rv = ref.bipred_average(px[0], px[1], im[0], im[1], 0, 0, w, h, f0, f1, impl)
                    oy = orc.bipred_average_plane(im[0]["y"] if f0 & 1 else px[0]["y"], im[1]["y"] if f1 & 1 else px[1]["y"],
                                                  f0 & 1, f1 & 1, w, h, 64)
                    assert np.array_equal(oy[: h * 64].reshape(h, 64)[:, :w], ry.reshape(64, 64)[:h, :w]), (w, h, f0, f1)
                    ou = orc.bipred_average_plane(im[0]["u"] if f0 & 2 else px[0]["u"], im[1]["u"] if f1 & 2 else px[1]["u"],
                                                  (f0 >> 1) & 1, (f1 >> 1) & 1, w // 2, h // 2, 32)
                    assert np.array_equal(ou[: h // 2 * 32].reshape(h // 2, 32)[:, : w // 2],
                                          ru.reshape(32, 32)[: h // 2, : w // 2])


# ------------------------------------------------------------------ dct group
@pytest.mark.parametrize("n", (4, 8, 16, 32))
def test_transforms_vs_reference(orc, ref, n):
    r = cs.rng(20 + n)
    grad = cs.dct_test_buf()[: n * n]       # the reference suite's own input (tests/dct_tests.c:68-90)
    kinds = ("residual", "full", "sparse", "small")
    inputs = [("grad", grad)] + [(k, cs.rand_coeffs(r, n * n, k)) for k in kinds]
    # The AVX2 transforms use 16-bit saturating/madd arithmetic and only agree with generic C (the
    # bit-exactness gate) on inputs an encoder can produce: 9-bit residuals forward, dequantised
    # coefficients of such residuals inverse.  Out-of-range noise is compared against generic only.
    fwd_ok = {"grad", "residual", "small"}
    for kind, x in inputs:
        for impl in (IMPLS if kind in fwd_ok else ("generic",)):
            assert np.array_equal(orc.dct(n, 8, x), ref.transform(f"dct_{n}x{n}", 8, x, n, impl)), (n, kind, impl)
            if n == 4:
                assert np.array_equal(orc.dst4(8, x), ref.transform("fast_forward_dst_4x4", 8, x, 4, impl))
        inv_inputs = [x] if kind not in fwd_ok else [x, orc.dct(n, 8, x)]
        for y in inv_inputs:
            for impl in (IMPLS if kind in fwd_ok else ("generic",)):
                assert np.array_equal(orc.idct(n, 8, y), ref.transform(f"idct_{n}x{n}", 8, y, n, impl)), (n, kind, impl)
                if n == 4:
                    assert np.array_equal(orc.idst4(8, y), ref.transform("fast_inverse_dst_4x4", 8, y, 4, impl))


# ---------------------------------------------------------------- quant group
@pytest.mark.parametrize("signhide", (0, 1))
@pytest.mark.parametrize("qp", (17, 22, 27, 32, 37, 51))
def test_quant_dequant_vs_reference(orc, ref, qp, signhide):
    r = cs.rng(30 + qp)
    for n in (4, 8, 16, 32):
        for kind in ("residual", "full", "sparse", "small"):
            coef = cs.rand_coeffs(r, n * n, kind)
            if kind == "residual":
                coef = orc.dct(n, 8, coef)
            for type_ in ((0, 2) if n < 32 else (0,)):      # 4:2:0 chroma TUs are at most 16x16
                for scan in (0, 1, 2):
                    for intra in (0, 1):
                        q = orc.quant(qp, coef, n, n, type_, scan, 1, intra, signhide)
                        for impl in IMPLS:
                            assert np.array_equal(q, ref.quant(qp, coef, n, n, type_, scan, 1, intra, signhide, impl)), \
                                (n, kind, type_, scan, intra, impl)
                for dq_type in ((0, 2, 3) if n < 32 else (0,)):
                    dq = orc.dequant(qp, q, n, n, dq_type, 1)
                    for impl in IMPLS:
                        assert np.array_equal(dq, ref.dequant(qp, q, n, n, dq_type, 1, impl))


@pytest.mark.parametrize("qp", (22, 27, 32))
def test_quantize_residual_vs_reference(orc, ref, qp):
    r = cs.rng(40 + qp)
    for n in (4, 8, 16, 32):
        for kind in cs.KINDS:
            src = cs.rand_pix(r, n * 64, kind=kind)
            pred = cs.rand_pix(r, n * 64, kind="smooth")
            for color in ((0, 1, 2) if n < 32 else (0,)):
                for cu_intra in (0, 1):
                    for trskip in ((0, 1) if n == 4 else (0,)):
                        for signhide in (0, 1):
                            o = orc.quantize_residual(qp, n, color, 0, trskip, cu_intra, 64, src, pred, 1, signhide)
                            for impl in IMPLS:
                                g = ref.quantize_residual(qp, n, color, 0, trskip, cu_intra, 64, src, pred, 1, signhide,
                                                          0, impl)
                                assert o[0] == g[0]
                                assert np.array_equal(o[2], g[2]), (n, kind, color, cu_intra, trskip, impl)
                                assert np.array_equal(o[1].reshape(n, 64)[:, :n], g[1].reshape(n, 64)[:, :n])


def test_coeff_helpers_vs_reference(orc, ref):
    r = cs.rng(50)
    for n in (4, 8, 16, 32):
        for kind in ("full", "sparse", "small"):
            c = cs.rand_coeffs(r, n * n, kind)
            for impl in IMPLS:
                assert orc.coeff_abs_sum(c) == ref.coeff_abs_sum(c, impl)
                w = int(r.integers(0, 2 ** 63))
                assert orc.fast_coeff_cost(c, n, w) == ref.fast_coeff_cost(c, n, w, impl)


# ---------------------------------------------------------------- intra group
@pytest.mark.parametrize("log2w", (2, 3, 4, 5))
def test_intra_predictors_vs_reference(orc, ref, log2w):
    r = cs.rng(60 + log2w)
    for kind in cs.KINDS:
        top, left = cs.rand_refs(r, log2w, kind=kind)
        for impl in IMPLS:
            assert np.array_equal(orc.planar(log2w, top, left), ref.planar(log2w, top, left, impl))
            assert np.array_equal(orc.filtered_dc(log2w, top, left), ref.filtered_dc(log2w, top, left, impl))
            for mode in range(2, 35):
                assert np.array_equal(orc.angular(log2w, mode, top, left), ref.angular(log2w, mode, top, left, impl)), \
                    (log2w, mode, impl)
        for mode in range(35):
            for color in (0, 1):
                for fb in (0, 1):
                    assert np.array_equal(orc.intra_predict(log2w, mode, color, top, left, fb),
                                          ref.intra_predict(log2w, mode, color, top, left, fb)), (log2w, mode, color, fb)


def test_intra_build_reference_vs_reference(orc, ref):
    r = cs.rng(70)
    pic_w, pic_h = 200, 136          # not CTU aligned: exercises the right/bottom clamps
    planes = {0: cs.rand_pix(r, pic_w * pic_h), 1: cs.rand_pix(r, pic_w * pic_h // 4)}
    for log2w in (2, 3, 4, 5):
        w = 1 << log2w
        for color in (0, 1):
            step = w << (1 if color else 0)          # luma-coordinate step of a block of this size
            stride = pic_w >> (1 if color else 0)
            for ly in range(0, pic_h - step + 1, step):
                for lx in range(0, pic_w - step + 1, step):
                    o = orc.intra_build_reference(log2w, color, lx, ly, pic_w, pic_h, planes[color], stride)
                    g = ref.intra_build_reference(log2w, color, lx, ly, pic_w, pic_h, planes[color], stride)
                    assert np.array_equal(o[0], g[0]) and np.array_equal(o[1], g[1]), (log2w, color, lx, ly)


# ----------------------------------------------------------------- ipol group
def _padded_src(r, w, h, pad=8, kind="uniform"):
    stride = w + 2 * pad + 5
    arr = cs.rand_pix(r, stride * (h + 2 * pad + 2), kind=kind)
    return arr, stride, pad * stride + pad


def test_sample_interpolation_vs_reference(orc, ref):
    r = cs.rng(80)
    for (w, h) in [(4, 4), (8, 4), (8, 8), (16, 12), (24, 32), (32, 32), (64, 64), (48, 64)]:
        for kind in ("uniform", "extreme"):
            src, stride, org = _padded_src(r, w, h, kind=kind)
            for mvx in range(4):
                for mvy in range(4):
                    for k in ("luma", "luma_hi"):
                        o = orc.sample(k, src, org, stride, w, h, mvx, mvy)
                        for impl in IMPLS:
                            assert np.array_equal(o, ref.sample(k, src, org, stride, w, h, mvx, mvy, impl=impl)), \
                                (k, w, h, mvx, mvy, impl)
            if w <= 32 and h <= 32:
                for mvx in range(8):
                    for mvy in range(8):
                        for k in ("chroma", "chroma_hi"):
                            o = orc.sample(k, src, org, stride, w, h, mvx, mvy)
                            for impl in IMPLS:
                                assert np.array_equal(o, ref.sample(k, src, org, stride, w, h, mvx, mvy, impl=impl)), \
                                    (k, w, h, mvx, mvy, impl)


def _fme_compare(orc, ref, so, sr, w, h, what):
    for k in range(4):
        a = so[0][k * 4096:(k + 1) * 4096].reshape(64, 64)[:h, :w]
        b = sr[0][k * 4096:(k + 1) * 4096].reshape(64, 64)[:h, :w]
        assert np.array_equal(a, b), (what, "filtered", k)


def test_fme_filters_vs_reference(orc, ref):
    r = cs.rng(90)
    for (w, h) in [(8, 8), (16, 16), (16, 8), (32, 32), (64, 64), (8, 16), (24, 32)]:
        for kind in ("uniform", "extreme"):
            src, stride, org = _padded_src(r, w + 1, h + 1, kind=kind)
            for impl in IMPLS:
                for (ox, oy) in [(0, 0), (-1, 0), (1, 0), (0, -1), (0, 1), (-1, -1), (1, 1), (1, -1), (-1, 1)]:
                    so, sr = orc.fme_state(), ref.fme_state()
                    for stage in range(4):
                        orc.filter_fme(stage, src, org, stride, w, h, so, 2, ox, oy)
                        ref.filter_fme(stage, src, org, stride, w, h, sr, 2, ox, oy, impl)
                        _fme_compare(orc, ref, so, sr, w, h, (w, h, kind, impl, ox, oy, stage))


def test_get_extended_block_vs_reference(orc, ref):
    r = cs.rng(95)
    sw, sh, ss = 80, 48, 88
    src = cs.rand_pix(r, ss * sh)
    for (bx, by) in [(8, 8), (-6, 4), (70, 10), (20, -5), (20, 44), (-10, -10), (76, 46), (-40, 8), (100, 60)]:
        for (bw, bh) in [(8, 8), (16, 4), (32, 16)]:
            for pads in [(3, 4, 3, 4, 0), (3, 4, 3, 4, 1), (0, 0, 0, 0, 0), (1, 2, 1, 2, 3)]:
                o = orc.get_extended_block(src, sw, sh, ss, bx, by, bw, bh, *pads)
                g = ref.get_extended_block(src, sw, sh, ss, bx, by, bw, bh, *pads)
                assert o[0] == g[0] and o[2] == g[2], (bx, by, bw, bh, pads)
                if o[0]:
                    assert np.array_equal(o[1], g[1]), (bx, by, bw, bh, pads)


# ------------------------------------------------------------------ sao group
def test_sao_vs_reference(orc, ref):
    r = cs.rng(110)
    for (bw, bh) in [(64, 64), (32, 32), (56, 64), (64, 24), (8, 8), (16, 40)]:
        for kind in cs.KINDS:
            orig = cs.rand_pix(r, bw * bh, kind=kind)
            rec = np.clip(orig.astype(int) + r.integers(-3, 4, bw * bh), 0, 255).astype(np.uint8)
            for eo in range(4):
                o = orc.calc_sao_edge_dir(8, orig, rec, eo, bw, bh)
                offs = r.integers(-7, 8, 5).astype(np.int32)
                offs[0] = 0 if eo % 2 else offs[0]
                for impl in IMPLS:
                    assert np.array_equal(o, ref.calc_sao_edge_dir(orig, rec, eo, bw, bh, impl)), (bw, bh, eo, impl)
                    assert orc.sao_edge_ddistortion(8, orig, rec, bw, bh, eo, offs) == \
                        ref.sao_edge_ddistortion(orig, rec, bw, bh, eo, offs, impl)
            for band_pos in (0, 5, 13, 28, 31):
                bands = r.integers(-7, 8, 4).astype(np.int32)
                for impl in IMPLS:
                    assert orc.sao_band_ddistortion(8, orig, rec, bw, bh, band_pos, bands) == \
                        ref.sao_band_ddistortion(orig, rec, bw, bh, band_pos, bands, impl)
    # reconstruct: strided in/out with a 1-pixel halo available around the block
    stride, new_stride = 70, 66
    for (bw, bh) in [(64, 64), (32, 20), (10, 54), (54, 10), (1, 1), (3, 64)]:
        recbuf = cs.rand_pix(r, stride * (bh + 2) + 2)
        org = stride + 1
        for color in (0, 1, 2):
            offsets = r.integers(-7, 8, 10).astype(np.int32)
            for eo in range(4):
                o = orc.sao_reconstruct_color(8, recbuf, org, 2, eo, [0, 0], offsets, stride, new_stride, bw, bh, color)
                for impl in IMPLS:
                    g = ref.sao_reconstruct_color(recbuf, org, 2, eo, [0, 0], offsets, stride, new_stride, bw, bh, color, impl)
                    assert np.array_equal(o.reshape(bh, new_stride)[:, :bw], g.reshape(bh, new_stride)[:, :bw]), (bw, bh, eo, impl)
            for bp in ([0, 3], [12, 28], [28, 1]):
                o = orc.sao_reconstruct_color(8, recbuf, org, 1, 0, bp, offsets, stride, new_stride, bw, bh, color)
                for impl in IMPLS:
                    g = ref.sao_reconstruct_color(recbuf, org, 1, 0, bp, offsets, stride, new_stride, bw, bh, color, impl)
                    assert np.array_equal(o.reshape(bh, new_stride)[:, :bw], g.reshape(bh, new_stride)[:, :bw])


# ------------------------------------------------------------------ nal group
def test_array_checksum_vs_reference(orc, ref):
    r = cs.rng(120)
    for (w, h) in [(64, 64), (1920, 8), (3, 5), (520, 300), (960, 540)]:
        data = cs.rand_pix(r, w * h)
        o = orc.array_checksum(data, h, w, w)
        for impl in ("generic", "generic4", "generic8"):
            assert np.array_equal(o, ref.array_checksum(data, h, w, w, impl)), (w, h, impl)
