"""RDOQ on device (ref: kvz_rdoq, src/rdo.c:661-977) against the compiled reference, bit-exact.

The reference function is called through oracle/ref_shim.c with the same context models, lambda and QP; the plain-C
restatement oracle/kvz_oracle.c orc_rdoq is pinned against it on the CPU.
"""
import ctypes as C

import numpy as np
import pytest


def synth_coeffs(rng, n, count, energy):
    """Transform-coefficient-like blocks: Laplacian magnitudes decaying with frequency, a few outliers."""
    fy, fx = np.mgrid[0:n, 0:n]
    decay = np.exp(-(fx + fy) / (n * rng.uniform(0.08, 0.6, (count, 1, 1))))
    c = rng.laplace(0, 1, (count, n, n)) * decay * energy * rng.uniform(0.05, 2.0, (count, 1, 1))
    c[rng.random(count) < 0.1] = 0                                    # empty blocks
    hot = rng.random((count, n, n)) < 0.002
    c = np.where(hot, rng.integers(-32768, 32768, (count, n, n)), c)
    return np.clip(np.rint(c), -32768, 32767).astype(np.int16)


def test_cabac_ctx_layout(ref):
    """kvz_cuda_cabac_ctx (include/kvz_cuda.h) has the reference's member offsets (src/cabac.h:66-102)."""
    from kvazaar_b200 import api
    assert ref.cabac_ctx_size() == api.CABAC_CTX_BYTES
    sizes = [("sao_merge_flag_model", 1), ("sao_type_idx_model", 1), ("split_flag_model", 3), ("intra_mode_model", 1),
             ("chroma_pred_model", 2), ("inter_dir", 5), ("trans_subdiv_model", 3), ("qt_cbf_model_luma", 4),
             ("qt_cbf_model_chroma", 4), ("cu_qp_delta_abs", 4), ("part_size_model", 4), ("cu_sig_coeff_group_model", 4),
             ("cu_sig_model_luma", 27), ("cu_sig_model_chroma", 15), ("cu_ctx_last_y_luma", 15), ("cu_ctx_last_y_chroma", 15),
             ("cu_ctx_last_x_luma", 15), ("cu_ctx_last_x_chroma", 15), ("cu_one_model_luma", 16), ("cu_one_model_chroma", 8),
             ("cu_abs_model_luma", 4), ("cu_abs_model_chroma", 2), ("cu_pred_mode_model", 1), ("cu_skip_flag_model", 3),
             ("cu_merge_idx_ext_model", 1), ("cu_merge_flag_ext_model", 1), ("cu_transquant_bypass", 1), ("cu_mvd_model", 2),
             ("cu_ref_pic_model", 2), ("mvp_idx_model", 2), ("cu_qt_root_cbf_model", 1), ("transform_skip_model_luma", 1),
             ("transform_skip_model_chroma", 1)]
    off, o = {}, 0
    for name, sz in sizes:
        off[name] = o
        o += sz
    assert o == api.CABAC_CTX_BYTES
    names = ["qt_cbf_model_luma", "qt_cbf_model_chroma", "cu_sig_coeff_group_model", "cu_sig_model_luma", "cu_sig_model_chroma",
             "cu_ctx_last_y_luma", "cu_ctx_last_y_chroma", "cu_ctx_last_x_luma", "cu_ctx_last_x_chroma", "cu_one_model_luma",
             "cu_one_model_chroma", "cu_abs_model_luma", "cu_abs_model_chroma", "cu_qt_root_cbf_model"]
    assert [off[k] for k in names] == list(ref.cabac_ctx_offsets())


def test_cabac_ctx_init_matches_reference(ref):
    """kvz_cuda_cabac_ctx_init (csrc/cabac_init.cu, host code: needs no GPU) == kvz_init_contexts for every QP / slice type."""
    from kvazaar_b200 import api, lib
    skip = [26, 27]                                  # cu_qp_delta_abs[2..3]: never initialised by the reference
    for slice_type in (0, 1, 2):
        for qp in range(0, 52):
            out = np.zeros(api.CABAC_CTX_BYTES, np.uint8)
            assert lib().kvz_cuda_cabac_ctx_init(qp, slice_type, C.c_void_p(out.ctypes.data)) == 0
            want = ref.init_contexts(qp, slice_type)
            keep = np.ones(api.CABAC_CTX_BYTES, bool)
            keep[skip] = False
            assert np.array_equal(out[keep], want[keep]), (slice_type, qp)


def lambda_for(qp):
    return 0.57 * 2.0 ** ((qp - 12) / 3.0)


def test_oracle_rdoq_vs_reference(orc, ref):
    """oracle/kvz_oracle.c orc_rdoq (plain-C restatement of rdo.c:661-977) == kvz_rdoq of the compiled reference, on the CPU."""
    rng = np.random.default_rng(5)
    for n, count in ((4, 120), (8, 80), (16, 30), (32, 10)):
        for qp in (22, 32):
            for type_ in (0, 2):
                if n == 32 and type_ == 2:
                    continue
                coef = synth_coeffs(rng, n, count, energy=6.0 * 2.0 ** ((qp - 4) / 6.0))
                for signhide in (0, 1):
                    cabac = ref.init_contexts(qp, 2) if signhide else rng.integers(0, 126, ref.cabac_ctx_size()).astype(np.uint8)
                    lam = lambda_for(qp) * float(rng.uniform(0.5, 2.0))
                    for i in range(count):
                        scan = int(rng.integers(0, 3)) if n <= 8 else 0
                        bt, trd = int(rng.integers(1, 3)), int(rng.integers(0, 3))
                        want = ref.rdoq(coef[i].ravel(), n, qp, lam, cabac, type_, scan, bt, trd, signhide)
                        got = orc.rdoq(coef[i].ravel(), n, qp, lam, cabac, type_, scan, bt, trd, signhide)
                        assert np.array_equal(want, got), (n, qp, type_, signhide, i)


CASES = [(n, qp, type_, signhide) for n in (4, 8, 16, 32) for qp in (22, 27, 37) for type_ in (0, 2) for signhide in (0, 1)
         if not (n == 32 and type_ == 2)]


@pytest.mark.gpu
@pytest.mark.parametrize("n,qp,type_,signhide", CASES)
def test_cuda_rdoq_vs_reference(cuda_lib, ref, n, qp, type_, signhide):
    from kvazaar_b200 import api
    rng = np.random.default_rng(1000 * n + 10 * qp + type_ + signhide)
    count = {4: 400, 8: 300, 16: 120, 32: 40}[n]
    coef = synth_coeffs(rng, n, count, energy=6.0 * 2.0 ** ((qp - 4) / 6.0))
    tus = np.zeros(count, api.RDOQ_TU)
    tus["off_coef"] = tus["off_dest"] = np.arange(count) * n * n
    tus["type"] = type_
    tus["scan_idx"] = rng.integers(0, 3, count) if n <= 8 else 0
    tus["block_type"] = rng.integers(1, 3, count)
    tus["tr_depth"] = rng.integers(0, 3, count)
    # half of the cases use the slice-initial context models, half random states
    ctxs = [ref.init_contexts(qp, 2), ref.init_contexts(qp, 1), rng.integers(0, 126, api.CABAC_CTX_BYTES).astype(np.uint8)]
    for ci, cabac in enumerate(ctxs):
        lam = lambda_for(qp) * (1.0 if ci == 0 else float(rng.uniform(0.3, 3.0)))
        got = api.rdoq_batch(api.to_dev(coef.ravel()), n, tus, cabac, qp, lam, 8, signhide).cpu().numpy().reshape(count, n * n)
        bad = []
        nonzero = 0
        for i in range(count):
            want = ref.rdoq(coef[i].ravel(), n, qp, lam, cabac, type_, int(tus["scan_idx"][i]), int(tus["block_type"][i]),
                            int(tus["tr_depth"][i]), signhide)
            nonzero += int(np.count_nonzero(want))
            if not np.array_equal(want, got[i]):
                bad.append(i)
        assert not bad, f"ctx {ci}: {len(bad)} of {count} TUs differ, first {bad[:5]}"
        assert nonzero > 0


# ------------------------------------------------------------------------------------------ coefficient bit cost
def synth_levels(rng, n, count):
    """Quantised-level-like blocks: sparse, small magnitudes, a few large ones, low frequencies denser."""
    fy, fx = np.mgrid[0:n, 0:n]
    dens = np.exp(-(fx + fy) / (n * rng.uniform(0.05, 0.8, (count, 1, 1))))
    nz = rng.random((count, n, n)) < dens * rng.uniform(0.05, 1.0, (count, 1, 1))
    mag = np.rint(np.abs(rng.laplace(0, 1.2, (count, n, n)))).astype(np.int64) + 1
    mag = np.where(rng.random((count, n, n)) < 0.01, rng.integers(1, 3000, (count, n, n)), mag)
    lv = np.where(nz, mag * rng.choice([-1, 1], (count, n, n)), 0)
    lv[rng.random(count) < 0.08] = 0
    return np.clip(lv, -32768, 32767).astype(np.int16)


def test_oracle_mode_bits_vs_reference(ref, orc):
    """MPM derivation and luma / chroma mode-bit estimates (groundwork for the CTU search driver row): oracle == reference."""
    import ctypes as C
    L, O = ref.lib, orc.lib
    L.kvzref_luma_mode_bits.restype = L.kvzref_chroma_mode_bits.restype = C.c_double
    O.orc_luma_mode_bits.restype = O.orc_chroma_mode_bits.restype = C.c_double
    rng = np.random.default_rng(21)
    ctx = ref.ctx(27)
    for _ in range(400):
        left, above, y = int(rng.integers(-1, 35)), int(rng.integers(-1, 35)), int(rng.choice([0, 8, 64, 72, 128]))
        pw, pg = np.zeros(3, np.int8), np.zeros(3, np.int8)
        L.kvzref_intra_mpm(left, above, y, C.c_void_p(pw.ctypes.data))
        O.orc_intra_mpm(left, above, y, C.c_void_p(pg.ctypes.data))
        assert np.array_equal(pw, pg), (left, above, y)
        cabac = rng.integers(0, 126, ref.cabac_ctx_size()).astype(np.uint8)
        mode, cmode = int(rng.integers(0, 35)), int(rng.choice([0, 1, 10, 26, 34, int(rng.integers(0, 35))]))
        cb = np.ascontiguousarray(cabac)
        want = L.kvzref_luma_mode_bits(ctx, C.c_void_p(cb.ctypes.data), mode, C.c_void_p(pw.ctypes.data))
        got = O.orc_luma_mode_bits(C.c_void_p(cb.ctypes.data), mode, C.c_void_p(pg.ctypes.data))
        assert want == got
        assert L.kvzref_chroma_mode_bits(ctx, C.c_void_p(cb.ctypes.data), cmode, mode) == O.orc_chroma_mode_bits(C.c_void_p(cb.ctypes.data), cmode, mode)


def test_python_coeff_cost_port_vs_reference(ref, orc):
    """oracle/coeff_cost_port.py (plain restatement of the bit count) == the compiled reference, on the CPU."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("coeff_cost_port", os.path.join(os.path.dirname(__file__), "..", "oracle", "coeff_cost_port.py"))
    port = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(port)
    rng = np.random.default_rng(11)
    for n, count in ((4, 60), (8, 40), (16, 12), (32, 5)):
        lv = synth_levels(rng, n, count)
        for type_ in (0, 2):
            if n == 32 and type_ == 2:
                continue
            for cabac in (ref.init_contexts(27, 2), rng.integers(0, 126, ref.cabac_ctx_size()).astype(np.uint8)):
                for i in range(count):
                    scan_idx = int(rng.integers(0, 3)) if n <= 8 else 0
                    signhide, tr_skip = int(rng.integers(0, 2)), int(rng.integers(0, 2))
                    want, _ = ref.coeff_cost(lv[i].ravel(), n, cabac, type_, scan_idx, tr_skip, signhide, 1, 0)
                    scan = orc.scan_table(scan_idx, n.bit_length() - 1)
                    assert port.cost(lv[i], n, type_, scan_idx, cabac, scan, signhide, 1, tr_skip) == want, (n, type_, i)


@pytest.mark.gpu
@pytest.mark.parametrize("n,type_,signhide,update", [(n, t, s, u) for n in (4, 8, 16, 32) for t in (0, 2) for s in (0, 1) for u in (0, 1)
                                                     if not (n == 32 and t == 2)])
def test_cuda_coeff_cost_vs_reference(cuda_lib, ref, n, type_, signhide, update):
    """kvz_cuda_coeff_cost_batch == kvz_encode_coeff_nxn in only_count mode (generic and the selected AVX2 version):
    the double bit count and, with update = 1, the adapted context models, bit for bit."""
    from kvazaar_b200 import api
    rng = np.random.default_rng(77 * n + type_ + 3 * signhide + 5 * update)
    count = {4: 300, 8: 200, 16: 80, 32: 30}[n]
    lv = synth_levels(rng, n, count)
    tus = np.zeros(count, api.RDOQ_TU)
    tus["off_coef"] = np.arange(count) * n * n
    tus["type"] = type_
    tus["scan_idx"] = rng.integers(0, 3, count) if n <= 8 else 0
    tus["block_type"] = rng.integers(0, 2, count)                         # transform_skip flag
    trskip = 1
    for cabac in (ref.init_contexts(32, 2), rng.integers(0, 126, api.CABAC_CTX_BYTES).astype(np.uint8)):
        bits, ctx_out = api.coeff_cost_batch(api.to_dev(lv.ravel()), n, tus, cabac, signhide, trskip, update, want_ctx=True)
        bits = bits.cpu().numpy()
        ctx_out = ctx_out.cpu().numpy().reshape(count, -1)
        some = 0
        for i in range(count):
            want, after = ref.coeff_cost(lv[i].ravel(), n, cabac, type_, int(tus["scan_idx"][i]), int(tus["block_type"][i]), signhide, trskip, update)
            assert want == bits[i], (i, want, bits[i])
            if update and want > 0:
                assert np.array_equal(after, ctx_out[i]), i
            some += want > 0
            if i % 7 == 0:                                                # the AVX2 strategy counts the same bits
                w2, _ = ref.coeff_cost(lv[i].ravel(), n, cabac, type_, int(tus["scan_idx"][i]), int(tus["block_type"][i]), signhide, trskip, update,
                                       impl=ref.selected_name("encode_coeff_nxn"))
                assert w2 == want
        assert some > count // 2
