"""Python face of the batched device API (include/kvz_cuda.h, layer 1).

Every function takes torch CUDA tensors (uint8 for 8-bit pixels, int16-viewed-uint16 for >8-bit, int16
coefficients), launches on torch's current stream and returns torch tensors.  Names and argument meaning follow
the reference's strategy functions (src/strategies/strategies-*.h); descriptors are numpy structured arrays with
the layout of the C structs.
"""
import ctypes as C

import numpy as np

from . import lib, KvzCudaError

# numpy dtypes matching the C structs of include/kvz_cuda.h
BLK = np.dtype([("off_a", "<i4"), ("off_b", "<i4"), ("w", "<i2"), ("h", "<i2"), ("left", "<i2"), ("right", "<i2")])
QUAD = np.dtype([("off_pred", "<i4", 4), ("off_orig", "<i4"), ("w", "<i2"), ("h", "<i2")])
TU = np.dtype([("off_ref", "<i4"), ("off_pred", "<i4"), ("off_rec", "<i4"), ("off_coeff", "<i4"), ("width", "u1"),
               ("color", "u1"), ("scan_idx", "u1"), ("use_trskip", "u1"), ("cu_is_intra", "u1"), ("early_skip", "u1"),
               ("phase", "u1"), ("pad", "u1")])
IPOL = np.dtype([("off_src", "<i4"), ("off_dst", "<i4"), ("w", "<i2"), ("h", "<i2"), ("mvx", "<i2"), ("mvy", "<i2")])
SAO_BLK = np.dtype([("off_orig", "<i4"), ("off_rec", "<i4"), ("bw", "<i2"), ("bh", "<i2"), ("stride_orig", "<i4"),
                    ("stride_rec", "<i4")])
SAO_REC = np.dtype([("off_rec", "<i4"), ("off_new", "<i4"), ("bw", "<i2"), ("bh", "<i2"), ("type", "i1"),
                    ("eo_class", "i1"), ("color", "i1"), ("pad", "i1"), ("band_position", "<i4", 2),
                    ("offsets", "<i4", 10)])
assert BLK.itemsize == 16 and QUAD.itemsize == 24 and TU.itemsize == 24 and IPOL.itemsize == 16
assert SAO_BLK.itemsize == 20 and SAO_REC.itemsize == 64

OP_REG_SAD, OP_SATD_ANY, OP_SSD, OP_VER_SAD, OP_HOR_SAD = range(5)
TR_DCT, TR_IDCT, TR_DST, TR_IDST = range(4)
IPOL_LUMA, IPOL_LUMA_HI, IPOL_CHROMA, IPOL_CHROMA_HI = range(4)
IPOL_IM_SIZE = (71 + 1) * 64 + 1
IPOL_FIRST_COLS = 71 + 1


class QuantParams(C.Structure):
    _fields_ = [("qp", C.c_int32), ("bitdepth", C.c_int32), ("slice_is_intra", C.c_int32),
                ("signhide_enable", C.c_int32), ("scaling_list_enable", C.c_int32)]


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise KvzCudaError("no CUDA device: the cuda strategy has no CPU fallback")
    return torch


def _ck(rc):
    if rc != 0:
        raise KvzCudaError(f"libkvzcuda error {rc}: {lib().kvz_cuda_last_error().decode()}")


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(_torch().cuda.current_stream().cuda_stream)


def _bits(t):
    torch = _torch()
    return 8 if t.dtype == torch.uint8 else 10


def to_dev(a):
    """numpy array (incl. structured descriptor arrays) -> CUDA tensor (uint16 is carried as int16)."""
    torch = _torch()
    a = np.ascontiguousarray(a)
    if a.dtype.fields is not None:
        return torch.from_numpy(a.view(np.uint8).copy()).cuda()
    if a.dtype == np.uint16:
        a = a.view(np.int16)
    if a.dtype == np.uint32:
        a = a.view(np.int32)
    return torch.from_numpy(a.copy()).cuda()


def init(device=-1):
    _ck(lib().kvz_cuda_init(device))


def launch_count():
    return int(lib().kvz_cuda_launch_count())


# ------------------------------------------------------------------ picture group
def _nxn(fn, n, a, b, count, out=None):
    torch = _torch()
    if out is None:
        out = torch.empty(count, dtype=torch.int32, device=a.device)
    _ck(fn(n, _bits(a), _p(a), _p(b), count, _p(out), _stream()))
    return out


def sad_nxn_batch(n, a, b, count):
    return _nxn(lib().kvz_cuda_sad_nxn_batch, n, a, b, count)


def satd_nxn_batch(n, a, b, count, out=None):
    return _nxn(lib().kvz_cuda_satd_nxn_batch, n, a, b, count, out)


def cost_nxn_multi_batch(use_satd, n, preds, block_pitch, mode_pitch, num_modes, orig, count):
    torch = _torch()
    out = torch.empty(count * num_modes, dtype=torch.int32, device=preds.device)
    _ck(lib().kvz_cuda_cost_nxn_multi_batch(int(use_satd), n, _bits(preds), _p(preds), C.c_int64(block_pitch),
                                            mode_pitch, num_modes, _p(orig), count, _p(out), _stream()))
    return out.view(count, num_modes)


def block_cost_batch(op, plane_a, stride_a, plane_b, stride_b, descs):
    torch = _torch()
    d = to_dev(descs)
    out = torch.empty(len(descs), dtype=torch.int32, device=plane_a.device)
    _ck(lib().kvz_cuda_block_cost_batch(op, _bits(plane_a), _p(plane_a), stride_a, _p(plane_b), stride_b, _p(d),
                                        len(descs), _p(out), _stream()))
    return out


def satd_any_size_quad_batch(pred_base, pred_stride, orig_base, orig_stride, descs):
    torch = _torch()
    d = to_dev(descs)
    out = torch.empty(len(descs) * 4, dtype=torch.int32, device=pred_base.device)
    _ck(lib().kvz_cuda_satd_any_size_quad_batch(_bits(pred_base), _p(pred_base), pred_stride, _p(orig_base),
                                                orig_stride, _p(d), len(descs), _p(out), _stream()))
    return out.view(-1, 4)


def bipred_average_plane(dst, dst_stride, l0, l1, l0_is_im, l1_is_im, w, h, bitdepth=8):
    _ck(lib().kvz_cuda_bipred_average_plane(bitdepth, _p(dst), dst_stride, _p(l0), _p(l1), int(l0_is_im),
                                            int(l1_is_im), w, h, _stream()))
    return dst


def pixel_var_batch(buf, length, count):
    torch = _torch()
    out = torch.empty(count, dtype=torch.float64, device=buf.device)
    _ck(lib().kvz_cuda_pixel_var_batch(_bits(buf), _p(buf), C.c_uint32(length), count, _p(out), _stream()))
    return out


# ------------------------------------------------------------------ dct / quant groups
def transform_batch(kind, n, bitdepth, inp, count):
    torch = _torch()
    out = torch.empty_like(inp)
    _ck(lib().kvz_cuda_transform_batch(kind, n, bitdepth, _p(inp), _p(out), count, _stream()))
    return out


def quant_params(qp, bitdepth=8, slice_is_intra=1, signhide=0):
    return QuantParams(qp, bitdepth, slice_is_intra, signhide, 0)


def quant_batch(params, coef, n, type_, scan_idx, count):
    torch = _torch()
    out = torch.empty_like(coef)
    _ck(lib().kvz_cuda_quant_batch(C.byref(params), _p(coef), _p(out), n, type_, _p(scan_idx), count, _stream()))
    return out


def dequant_batch(params, q_coef, n, type_, count):
    torch = _torch()
    out = torch.empty_like(q_coef)
    _ck(lib().kvz_cuda_dequant_batch(C.byref(params), _p(q_coef), _p(out), n, type_, count, _stream()))
    return out


def quantize_residual_batch(params, ref_plane, pred_plane, in_stride, rec_plane, out_stride, coeff_out, tus):
    torch = _torch()
    d = to_dev(tus)
    has = torch.empty(len(tus), dtype=torch.int32, device=ref_plane.device)
    _ck(lib().kvz_cuda_quantize_residual_batch(C.byref(params), _p(ref_plane), _p(pred_plane), in_stride,
                                               _p(rec_plane), out_stride, _p(coeff_out), _p(d), len(tus), _p(has),
                                               _stream()))
    return has


def coeff_abs_sum_batch(coeffs, length, count):
    torch = _torch()
    out = torch.empty(count, dtype=torch.int32, device=coeffs.device)
    _ck(lib().kvz_cuda_coeff_abs_sum_batch(_p(coeffs), C.c_size_t(length), count, _p(out), _stream()))
    return out


def fast_coeff_cost_batch(coeffs, width, weights, count):
    torch = _torch()
    out = torch.empty(count, dtype=torch.int32, device=coeffs.device)
    _ck(lib().kvz_cuda_fast_coeff_cost_batch(_p(coeffs), width, C.c_uint64(weights), count, _p(out), _stream()))
    return out


# ------------------------------------------------------------------ intra group
def intra_predict_batch(level, log2w, color, filter_boundary, ref_top, ref_left, modes, count):
    torch = _torch()
    out = torch.empty(count << (2 * log2w), dtype=ref_top.dtype, device=ref_top.device)
    _ck(lib().kvz_cuda_intra_predict_batch(level, log2w, color, int(filter_boundary), _bits(ref_top), _p(ref_top),
                                           _p(ref_left), _p(modes), count, _p(out), _stream()))
    return out


def intra_build_reference_batch(log2w, color, rec_plane, stride, pic_w, pic_h, luma_xy):
    torch = _torch()
    xy = to_dev(np.ascontiguousarray(luma_xy, np.int32))
    count = len(luma_xy)
    n = 2 * (1 << log2w) + 1
    top = torch.empty(count * n, dtype=rec_plane.dtype, device=rec_plane.device)
    left = torch.empty_like(top)
    _ck(lib().kvz_cuda_intra_build_reference_batch(log2w, color, _bits(rec_plane), _p(rec_plane), stride, pic_w, pic_h,
                                                   _p(xy), count, _p(top), _p(left), _stream()))
    return top.view(count, n), left.view(count, n)


def intra_rough_search_frame(log2w, src_plane, rec_plane, stride, pic_w, pic_h, out=None):
    torch = _torch()
    w = 1 << log2w
    nblk = (pic_w // w) * (pic_h // w)
    if out is None:
        out = torch.empty(nblk * 35, dtype=torch.int32, device=src_plane.device)
    _ck(lib().kvz_cuda_intra_rough_search_frame(log2w, _bits(src_plane), _p(src_plane), _p(rec_plane), stride, pic_w,
                                                pic_h, _p(out), _stream()))
    return out.view(nblk, 35)


# ------------------------------------------------------------------ ipol group
def sample_batch(kind, src_plane, src_stride, dst, dst_stride, descs):
    d = to_dev(descs)
    _ck(lib().kvz_cuda_sample_batch(kind, _bits(src_plane), _p(src_plane), src_stride, _p(dst), dst_stride, _p(d),
                                    len(descs), _stream()))
    return dst


def filter_fme_batch(stage, src_plane, src_stride, src_off, w, h, filtered, hor_intermediate, fme_level,
                     hor_first_cols, hpel_off):
    so = to_dev(np.ascontiguousarray(src_off, np.int32))
    ho = to_dev(np.ascontiguousarray(hpel_off, np.int8)) if hpel_off is not None else None
    _ck(lib().kvz_cuda_filter_fme_batch(stage, _bits(src_plane), _p(src_plane), src_stride, _p(so), w, h, _p(filtered),
                                        _p(hor_intermediate), fme_level, _p(hor_first_cols), _p(ho), len(src_off),
                                        _stream()))


def extend_block(src, src_w, src_h, src_s, bx, by, bw, bh, pl, pr, pt, pb, pbs):
    torch = _torch()
    buf = torch.empty((pl + bw + pr) * (pt + bh + pb + pbs) + 1, dtype=src.dtype, device=src.device)
    _ck(lib().kvz_cuda_extend_block(_bits(src), _p(src), src_w, src_h, src_s, bx, by, bw, bh, pl, pr, pt, pb, pbs,
                                    _p(buf), _stream()))
    return buf


# ------------------------------------------------------------------ sao group
def sao_edge_stats_batch(bitdepth, orig, rec, blks):
    torch = _torch()
    d = to_dev(blks)
    out = torch.empty(len(blks) * 40, dtype=torch.int32, device=orig.device)
    _ck(lib().kvz_cuda_sao_edge_stats_batch(bitdepth, _p(orig), _p(rec), _p(d), len(blks), _p(out), _stream()))
    return out.view(len(blks), 4, 2, 5)


def sao_edge_ddistortion_batch(bitdepth, orig, rec, blks, eo_class, offsets):
    torch = _torch()
    d = to_dev(blks)
    eo = to_dev(np.ascontiguousarray(eo_class, np.int8))
    of = to_dev(np.ascontiguousarray(offsets, np.int32))
    out = torch.empty(len(blks), dtype=torch.int32, device=orig.device)
    _ck(lib().kvz_cuda_sao_edge_ddistortion_batch(bitdepth, _p(orig), _p(rec), _p(d), _p(eo), _p(of), len(blks),
                                                  _p(out), _stream()))
    return out


def sao_band_ddistortion_batch(bitdepth, orig, rec, blks, band_pos, bands):
    torch = _torch()
    d = to_dev(blks)
    bp = to_dev(np.ascontiguousarray(band_pos, np.int32))
    bd = to_dev(np.ascontiguousarray(bands, np.int32))
    out = torch.empty(len(blks), dtype=torch.int32, device=orig.device)
    _ck(lib().kvz_cuda_sao_band_ddistortion_batch(bitdepth, _p(orig), _p(rec), _p(d), _p(bp), _p(bd), len(blks),
                                                  _p(out), _stream()))
    return out


def sao_reconstruct_batch(bitdepth, rec, stride, new_rec, new_stride, descs):
    d = to_dev(descs)
    _ck(lib().kvz_cuda_sao_reconstruct_batch(bitdepth, _p(rec), stride, _p(new_rec), new_stride, _p(d), len(descs),
                                             _stream()))
    return new_rec


# ------------------------------------------------------------------ nal group
def array_checksum(data, height, width, stride, out=None):
    torch = _torch()
    if out is None:
        out = torch.empty(4, dtype=torch.uint8, device=data.device)
    _ck(lib().kvz_cuda_array_checksum(_bits(data), _p(data), height, width, stride, _p(out), _stream()))
    return out


# ------------------------------------------------------------------ RDOQ (rdoq.cu)
RDOQ_TU = np.dtype([("off_coef", "<i4"), ("off_dest", "<i4"), ("type", "u1"), ("scan_idx", "u1"), ("block_type", "u1"),
                    ("tr_depth", "u1")])
CABAC_CTX_BYTES = 184          # sizeof(cabac_data_t.ctx), src/cabac.h:66-102 (pinned by tests/test_rdoq.py)


class RdoqParams(C.Structure):
    _fields_ = [("lambda_", C.c_double), ("qp", C.c_int32), ("bitdepth", C.c_int32), ("signhide_enable", C.c_int32),
                ("pad", C.c_int32)]


def rdoq_batch(coef, n, tus, cabac_ctx, qp, lambda_, bitdepth=8, signhide=0, dest=None):
    """kvz_rdoq (src/rdo.c:661) for `len(tus)` TUs of width n.  coef: int16 CUDA tensor; tus: RDOQ_TU array;
    cabac_ctx: uint8[CABAC_CTX_BYTES] image of state->cabac.ctx.  Returns the quantised levels (int16 tensor)."""
    torch = _torch()
    if dest is None:
        dest = torch.zeros_like(coef)
    prm = RdoqParams(lambda_, qp, bitdepth, signhide, 0)
    ctx = cabac_ctx if hasattr(cabac_ctx, "data_ptr") else to_dev(np.asarray(cabac_ctx, np.uint8))
    assert ctx.numel() == CABAC_CTX_BYTES
    tus_d = tus if hasattr(tus, "data_ptr") else to_dev(tus)
    count = tus_d.numel() // RDOQ_TU.itemsize
    _ck(lib().kvz_cuda_rdoq_batch(C.byref(prm), _p(ctx), _p(coef), _p(dest), n, _p(tus_d), count, _stream()))
    return dest


class CoeffCostParams(C.Structure):
    _fields_ = [("signhide_enable", C.c_int32), ("trskip_enable", C.c_int32), ("update", C.c_int32), ("pad", C.c_int32)]


def coeff_cost_batch(coeff, n, tus, cabac_ctx, signhide=0, trskip_enable=0, update=0, want_ctx=False):
    """CABAC bit cost per TU (kvz_get_coeff_cost's CABAC branch, src/rdo.c:291-330).  tus: RDOQ_TU array (off_coef, type,
    scan_idx, block_type = transform_skip flag).  Returns float64 tensor [count] (and the adapted context models)."""
    torch = _torch()
    ctx = cabac_ctx if hasattr(cabac_ctx, "data_ptr") else to_dev(np.asarray(cabac_ctx, np.uint8))
    tus_d = tus if hasattr(tus, "data_ptr") else to_dev(tus)
    count = tus_d.numel() // RDOQ_TU.itemsize
    bits = torch.zeros(count, dtype=torch.float64, device=coeff.device)
    ctx_out = torch.zeros(count * CABAC_CTX_BYTES, dtype=torch.uint8, device=coeff.device) if want_ctx else None
    prm = CoeffCostParams(signhide, trskip_enable, update, 0)
    _ck(lib().kvz_cuda_coeff_cost_batch(C.byref(prm), _p(ctx), _p(coeff), n, _p(tus_d), count, _p(bits), _p(ctx_out), _stream()))
    return (bits, ctx_out) if want_ctx else bits


# ------------------------------------------------------------------ deblocking (deblock.cu)
class DbkParams(C.Structure):
    """kvz_cuda_dbk_params."""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("qp", C.c_int32), ("beta_offset_div2", C.c_int32),
                ("tc_offset_div2", C.c_int32), ("slice_is_b", C.c_int32), ("per_cu_qp", C.c_int32),
                ("cu_stride_scu", C.c_int32), ("ref_LX", C.c_uint8 * 32)]


def deblock_frame(y, u, v, cus, width, height, qp, beta_offset_div2=0, tc_offset_div2=0, slice_is_b=0, per_cu_qp=0,
                  ref_lx=None):
    """kvz_filter_deblock_lcu over a whole frame (src/filter.c:783), in place on the CUDA plane tensors.
    `cus`: uint8 CUDA tensor [rows_scu, stride_scu, 20] -- the reference's cu_info_t records."""
    prm = DbkParams(width, height, qp, beta_offset_div2, tc_offset_div2, slice_is_b, per_cu_qp, int(cus.shape[1]))
    if ref_lx is not None:
        C.memmove(prm.ref_LX, np.ascontiguousarray(ref_lx, np.uint8).ctypes.data, 32)
    _ck(lib().kvz_cuda_deblock_frame(C.byref(prm), _bits(y), _p(y), _p(u), _p(v), _p(cus), _stream()))
    return y, u, v


def call_deblock_frame(y, u, v, stride, cus, width, height, qp, beta_offset_div2=0, tc_offset_div2=0, slice_is_b=0,
                       per_cu_qp=0, ref_lx=None, bitdepth=8):
    """Host-buffer form (numpy arrays, in place) -- what the binding in INTEGRATION.md calls."""
    prm = DbkParams(width, height, qp, beta_offset_div2, tc_offset_div2, slice_is_b, per_cu_qp, int(cus.shape[1]))
    if ref_lx is not None:
        C.memmove(prm.ref_LX, np.ascontiguousarray(ref_lx, np.uint8).ctypes.data, 32)
    _ck(lib().kvz_cuda_call_deblock_frame(C.byref(prm), bitdepth, C.c_void_p(y.ctypes.data), C.c_void_p(u.ctypes.data),
                                          C.c_void_p(v.ctypes.data), stride, C.c_void_p(cus.ctypes.data)))


# ------------------------------------------------------------------ frame-level pass (framepass.cu)
class FpParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("bitdepth", C.c_int32), ("qp", C.c_int32),
                ("signhide", C.c_int32), ("rdoq", C.c_int32), ("trskip", C.c_int32), ("pad", C.c_int32), ("lambda_", C.c_double)]


class FpLayout(C.Structure):
    _fields_ = [("nblk", C.c_int32 * 4), ("nctu", C.c_int32), ("host_bytes", C.c_uint64),
                ("mode_y", C.c_uint64 * 4), ("cost_y", C.c_uint64 * 4), ("has_y", C.c_uint64 * 4),
                ("ssd_y", C.c_uint64 * 4), ("coeff_y", C.c_uint64 * 4),
                ("has_u", C.c_uint64 * 3), ("has_v", C.c_uint64 * 3), ("ssd_u", C.c_uint64 * 3),
                ("ssd_v", C.c_uint64 * 3), ("coeff_u", C.c_uint64 * 3), ("coeff_v", C.c_uint64 * 3),
                ("sao_stats", C.c_uint64), ("sao_dd", C.c_uint64), ("sao_band_dd", C.c_uint64),
                ("sao_best", C.c_uint64), ("sao_rec", C.c_uint64), ("checksum", C.c_uint64),
                ("bits_y", C.c_uint64 * 4), ("bits_u", C.c_uint64 * 3), ("bits_v", C.c_uint64 * 3), ("trskip_y", C.c_uint64),
                ("coeff_begin", C.c_uint64), ("n_chunks", C.c_uint64), ("compact_header_bytes", C.c_uint64)]


def fp_layout_for(width, height, qp=27, signhide=0, bitdepth=8):
    """Result-blob layout; needs no GPU."""
    lay = FpLayout()
    prm = FpParams(width, height, bitdepth, qp, signhide, 0, 0, 0, 0.0)
    _ck(lib().kvz_cuda_fp_layout_for(C.byref(prm), C.byref(lay)))
    return lay


class FramePass:
    """One in-flight frame of the frame-level pass (device buffers owned by the library)."""

    def __init__(self, width, height, qp=27, signhide=0, rdoq=0, lambda_=0.0, trskip=0, bitdepth=8):
        _torch()
        L = lib()
        L.kvz_cuda_fp_create.restype = C.c_void_p
        L.kvz_cuda_fp_result_dev.restype = C.c_void_p
        L.kvz_cuda_fp_result_dev.argtypes = [C.c_void_p]
        L.kvz_cuda_fp_frame_bytes.restype = C.c_size_t
        L.kvz_cuda_fp_frame_bytes.argtypes = [C.c_void_p]
        self.params = FpParams(width, height, bitdepth, qp, signhide, rdoq, trskip, 0, lambda_)
        self.bitdepth = bitdepth
        h = L.kvz_cuda_fp_create(C.byref(self.params))
        if not h:
            raise KvzCudaError(f"kvz_cuda_fp_create failed: {L.kvz_cuda_last_error().decode()}")
        self.h = C.c_void_p(h)
        self.layout = FpLayout()
        _ck(L.kvz_cuda_fp_layout_get(self.h, C.byref(self.layout)))
        self.host_bytes = int(self.layout.host_bytes)
        self.frame_bytes = int(L.kvz_cuda_fp_frame_bytes(self.h))
        self.width, self.height = width, height

    def close(self):
        if self.h:
            lib().kvz_cuda_fp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_dev(self, src, rec_in=None):
        _ck(lib().kvz_cuda_fp_run_dev(self.h, _p(src), _p(rec_in), _stream()))

    def run_host(self, src_host, result_host):
        """src_host / result_host: pinned CPU uint8 tensors (frame_bytes / host_bytes)."""
        _ck(lib().kvz_cuda_fp_run_host(self.h, C.c_void_p(src_host.data_ptr()), C.c_void_p(result_host.data_ptr()),
                                       _stream()))

    def run_host_compact(self, src_host, small_host, compact_host, budget_chunks):
        """Like run_host, but the coefficient region comes back as bitmap + non-zero 32-byte chunks (see
        kvz_cuda_fp_run_host_compact).  small_host: layout.coeff_begin bytes; compact_host: layout.compact_header_bytes +
        32 * budget_chunks bytes (pinned uint8 tensors)."""
        _ck(lib().kvz_cuda_fp_run_host_compact(self.h, C.c_void_p(src_host.data_ptr()), C.c_void_p(small_host.data_ptr()),
                                               C.c_void_p(compact_host.data_ptr()), C.c_uint32(budget_chunks), _stream()))

    def result_host(self):
        """Synchronous copy of the result blob to a numpy array."""
        torch = _torch()
        out = np.empty(self.host_bytes, np.uint8)
        torch.cuda.current_stream().synchronize()
        _ck(lib().kvz_cuda_memcpy_d2h(C.c_void_p(out.ctypes.data), C.c_void_p(lib().kvz_cuda_fp_result_dev(self.h)),
                                      C.c_size_t(self.host_bytes), _stream()))
        torch.cuda.current_stream().synchronize()
        return out


def fp_sections(layout, width, height, bitdepth=8):
    """name -> (offset, dtype, count) for every section of the result blob."""
    out = {}
    for d in range(4):
        w = 32 >> d
        nb = layout.nblk[d]
        out[f"mode_y{d}"] = (layout.mode_y[d], np.int8, nb)
        out[f"cost_y{d}"] = (layout.cost_y[d], np.uint32, nb)
        out[f"has_y{d}"] = (layout.has_y[d], np.uint8, nb)
        out[f"ssd_y{d}"] = (layout.ssd_y[d], np.uint32, nb)
        out[f"coeff_y{d}"] = (layout.coeff_y[d], np.int16, nb * w * w)
        out[f"bits_y{d}"] = (layout.bits_y[d], np.float64, nb)
        if d == 3:
            out["trskip_y"] = (layout.trskip_y, np.uint8, nb)
        if d < 3:
            wc = w // 2
            for c in "uv":
                out[f"has_{c}{d}"] = (getattr(layout, f"has_{c}")[d], np.uint8, nb)
                out[f"ssd_{c}{d}"] = (getattr(layout, f"ssd_{c}")[d], np.uint32, nb)
                out[f"coeff_{c}{d}"] = (getattr(layout, f"coeff_{c}")[d], np.int16, nb * wc * wc)
                out[f"bits_{c}{d}"] = (getattr(layout, f"bits_{c}")[d], np.float64, nb)
    n3 = 3 * layout.nctu
    out["sao_stats"] = (layout.sao_stats, np.int32, n3 * 40)
    out["sao_dd"] = (layout.sao_dd, np.int32, n3 * 4)
    out["sao_band_dd"] = (layout.sao_band_dd, np.int32, n3)
    out["sao_best"] = (layout.sao_best, np.int8, n3)
    out["sao_rec"] = (layout.sao_rec, np.uint8 if bitdepth == 8 else np.uint16, width * height * 3 // 2)
    out["checksum"] = (layout.checksum, np.uint8, 12)
    return out


def fp_expand_compact(layout, small, compact):
    """Rebuild the full result blob from the compact form (numpy uint8 arrays).  Host-side inverse of the device
    compaction: bit c of the bitmap set <=> the next 32 bytes of the packed stream belong at chunk c."""
    n_chunks, hdr = int(layout.n_chunks), int(layout.compact_header_bytes)
    head = compact[:12].view(np.uint32)
    nonzero, copied = int(head[0]), (len(compact) - hdr) // 32
    assert int(head[1]) == n_chunks and nonzero <= copied, "compact buffer truncated: fetch the tail with kvz_cuda_fp_compact_fetch"
    bits = np.unpackbits(compact[256:256 + (n_chunks + 7) // 8], bitorder="little")[:n_chunks].astype(bool)
    assert int(bits.sum()) == nonzero
    region = np.zeros((n_chunks, 32), np.uint8)
    region[bits] = compact[hdr:hdr + nonzero * 32].reshape(nonzero, 32)
    return np.concatenate([small[:int(layout.coeff_begin)], region.ravel()])


def fp_section(blob, sections, name):
    off, dt, n = sections[name]
    return blob[off: off + n * np.dtype(dt).itemsize].view(dt)


# ------------------------------------------------------------------ frame-level INTER pass (interpass.cu)
class IpParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("bitdepth", C.c_int32), ("qp", C.c_int32),
                ("search_range", C.c_int32)]


class IpLayout(C.Structure):
    _fields_ = [("npu", C.c_int32), ("pus_x", C.c_int32), ("pus_y", C.c_int32), ("host_bytes", C.c_uint64)] + \
               [(n, C.c_uint64) for n in ("mv_int", "sad_int", "mv_final", "satd_best", "has_y", "ssd_y", "coeff_y",
                                          "has_u", "has_v", "ssd_u", "ssd_v", "coeff_u", "coeff_v", "rec")]


def ip_layout_for(width, height, qp=27, search_range=8):
    lay = IpLayout()
    prm = IpParams(width, height, 8, qp, search_range)
    _ck(lib().kvz_cuda_ip_layout_for(C.byref(prm), C.byref(lay)))
    return lay


def ip_sections(layout, width, height):
    n = layout.npu
    out = {"mv_int": (layout.mv_int, np.int16, 2 * n), "sad_int": (layout.sad_int, np.uint32, n),
           "mv_final": (layout.mv_final, np.int16, 2 * n), "satd_best": (layout.satd_best, np.uint32, n),
           "has_y": (layout.has_y, np.int32, n), "ssd_y": (layout.ssd_y, np.uint32, n), "coeff_y": (layout.coeff_y, np.int16, n * 256),
           "rec": (layout.rec, np.uint8, width * height * 3 // 2)}
    for c in "uv":
        out[f"has_{c}"] = (getattr(layout, f"has_{c}"), np.int32, n)
        out[f"ssd_{c}"] = (getattr(layout, f"ssd_{c}"), np.uint32, n)
        out[f"coeff_{c}"] = (getattr(layout, f"coeff_{c}"), np.int16, n * 64)
    return out


class InterPass:
    """One in-flight frame of the frame-level inter pass."""

    def __init__(self, width, height, qp=27, search_range=8):
        _torch()
        L = lib()
        L.kvz_cuda_ip_create.restype = C.c_void_p
        L.kvz_cuda_ip_result_dev.restype = C.c_void_p
        L.kvz_cuda_ip_result_dev.argtypes = [C.c_void_p]
        self.params = IpParams(width, height, 8, qp, search_range)
        h = L.kvz_cuda_ip_create(C.byref(self.params))
        if not h:
            raise KvzCudaError(f"kvz_cuda_ip_create failed: {L.kvz_cuda_last_error().decode()}")
        self.h = C.c_void_p(h)
        self.layout = ip_layout_for(width, height, qp, search_range)
        self.host_bytes = int(self.layout.host_bytes)
        self.frame_bytes = width * height * 3 // 2

    def close(self):
        if self.h:
            lib().kvz_cuda_ip_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_dev(self, cur, ref):
        _ck(lib().kvz_cuda_ip_run_dev(self.h, _p(cur), _p(ref), _stream()))

    def run_host(self, cur_host, ref_host, result_host):
        _ck(lib().kvz_cuda_ip_run_host(self.h, C.c_void_p(cur_host.data_ptr()), C.c_void_p(ref_host.data_ptr()),
                                       C.c_void_p(result_host.data_ptr()), _stream()))

    def result_host(self):
        torch = _torch()
        out = np.empty(self.host_bytes, np.uint8)
        torch.cuda.current_stream().synchronize()
        _ck(lib().kvz_cuda_memcpy_d2h(C.c_void_p(out.ctypes.data), C.c_void_p(lib().kvz_cuda_ip_result_dev(self.h)),
                                      C.c_size_t(self.host_bytes), _stream()))
        torch.cuda.current_stream().synchronize()
        return out


# ------------------------------------------------------------------ integer motion estimation (me_search.cu)
# record layouts of include/kvz_cuda.h (kvz_cuda_me_merge / kvz_cuda_me_pu / kvz_cuda_me_result / kvz_cuda_me_params)
ME_MERGE = np.dtype([("mv", "<i2", (2, 2)), ("dir", "u1"), ("ref", "u1", (2,)), ("pad", "u1")])
ME_PU = np.dtype([("x", "<i2"), ("y", "<i2"), ("w", "<i2"), ("h", "<i2"), ("mv_cand", "<i2", (2, 2)), ("start_mv", "<i2", (2,)),
                  ("num_merge", "<i2"), ("pad", "<i2"), ("merge", ME_MERGE, (5,))])
ME_RESULT = np.dtype([("cost", "<f8"), ("bits", "<i4"), ("mv", "<i2", (2,)), ("points", "<i4"), ("pad", "<i4")])


class MeParams(C.Structure):
    """kvz_cuda_me_params: the configuration fields the reference's integer search reads (search_inter.c:94-247, 436-888)"""
    _fields_ = [(n, C.c_int32) for n in ("width", "height", "bitdepth", "ime_algorithm", "me_max_steps", "me_early_termination", "mv_constraint",
                                         "wpp_owf", "delay_px", "max_ref_lcu_right", "max_ref_lcu_down", "satd_final")] + [("lambda_sqrt", C.c_double)]


def me_search_batch(params, cur, ref, pus, out=None):
    """Integer motion search of `pus` (CUDA byte tensor holding ME_PU records) on device planes `cur` / `ref` (2-D tensors,
    uint8 or int16-carried 10-bit); returns a CUDA byte tensor of ME_RESULT records.  One launch, current stream."""
    torch = _torch()
    count = pus.numel() // ME_PU.itemsize
    if out is None:
        out = torch.empty(count * ME_RESULT.itemsize, dtype=torch.uint8, device=cur.device)
    _ck(lib().kvz_cuda_me_search_batch(C.byref(params), _p(cur), C.c_int(cur.stride(0)), _p(ref), C.c_int(ref.stride(0)), _p(pus), C.c_int(count),
                                       _p(out), _stream()))
    return out


# AMVP / merge candidate derivation: kvz_cuda_me_cu / kvz_cuda_me_cand_pu / kvz_cuda_me_cand_out / kvz_cuda_me_frame
ME_CU = np.dtype([("mv", "<i2", (2, 2)), ("type", "u1"), ("mv_dir", "u1"), ("mv_ref", "u1", (2,))])
ME_CAND_PU = np.dtype([("x", "<i2"), ("y", "<i2"), ("w", "<i2"), ("h", "<i2"), ("mv_ref", "u1", (2,)), ("use_a1", "u1"), ("use_b1", "u1")])
ME_CAND_OUT = np.dtype([("mv_cand", "<i2", (2, 2, 2)), ("num_merge", "<i4"), ("merge", ME_MERGE, (5,))])


class MeFrame(C.Structure):
    """kvz_cuda_me_frame: reference lists and POCs as src/inter.c:836-1572 reads them from state->frame"""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("poc", C.c_int32), ("slice_b", C.c_int32), ("tmvp_enable", C.c_int32),
                ("max_merge", C.c_int32), ("used_size", C.c_int32), ("ref_LX_size", C.c_int32 * 2), ("pocs", C.c_int32 * 16),
                ("col_ref_pocs", (C.c_int32 * 16) * 2), ("ref_LX", (C.c_uint8 * 16) * 2)]


def me_candidates_batch(frame, cus, col_cus, pus, out=None):
    """AMVP + merge candidates of `pus` (CUDA byte tensor of ME_CAND_PU records) from the CU records `cus` / `col_cus`
    (2-D CUDA byte tensors, rows x (stride * 12) bytes); returns a CUDA byte tensor of ME_CAND_OUT records."""
    torch = _torch()
    count = pus.numel() // ME_CAND_PU.itemsize
    if out is None:
        out = torch.empty(count * ME_CAND_OUT.itemsize, dtype=torch.uint8, device=cus.device)
    _ck(lib().kvz_cuda_me_candidates_batch(C.byref(frame), _p(cus), C.c_int(cus.stride(0) // ME_CU.itemsize), _p(col_cus),
                                           C.c_int(col_cus.stride(0) // ME_CU.itemsize), _p(pus), C.c_int(count), _p(out), _stream()))
    return out


def me_frac_search_batch(params, fme_level, cur, ref, pus, out=None):
    """Fractional search (search_frac) of `pus` around their start_mv (the integer search's best MV); layouts as me_search_batch."""
    torch = _torch()
    count = pus.numel() // ME_PU.itemsize
    if out is None:
        out = torch.empty(count * ME_RESULT.itemsize, dtype=torch.uint8, device=cur.device)
    _ck(lib().kvz_cuda_me_frac_search_batch(C.byref(params), C.c_int(fme_level), _p(cur), C.c_int(cur.stride(0)), _p(ref), C.c_int(ref.stride(0)),
                                            _p(pus), C.c_int(count), _p(out), _stream()))
    return out


# merge analysis: kvz_cuda_me_refs / kvz_cuda_me_merge_cost
ME_MERGE_COST = np.dtype([("cost", "<f8", (5,)), ("bits", "<f8", (5,)), ("size", "<i4"), ("keys", "i1", (5,)), ("merge_idx", "i1", (5,)), ("pad", "i1", (2,))])


class MeRefs(C.Structure):
    """kvz_cuda_me_refs: the reference pictures' luma planes (device pointers), the reference lists, cfg.bipred and the merge bits"""
    _fields_ = [("plane", C.c_void_p * 16), ("stride", C.c_int32 * 16), ("ref_LX", (C.c_uint8 * 16) * 2), ("bipred", C.c_int32), ("pad", C.c_int32),
                ("merge_flag_bits", C.c_double), ("merge_idx_bits", C.c_double * 2)]


def me_merge_cost_batch(params, refs, cur, pus, out=None):
    """Merge analysis of `pus` (their merge candidates) against the pictures of `refs`; returns a CUDA byte tensor of ME_MERGE_COST records."""
    torch = _torch()
    count = pus.numel() // ME_PU.itemsize
    if out is None:
        out = torch.empty(count * ME_MERGE_COST.itemsize, dtype=torch.uint8, device=cur.device)
    _ck(lib().kvz_cuda_me_merge_cost_batch(C.byref(params), C.byref(refs), _p(cur), C.c_int(cur.stride(0)), _p(pus), C.c_int(count), _p(out), _stream()))
    return out


# bi-prediction from two uni-predictions: kvz_cuda_me_bipred_pu / kvz_cuda_me_bipred_result
ME_BIPRED_PU = np.dtype([("x", "<i2"), ("y", "<i2"), ("w", "<i2"), ("h", "<i2"), ("mv", "<i2", (2, 2)), ("mv_ref", "u1", (2,)), ("pad", "u1", (2,)),
                         ("mv_cand", "<i2", (2, 2))])
ME_BIPRED_RESULT = np.dtype([("cost", "<f8"), ("bits", "<i4"), ("mv_cand_idx", "u1", (2,)), ("valid", "u1"), ("pad", "u1")])


def me_bipred_batch(params, refs, cur, pus, out=None):
    """Cost of bi-predicting `pus` (CUDA byte tensor of ME_BIPRED_PU records) from their two uni-predictions; ME_BIPRED_RESULT records."""
    torch = _torch()
    count = pus.numel() // ME_BIPRED_PU.itemsize
    if out is None:
        out = torch.empty(count * ME_BIPRED_RESULT.itemsize, dtype=torch.uint8, device=cur.device)
    _ck(lib().kvz_cuda_me_bipred_batch(C.byref(params), C.byref(refs), _p(cur), C.c_int(cur.stride(0)), _p(pus), C.c_int(count), _p(out), _stream()))
    return out


# motion compensation: kvz_cuda_me_mc_refs / kvz_cuda_me_mc_pu
ME_MC_PU = np.dtype([("x", "<i2"), ("y", "<i2"), ("w", "<i2"), ("h", "<i2"), ("mv", "<i2", (2, 2)), ("mv_ref", "u1", (2,)), ("dir", "u1"), ("pad", "u1")])


class MeMcRefs(C.Structure):
    """kvz_cuda_me_mc_refs: Y / U / V planes of the reference pictures (device pointers) and the reference lists"""
    _fields_ = [("y", C.c_void_p * 16), ("u", C.c_void_p * 16), ("v", C.c_void_p * 16), ("ref_LX", (C.c_uint8 * 16) * 2)]


def me_predict_batch(params, refs, pus, pred_y, pred_u, pred_v):
    """Motion compensation of `pus` (CUDA byte tensor of ME_MC_PU records, not overlapping) into the prediction planes (CUDA tensors)."""
    count = pus.numel() // ME_MC_PU.itemsize
    _ck(lib().kvz_cuda_me_predict_batch(C.byref(params), C.byref(refs), _p(pus), C.c_int(count), _p(pred_y), _p(pred_u), _p(pred_v), _stream()))
