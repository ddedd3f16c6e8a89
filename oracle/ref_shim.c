/*
 * ref_shim.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Small adapter compiled against the UNMODIFIED reference headers and linked to
 * oracle/_ref/libkvazaar_ref.so (built by oracle/Makefile from /root/reference/src).
 * It exposes the reference's strategy registry and the handful of struct-taking
 * strategy functions through plain-C entry points that ctypes can call, so the
 * tests can use the real reference as the parity checker (mode 1).
 * It contains no kernel arithmetic of its own.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "global.h"
#include "kvazaar.h"
#include "kvazaar_internal.h"
#include "encoder.h"
#include "encoderstate.h"
#include "strategyselector.h"
#include "cu.h"
#include "image.h"
#include "intra.h"
#include "sao.h"
#include "tables.h"

static strategy_list_t g_list;
static int g_ready = 0;

int kvzref_bitdepth(void) { return KVZ_BIT_DEPTH; }

/* Same sequence as the reference's own test harness (tests/test_strategies.c:41-65),
 * extended to all eight groups. */
int kvzref_init(void)
{
  if (g_ready) return 1;
  memset(&g_list, 0, sizeof(g_list));
  if (!kvz_strategyselector_init(1, KVZ_BIT_DEPTH, 0)) return 0;
  int ok = 1;
  ok &= kvz_strategy_register_picture(&g_list, KVZ_BIT_DEPTH);
  ok &= kvz_strategy_register_nal(&g_list, KVZ_BIT_DEPTH);
  ok &= kvz_strategy_register_dct(&g_list, KVZ_BIT_DEPTH);
  ok &= kvz_strategy_register_ipol(&g_list, KVZ_BIT_DEPTH);
  ok &= kvz_strategy_register_quant(&g_list, KVZ_BIT_DEPTH);
  ok &= kvz_strategy_register_intra(&g_list, KVZ_BIT_DEPTH);
  ok &= kvz_strategy_register_sao(&g_list, KVZ_BIT_DEPTH);
  ok &= kvz_strategy_register_encode(&g_list, KVZ_BIT_DEPTH);
  g_ready = ok;
  return ok;
}

/* (type, strategy_name) -> function pointer, e.g. ("satd_8x8", "generic"). */
void *kvzref_find(const char *type, const char *name)
{
  if (!kvzref_init()) return NULL;
  void *best = NULL;
  unsigned best_prio = 0;
  for (unsigned i = 0; i < g_list.count; ++i) {
    const strategy_t *s = &g_list.strategies[i];
    if (strcmp(s->type, type) != 0) continue;
    if (name && name[0]) {
      if (strcmp(s->strategy_name, name) == 0) return s->fptr;
    } else if (!best || s->priority >= best_prio) {
      best = s->fptr; best_prio = s->priority;
    }
  }
  return best;
}

/* The implementation the reference itself selected (highest priority; AVX2 where present). */
void *kvzref_selected(const char *type)
{
  if (!kvzref_init()) return NULL;
  for (const strategy_to_select_t *s = strategies_to_select; s->strategy_type; ++s)
    if (strcmp(s->strategy_type, type) == 0) return *s->fptr;
  return NULL;
}

const char *kvzref_selected_name(const char *type)
{
  void *sel = kvzref_selected(type);
  for (unsigned i = 0; i < g_list.count; ++i)
    if (g_list.strategies[i].fptr == sel && strcmp(g_list.strategies[i].type, type) == 0)
      return g_list.strategies[i].strategy_name;
  return "?";
}

int kvzref_count(void) { return kvzref_init() ? (int)g_list.count : 0; }
const char *kvzref_entry(int i, const char **name, int *prio)
{
  *name = g_list.strategies[i].strategy_name;
  *prio = (int)g_list.strategies[i].priority;
  return g_list.strategies[i].type;
}

/* Tables the oracle regenerates from rules; exported so tests can pin them. */
const uint32_t *kvzref_scan_table(int scan_idx, int log2_size) { return kvz_g_sig_last_scan[scan_idx][log2_size - 1]; }
extern const int16_t kvz_g_dct_4[4][4], kvz_g_dct_8[8][8], kvz_g_dct_16[16][16], kvz_g_dct_32[32][32];
int kvzref_dct_coef(int n, int k, int i)
{
  switch (n) { case 4: return kvz_g_dct_4[k][i]; case 8: return kvz_g_dct_8[k][i];
               case 16: return kvz_g_dct_16[k][i]; default: return kvz_g_dct_32[k][i]; }
}
int32_t kvzref_get_scaled_qp(int type, int qp, int off) { return kvz_get_scaled_qp((int8_t)type, (int8_t)qp, (int8_t)off); }

/* ------------------------------------------------------------------------- */
/* An encoder instance, only to obtain a correctly initialised encoder_state_t /
 * encoder_control_t (scaling lists, bitdepth) for quant / sao / ipol calls.   */
typedef struct {
  const kvz_api *api;
  kvz_config *cfg;
  kvz_encoder *enc;
} kvzref_ctx;

kvzref_ctx *kvzref_ctx_open(int width, int height, int qp, int signhide, int rdoq)
{
  kvzref_ctx *c = calloc(1, sizeof(*c));
  c->api = kvz_api_get(KVZ_BIT_DEPTH);
  c->cfg = c->api->config_alloc();
  c->api->config_init(c->cfg);
  c->cfg->width = width; c->cfg->height = height; c->cfg->qp = qp;
  c->cfg->threads = 0; c->cfg->owf = 0; c->cfg->wpp = 0;
  c->cfg->signhide_enable = signhide; c->cfg->rdoq_enable = rdoq; c->cfg->rdoq_skip = 0;
  c->cfg->hash = KVZ_HASH_NONE;
  c->cfg->enable_logging_output = 0;
  c->enc = c->api->encoder_open(c->cfg);
  if (!c->enc) { free(c); return NULL; }
  return c;
}
void kvzref_ctx_close(kvzref_ctx *c)
{
  if (!c) return;
  c->api->encoder_close(c->enc);
  c->api->config_destroy(c->cfg);
  free(c);
}
static encoder_state_t *ctx_state(kvzref_ctx *c, int qp, int intra_slice)
{
  encoder_state_t *st = &c->enc->states[0];
  st->qp = (int8_t)qp;
  st->frame->slicetype = intra_slice ? KVZ_SLICE_I : KVZ_SLICE_P;
  return st;
}

typedef void (quant_fn)(const encoder_state_t *, coeff_t *, coeff_t *, int32_t, int32_t, int8_t, int8_t, int8_t);
typedef void (dequant_fn)(const encoder_state_t *, coeff_t *, coeff_t *, int32_t, int32_t, int8_t, int8_t);
typedef int (qres_fn)(encoder_state_t *, const cu_info_t *, int, color_t, coeff_scan_order_t, int, int, int,
                      const kvz_pixel *, const kvz_pixel *, kvz_pixel *, coeff_t *, bool);

void kvzref_quant(kvzref_ctx *c, const char *impl, int qp, int intra_slice, coeff_t *coef, coeff_t *q_coef,
                  int w, int h, int type, int scan_idx, int block_type)
{
  ((quant_fn *)kvzref_find("quant", impl))(ctx_state(c, qp, intra_slice), coef, q_coef, w, h,
                                           (int8_t)type, (int8_t)scan_idx, (int8_t)block_type);
}
void kvzref_dequant(kvzref_ctx *c, const char *impl, int qp, coeff_t *q_coef, coeff_t *coef, int w, int h,
                    int type, int block_type)
{
  ((dequant_fn *)kvzref_find("dequant", impl))(ctx_state(c, qp, 1), q_coef, coef, w, h, (int8_t)type, (int8_t)block_type);
}
int kvzref_quantize_residual(kvzref_ctx *c, const char *impl, int qp, int intra_slice, int width, int color,
                             int scan_idx, int use_trskip, int cu_is_intra, int in_stride, int out_stride,
                             const kvz_pixel *ref_in, const kvz_pixel *pred_in, kvz_pixel *rec_out,
                             coeff_t *coeff_out, int early_skip)
{
  cu_info_t cu; memset(&cu, 0, sizeof(cu));
  cu.type = cu_is_intra ? CU_INTRA : CU_INTER;
  cu.part_size = SIZE_2Nx2N;
  return ((qres_fn *)kvzref_find("quantize_residual", impl))(ctx_state(c, qp, intra_slice), &cu, width, (color_t)color,
          (coeff_scan_order_t)scan_idx, use_trskip, in_stride, out_stride, ref_in, pred_in, rec_out, coeff_out,
          early_skip != 0);
}

/* ---- sao ---- */
void kvzref_calc_sao_edge_dir(kvzref_ctx *c, const char *impl, const kvz_pixel *orig, const kvz_pixel *rec,
                              int eo_class, int bw, int bh, int *cat_sum_cnt /* [2][5] */)
{
  ((calc_sao_edge_dir_func *)kvzref_find("calc_sao_edge_dir", impl))(c->enc->control, orig, rec, eo_class, bw, bh,
                                                                      (int (*)[NUM_SAO_EDGE_CATEGORIES])cat_sum_cnt);
}
int kvzref_sao_edge_ddistortion(kvzref_ctx *c, const char *impl, const kvz_pixel *orig, const kvz_pixel *rec,
                                int bw, int bh, int eo_class, int *offsets)
{
  return ((sao_edge_ddistortion_func *)kvzref_find("sao_edge_ddistortion", impl))(c->enc->control, orig, rec, bw, bh,
                                                                                    eo_class, offsets);
}
int kvzref_sao_band_ddistortion(kvzref_ctx *c, const char *impl, const kvz_pixel *orig, const kvz_pixel *rec,
                                int bw, int bh, int band_pos, const int *bands)
{
  return ((sao_band_ddistortion_func *)kvzref_find("sao_band_ddistortion", impl))(ctx_state(c, 22, 1), orig, rec, bw, bh,
                                                                                    band_pos, bands);
}
void kvzref_sao_reconstruct_color(kvzref_ctx *c, const char *impl, const kvz_pixel *rec, kvz_pixel *new_rec,
                                  int sao_type_i, int eo_class, const int *band_position, const int *offsets,
                                  int stride, int new_stride, int bw, int bh, int color)
{
  sao_info_t sao; memset(&sao, 0, sizeof(sao));
  sao.type = (sao_type)sao_type_i; sao.eo_class = (sao_eo_class)eo_class;
  sao.band_position[0] = band_position[0]; sao.band_position[1] = band_position[1];
  memcpy(sao.offsets, offsets, sizeof(sao.offsets));
  ((sao_reconstruct_color_func *)kvzref_find("sao_reconstruct_color", impl))(c->enc->control, rec, new_rec, &sao, stride,
                                                                              new_stride, bw, bh, (color_t)color);
}

/* ---- intra: the non-dispatched but inseparable part (src/intra.c) ---- */
void kvzref_intra_predict(int log2_width, int mode, int color, const kvz_pixel *top, const kvz_pixel *left,
                          kvz_pixel *dst, int filter_boundary)
{
  kvzref_init();
  kvz_intra_references refs; memset(&refs, 0, sizeof(refs));
  const int n = 2 * (1 << log2_width) + 1;
  memcpy(refs.ref.top, top, n * sizeof(kvz_pixel));
  memcpy(refs.ref.left, left, n * sizeof(kvz_pixel));
  refs.filtered_initialized = false;
  kvz_intra_predict(&refs, (int_fast8_t)log2_width, (int_fast8_t)mode, (color_t)color, dst, filter_boundary != 0);
}

/* Build an lcu_t view (rec + top/left border buffers) of a frame-level reconstruction
 * plane exactly the way init_lcu_t does (src/search.c:1077-1174), then call the
 * reference's kvz_intra_build_reference. */
void kvzref_intra_build_reference(int log2_width, int color, int luma_x, int luma_y, int pic_w, int pic_h,
                                  const kvz_pixel *plane, int stride, kvz_pixel *out_top, kvz_pixel *out_left)
{
  static __thread lcu_t lcu;
  const int is_c = color != 0;
  const int lw = 64 >> is_c;                         /* LCU width in this plane */
  const int pw = pic_w >> is_c, ph = pic_h >> is_c;  /* plane dimensions */
  const int ox = (luma_x / 64) * lw, oy = (luma_y / 64) * lw;
  kvz_pixel *rec = !color ? lcu.rec.y : (color == 1 ? lcu.rec.u : lcu.rec.v);
  kvz_pixel *top = !color ? lcu.top_ref.y : (color == 1 ? lcu.top_ref.u : lcu.top_ref.v);
  kvz_pixel *left = !color ? lcu.left_ref.y : (color == 1 ? lcu.left_ref.u : lcu.left_ref.v);
  memset(&lcu, 0, sizeof(lcu));
  #define PL(x, y) plane[MIN(MAX((y), 0), ph - 1) * stride + MIN(MAX((x), 0), pw - 1)]
  for (int y = 0; y < lw; ++y) for (int x = 0; x < lw; ++x) rec[y * lw + x] = PL(ox + x, oy + y);
  const int nref = (LCU_REF_PX_WIDTH >> is_c);
  for (int i = 0; i <= nref; ++i) { top[i] = PL(ox - 1 + i, oy - 1); left[i] = PL(ox - 1, oy - 1 + i); }
  #undef PL
  kvz_intra_references refs; memset(&refs, 0, sizeof(refs));
  vector2d_t lpx = { luma_x, luma_y }, ppx = { pic_w, pic_h };
  kvz_intra_build_reference((int_fast8_t)log2_width, (color_t)color, &lpx, &ppx, &lcu, &refs);
  const int n = 2 * (1 << log2_width) + 1;
  memcpy(out_top, refs.ref.top, n * sizeof(kvz_pixel));
  memcpy(out_left, refs.ref.left, n * sizeof(kvz_pixel));
}

/* ---- ipol ---- */
typedef void (smp_px_fn)(const encoder_control_t *, kvz_pixel *, int16_t, int, int, kvz_pixel *, int16_t, int8_t, int8_t, const int16_t[2]);
typedef void (smp_im_fn)(const encoder_control_t *, kvz_pixel *, int16_t, int, int, int16_t *, int16_t, int8_t, int8_t, const int16_t[2]);

void kvzref_sample(kvzref_ctx *c, const char *type, const char *impl, const kvz_pixel *src, int src_stride,
                   int w, int h, void *dst, int dst_stride, int mvx, int mvy)
{
  const int16_t mv[2] = { (int16_t)mvx, (int16_t)mvy };
  void *f = kvzref_find(type, impl);
  if (strstr(type, "_hi")) ((smp_im_fn *)f)(c->enc->control, (kvz_pixel *)src, (int16_t)src_stride, w, h, (int16_t *)dst, (int16_t)dst_stride, 0, 0, mv);
  else ((smp_px_fn *)f)(c->enc->control, (kvz_pixel *)src, (int16_t)src_stride, w, h, (kvz_pixel *)dst, (int16_t)dst_stride, 0, 0, mv);
}

int kvzref_ipol_im_size(void) { return KVZ_IPOL_MAX_IM_SIZE_LUMA_SIMD; }
int kvzref_ipol_first_cols(void) { return KVZ_EXT_BLOCK_W_LUMA + 1; }

void kvzref_filter_fme(kvzref_ctx *c, const char *impl, int stage, const kvz_pixel *src, int src_stride, int w, int h,
                       kvz_pixel *filtered, int16_t *hor_intermediate, int fme_level, int16_t *hor_first_cols,
                       int off_x, int off_y)
{
  static const char *names[4] = { "filter_hpel_blocks_hor_ver_luma", "filter_hpel_blocks_diag_luma",
                                  "filter_qpel_blocks_hor_ver_luma", "filter_qpel_blocks_diag_luma" };
  ipol_blocks_func *f = (ipol_blocks_func *)kvzref_find(names[stage], impl);
  f(c->enc->control, (kvz_pixel *)src, (int16_t)src_stride, w, h, (kvz_pixel (*)[LCU_LUMA_SIZE])filtered,
    (int16_t (*)[KVZ_IPOL_MAX_IM_SIZE_LUMA_SIMD])hor_intermediate, (int8_t)fme_level,
    (int16_t (*)[KVZ_EXT_BLOCK_W_LUMA + 1])hor_first_cols, (int8_t)off_x, (int8_t)off_y);
}

int kvzref_get_extended_block(const char *impl, const kvz_pixel *src, int src_w, int src_h, int src_s, int blk_x,
                              int blk_y, int blk_w, int blk_h, int pad_l, int pad_r, int pad_t, int pad_b,
                              int pad_b_simd, kvz_pixel *buf, int *ext_off, int *ext_origin_off, int *ext_s)
{
  kvz_pixel *ext = NULL, *ext_origin = NULL;
  kvz_epol_args a = { .src = (kvz_pixel *)src, .src_w = src_w, .src_h = src_h, .src_s = src_s, .blk_x = blk_x,
                      .blk_y = blk_y, .blk_w = blk_w, .blk_h = blk_h, .pad_l = pad_l, .pad_r = pad_r, .pad_t = pad_t,
                      .pad_b = pad_b, .pad_b_simd = pad_b_simd, .buf = buf, .ext = &ext, .ext_origin = &ext_origin,
                      .ext_s = ext_s };
  ((epol_func *)kvzref_find("get_extended_block", impl))(&a);
  int in_buf = (ext == buf);
  const kvz_pixel *base = in_buf ? buf : src;
  *ext_off = (int)(ext - base); *ext_origin_off = (int)(ext_origin - base);
  return in_buf;
}

/* ---- bipred_average: one call on an lcu_t, luma+chroma copied out ---- */
void kvzref_bipred_average(const char *impl, const kvz_pixel *px0_y, const kvz_pixel *px1_y,
                           const int16_t *im0_y, const int16_t *im1_y,
                           const kvz_pixel *px0_u, const kvz_pixel *px1_u, const int16_t *im0_u, const int16_t *im1_u,
                           const kvz_pixel *px0_v, const kvz_pixel *px1_v, const int16_t *im0_v, const int16_t *im1_v,
                           int pu_x, int pu_y, int pu_w, int pu_h, int flags0, int flags1,
                           kvz_pixel *out_y /*64*64*/, kvz_pixel *out_u /*32*32*/, kvz_pixel *out_v)
{
  static __thread lcu_t lcu;
  memset(&lcu, 0, sizeof(lcu));
  yuv_t p0 = { pu_w * pu_h, (kvz_pixel *)px0_y, (kvz_pixel *)px0_u, (kvz_pixel *)px0_v };
  yuv_t p1 = { pu_w * pu_h, (kvz_pixel *)px1_y, (kvz_pixel *)px1_u, (kvz_pixel *)px1_v };
  yuv_im_t i0 = { pu_w * pu_h, (kvz_pixel_im *)im0_y, (kvz_pixel_im *)im0_u, (kvz_pixel_im *)im0_v };
  yuv_im_t i1 = { pu_w * pu_h, (kvz_pixel_im *)im1_y, (kvz_pixel_im *)im1_u, (kvz_pixel_im *)im1_v };
  ((inter_recon_bipred_func *)kvzref_find("bipred_average", impl))(&lcu, &p0, &p1, &i0, &i1, pu_x, pu_y, pu_w, pu_h,
                                                                  flags0, flags1, true, true);
  memcpy(out_y, lcu.rec.y, sizeof(lcu.rec.y));
  memcpy(out_u, lcu.rec.u, sizeof(lcu.rec.u));
  memcpy(out_v, lcu.rec.v, sizeof(lcu.rec.v));
}

/* ---- nal ---- */
typedef void (cksum_fn)(const kvz_pixel *, const int, const int, const int, unsigned char[SEI_HASH_MAX_LENGTH], const uint8_t);
void kvzref_array_checksum(const char *impl, const kvz_pixel *data, int height, int width, int stride, unsigned char *out)
{
  unsigned char tmp[SEI_HASH_MAX_LENGTH] = { 0 };
  ((cksum_fn *)kvzref_find("array_checksum", impl))(data, height, width, stride, tmp, KVZ_BIT_DEPTH);
  memcpy(out, tmp, 4);
}

/* ---- deblocking (not a strategy: kvz_filter_deblock_lcu, filter.c:783) ---- */
#include "filter.h"
int kvzref_sizeof_cu_info(void) { return (int)sizeof(cu_info_t); }
/* build one cu_info_t through the reference's own bitfields, for pinning the 20-byte record layout */
void kvzref_make_cu_info(int type, int depth, int part_size, int tr_depth, int cbf, int qp, int mv_dir,
                         const int16_t mv[4], const uint8_t mv_ref[2], uint8_t *out)
{
  cu_info_t cu; memset(&cu, 0, sizeof(cu));
  cu.type = type; cu.depth = depth; cu.part_size = part_size; cu.tr_depth = tr_depth; cu.cbf = (uint16_t)cbf; cu.qp = (uint8_t)qp;
  if (type != CU_INTRA) {
    cu.inter.mv[0][0] = mv[0]; cu.inter.mv[0][1] = mv[1]; cu.inter.mv[1][0] = mv[2]; cu.inter.mv[1][1] = mv[3];
    cu.inter.mv_ref[0] = mv_ref[0]; cu.inter.mv_ref[1] = mv_ref[1]; cu.inter.mv_dir = mv_dir;
  }
  memcpy(out, &cu, sizeof(cu));
}
int kvzref_deblock_frame(kvzref_ctx *c, kvz_pixel *y, kvz_pixel *u, kvz_pixel *v, const uint8_t *cus, int cu_stride_scu,
                         int qp, int beta_offset_div2, int tc_offset_div2, int slice_type, int per_cu_qp, const uint8_t *ref_LX)
{
  encoder_state_t *st = &c->enc->states[0];
  videoframe_t *frame = st->tile->frame;
  encoder_control_t *ctrl = (encoder_control_t *)st->encoder_control;
  const int W = frame->width, H = frame->height;
  if (!frame->rec) frame->rec = kvz_image_alloc(KVZ_CSP_420, W, H);
  if (!frame->cu_array) frame->cu_array = kvz_cu_array_alloc(W, H);
  if (!frame->rec || !frame->cu_array) return -1;
  cu_array_t *cua = frame->cu_array;
  const int scu_w = cua->stride / 4, scu_h = cua->height / 4;
  for (int r = 0; r < scu_h && r * 4 < ((H + 3) & ~3); ++r)
    memcpy(&cua->data[(size_t)r * scu_w], cus + (size_t)r * cu_stride_scu * sizeof(cu_info_t), (size_t)((W + 3) / 4) * sizeof(cu_info_t));
  kvz_picture *rec = frame->rec;
  for (int r = 0; r < H; ++r) memcpy(rec->y + (size_t)r * rec->stride, y + (size_t)r * W, (size_t)W * sizeof(kvz_pixel));
  for (int r = 0; r < H / 2; ++r) {
    memcpy(rec->u + (size_t)r * (rec->stride / 2), u + (size_t)r * (W / 2), (size_t)(W / 2) * sizeof(kvz_pixel));
    memcpy(rec->v + (size_t)r * (rec->stride / 2), v + (size_t)r * (W / 2), (size_t)(W / 2) * sizeof(kvz_pixel));
  }
  st->qp = (int8_t)qp;
  st->frame->slicetype = (enum kvz_slice_type)slice_type;
  st->frame->max_qp_delta_depth = per_cu_qp ? 0 : -1;
  if (ref_LX) memcpy(st->frame->ref_LX, ref_LX, 32);
  ctrl->cfg.deblock_beta = (int8_t)beta_offset_div2; ctrl->cfg.deblock_tc = (int8_t)tc_offset_div2;
  for (int ly = 0; ly < H; ly += LCU_WIDTH) for (int lx = 0; lx < W; lx += LCU_WIDTH) kvz_filter_deblock_lcu(st, lx, ly);
  for (int r = 0; r < H; ++r) memcpy(y + (size_t)r * W, rec->y + (size_t)r * rec->stride, (size_t)W * sizeof(kvz_pixel));
  for (int r = 0; r < H / 2; ++r) {
    memcpy(u + (size_t)r * (W / 2), rec->u + (size_t)r * (rec->stride / 2), (size_t)(W / 2) * sizeof(kvz_pixel));
    memcpy(v + (size_t)r * (W / 2), rec->v + (size_t)r * (rec->stride / 2), (size_t)(W / 2) * sizeof(kvz_pixel));
  }
  st->frame->max_qp_delta_depth = -1;
  return 0;
}

/* ---- RDOQ (not a strategy: kvz_rdoq, rdo.c:661) ---- */
#include "rdo.h"
#include "context.h"
int kvzref_cabac_ctx_size(void) { return (int)sizeof(((cabac_data_t *)0)->ctx); }
/* the context models after kvz_init_contexts for (qp, slice_type) -> out[kvzref_cabac_ctx_size()] */
void kvzref_init_contexts(kvzref_ctx *c, int qp, int slice_type, uint8_t *out)
{
  encoder_state_t *st = &c->enc->states[0];
  kvz_init_contexts(st, (int8_t)qp, (int8_t)slice_type);
  memcpy(out, &st->cabac.ctx, sizeof(st->cabac.ctx));
}
/* byte offsets of the members RDOQ reads, for pinning kvz_cuda_cabac_ctx */
void kvzref_cabac_ctx_offsets(int32_t *out)
{
  typedef cabac_data_t T;
  const size_t base = offsetof(T, ctx);
  int i = 0;
  out[i++] = (int32_t)(offsetof(T, ctx.qt_cbf_model_luma) - base);
  out[i++] = (int32_t)(offsetof(T, ctx.qt_cbf_model_chroma) - base);
  out[i++] = (int32_t)(offsetof(T, ctx.cu_sig_coeff_group_model) - base);
  out[i++] = (int32_t)(offsetof(T, ctx.cu_sig_model_luma) - base);
  out[i++] = (int32_t)(offsetof(T, ctx.cu_sig_model_chroma) - base);
  out[i++] = (int32_t)(offsetof(T, ctx.cu_ctx_last_y_luma) - base);
  out[i++] = (int32_t)(offsetof(T, ctx.cu_ctx_last_y_chroma) - base);
  out[i++] = (int32_t)(offsetof(T, ctx.cu_ctx_last_x_luma) - base);
  out[i++] = (int32_t)(offsetof(T, ctx.cu_ctx_last_x_chroma) - base);
  out[i++] = (int32_t)(offsetof(T, ctx.cu_one_model_luma) - base);
  out[i++] = (int32_t)(offsetof(T, ctx.cu_one_model_chroma) - base);
  out[i++] = (int32_t)(offsetof(T, ctx.cu_abs_model_luma) - base);
  out[i++] = (int32_t)(offsetof(T, ctx.cu_abs_model_chroma) - base);
  out[i++] = (int32_t)(offsetof(T, ctx.cu_qt_root_cbf_model) - base);
}
/* kvz_rdoq on one TU with the given context models, lambda and qp.  ctx must be opened with the wanted signhide. */
void kvzref_rdoq(kvzref_ctx *c, int qp, double lambda, const uint8_t *cabac_ctx, coeff_t *coef, coeff_t *dest, int width,
                 int type, int scan_mode, int block_type, int tr_depth)
{
  encoder_state_t *st = &c->enc->states[0];
  st->qp = (int8_t)qp;
  st->lambda = lambda;
  memcpy(&st->cabac.ctx, cabac_ctx, sizeof(st->cabac.ctx));
  kvz_rdoq(st, coef, dest, width, width, (int8_t)type, (int8_t)scan_mode, (int8_t)block_type, (int8_t)tr_depth);
}

/* ---- coefficient bit cost: kvz_encode_coeff_nxn in only_count mode (what get_coeff_cabac_cost runs, rdo.c:223-264) ---- */
typedef void (encode_coeff_fn)(encoder_state_t *, cabac_data_t *, const coeff_t *, uint8_t, uint8_t, int8_t, int8_t, double *);
double kvzref_coeff_cost(kvzref_ctx *c, const char *impl, const uint8_t *cabac_ctx, int update, int trskip_enable, const coeff_t *coeff,
                         int width, int type, int scan_mode, int tr_skip, uint8_t *ctx_after)
{
  encoder_state_t *st = &c->enc->states[0];
  encoder_control_t *ctrl = (encoder_control_t *)st->encoder_control;
  ctrl->cfg.trskip_enable = trskip_enable;
  int found = 0;
  for (int i = 0; i < width * width; ++i) if (coeff[i]) { found = 1; break; }
  if (!found) return 0;                                     /* rdo.c:231-238 */
  cabac_data_t cabac_copy;
  memcpy(&cabac_copy, &st->search_cabac, sizeof(cabac_copy));
  memcpy(&cabac_copy.ctx, cabac_ctx, sizeof(cabac_copy.ctx));
  cabac_copy.only_count = 1;
  cabac_copy.update = update ? 1 : 0;
  double bits = 0;
  ((encode_coeff_fn *)kvzref_find("encode_coeff_nxn", impl))(st, &cabac_copy, coeff, (uint8_t)width, (uint8_t)type, (int8_t)scan_mode,
                                                            (int8_t)tr_skip, &bits);
  if (ctx_after) memcpy(ctx_after, &cabac_copy.ctx, sizeof(cabac_copy.ctx));
  return bits;
}

/* cfg.trskip_enable of an opened context (the frame pass tries transform skip on 4x4 luma TUs when set) */
void kvzref_set_trskip(kvzref_ctx *c, int enable)
{
  ((encoder_control_t *)c->enc->states[0].encoder_control)->cfg.trskip_enable = enable;
}

/* ---- intra mode signalling cost (not strategies): MPM derivation and the mode-bit estimates of the rough search ---- */
#include "intra.h"
#include "search_intra.h"
void kvzref_intra_mpm(int left_mode, int above_mode, int y, int8_t *preds)
{
  cu_info_t left, above, cur; memset(&left, 0, sizeof(left)); memset(&above, 0, sizeof(above)); memset(&cur, 0, sizeof(cur));
  left.type = CU_INTRA; left.intra.mode = (int8_t)left_mode;
  above.type = CU_INTRA; above.intra.mode = (int8_t)above_mode;
  kvz_intra_get_dir_luma_predictor(0, (uint32_t)y, preds, &cur, left_mode >= 0 ? &left : NULL, above_mode >= 0 ? &above : NULL);
}
double kvzref_luma_mode_bits(kvzref_ctx *c, const uint8_t *cabac_ctx, int luma_mode, const int8_t *preds)
{
  encoder_state_t *st = &c->enc->states[0];
  memcpy(&st->search_cabac.ctx, cabac_ctx, sizeof(st->search_cabac.ctx));
  st->search_cabac.only_count = 1; st->search_cabac.update = 0;
  return kvz_luma_mode_bits(st, (int8_t)luma_mode, preds);
}
double kvzref_chroma_mode_bits(kvzref_ctx *c, const uint8_t *cabac_ctx, int chroma_mode, int luma_mode)
{
  encoder_state_t *st = &c->enc->states[0];
  memcpy(&st->search_cabac.ctx, cabac_ctx, sizeof(st->search_cabac.ctx));
  st->search_cabac.only_count = 1; st->search_cabac.update = 0;
  return kvz_chroma_mode_bits(st, (int8_t)chroma_mode, (int8_t)luma_mode);
}
