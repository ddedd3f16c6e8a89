/*
 * kvz_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's generic strategy kernels
 * (/root/reference/src/strategies/generic/<group>-generic.c) used as the parity checker for the
 * CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (libkvzcuda.so) never
 * links or calls it.
 *
 * Parity pinning: every function here is checked (tests/test_oracle_*.py) against
 *   (1) the golden constants of the reference's own unit tests
 *       (tests/satd_tests.c:122,140,159, tests/sad_tests.c, tests/intra_sad_tests.c,
 *        tests/coeff_sum_tests.c), and
 *   (2) the reference itself, compiled unmodified into oracle/_ref/ and called
 *       through oracle/ref_shim.c, on seeded random + extreme inputs, and
 *   (3) fixtures generated from (2) and committed under tests/golden/.
 *
 * Like the reference, the pixel type is a compile-time choice: build with
 * -DORC_BITDEPTH=8 (default, orc_pix = uint8_t) or -DORC_BITDEPTH=10 (uint16_t).
 */
#ifndef KVZ_ORACLE_H_
#define KVZ_ORACLE_H_

#include <stdint.h>
#include <stddef.h>

#ifndef ORC_BITDEPTH
#define ORC_BITDEPTH 8
#endif

#if ORC_BITDEPTH == 8
typedef uint8_t orc_pix;
#else
typedef uint16_t orc_pix;
#endif
#define ORC_PIXEL_MAX ((1 << ORC_BITDEPTH) - 1)

#ifdef __cplusplus
extern "C" {
#endif

int orc_bitdepth(void);

/* ---- picture group (picture-generic.c) ---- */
unsigned orc_reg_sad(const orc_pix *a, const orc_pix *b, int w, int h, unsigned s1, unsigned s2);
unsigned orc_sad_nxn(int n, const orc_pix *a, const orc_pix *b);
void     orc_sad_nxn_dual(int n, const orc_pix *preds /* [2][32*32] */, const orc_pix *orig, unsigned *costs);
unsigned orc_satd_nxn(int n, const orc_pix *a, const orc_pix *b);
void     orc_satd_nxn_dual(int n, const orc_pix *preds /* [2][32*32] */, const orc_pix *orig, unsigned *costs);
unsigned orc_satd_any_size(int w, int h, const orc_pix *b1, int s1, const orc_pix *b2, int s2);
void     orc_satd_any_size_quad(int w, int h, const orc_pix *const preds[4], int stride,
                                const orc_pix *orig, int orig_stride, unsigned num_modes,
                                unsigned *costs, int8_t *valid);
unsigned orc_pixels_calc_ssd(const orc_pix *ref, const orc_pix *rec, int ref_stride, int rec_stride, int width);
uint32_t orc_ver_sad(const orc_pix *pic, const orc_pix *ref, int w, int h, uint32_t pic_stride);
uint32_t orc_hor_sad(const orc_pix *pic, const orc_pix *ref, int w, int h, uint32_t pic_stride,
                     uint32_t ref_stride, uint32_t left, uint32_t right);
/* one plane of bipred_average: flags bit0 = L0 is 14-bit intermediate, bit1 = L1 is */
void     orc_bipred_average_plane(orc_pix *dst, unsigned dst_stride, const void *l0, const void *l1,
                                  int l0_is_im, int l1_is_im, unsigned w, unsigned h);
double   orc_pixel_var(const orc_pix *buf, uint32_t len);

/* ---- dct group (dct-generic.c) ---- */
void orc_dct_nxn(int n, int bitdepth, const int16_t *in, int16_t *out);
void orc_idct_nxn(int n, int bitdepth, const int16_t *in, int16_t *out);
void orc_dst_4x4(int bitdepth, const int16_t *in, int16_t *out);
void orc_idst_4x4(int bitdepth, const int16_t *in, int16_t *out);

/* ---- quant group (quant-generic.c); encoder_state_t flattened into plain params ---- */
typedef struct {
  int32_t qp;              /* state->qp */
  int32_t bitdepth;        /* encoder->bitdepth */
  int32_t slice_is_intra;  /* state->frame->slicetype == KVZ_SLICE_I */
  int32_t signhide_enable; /* encoder->cfg.signhide_enable */
  /* flat scaling list only (scaling_list.enable == 0): quant_coeff = quant_scales[qp%6] */
} orc_quant_params;

int32_t  orc_get_scaled_qp(int type, int qp, int qp_offset);
const uint32_t *orc_scan_table(int scan_idx, int log2_size); /* log2_size 1..5 */
void orc_quant(const orc_quant_params *p, const int16_t *coef, int16_t *q_coef, int w, int h,
               int type, int scan_idx, int block_type);
void orc_dequant(const orc_quant_params *p, const int16_t *q_coef, int16_t *coef, int w, int h,
                 int type, int block_type);
/* quantize_residual, non-RDOQ branch.  color: 0 Y, 1 U, 2 V.  luma_intra_4x4_dst: use DST when
 * width==4, color==Y, intra.  Returns has_coeffs. */
int  orc_quantize_residual(const orc_quant_params *p, int width, int color, int scan_idx,
                           int use_trskip, int cu_is_intra, int in_stride, int out_stride,
                           const orc_pix *ref_in, const orc_pix *pred_in, orc_pix *rec_out,
                           int16_t *coeff_out, int early_skip);
uint32_t orc_coeff_abs_sum(const int16_t *coeffs, size_t length);
double   orc_fast_coeff_cost(const int16_t *coeff, int32_t width, uint64_t weights);

/* ---- intra group (intra-generic.c + the inseparable part of intra.c) ---- */
void orc_angular_pred(int log2_width, int mode, const orc_pix *ref_top, const orc_pix *ref_left, orc_pix *dst);
void orc_intra_pred_planar(int log2_width, const orc_pix *ref_top, const orc_pix *ref_left, orc_pix *dst);
void orc_intra_pred_filtered_dc(int log2_width, const orc_pix *ref_top, const orc_pix *ref_left, orc_pix *dst);
/* kvz_intra_predict: refs are [2*width+1], index 0 = corner */
void orc_intra_predict(int log2_width, int mode, int color, const orc_pix *ref_top, const orc_pix *ref_left,
                       orc_pix *dst, int filter_boundary);
/* kvz_intra_build_reference restated over a frame-level reconstruction plane */
void orc_intra_build_reference(int log2_width, int color, int luma_x, int luma_y, int pic_w, int pic_h,
                               const orc_pix *rec_plane, int rec_stride,
                               orc_pix *out_top /* [2w+1] */, orc_pix *out_left /* [2w+1] */);

/* ---- ipol group (ipol-generic.c) ---- */
void orc_sample_quarterpel_luma(const orc_pix *src, int src_stride, int w, int h, orc_pix *dst, int dst_stride,
                                int mvx, int mvy);
void orc_sample_quarterpel_luma_hi(const orc_pix *src, int src_stride, int w, int h, int16_t *dst, int dst_stride,
                                   int mvx, int mvy);
void orc_sample_octpel_chroma(const orc_pix *src, int src_stride, int w, int h, orc_pix *dst, int dst_stride,
                              int mvx, int mvy);
void orc_sample_octpel_chroma_hi(const orc_pix *src, int src_stride, int w, int h, int16_t *dst, int dst_stride,
                                 int mvx, int mvy);
#define ORC_LCU_W 64
#define ORC_EXT_BLOCK_W_LUMA 71                          /* LCU_WIDTH + 7 */
#define ORC_IPOL_IM_SIZE ((ORC_EXT_BLOCK_W_LUMA + 1) * ORC_LCU_W + 1)
#define ORC_FIRST_COLS   (ORC_EXT_BLOCK_W_LUMA + 1)
/* stage: 0 hpel hor/ver, 1 hpel diag, 2 qpel hor/ver, 3 qpel diag.
 * filtered: [4][64*64]; hor_intermediate: [5][ORC_IPOL_IM_SIZE]; hor_first_cols: [5][ORC_FIRST_COLS] */
void orc_filter_fme(int stage, const orc_pix *src, int src_stride, int w, int h, orc_pix *filtered,
                    int16_t *hor_intermediate, int fme_level, int16_t *hor_first_cols,
                    int hpel_off_x, int hpel_off_y);
/* get_extended_block: returns 1 if the border-replicated copy in buf was built (ext=buf), else 0 */
int orc_get_extended_block(const orc_pix *src, int src_w, int src_h, int src_s, int blk_x, int blk_y,
                           int blk_w, int blk_h, int pad_l, int pad_r, int pad_t, int pad_b, int pad_b_simd,
                           orc_pix *buf, int *ext_off /* offset of ext from src or buf */,
                           int *ext_origin_off, int *ext_s);

/* ---- sao group (sao-generic.c, sao_shared_generics.h) ---- */
void orc_calc_sao_edge_dir(int bitdepth, const orc_pix *orig, const orc_pix *rec, int eo_class,
                           int block_w, int block_h, int cat_sum_cnt[2][5]);
int  orc_sao_edge_ddistortion(int bitdepth, const orc_pix *orig, const orc_pix *rec, int block_w, int block_h,
                              int eo_class, const int offsets[5]);
int  orc_sao_band_ddistortion(int bitdepth, const orc_pix *orig, const orc_pix *rec, int block_w, int block_h,
                              int band_pos, const int sao_bands[4]);
/* sao_type: 1 band, 2 edge.  offsets[10], band_position[2], color 0/1/2 */
void orc_sao_reconstruct_color(int bitdepth, const orc_pix *rec, orc_pix *new_rec, int sao_type, int eo_class,
                               const int band_position[2], const int offsets[10], int stride, int new_stride,
                               int block_w, int block_h, int color);

/* ---- nal group (nal-generic.c) ---- */
void orc_array_checksum(const orc_pix *data, int height, int width, int stride, unsigned char out[4]);

/* ---- RDOQ (kvz_rdoq, src/rdo.c:661-977) ---- */
typedef struct { double lambda; int32_t qp, bitdepth, signhide_enable, pad; } orc_rdoq_params;
/* cabac: image of cabac_data_t.ctx (184 state bytes); coef -> q, width x width; type 0 luma / 2 chroma */
void orc_rdoq(const orc_rdoq_params *p, const uint8_t *cabac, const int16_t *coef, int16_t *q, int width, int type, int scan_idx,
              int block_type, int tr_depth);

/* ---- intra mode signalling cost (intra.c:84-127, search_intra.c:641-698) ---- */
void orc_intra_mpm(int left_mode, int above_mode, int y, int8_t preds[3]);
double orc_luma_mode_bits(const uint8_t *cabac, int luma_mode, const int8_t preds[3]);
double orc_chroma_mode_bits(const uint8_t *cabac, int chroma_mode, int luma_mode);

/* ---- deblocking, frame level (filter.c:95-792) ---- */
typedef struct {
  int32_t width, height;          /* luma size, multiples of 8 */
  int32_t qp;                     /* state->qp, used when per_cu_qp == 0 (max_qp_delta_depth < 0) */
  int32_t beta_offset_div2, tc_offset_div2;
  int32_t slice_is_b;             /* frame->slicetype == KVZ_SLICE_B */
  int32_t per_cu_qp;
  int32_t cu_stride_scu;          /* cu_array stride in 4x4 SCUs */
  uint8_t ref_LX[2][16];          /* frame->ref_LX (encoderstate.h:125) */
} orc_dbk_params;
/* cus: 20-byte records per SCU in the reference's cu_info_t memory layout; planes are filtered in place */
void orc_deblock_frame(const orc_dbk_params *p, orc_pix *y, orc_pix *u, orc_pix *v, const uint8_t *cus);

#ifdef __cplusplus
}
#endif
#endif
