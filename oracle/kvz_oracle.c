/*
 * kvz_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see kvz_oracle.h).
 *
 * CPU restatement of the reference's generic strategies.  Written from the
 * behaviour of /root/reference/src/strategies/generic/<group>-generic.c (cited per function as
 * ref:<file>:<lines>), not copied from it: transforms are written as the matrix
 * products the butterflies compute, Hadamards as generic-N loops, tables are
 * generated from their defining rules.  tests/test_oracle_*.py pin every function
 * against the reference's golden constants and against the compiled reference.
 */
#include "kvz_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORC_MIN(a, b) ((a) < (b) ? (a) : (b))
#define ORC_MAX(a, b) ((a) > (b) ? (a) : (b))
#define ORC_CLIP(lo, hi, v) ((v) < (lo) ? (lo) : ((v) > (hi) ? (hi) : (v)))

int orc_bitdepth(void) { return ORC_BITDEPTH; }

/* ref:picture-generic.c:40-82 -- the "fast clip" bit tricks equal a plain clamp to
 * [0, PIXEL_MAX] for every int16/int32 input except that the 16-bit variant negates
 * in int (so -32768 clamps to 0 as expected). A plain clamp is the restatement. */
static inline orc_pix clip_pix(int v) { return (orc_pix)ORC_CLIP(0, ORC_PIXEL_MAX, v); }

/* ======================================================================= */
/* picture group                                                           */
/* ======================================================================= */

/* ref:picture-generic.c:98-111 -- strided SAD, no bit-depth shift */
unsigned orc_reg_sad(const orc_pix *a, const orc_pix *b, int w, int h, unsigned s1, unsigned s2)
{
  unsigned sad = 0;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x)
      sad += (unsigned)abs((int)a[y * s1 + x] - (int)b[y * s2 + x]);
  return sad;
}

/* ref:picture-generic.c:475-501 -- contiguous NxN SAD, >> (bitdepth-8) */
unsigned orc_sad_nxn(int n, const orc_pix *a, const orc_pix *b)
{
  unsigned sum = 0;
  for (int i = 0; i < n * n; ++i) sum += (unsigned)abs((int)a[i] - (int)b[i]);
  return sum >> (ORC_BITDEPTH - 8);
}

/* ref:picture-generic.c:512-534 -- preds is kvz_pixel[2][32*32]; only [0],[1] used */
void orc_sad_nxn_dual(int n, const orc_pix *preds, const orc_pix *orig, unsigned *costs)
{
  /* NOTE: pred_buffer rows are 32*32 pixels apart regardless of n (strategies-picture.h:48).
   * For n == 64 the reference indexes past one row into the next; keep that layout. */
  costs[0] = orc_sad_nxn(n, preds, orig);
  costs[1] = orc_sad_nxn(n, preds + 32 * 32, orig);
}

/* In-place 1-D Walsh-Hadamard butterflies over `len` values with stride `st`.
 * The reference (picture-generic.c:117-208, 252-340) hard-codes one particular
 * butterfly ordering whose output is a permutation (with identical magnitudes) of
 * the natural-ordered transform; SATD only sums |coeff|, so ordering is irrelevant. */
static void wht_1d(int32_t *v, int len, int st)
{
  for (int half = 1; half < len; half <<= 1)
    for (int base = 0; base < len; base += 2 * half)
      for (int k = 0; k < half; ++k) {
        int32_t p = v[(base + k) * st], q = v[(base + k + half) * st];
        v[(base + k) * st] = p + q;
        v[(base + k + half) * st] = p - q;
      }
}

static unsigned hadamard_abs_sum(int32_t *d, int n)
{
  for (int r = 0; r < n; ++r) wht_1d(d + r * n, n, 1);
  for (int c = 0; c < n; ++c) wht_1d(d + c, n, n);
  unsigned s = 0;
  for (int i = 0; i < n * n; ++i) s += (unsigned)abs(d[i]);
  return s;
}

/* ref:picture-generic.c:117-208,213-236 -- 4x4: (sum+1)>>1 */
static unsigned satd4_sub(const orc_pix *a, int sa, const orc_pix *b, int sb)
{
  int32_t d[16];
  for (int y = 0; y < 4; ++y)
    for (int x = 0; x < 4; ++x) d[y * 4 + x] = (int)a[y * sa + x] - (int)b[y * sb + x];
  return (hadamard_abs_sum(d, 4) + 1) >> 1;
}

/* ref:picture-generic.c:252-340 -- 8x8: (sum+2)>>2 */
static unsigned satd8_sub(const orc_pix *a, int sa, const orc_pix *b, int sb)
{
  int32_t d[64];
  for (int y = 0; y < 8; ++y)
    for (int x = 0; x < 8; ++x) d[y * 8 + x] = (int)a[y * sa + x] - (int)b[y * sb + x];
  return (hadamard_abs_sum(d, 8) + 2) >> 2;
}

/* ref:picture-generic.c:213-221 (4x4, no bit-depth shift!) and
 * strategies-picture.h:53-69 (N>=8: sum of 8x8 sub-block SATDs, >> (bitdepth-8)) */
unsigned orc_satd_nxn(int n, const orc_pix *a, const orc_pix *b)
{
  if (n == 4) return satd4_sub(a, 4, b, 4);
  unsigned sum = 0;
  for (int y = 0; y < n; y += 8)
    for (int x = 0; x < n; x += 8) sum += satd8_sub(a + y * n + x, n, b + y * n + x, n);
  return sum >> (ORC_BITDEPTH - 8);
}

/* ref:picture-generic.c:363-402 */
void orc_satd_nxn_dual(int n, const orc_pix *preds, const orc_pix *orig, unsigned *costs)
{
  /* 4x4: satd_4x4_generic(orig, preds[k]); N>=8: satd_8x8_subblock(preds[k], orig).
   * |H(a-b)| == |H(b-a)| so argument order does not matter. */
  costs[0] = orc_satd_nxn(n, preds, orig);
  costs[1] = orc_satd_nxn(n, preds + 32 * 32, orig);
}

/* ref:strategies-picture.h:75-113 (SATD_ANY_SIZE) */
unsigned orc_satd_any_size(int w, int h, const orc_pix *b1, int s1, const orc_pix *b2, int s2)
{
  unsigned sum = 0;
  if (w % 8 != 0) {               /* first 4-wide column as 4x4 blocks */
    for (int y = 0; y < h; y += 4) sum += satd4_sub(b1 + y * s1, s1, b2 + y * s2, s2);
    b1 += 4; b2 += 4; w -= 4;
  }
  if (h % 8 != 0) {               /* first 4-high row as 4x4 blocks */
    for (int x = 0; x < w; x += 4) sum += satd4_sub(b1 + x, s1, b2 + x, s2);
    b1 += 4 * s1; b2 += 4 * s2; h -= 4;
  }
  for (int y = 0; y < h; y += 8)
    for (int x = 0; x < w; x += 8) sum += satd8_sub(b1 + y * s1 + x, s1, b2 + y * s2 + x, s2);
  return sum >> (ORC_BITDEPTH - 8);
}

/* ref:picture-generic.c:404-471 (SATD_ANY_SIZE_MULTI_GENERIC) -- including the quirk that
 * when height % 8 == 4 the 8x8 pass restarts from row 0 of the ORIGINAL pointers
 * (rows 0..3 counted twice, the last 4 rows never), see SURVEY.md H5. */
void orc_satd_any_size_quad(int w, int h, const orc_pix *const preds[4], int stride,
                            const orc_pix *orig, int orig_stride, unsigned num_modes,
                            unsigned *costs, int8_t *valid)
{
  (void)num_modes; (void)valid;
  const int wmod = w % 8;
  for (int k = 0; k < 4; ++k) costs[k] = 0;
  if (wmod != 0) {
    for (int y = 0; y < h; y += 4)
      for (int k = 0; k < 4; ++k)
        costs[k] += satd4_sub(orig + y * orig_stride, orig_stride, preds[k] + y * stride, stride);
    w -= 4;
  }
  if (h % 8 != 0) {
    /* first row of 4x4 blocks: starts at column 0 of preds AND orig, runs over the
     * (already reduced) width */
    for (int x = 0; x < w; x += 4)
      for (int k = 0; k < 4; ++k)
        costs[k] += satd4_sub(orig + x, orig_stride, preds[k] + x, stride);
    h -= 4;
  }
  for (int y = h % 8; y < h; y += 8)
    for (int x = wmod; x < w; x += 8)
      for (int k = 0; k < 4; ++k)
        costs[k] += satd8_sub(orig + y * orig_stride + x, orig_stride, preds[k] + y * stride + x, stride);
  for (int k = 0; k < 4; ++k) costs[k] >>= (ORC_BITDEPTH - 8);
}

/* ref:picture-generic.c:536-551 */
unsigned orc_pixels_calc_ssd(const orc_pix *ref, const orc_pix *rec, int ref_stride, int rec_stride, int width)
{
  int ssd = 0;
  for (int y = 0; y < width; ++y)
    for (int x = 0; x < width; ++x) {
      int d = (int)ref[x + y * ref_stride] - (int)rec[x + y * rec_stride];
      ssd += d * d;
    }
  return (unsigned)(ssd >> (2 * (ORC_BITDEPTH - 8)));
}

/* ref:picture-generic.c:687-701 */
uint32_t orc_ver_sad(const orc_pix *pic, const orc_pix *ref, int w, int h, uint32_t pic_stride)
{
  uint32_t sad = 0;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) sad += (uint32_t)abs((int)pic[y * pic_stride + x] - (int)ref[x]);
  return sad;
}

static uint32_t hor_sad_col(const orc_pix *pic, const orc_pix *ref, int w, int h, uint32_t ps, uint32_t rs)
{
  uint32_t sad = 0;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) sad += (uint32_t)abs((int)pic[y * ps + x] - (int)ref[y * rs]);
  return sad;
}

/* ref:picture-generic.c:714-752 */
uint32_t orc_hor_sad(const orc_pix *pic, const orc_pix *ref, int w, int h, uint32_t ps,
                     uint32_t rs, uint32_t left, uint32_t right)
{
  uint32_t r = 0;
  if (left) {
    r += hor_sad_col(pic, ref + left, (int)left, h, ps, rs);
    r += orc_reg_sad(pic + left, ref + left, w - (int)left, h, ps, rs);
  } else if (right) {
    r += orc_reg_sad(pic, ref, w - (int)right, h, ps, rs);
    r += hor_sad_col(pic + w - right, ref + w - right - 1, (int)right, h, ps, rs);
  } else {
    r += orc_reg_sad(pic, ref, w, h, ps, rs);
  }
  return r;
}

/* ref:picture-generic.c:553-668 -- one plane; inputs are contiguous pu_w*pu_h */
void orc_bipred_average_plane(orc_pix *dst, unsigned dst_stride, const void *l0, const void *l1,
                              int l0_is_im, int l1_is_im, unsigned w, unsigned h)
{
  const int shift = 15 - ORC_BITDEPTH;
  const int offset = 1 << (shift - 1);
  for (unsigned i = 0; i < w * h; ++i) {
    int16_t s0 = l0_is_im ? ((const int16_t *)l0)[i] : (int16_t)(((const orc_pix *)l0)[i] << (14 - ORC_BITDEPTH));
    int16_t s1 = l1_is_im ? ((const int16_t *)l1)[i] : (int16_t)(((const orc_pix *)l1)[i] << (14 - ORC_BITDEPTH));
    int32_t r = ((int32_t)s0 + (int32_t)s1 + offset) >> shift;
    dst[(i / w) * dst_stride + (i % w)] = clip_pix(r);
  }
}

/* ref:picture-generic.c:755-778 */
double orc_pixel_var(const orc_pix *arr, uint32_t len)
{
  double sum = 0, var = 0;
  for (uint32_t i = 0; i < len; ++i) sum += arr[i];
  double mean = sum / (double)len;
  for (uint32_t i = 0; i < len; ++i) { double t = (double)arr[i] - mean; var += t * t; }
  return var / len;
}

/* ======================================================================= */
/* dct group                                                               */
/* ======================================================================= */

/* HEVC core transform matrix entries.  The 32-point matrix is
 * M32[k][i] = C[(k*(2i+1)) mod 128] where C follows the cosine symmetries
 * C[64-m] = -C[m], C[128-m] = C[m] and C[0..32] is the spec's coefficient list;
 * the N-point matrix is rows 0, 32/N, 2*32/N... truncated to N columns.
 * tests pin this against the reference's kvz_g_dct_{4,8,16,32} tables
 * (ref:dct-generic.c:38-120). */
static int dct_coef(int n, int k, int i)
{
  static const int c32[33] = { 64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
                               61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4, 0 };
  int m = ((k * (32 / n)) * (2 * i + 1)) % 128;
  if (m > 64) m = 128 - m;          /* C[128-m] = C[m] */
  return m <= 32 ? c32[m] : -c32[64 - m];
}

static const int dst4_mat[4][4] = { /* DST-VII, ref:dct-generic.c:38-44 */
  { 29, 55, 74, 84 }, { 74, 74, 0, -74 }, { 84, -29, -74, 55 }, { 55, -84, 74, -29 } };

static int tr_coef(int n, int is_dst, int k, int i) { return is_dst ? dst4_mat[k][i] : dct_coef(n, k, i); }

/* One forward pass: dst[k*n + j] = (short)((sum_i M[k][i]*src[j*n+i] + add) >> shift)
 * (what partial_butterfly_N computes, ref:dct-generic.c:255-279 etc.; result is
 * TRUNCATED to int16, not clipped). */
static void fwd_pass(int n, int is_dst, const int16_t *src, int16_t *dst, int shift)
{
  const int32_t add = 1 << (shift - 1);
  for (int j = 0; j < n; ++j)
    for (int k = 0; k < n; ++k) {
      int32_t acc = 0;
      for (int i = 0; i < n; ++i) acc += tr_coef(n, is_dst, k, i) * src[j * n + i];
      dst[k * n + j] = (int16_t)((acc + add) >> shift);
    }
}

/* One inverse pass: dst[j*n + k] = clip16((sum_i M[i][k]*src[i*n+j] + add) >> shift)
 * (ref:dct-generic.c:281-303 etc.; CLIPPED to int16). */
static void inv_pass(int n, int is_dst, const int16_t *src, int16_t *dst, int shift)
{
  const int32_t add = 1 << (shift - 1);
  for (int j = 0; j < n; ++j)
    for (int k = 0; k < n; ++k) {
      int32_t acc = 0;
      for (int i = 0; i < n; ++i) acc += tr_coef(n, is_dst, i, k) * src[i * n + j];
      int32_t v = (acc + add) >> shift;
      dst[j * n + k] = (int16_t)ORC_CLIP(-32768, 32767, v);
    }
}

static int ilog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }

/* ref:dct-generic.c:579-588 */
void orc_dct_nxn(int n, int bitdepth, const int16_t *in, int16_t *out)
{
  int16_t tmp[32 * 32];
  fwd_pass(n, 0, in, tmp, ilog2(n) - 1 + (bitdepth - 8));
  fwd_pass(n, 0, tmp, out, ilog2(n) + 6);
}
/* ref:dct-generic.c:590-599 */
void orc_idct_nxn(int n, int bitdepth, const int16_t *in, int16_t *out)
{
  int16_t tmp[32 * 32];
  inv_pass(n, 0, in, tmp, 7);
  inv_pass(n, 0, tmp, out, 12 - (bitdepth - 8));
}
/* ref:dct-generic.c:611-619 */
void orc_dst_4x4(int bitdepth, const int16_t *in, int16_t *out)
{
  int16_t tmp[16];
  fwd_pass(4, 1, in, tmp, 1 + (bitdepth - 8));
  fwd_pass(4, 1, tmp, out, 8);
}
/* ref:dct-generic.c:621-629 */
void orc_idst_4x4(int bitdepth, const int16_t *in, int16_t *out)
{
  int16_t tmp[16];
  inv_pass(4, 1, in, tmp, 7);
  inv_pass(4, 1, tmp, out, 12 - (bitdepth - 8));
}

/* ======================================================================= */
/* quant group                                                             */
/* ======================================================================= */

/* ref:transform.c:56-62 (chroma QP mapping table) built from its defining rule:
 * identity below 30, then the HEVC table 29,30,31,32,33,33,34,34,35,35,36,36,37,37, then qp-6. */
static int chroma_scale(int qp)
{
  static const int mid[14] = { 29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37 };
  if (qp < 30) return qp;
  if (qp < 44) return mid[qp - 30];
  return qp - 6;
}

/* ref:transform.c:88-102 */
int32_t orc_get_scaled_qp(int type, int qp, int qp_offset)
{
  if (type == 0) return qp + qp_offset;
  int q = ORC_CLIP(-qp_offset, 57, qp);
  return q < 0 ? q + qp_offset : chroma_scale(q) + qp_offset;
}

/* Coefficient scan tables (ref:tables.c:10-70 kvz_g_sig_last_scan[scan_idx][log2-1]),
 * generated: 4x4 coefficient groups visited in the scan order, and the same order
 * inside each group.  scan_idx 0 = up-right diagonal, 1 = horizontal, 2 = vertical. */
static uint32_t scan_tab[3][6][32 * 32];
static int scan_ready = 0;

static int order_small(int scan_idx, int dim, int *xs, int *ys)
{
  int n = 0;
  if (scan_idx == 1) { for (int y = 0; y < dim; ++y) for (int x = 0; x < dim; ++x) { xs[n] = x; ys[n++] = y; } }
  else if (scan_idx == 2) { for (int x = 0; x < dim; ++x) for (int y = 0; y < dim; ++y) { xs[n] = x; ys[n++] = y; } }
  else {
    for (int d = 0; d < 2 * dim - 1; ++d)           /* anti-diagonals, bottom-left to top-right */
      for (int y = ORC_MIN(d, dim - 1); y >= 0 && d - y < dim; --y) { xs[n] = d - y; ys[n++] = y; }
  }
  return n;
}

static void build_scans(void)
{
  int xs[64], ys[64], cx[64], cy[64];
  for (int s = 0; s < 3; ++s)
    for (int l = 1; l <= 5; ++l) {
      int w = 1 << l, n = 0;
      if (l <= 2) {
        int cnt = order_small(s, w, xs, ys);
        for (int i = 0; i < cnt; ++i) scan_tab[s][l][n++] = (uint32_t)(ys[i] * w + xs[i]);
      } else {
        int ncg = order_small(s, w / 4, cx, cy);
        int nin = order_small(s, 4, xs, ys);
        for (int g = 0; g < ncg; ++g)
          for (int i = 0; i < nin; ++i)
            scan_tab[s][l][n++] = (uint32_t)((cy[g] * 4 + ys[i]) * w + cx[g] * 4 + xs[i]);
      }
    }
  scan_ready = 1;
}

const uint32_t *orc_scan_table(int scan_idx, int log2_size)
{
  if (!scan_ready) build_scans();
  return scan_tab[scan_idx][log2_size];
}

static const int quant_scales[6] = { 26214, 23302, 20560, 18396, 16384, 14564 };  /* ref:scalinglist.c:78 */
static const int inv_quant_scales[6] = { 40, 45, 51, 57, 64, 72 };                /* ref:scalinglist.c:79 */

/* ref:quant-generic.c:50-180 (kvz_quant_generic), flat scaling list */
void orc_quant(const orc_quant_params *p, const int16_t *coef, int16_t *q_coef, int w, int h,
               int type, int scan_idx, int block_type)
{
  (void)block_type;
  const int log2_tr = ilog2(w);
  const uint32_t *scan = orc_scan_table(scan_idx, log2_tr);
  const int qp_scaled = orc_get_scaled_qp(type, p->qp, (p->bitdepth - 8) * 6);
  const int qc = quant_scales[qp_scaled % 6];
  const int transform_shift = 15 - p->bitdepth - log2_tr;
  const int q_bits = 14 + qp_scaled / 6 + transform_shift;
  const int32_t add = (p->slice_is_intra ? 171 : 85) << (q_bits - 9);
  const int q_bits8 = q_bits - 8;
  uint32_t ac_sum = 0;

  for (int n = 0; n < w * h; ++n) {
    int32_t level = coef[n];
    int64_t abs_level = (int64_t)abs(level);
    int32_t sign = level < 0 ? -1 : 1;
    level = (int32_t)((abs_level * qc + add) >> q_bits);
    ac_sum += (uint32_t)level;
    level *= sign;
    q_coef[n] = (int16_t)ORC_CLIP(-32768, 32767, level);
  }
  if (!p->signhide_enable || ac_sum < 2) return;

  int32_t delta_u[32 * 32];
  for (int n = 0; n < w * h; ++n) {
    int64_t abs_level = (int64_t)abs((int32_t)coef[n]);
    int32_t level = (int32_t)((abs_level * qc + add) >> q_bits);
    delta_u[n] = (int32_t)((abs_level * qc - ((int64_t)level << q_bits)) >> q_bits8);
  }

  int last_cg = -1;
  for (int subset = (w * h - 1) >> 4; subset >= 0; --subset) {
    const int subpos = subset << 4;
    int first_nz = 16, last_nz = -1, abssum = 0;
    for (int n = 15; n >= 0; --n) if (q_coef[scan[n + subpos]]) { last_nz = n; break; }
    for (int n = 0; n < 16; ++n) if (q_coef[scan[n + subpos]]) { first_nz = n; break; }
    for (int n = first_nz; n <= last_nz; ++n) abssum += q_coef[scan[n + subpos]];
    if (last_nz >= 0 && last_cg == -1) last_cg = 1;

    if (last_nz - first_nz >= 4) {
      int signbit = q_coef[scan[subpos + first_nz]] > 0 ? 0 : 1;
      if (signbit != (abssum & 1)) {
        int32_t min_cost = 0x7fffffff, cur_cost = 0x7fffffff;
        int min_pos = -1;
        int16_t final_change = 0, cur_change = 0;
        for (int n = (last_cg == 1 ? last_nz : 15); n >= 0; --n) {
          uint32_t blk = scan[n + subpos];
          if (q_coef[blk] != 0) {
            if (delta_u[blk] > 0) { cur_cost = -delta_u[blk]; cur_change = 1; }
            else if (n == first_nz && abs(q_coef[blk]) == 1) { cur_cost = 0x7fffffff; }
            else { cur_cost = delta_u[blk]; cur_change = -1; }
          } else if (n < first_nz && ((coef[blk] >= 0) ? 0 : 1) != signbit) {
            cur_cost = 0x7fffffff;
          } else { cur_cost = -delta_u[blk]; cur_change = 1; }
          if (cur_cost < min_cost) { min_cost = cur_cost; final_change = cur_change; min_pos = (int)blk; }
        }
        if (q_coef[min_pos] == 32767 || q_coef[min_pos] == -32768) final_change = -1;
        if (coef[min_pos] >= 0) q_coef[min_pos] = (int16_t)(q_coef[min_pos] + final_change);
        else q_coef[min_pos] = (int16_t)(q_coef[min_pos] - final_change);
      }
    }
    if (last_cg == 1) last_cg = 0;
  }
}

/* ref:quant-generic.c:298-340 (kvz_dequant_generic), scaling_list.enable == 0 branch */
void orc_dequant(const orc_quant_params *p, const int16_t *q_coef, int16_t *coef, int w, int h,
                 int type, int block_type)
{
  (void)block_type;
  const int transform_shift = 15 - p->bitdepth - ilog2(w);
  const int qp_scaled = orc_get_scaled_qp(type, p->qp, (p->bitdepth - 8) * 6);
  const int shift = 20 - 14 - transform_shift;
  const int32_t scale = inv_quant_scales[qp_scaled % 6] << (qp_scaled / 6);
  const int32_t add = 1 << (shift - 1);
  for (int n = 0; n < w * h; ++n) {
    int32_t v = (q_coef[n] * scale + add) >> shift;
    coef[n] = (int16_t)ORC_CLIP(-32768, 32767, v);
  }
}

/* ref:quant-generic.c:198-292 (kvz_quantize_residual_generic), RDOQ-off branch;
 * transform choice ref:strategies-dct.c:78-96 (DST for 4x4 intra luma);
 * transform skip ref:transform.c:150-185 */
int orc_quantize_residual(const orc_quant_params *p, int width, int color, int scan_idx,
                          int use_trskip, int cu_is_intra, int in_stride, int out_stride,
                          const orc_pix *ref_in, const orc_pix *pred_in, orc_pix *rec_out,
                          int16_t *coeff_out, int early_skip)
{
  int16_t residual[32 * 32], coeff[32 * 32] = { 0 };
  const int use_dst = (width == 4 && color == 0 && cu_is_intra);
  const int ts_shift = 15 - p->bitdepth - ilog2(width);
  int has_coeffs = 0;

  for (int y = 0; y < width; ++y)
    for (int x = 0; x < width; ++x)
      residual[x + y * width] = (int16_t)((int)ref_in[x + y * in_stride] - (int)pred_in[x + y * in_stride]);

  if (use_trskip) {
    for (int i = 0; i < width * width; ++i) coeff[i] = (int16_t)((uint16_t)residual[i] << ts_shift);
  } else if (use_dst) orc_dst_4x4(p->bitdepth, residual, coeff);
  else orc_dct_nxn(width, p->bitdepth, residual, coeff);

  orc_quant(p, coeff, coeff_out, width, width, color == 0 ? 0 : 2, scan_idx, cu_is_intra ? 1 : 2);

  for (int i = 0; i < width * width; ++i) if (coeff_out[i] != 0) { has_coeffs = 1; break; }

  if (has_coeffs && !early_skip) {
    orc_dequant(p, coeff_out, coeff, width, width, color == 0 ? 0 : (color == 1 ? 2 : 3), cu_is_intra ? 1 : 2);
    if (use_trskip) {
      const int32_t off = 1 << (ts_shift - 1);
      for (int i = 0; i < width * width; ++i) residual[i] = (int16_t)((coeff[i] + off) >> ts_shift);
    } else if (use_dst) orc_idst_4x4(p->bitdepth, coeff, residual);
    else orc_idct_nxn(width, p->bitdepth, coeff, residual);
    for (int y = 0; y < width; ++y)
      for (int x = 0; x < width; ++x) {
        int16_t val = (int16_t)(residual[x + y * width] + pred_in[x + y * in_stride]);
        rec_out[x + y * out_stride] = (orc_pix)ORC_CLIP(0, ORC_PIXEL_MAX, val);
      }
  } else if (rec_out != pred_in) {
    for (int y = 0; y < width; ++y)
      for (int x = 0; x < width; ++x) rec_out[x + y * out_stride] = pred_in[x + y * in_stride];
  }
  return has_coeffs;
}

/* ref:quant-generic.c:342-349 */
uint32_t orc_coeff_abs_sum(const int16_t *c, size_t length)
{
  uint32_t s = 0;
  for (size_t i = 0; i < length; ++i) s += (uint32_t)abs((int)c[i]);
  return s;
}

/* ref:quant-generic.c:351-375 */
double orc_fast_coeff_cost(const int16_t *coeff, int32_t width, uint64_t weights)
{
  uint32_t sum = 0;
  for (int i = 0; i < width * width; ++i) {
    uint32_t a = (uint32_t)abs((int)coeff[i]);
    if (a > 3) a = 3;
    sum += (uint32_t)((weights >> (16 * a)) & 0xffff);
  }
  return (double)sum / 256.0;
}

/* ======================================================================= */
/* intra group                                                             */
/* ======================================================================= */

/* ref:intra-generic.c:49-155 */
void orc_angular_pred(int log2_width, int mode, const orc_pix *ref_top, const orc_pix *ref_left, orc_pix *dst)
{
  static const int disp_tab[9] = { 0, 2, 5, 9, 13, 17, 21, 26, 32 };
  static const int inv_tab[9] = { 0, 4096, 1638, 910, 630, 482, 390, 315, 256 };
  const int w = 1 << log2_width;
  const int vertical = mode >= 18;
  const int mdisp = vertical ? mode - 26 : 10 - mode;
  const int adisp = abs(mdisp);
  const int sdisp = (mdisp < 0 ? -1 : 1) * disp_tab[adisp];
  const orc_pix *main_in = (vertical ? ref_top : ref_left) + 1;   /* index 0 = block coord 0 */
  const orc_pix *side_in = (vertical ? ref_left : ref_top) + 1;
  orc_pix ext[2 * 32 + 1 + 32];
  const orc_pix *rmain = main_in;

  if (sdisp < 0) {
    orc_pix *m = ext + w;                                           /* m[-w .. w-1] */
    for (int x = -1; x < w; ++x) m[x] = main_in[x];
    int acc = 128;
    const int last = (w * sdisp) >> 5;
    for (int x = -2; x >= last; --x) {
      acc += inv_tab[adisp];
      m[x] = side_in[(acc >> 8) - 1];
    }
    rmain = m;
  }

  /* compute in the "vertical" orientation, transpose on write for horizontal modes */
  int pos = 0;
  for (int y = 0; y < w; ++y) {
    pos += sdisp;
    const int di = pos >> 5, df = pos & 31;
    for (int x = 0; x < w; ++x) {
      int v;
      if (sdisp == 0) v = rmain[x];
      else if (df) v = ((32 - df) * rmain[x + di] + df * rmain[x + di + 1] + 16) >> 5;
      else v = rmain[x + di];
      if (vertical) dst[y * w + x] = (orc_pix)v; else dst[x * w + y] = (orc_pix)v;
    }
  }
}

/* ref:intra-generic.c:165-201 */
void orc_intra_pred_planar(int log2_width, const orc_pix *ref_top, const orc_pix *ref_left, orc_pix *dst)
{
  const int w = 1 << log2_width;
  const int tr = ref_top[w + 1], bl = ref_left[w + 1];
  for (int y = 0; y < w; ++y)
    for (int x = 0; x < w; ++x) {
      int hor = (w - 1 - x) * ref_left[y + 1] + (x + 1) * tr;
      int ver = (w - 1 - y) * ref_top[x + 1] + (y + 1) * bl;
      dst[y * w + x] = (orc_pix)((ver + hor + w) >> (log2_width + 1));
    }
}

static int dc_value(int log2_width, const orc_pix *ref_top, const orc_pix *ref_left)
{
  const int w = 1 << log2_width;
  int sum = 0;
  for (int i = 0; i < w; ++i) sum += ref_top[i + 1] + ref_left[i + 1];
  return (orc_pix)((sum + w) >> (log2_width + 1));
}

/* ref:intra-generic.c:210-241 */
void orc_intra_pred_filtered_dc(int log2_width, const orc_pix *ref_top, const orc_pix *ref_left, orc_pix *dst)
{
  const int w = 1 << log2_width;
  const int dc = dc_value(log2_width, ref_top, ref_left);
  for (int y = 0; y < w; ++y)
    for (int x = 0; x < w; ++x) dst[y * w + x] = (orc_pix)dc;
  dst[0] = (orc_pix)((ref_left[1] + 2 * dc + ref_top[1] + 2) / 4);
  for (int x = 1; x < w; ++x) dst[x] = (orc_pix)((ref_top[x + 1] + 3 * dc + 2) / 4);
  for (int y = 1; y < w; ++y) dst[y * w] = (orc_pix)((ref_left[y + 1] + 3 * dc + 2) / 4);
}

/* ref:intra.c:176-204 ([1 2 1] smoothing of both reference arrays) */
static void filter_refs(int log2_width, const orc_pix *top, const orc_pix *left, orc_pix *ftop, orc_pix *fleft)
{
  const int rw = 2 * (1 << log2_width) + 1;
  fleft[0] = (orc_pix)((left[1] + 2 * left[0] + top[1] + 2) / 4);
  ftop[0] = fleft[0];
  for (int i = 1; i < rw - 1; ++i) {
    fleft[i] = (orc_pix)((left[i - 1] + 2 * left[i] + left[i + 1] + 2) / 4);
    ftop[i] = (orc_pix)((top[i - 1] + 2 * top[i] + top[i + 1] + 2) / 4);
  }
  fleft[rw - 1] = left[rw - 1];
  ftop[rw - 1] = top[rw - 1];
}

/* ref:intra.c:252-302 (kvz_intra_predict) incl. intra_pred_dc :229-249 and
 * intra_post_process_angular :207-219 */
void orc_intra_predict(int log2_width, int mode, int color, const orc_pix *ref_top, const orc_pix *ref_left,
                       orc_pix *dst, int filter_boundary)
{
  static const int thres[4] = { 0, 7, 1, 0 };    /* by log2_width-2; ref:intra.c:271 */
  const int w = 1 << log2_width;
  orc_pix ftop[65], fleft[65];
  const orc_pix *top = ref_top, *left = ref_left;
  int use_filtered = 0;
  if (color != 0 || mode == 1 || w == 4) use_filtered = 0;
  else if (mode == 0) use_filtered = 1;
  else {
    int dist = ORC_MIN(abs(mode - 26), abs(mode - 10));
    use_filtered = dist > thres[log2_width - 2];
  }
  if (use_filtered) { filter_refs(log2_width, ref_top, ref_left, ftop, fleft); top = ftop; left = fleft; }

  if (mode == 0) orc_intra_pred_planar(log2_width, top, left, dst);
  else if (mode == 1) {
    if (color == 0 && w < 32) orc_intra_pred_filtered_dc(log2_width, top, left, dst);
    else { int dc = dc_value(log2_width, top, left); for (int i = 0; i < w * w; ++i) dst[i] = (orc_pix)dc; }
  } else {
    orc_angular_pred(log2_width, mode, top, left, dst);
    if (color == 0 && w < 32 && filter_boundary && (mode == 10 || mode == 26)) {
      const orc_pix *r = (mode == 10) ? top : left;
      const int st = (mode == 10) ? 1 : w;
      for (int i = 0; i < w; ++i) {
        int v = dst[i * st] + ((r[i + 1] - r[0]) >> 1);
        dst[i * st] = clip_pix(v);
      }
    }
  }
}

/* Availability tables ref:intra.c:47-82 (num_ref_pixels_top/left[uy][ux]) regenerated from
 * their rule: 4x4 units inside a 64x64 CTU are coded in z-order, so a neighbouring unit is
 * available iff its z-order index is smaller; the CTU above (and above-right) and the CTU to
 * the left are complete, the CTUs to the right / below-left are not. */
static int zidx(int ux, int uy)
{
  int z = 0;
  for (int b = 0; b < 4; ++b) z |= (((ux >> b) & 1) << (2 * b)) | (((uy >> b) & 1) << (2 * b + 1));
  return z;
}
static int ref_px_top(int uy, int ux)
{
  if (uy == 0) return 64;
  int n = 0;
  while (ux + n < 16 && zidx(ux + n, uy - 1) < zidx(ux, uy)) ++n;
  return 4 * n;
}
static int ref_px_left(int uy, int ux)
{
  if (ux == 0) return 4 * (16 - uy);
  int n = 0;
  while (uy + n < 16 && zidx(ux - 1, uy + n) < zidx(ux, uy)) ++n;
  return 4 * n;
}

/* ref:intra.c:305-559 (kvz_intra_build_reference{,_any,_inner}) over a frame plane:
 * lcu->rec / top_ref / left_ref are all views of the same reconstruction, so
 * top_border = rec[(y-1)*stride + x], left_border = rec[y*stride + x-1]. */
void orc_intra_build_reference(int log2_width, int color, int luma_x, int luma_y, int pic_w, int pic_h,
                               const orc_pix *rec, int stride, orc_pix *out_top, orc_pix *out_left)
{
  const int is_c = color != 0;
  const int w = 1 << log2_width;
  const int lx = luma_x % 64, ly = luma_y % 64;
  const int px = luma_x >> is_c, py = luma_y >> is_c;       /* plane coordinates */
  const orc_pix dc = (orc_pix)(1 << (ORC_BITDEPTH - 1));
  const int inner = luma_x > 0 && luma_y > 0;
  #define REC(xx, yy) rec[(yy) * stride + (xx)]

  int avail_l = 0, avail_t = 0;
  if (luma_x > 0) {
    avail_l = ref_px_left(ly / 4, lx / 4) >> is_c;
    avail_l = ORC_MIN(avail_l, 2 * w);
    avail_l = ORC_MIN(avail_l, (pic_h - luma_y) >> is_c);
  }
  if (luma_y > 0) {
    avail_t = ref_px_top(ly / 4, lx / 4) >> is_c;
    avail_t = ORC_MIN(avail_t, 2 * w);
    avail_t = ORC_MIN(avail_t, (pic_w - luma_x) >> is_c);
  }

  if (inner) {
    /* _inner copies in groups of 4 (at least one group), then extends the last copied value */
    out_left[0] = out_top[0] = REC(px - 1, py - 1);
    int i = 0;
    do { for (int k = 0; k < 4; ++k) out_left[i + 1 + k] = REC(px - 1, py + i + k); i += 4; } while (i < avail_l);
    orc_pix near = out_left[i];
    for (; i < 2 * w; ++i) out_left[i + 1] = near;
    i = 0;
    do { for (int k = 0; k < 4; ++k) out_top[i + 1 + k] = REC(px + i + k, py - 1); i += 4; } while (i < avail_t);
    near = out_top[i];
    for (; i < 2 * w; ++i) out_top[i + 1] = near;
  } else {
    if (luma_x > 0) {
      for (int i = 0; i < avail_l; ++i) out_left[i + 1] = REC(px - 1, py + i);
      orc_pix near = out_left[avail_l];
      for (int i = avail_l; i < 2 * w; ++i) out_left[i + 1] = near;
    } else {
      orc_pix near = luma_y > 0 ? REC(px, py - 1) : dc;
      for (int i = 0; i < 2 * w; ++i) out_left[i + 1] = near;
    }
    /* (luma_x > 0 && luma_y > 0) is false here */
    out_left[0] = out_left[1];
    out_top[0] = out_left[1];
    if (luma_y > 0) {
      for (int i = 0; i < avail_t; ++i) out_top[i + 1] = REC(px + i, py - 1);
      orc_pix near = REC(px + avail_t - 1, py - 1);
      for (int i = avail_t; i < 2 * w; ++i) out_top[i + 1] = near;
    } else {
      orc_pix near = luma_x > 0 ? REC(px - 1, py) : dc;
      for (int i = 0; i < 2 * w; ++i) out_top[i + 1] = near;
    }
  }
  #undef REC
}

/* ======================================================================= */
/* ipol group                                                              */
/* ======================================================================= */

static const int8_t luma_fir[4][8] = {          /* ref:filter.c:66-72 */
  { 0, 0, 0, 64, 0, 0, 0, 0 }, { -1, 4, -10, 58, 17, -5, 1, 0 },
  { -1, 4, -11, 40, 40, -11, 4, -1 }, { 0, 1, -5, 17, 58, -10, 4, -1 } };
static const int8_t chroma_fir[8][4] = {        /* ref:filter.c:74-84 */
  { 0, 64, 0, 0 }, { -2, 58, 10, -2 }, { -4, 54, 16, -2 }, { -6, 46, 28, -4 },
  { -4, 36, 36, -4 }, { -4, 28, 46, -6 }, { -2, 16, 54, -4 }, { -2, 10, 58, -2 } };

static int32_t fir_px(const int8_t *f, int taps, const orc_pix *d, int st)
{ int32_t t = 0; for (int i = 0; i < taps; ++i) t += f[i] * d[i * st]; return t; }
static int32_t fir_im(const int8_t *f, int taps, const int16_t *d, int st)
{ int32_t t = 0; for (int i = 0; i < taps; ++i) t += f[i] * d[i * st]; return t; }

/* shared body of ref:ipol-generic.c:134-211 (luma) and :681-758 (chroma) */
static void sample_sep(const orc_pix *src, int ss, int w, int h, orc_pix *dst_px, int16_t *dst_im, int ds,
                       const int8_t *hf, const int8_t *vf, int taps)
{
  const int off = taps / 2 - 1;                 /* 3 luma, 1 chroma */
  const int shift1 = ORC_BITDEPTH - 8, shift2 = 6;
  const int wp_shift = 14 - ORC_BITDEPTH, wp_off = 1 << (wp_shift - 1);
  int16_t tmp[(64 + 7) * 64];
  for (int y = 0; y < h + taps - 1; ++y)
    for (int x = 0; x < w; ++x)
      tmp[y * 64 + x] = (int16_t)(fir_px(hf, taps, &src[ss * (y - off) + (x - off)], 1) >> shift1);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int32_t v = fir_im(vf, taps, &tmp[y * 64 + x], 64) >> shift2;
      if (dst_px) dst_px[y * ds + x] = clip_pix((v + wp_off) >> wp_shift);
      else dst_im[y * ds + x] = (int16_t)v;
    }
}

void orc_sample_quarterpel_luma(const orc_pix *src, int ss, int w, int h, orc_pix *dst, int ds, int mvx, int mvy)
{ sample_sep(src, ss, w, h, dst, NULL, ds, luma_fir[mvx & 3], luma_fir[mvy & 3], 8); }
void orc_sample_quarterpel_luma_hi(const orc_pix *src, int ss, int w, int h, int16_t *dst, int ds, int mvx, int mvy)
{ sample_sep(src, ss, w, h, NULL, dst, ds, luma_fir[mvx & 3], luma_fir[mvy & 3], 8); }
void orc_sample_octpel_chroma(const orc_pix *src, int ss, int w, int h, orc_pix *dst, int ds, int mvx, int mvy)
{ sample_sep(src, ss, w, h, dst, NULL, ds, chroma_fir[mvx & 7], chroma_fir[mvy & 7], 4); }
void orc_sample_octpel_chroma_hi(const orc_pix *src, int ss, int w, int h, int16_t *dst, int ds, int mvx, int mvy)
{ sample_sep(src, ss, w, h, NULL, dst, ds, chroma_fir[mvx & 7], chroma_fir[mvy & 7], 4); }

/* --- fractional motion estimation filters, ref:ipol-generic.c:213-679 ---
 * State shared between the four stages:
 *   im[k][y*64+x]  : horizontal 8-tap (phase k: im[0]=0, im[1]=2/4, im[3]=left qpel, im[4]=right qpel)
 *                    of source row (y-3), column window starting at x-2  (i.e. x - 3 + 1)
 *   col[k][y]      : the same filter for the column one to the left (x = -3 window), rows y-3
 * Every output sample is  clip(( (int16)(vertical_8tap >> 6) + 32) >> 6)  (8-bit). */
#define IM(k) (im + (k) * ORC_IPOL_IM_SIZE)
#define COL(k) (cols + (k) * ORC_FIRST_COLS)
#define FLT(k) (filtered + (k) * 64 * 64)

static orc_pix fme_round(int16_t s)
{
  const int wp_shift = 14 - ORC_BITDEPTH, wp_off = 1 << (wp_shift - 1);
  return clip_pix((s + wp_off) >> wp_shift);
}

static void fme_hor(const orc_pix *src, int ss, int w, int rows, int first_y, const int8_t *f,
                    int16_t *im_k, int16_t *col_k)
{
  const int shift1 = ORC_BITDEPTH - 8;
  for (int y = first_y; y < rows; ++y) {
    for (int x = 0; x < w; ++x)
      im_k[y * 64 + x] = (int16_t)(fir_px(f, 8, &src[ss * (y - 3) + (x - 3 + 1)], 1) >> shift1);
    col_k[y] = (int16_t)(fir_px(f, 8, &src[ss * (y - 3) + (0 - 3)], 1) >> shift1);
  }
}

/* vertical stage helper used by the qpel functions: one output plane */
static void fme_ver_plane(orc_pix *out, int w, int h, const int8_t *vf, const int16_t *im_k,
                          const int16_t *col_k, int use_col_for_x0, int yoff)
{
  for (int y = 0; y < h; ++y) {
    if (use_col_for_x0)
      out[y * 64 + 0] = fme_round((int16_t)(fir_im(vf, 8, &col_k[y + yoff], 1) >> 6));
    for (int x = use_col_for_x0; x < w; ++x)
      out[y * 64 + x] = fme_round((int16_t)(fir_im(vf, 8, &im_k[(y + yoff) * 64 + x - use_col_for_x0], 64) >> 6));
  }
}

void orc_filter_fme(int stage, const orc_pix *src, int ss, int w, int h, orc_pix *filtered,
                    int16_t *im, int fme_level, int16_t *cols, int hpel_off_x, int hpel_off_y)
{
  const int shift1 = ORC_BITDEPTH - 8;
  const int rows = h + 7 + 1;
  if (stage == 0) {
    /* ref:ipol-generic.c:213-330 */
    fme_hor(src, ss, w, rows, 0, luma_fir[0], IM(0), COL(0));
    fme_hor(src, ss, w, rows, fme_level > 1 ? 0 : 1, luma_fir[2], IM(1), COL(2));
    for (int y = 0; y < h; ++y)                        /* right: hpel horizontal only */
      for (int x = 0; x < w; ++x) FLT(1)[y * 64 + x] = fme_round(IM(1)[(3 + 1) * 64 + y * 64 + x]);
    for (int y = 0; y < h; ++y) {                      /* left: shifted copy + extra column */
      FLT(0)[y * 64] = fme_round(COL(2)[y + 3 + 1]);
      for (int x = 1; x < w; ++x) FLT(0)[y * 64 + x] = FLT(1)[y * 64 + x - 1];
    }
    for (int y = 0; y <= h; ++y)                       /* top (rows 0..h-1) and bottom (rows 1..h) */
      for (int x = 0; x < w; ++x) {
        int16_t s = (int16_t)(fir_px(luma_fir[2], 8, &src[ss * (y - 3) + x + 1], ss) >> shift1);
        orc_pix v = fme_round(s);
        if (y < h) FLT(2)[y * 64 + x] = v;
        if (y > 0) FLT(3)[(y - 1) * 64 + x] = v;
      }
  } else if (stage == 1) {
    /* ref:ipol-generic.c:332-405 : diagonal hpel = vertical fir2 over im[1] / col[2] */
    for (int y = 0; y <= h; ++y) {
      /* column x of "right" blocks from im[1]; column 0 of "left" blocks from col[2] */
      orc_pix c0 = fme_round((int16_t)(fir_im(luma_fir[2], 8, &COL(2)[y], 1) >> 6));
      if (y < h) FLT(0)[y * 64] = c0;                  /* top-left */
      if (y > 0) FLT(2)[(y - 1) * 64] = c0;            /* bottom-left */
      for (int x = 0; x < w; ++x) {
        orc_pix v = fme_round((int16_t)(fir_im(luma_fir[2], 8, &IM(1)[y * 64 + x], 64) >> 6));
        if (y < h) { FLT(1)[y * 64 + x] = v; if (x + 1 < w) FLT(0)[y * 64 + x + 1] = v; }
        if (y > 0) { FLT(3)[(y - 1) * 64 + x] = v; if (x + 1 < w) FLT(2)[(y - 1) * 64 + x + 1] = v; }
      }
    }
  } else if (stage == 2) {
    /* ref:ipol-generic.c:407-563 */
    const int8_t *hfl = hpel_off_x != 0 ? luma_fir[1] : luma_fir[3];
    const int8_t *hfr = hpel_off_x != 0 ? luma_fir[3] : luma_fir[1];
    fme_hor(src, ss, w, rows, 0, hfl, IM(3), COL(1));
    fme_hor(src, ss, w, rows, 0, hfr, IM(4), COL(3));
    const int off_x_l = hpel_off_x < 1 ? 0 : 1, off_x_r = hpel_off_x < 0 ? 0 : 1;
    const int off_y_t = hpel_off_y < 1 ? 0 : 1, off_y_b = hpel_off_y < 0 ? 0 : 1;
    const int sample_off_y = hpel_off_y < 0 ? 0 : 1;
    const int sample_off_x = hpel_off_x > -1 ? 1 : 0;
    const int8_t *vlr = hpel_off_y != 0 ? luma_fir[2] : luma_fir[0];
    const int8_t *vt = hpel_off_y != 0 ? luma_fir[1] : luma_fir[3];
    const int8_t *vb = hpel_off_y != 0 ? luma_fir[3] : luma_fir[1];
    const int16_t *hp_im = hpel_off_x != 0 ? IM(1) : IM(0);
    const int16_t *hp_col = hpel_off_x != 0 ? COL(2) : COL(0);
    fme_ver_plane(FLT(0), w, h, vlr, IM(3), COL(1), !off_x_l, sample_off_y);
    fme_ver_plane(FLT(1), w, h, vlr, IM(4), COL(3), !off_x_r, sample_off_y);
    fme_ver_plane(FLT(2), w, h, vt, hp_im, hp_col, !sample_off_x, off_y_t);
    fme_ver_plane(FLT(3), w, h, vb, hp_im, hp_col, !sample_off_x, off_y_b);
  } else {
    /* ref:ipol-generic.c:565-679 */
    const int off_x_l = hpel_off_x < 1 ? 0 : 1, off_x_r = hpel_off_x < 0 ? 0 : 1;
    const int off_y_t = hpel_off_y < 1 ? 0 : 1, off_y_b = hpel_off_y < 0 ? 0 : 1;
    const int8_t *vt = hpel_off_y != 0 ? luma_fir[1] : luma_fir[3];
    const int8_t *vb = hpel_off_y != 0 ? luma_fir[3] : luma_fir[1];
    fme_ver_plane(FLT(0), w, h, vt, IM(3), COL(1), !off_x_l, off_y_t);
    fme_ver_plane(FLT(1), w, h, vt, IM(4), COL(3), !off_x_r, off_y_t);
    fme_ver_plane(FLT(2), w, h, vb, IM(3), COL(1), !off_x_l, off_y_b);
    fme_ver_plane(FLT(3), w, h, vb, IM(4), COL(3), !off_x_r, off_y_b);
  }
}

/* ref:ipol-generic.c:761-814 */
int orc_get_extended_block(const orc_pix *src, int src_w, int src_h, int src_s, int blk_x, int blk_y,
                           int blk_w, int blk_h, int pad_l, int pad_r, int pad_t, int pad_b, int pad_b_simd,
                           orc_pix *buf, int *ext_off, int *ext_origin_off, int *ext_s)
{
  const int min_y = blk_y - pad_t, max_y = blk_y + blk_h + pad_b + pad_b_simd - 1;
  const int min_x = blk_x - pad_l, max_x = blk_x + blk_w + pad_r - 1;
  if (min_y >= 0 && max_y < src_h && min_x >= 0 && max_x < src_w) {
    *ext_off = (blk_y - pad_t) * src_s + (blk_x - pad_l);
    *ext_origin_off = blk_y * src_s + blk_x;
    *ext_s = src_s;
    return 0;
  }
  const int es = pad_l + blk_w + pad_r;
  *ext_off = 0; *ext_s = es; *ext_origin_off = pad_t * es + pad_l;
  int y;
  for (y = -pad_t; y < blk_h + pad_b; ++y) {
    const int cy = ORC_CLIP(0, src_h - 1, blk_y + y);
    for (int x = 0; x < es; ++x) {
      const int cx = ORC_CLIP(0, src_w - 1, min_x + x);
      buf[(y + pad_t) * es + x] = src[cy * src_s + cx];
    }
  }
  for (int ys = 0; ys < pad_b_simd; ++ys) memset(buf + (y + pad_t + ys) * es, 0, (size_t)es * sizeof(orc_pix));
  buf[(blk_h + pad_b + pad_t + pad_b_simd - 1) * es + pad_l + blk_w + pad_r] = 0;
  return 1;
}

/* ======================================================================= */
/* sao group                                                               */
/* ======================================================================= */

static const int eo_dx[4][2] = { { -1, 1 }, { 0, 0 }, { -1, 1 }, { 1, -1 } };   /* ref:sao.h:71-76 */
static const int eo_dy[4][2] = { { 0, 0 }, { -1, 1 }, { -1, 1 }, { -1, 1 } };

static int sgn(int v) { return (v > 0) - (v < 0); }
/* ref:sao_shared_generics.h:41-50 */
static int eo_cat(int a, int b, int c)
{
  static const int map[5] = { 1, 2, 0, 3, 4 };
  return map[2 + sgn(c - a) + sgn(c - b)];
}

/* ref:sao-generic.c:50-81 */
void orc_calc_sao_edge_dir(int bitdepth, const orc_pix *orig, const orc_pix *rec, int eo_class,
                           int bw, int bh, int cat_sum_cnt[2][5])
{
  const int offset = bitdepth != 8 ? 1 << (bitdepth - 9) : 0;
  for (int y = 1; y < bh - 1; ++y)
    for (int x = 1; x < bw - 1; ++x) {
      int c = rec[y * bw + x];
      int a = rec[(y + eo_dy[eo_class][0]) * bw + x + eo_dx[eo_class][0]];
      int b = rec[(y + eo_dy[eo_class][1]) * bw + x + eo_dx[eo_class][1]];
      int cat = eo_cat(a, b, c);
      cat_sum_cnt[0][cat] += (orig[y * bw + x] - c + offset) >> (bitdepth - 8);
      cat_sum_cnt[1][cat] += 1;
    }
}

/* ref:sao_shared_generics.h:52-91 */
int orc_sao_edge_ddistortion(int bitdepth, const orc_pix *orig, const orc_pix *rec, int bw, int bh,
                             int eo_class, const int offsets[5])
{
  const int bit_offset = bitdepth != 8 ? 1 << (bitdepth - 9) : 0;
  int sum = 0;
  for (int y = 1; y < bh - 1; ++y)
    for (int x = 1; x < bw - 1; ++x) {
      int c = rec[y * bw + x];
      int a = rec[(y + eo_dy[eo_class][0]) * bw + x + eo_dx[eo_class][0]];
      int b = rec[(y + eo_dy[eo_class][1]) * bw + x + eo_dx[eo_class][1]];
      int off = offsets[eo_cat(a, b, c)];
      if (off != 0) {
        int diff = (orig[y * bw + x] - c + bit_offset) >> (bitdepth - 8);
        int delta = diff - off;
        sum += delta * delta - diff * diff;
      }
    }
  return sum;
}

/* ref:sao_shared_generics.h:93-130 */
int orc_sao_band_ddistortion(int bitdepth, const orc_pix *orig, const orc_pix *rec, int bw, int bh,
                             int band_pos, const int sao_bands[4])
{
  const int shift = bitdepth - 5;
  int sum = 0;
  for (int i = 0; i < bw * bh; ++i) {
    int band = (rec[i] >> shift) - band_pos;
    int off = (band >= 0 && band <= 3) ? sao_bands[band] : 0;
    if (off != 0) {
      int diff = orig[i] - rec[i];
      int delta = diff - off;
      sum += delta * delta - diff * diff;
    }
  }
  return sum;
}

/* ref:sao-generic.c:84-124 + kvz_calc_sao_offset_array ref:sao.c:180-202 */
void orc_sao_reconstruct_color(int bitdepth, const orc_pix *rec, orc_pix *new_rec, int sao_type, int eo_class,
                               const int band_position[2], const int offsets[10], int stride, int new_stride,
                               int bw, int bh, int color)
{
  const int offset_v = color == 2 ? 5 : 0;
  if (sao_type == 1) {
    const int values = 1 << bitdepth, shift = bitdepth - 5;
    const int bp = band_position[color == 2 ? 1 : 0];
    for (int y = 0; y < bh; ++y)
      for (int x = 0; x < bw; ++x) {
        int val = rec[y * stride + x];
        int d = (val >> shift) - bp;
        if (d >= 0 && d <= 3) val = ORC_CLIP(0, values - 1, val + offsets[d + 1 + offset_v]);
        new_rec[y * new_stride + x] = (orc_pix)val;
      }
  } else {
    for (int y = 0; y < bh; ++y)
      for (int x = 0; x < bw; ++x) {
        const orc_pix *c = &rec[y * stride + x];
        int a = c[eo_dy[eo_class][0] * stride + eo_dx[eo_class][0]];
        int b = c[eo_dy[eo_class][1] * stride + eo_dx[eo_class][1]];
        int v = c[0] + offsets[eo_cat(a, b, c[0]) + offset_v];
        new_rec[y * new_stride + x] = (orc_pix)ORC_CLIP(0, (1 << ORC_BITDEPTH) - 1, v);
      }
  }
}

/* ======================================================================= */
/* nal group                                                               */
/* ======================================================================= */

/* ref:nal-generic.c:57-82 (the generic4/generic8 variants :84-184 compute the same sum) */
void orc_array_checksum(const orc_pix *data, int height, int width, int stride, unsigned char out[4])
{
  uint32_t sum = 0;
  for (int y = 0; y < height; ++y)
    for (int x = 0; x < width; ++x) {
      const uint8_t mask = (uint8_t)((x & 0xff) ^ (y & 0xff) ^ (x >> 8) ^ (y >> 8));
      sum += (uint32_t)((data[y * stride + x] & 0xff) ^ mask);
#if ORC_BITDEPTH > 8
      sum += (uint32_t)(((data[y * stride + x] >> 8) & 0xff) ^ mask);
#endif
    }
  out[0] = (unsigned char)(sum >> 24); out[1] = (unsigned char)(sum >> 16);
  out[2] = (unsigned char)(sum >> 8); out[3] = (unsigned char)sum;
}

/* ======================================================================================================
 * Deblocking filter, frame level (ref: src/filter.c:95-792).
 *
 * The reference filters LCU by LCU (kvz_filter_deblock_lcu, filter.c:783-792) and delays the rightmost four
 * columns of every horizontal edge to the next LCU so that horizontal-edge filtering only ever sees columns whose
 * vertical edges are done.  Per frame this equals the two-pass form restated here: every vertical edge of the
 * 8x8 grid first, then every horizontal edge on the result.  An edge unit is the left (top) edge of one 8x8
 * luma block; whether it is filtered is decided from the SCU at the unit's origin (TU or PU boundary,
 * filter.c:194-246, 711-727); the strength is decided per 4-sample part from the two SCUs across the edge
 * (filter.c:385-470).
 *
 * CU records: 20 bytes per 4x4 SCU, byte-compatible with the reference's cu_info_t on x86-64 SysV
 * (src/cu.h:126-165); pinned by tests/test_deblock.py::test_cu_record_layout.
 * ====================================================================================================== */
static int dbk_tc_table(int i)
{
  /* run-length form of the HEVC tc table (filter.c:41-49) */
  static const unsigned char runs[][2] = { {18,0},{9,1},{4,2},{4,3},{3,4},{2,5},{2,6},{1,7},{1,8},{1,9},{1,10},{1,11},
                                           {1,13},{1,14},{1,16},{1,18},{1,20},{1,22},{1,24} };
  for (unsigned r = 0; r < sizeof(runs) / sizeof(runs[0]); ++r) {
    if (i < runs[r][0]) return runs[r][1];
    i -= runs[r][0];
  }
  return 24;
}
static int dbk_beta_table(int i) { return i < 16 ? 0 : (i < 29 ? i - 10 : 2 * i - 38); }   /* filter.c:51-59 */

typedef struct { int type, depth, part_size, tr_depth, cbf, qp, mv_dir; int mv[2][2]; int mv_ref[2]; } dbk_cu;
static dbk_cu dbk_cu_at(const orc_dbk_params *p, const uint8_t *cus, int x, int y)
{
  const uint8_t *r = cus + 20 * ((size_t)(x >> 2) + (size_t)(y >> 2) * p->cu_stride_scu);
  dbk_cu c;
  c.type = r[0] & 3; c.depth = (r[0] >> 2) & 7; c.part_size = (r[0] >> 5) & 7;
  c.tr_depth = r[1] & 7;
  c.cbf = r[4] | (r[5] << 8);
  c.qp = r[6];
  for (int l = 0; l < 2; ++l) for (int k = 0; k < 2; ++k) c.mv[l][k] = (int16_t)(r[8 + 4 * l + 2 * k] | (r[9 + 4 * l + 2 * k] << 8));
  c.mv_ref[0] = r[16]; c.mv_ref[1] = r[17];
  c.mv_dir = (r[18] >> 6) & 3;
  return c;
}
static int dbk_edge_wanted(const orc_dbk_params *p, const uint8_t *cus, int x, int y, int hor, int *tu_boundary)
{
  const dbk_cu s = dbk_cu_at(p, cus, x, y);
  const int tu_w = 64 >> s.tr_depth, cu_w = 64 >> s.depth;
  const int pos = hor ? y : x;
  *tu_boundary = (pos & (tu_w - 1)) == 0;
  if (*tu_boundary) return 1;
  /* PU boundary of the containing CU (filter.c:216-246): the CU origin or the split position */
  const dbk_cu cu = dbk_cu_at(p, cus, x & ~(cu_w - 1), y & ~(cu_w - 1));
  const int cu_pos = pos & ~(cu_w - 1);
  static const int split_x[8] = { 0, 0, 2, 2, 0, 0, 1, 3 }, split_y[8] = { 0, 2, 0, 2, 1, 3, 0, 0 };
  const int q = hor ? split_y[cu.part_size] : split_x[cu.part_size];
  return pos == cu_pos || (q && pos == cu_pos + q * cu_w / 4);
}
static int dbk_qp(const orc_dbk_params *p, const uint8_t *cus, int x, int y, int hor)
{
  if (!p->per_cu_qp) return p->qp;                                     /* filter.c:262-266 */
  const int qp_p = hor ? dbk_cu_at(p, cus, x, y - 1).qp : dbk_cu_at(p, cus, x - 1, y).qp;
  return (qp_p + dbk_cu_at(p, cus, x, y).qp + 1) >> 1;
}
static int dbk_cbf_y(const dbk_cu *c) { static const int m[5] = { 0x1f, 0x0f, 0x07, 0x03, 0x01 }; return (c->cbf & m[c->tr_depth > 4 ? 4 : c->tr_depth]) != 0; }
static int dbk_far(const int *a, const int *b) { return abs(a[0] - b[0]) >= 4 || abs(a[1] - b[1]) >= 4; }
static int dbk_strength(const orc_dbk_params *p, const dbk_cu *P, const dbk_cu *Q, int tu_boundary)
{
  if (Q->type == 1 || P->type == 1) return 2;
  if (tu_boundary && (dbk_cbf_y(Q) || dbk_cbf_y(P))) return 1;
  if (P->mv_dir != 3 && Q->mv_dir != 3) {
    /* mv_dir 0 never occurs for a coded inter CU; keep the reference's (dir - 1) & 1 wrap harmless */
    const int lp = (P->mv_dir - 1) & 1, lq = (Q->mv_dir - 1) & 1;
    if (dbk_far(Q->mv[lq], P->mv[lp])) return 1;
    if (Q->mv_ref[lq] != P->mv_ref[lp]) return 1;
  }
  if (!p->slice_is_b) return 0;
  int mvp[2][2], mvq[2][2];
  for (int l = 0; l < 2; ++l) for (int k = 0; k < 2; ++k) {
    mvp[l][k] = (P->mv_dir & (1 << l)) ? P->mv[l][k] : 0;
    mvq[l][k] = (Q->mv_dir & (1 << l)) ? Q->mv[l][k] : 0;
  }
  const int rp0 = (P->mv_dir & 1) ? p->ref_LX[0][P->mv_ref[0] & 15] : -1, rp1 = (P->mv_dir & 2) ? p->ref_LX[1][P->mv_ref[1] & 15] : -1;
  const int rq0 = (Q->mv_dir & 1) ? p->ref_LX[0][Q->mv_ref[0] & 15] : -1, rq1 = (Q->mv_dir & 2) ? p->ref_LX[1][Q->mv_ref[1] & 15] : -1;
  if (!((rp0 == rq0 && rp1 == rq1) || (rp0 == rq1 && rp1 == rq0))) return 1;
  const int straight = dbk_far(mvq[0], mvp[0]) || dbk_far(mvq[1], mvp[1]);
  const int crossed = dbk_far(mvq[1], mvp[0]) || dbk_far(mvq[0], mvp[1]);
  if (rp0 != rp1) return rp0 == rq0 ? straight : crossed;
  return straight && crossed;
}

/* one 4-sample luma part; px points at q0 of line 0, xs = step across the edge, ys = step along it */
static void dbk_luma_part(orc_pix *px, long xs, long ys, int beta, int tc)
{
  int b[4][8];
  for (int l = 0; l < 4; ++l) for (int i = -4; i < 4; ++i) b[l][i + 4] = px[l * ys + i * xs];
  const int dp0 = abs(b[0][1] - 2 * b[0][2] + b[0][3]), dq0 = abs(b[0][4] - 2 * b[0][5] + b[0][6]);
  const int dp3 = abs(b[3][1] - 2 * b[3][2] + b[3][3]), dq3 = abs(b[3][4] - 2 * b[3][5] + b[3][6]);
  const int dp = dp0 + dp3, dq = dq0 + dq3;
  if (dp + dq >= beta) return;
  const int strong = 2 * (dp0 + dq0) < (beta >> 2) && 2 * (dp3 + dq3) < (beta >> 2) &&
                     abs(b[0][3] - b[0][4]) < ((5 * tc + 1) >> 1) && abs(b[3][3] - b[3][4]) < ((5 * tc + 1) >> 1) &&
                     abs(b[0][0] - b[0][3]) + abs(b[0][4] - b[0][7]) < (beta >> 3) &&
                     abs(b[3][0] - b[3][3]) + abs(b[3][4] - b[3][7]) < (beta >> 3);
  const int side = (beta + (beta >> 1)) >> 3;
  for (int l = 0; l < 4; ++l) {
    const int *m = b[l];
    int o[8];
    for (int i = 0; i < 8; ++i) o[i] = m[i];
    if (strong) {                                                       /* filter.c:95-118 */
      o[1] = ORC_CLIP(m[1] - 2 * tc, m[1] + 2 * tc, (2 * m[0] + 3 * m[1] + m[2] + m[3] + m[4] + 4) >> 3);
      o[2] = ORC_CLIP(m[2] - 2 * tc, m[2] + 2 * tc, (m[1] + m[2] + m[3] + m[4] + 2) >> 2);
      o[3] = ORC_CLIP(m[3] - 2 * tc, m[3] + 2 * tc, (m[1] + 2 * m[2] + 2 * m[3] + 2 * m[4] + m[5] + 4) >> 3);
      o[4] = ORC_CLIP(m[4] - 2 * tc, m[4] + 2 * tc, (m[2] + 2 * m[3] + 2 * m[4] + 2 * m[5] + m[6] + 4) >> 3);
      o[5] = ORC_CLIP(m[5] - 2 * tc, m[5] + 2 * tc, (m[3] + m[4] + m[5] + m[6] + 2) >> 2);
      o[6] = ORC_CLIP(m[6] - 2 * tc, m[6] + 2 * tc, (m[3] + m[4] + m[5] + 3 * m[6] + 2 * m[7] + 4) >> 3);
    } else {                                                            /* filter.c:128-170 */
      int delta = (9 * (m[4] - m[3]) - 3 * (m[5] - m[2]) + 8) >> 4;
      if (abs(delta) >= tc * 10) continue;
      delta = ORC_CLIP(-tc, tc, delta);
      o[3] = ORC_CLIP(0, ORC_PIXEL_MAX, m[3] + delta);
      o[4] = ORC_CLIP(0, ORC_PIXEL_MAX, m[4] - delta);
      if (dp < side) o[2] = ORC_CLIP(0, ORC_PIXEL_MAX, m[2] + ORC_CLIP(-(tc >> 1), tc >> 1, (((m[1] + m[3] + 1) >> 1) - m[2] + delta) >> 1));
      if (dq < side) o[5] = ORC_CLIP(0, ORC_PIXEL_MAX, m[5] + ORC_CLIP(-(tc >> 1), tc >> 1, (((m[6] + m[4] + 1) >> 1) - m[5] - delta) >> 1));
    }
    for (int i = 1; i < 7; ++i) px[l * ys + (i - 4) * xs] = (orc_pix)o[i];
  }
}
static void dbk_chroma_part(orc_pix *px, long xs, long ys, int tc)    /* filter.c:175-192 */
{
  for (int l = 0; l < 4; ++l) {
    orc_pix *s = px + l * ys;
    const int m2 = s[-2 * xs], m3 = s[-xs], m4 = s[0], m5 = s[xs];
    const int delta = ORC_CLIP(-tc, tc, (((m4 - m3) * 4) + m2 - m5 + 4) >> 3);
    s[-xs] = (orc_pix)ORC_CLIP(0, ORC_PIXEL_MAX, m3 + delta);
    s[0] = (orc_pix)ORC_CLIP(0, ORC_PIXEL_MAX, m4 - delta);
  }
}

void orc_deblock_frame(const orc_dbk_params *p, orc_pix *y, orc_pix *u, orc_pix *v, const uint8_t *cus)
{
  const int W = p->width, H = p->height, Wc = W / 2;
  const int scale = 1 << (ORC_BITDEPTH - 8);
  for (int hor = 0; hor < 2; ++hor) {
    for (int ey = 0; ey < H; ey += 8) for (int ex = 0; ex < W; ex += 8) {
      if (hor ? ey == 0 : ex == 0) continue;                            /* filter.c:654-656 */
      int tu_boundary;
      if (!dbk_edge_wanted(p, cus, ex, ey, hor, &tu_boundary)) continue;
      const int qp = dbk_qp(p, cus, ex, ey, hor);
      const int beta = dbk_beta_table(ORC_CLIP(0, 51, qp + 2 * p->beta_offset_div2)) * scale;
      for (int part = 0; part < 2; ++part) {
        const int px = hor ? ex + 4 * part : ex, py = hor ? ey : ey + 4 * part;
        if (px >= W || py >= H) continue;
        const dbk_cu Q = dbk_cu_at(p, cus, px, py), P = dbk_cu_at(p, cus, hor ? px : px - 1, hor ? py - 1 : py);
        const int bs = dbk_strength(p, &P, &Q, tu_boundary);
        if (!bs) continue;
        const int tc = dbk_tc_table(ORC_CLIP(0, 53, qp + 2 * (bs - 1) + 2 * p->tc_offset_div2)) * scale;
        dbk_luma_part(y + (size_t)py * W + px, hor ? W : 1, hor ? 1 : W, beta, tc);
      }
      /* chroma: edges on the 8x8 chroma grid, only next to intra CUs (filter.c:560-626, 681-686) */
      if (u && v && ((hor ? ey : ex) & 15) == 0) {
        const int cx = ex / 2, cy = ey / 2;
        if (cx >= Wc || cy >= H / 2) continue;
        const dbk_cu Q = dbk_cu_at(p, cus, ex, ey), P = dbk_cu_at(p, cus, hor ? ex : ex - 2, hor ? ey - 2 : ey);
        if (Q.type != 1 && P.type != 1) continue;
        const int qpc = chroma_scale(dbk_qp(p, cus, ex, ey, hor));
        const int tc = dbk_tc_table(ORC_CLIP(0, 53, qpc + 2 + 2 * p->tc_offset_div2)) * scale;
        dbk_chroma_part(u + (size_t)cy * Wc + cx, hor ? Wc : 1, hor ? 1 : Wc, tc);
        dbk_chroma_part(v + (size_t)cy * Wc + cx, hor ? Wc : 1, hor ? 1 : Wc, tc);
      }
    }
  }
}

/* ======================================================================================================
 * RDOQ (ref: kvz_rdoq src/rdo.c:661-977, kvz_get_coded_level :395-452, kvz_get_ic_rate :346-393,
 * calc_last_bits / get_rate_last :465-508, kvz_rdoq_sign_hiding :518-653, find_last_scanpos_generic
 * quant-generic.c:376-399, context derivation src/context.c:315-397).  Flat scaling lists.
 * `cabac`: memory image of cabac_data_t.ctx (src/cabac.h:66-102), one state byte per context model.
 * Pinned against the compiled reference by tests/test_rdoq.py::test_oracle_rdoq_vs_reference (CPU).
 * ====================================================================================================== */
#define ORC_MIN(a, b) ((a) < (b) ? (a) : (b))
#define ORC_MAX(a, b) ((a) > (b) ? (a) : (b))
#define SCANPOS(sp) ((int)s->blk[sp])
/* kvz_entropy_bits (rdo.c:69-79) by MPS / LPS symbol of each of the 64 probability states */
static const int32_t ebits_mps[64] = {
  32768, 30426, 28306, 26378, 24617, 23005, 21523, 20159, 18899, 17734, 16653, 15650, 14717, 13849, 13038, 12282,
  11575, 10914, 10294, 9714, 9169, 8658, 8178, 7727, 7303, 6903, 6527, 6173, 5840, 5525, 5228, 4948,
  4684, 4435, 4199, 3977, 3767, 3568, 3380, 3202, 3034, 2876, 2725, 2583, 2448, 2321, 2200, 2086,
  1978, 1875, 1778, 1686, 1599, 1517, 1439, 1364, 1294, 1228, 1165, 1105, 1048, 994, 943, 895 };
static const int32_t ebits_lps[64] = {
  32768, 35232, 37696, 40159, 42623, 45087, 47551, 50015, 52479, 54942, 57406, 59870, 62334, 64798, 67262, 69725,
  72189, 74653, 77117, 79581, 82044, 84508, 86972, 89436, 91900, 94363, 96827, 99291, 101755, 104219, 106683, 109146,
  111610, 114074, 116538, 119002, 121465, 123929, 126393, 128857, 131321, 133785, 136248, 138712, 141176, 143640, 146104, 148568,
  151031, 153495, 155959, 158423, 160887, 163351, 165814, 168278, 170742, 173207, 175669, 178134, 180598, 183061, 185525, 187989 };
static int ebits_fn(uint8_t st, int bin) { return ((st ^ bin) & 1) ? ebits_lps[st >> 1] : ebits_mps[st >> 1]; }
#define ebits(st, bin) ebits_fn((st), (bin))
static int last_group(int x) { if (x < 4) return x; int l = 31 - __builtin_clz((unsigned)x); return 2 * l + ((x >> (l - 1)) & 1); }
/* byte offsets of the context models inside the image (field order of src/cabac.h:67-101) */
enum { CTXO_QT_CBF_LUMA = 16, CTXO_QT_CBF_CHROMA = 20, CTXO_SIG_CG = 32, CTXO_SIG_LUMA = 36, CTXO_SIG_CHROMA = 63, CTXO_LAST_Y_LUMA = 78,
       CTXO_LAST_Y_CHROMA = 93, CTXO_LAST_X_LUMA = 108, CTXO_LAST_X_CHROMA = 123, CTXO_ONE_LUMA = 138, CTXO_ONE_CHROMA = 154,
       CTXO_ABS_LUMA = 162, CTXO_ABS_CHROMA = 166, CTXO_ROOT_CBF = 181 };
typedef struct { const uint8_t *sig, *one, *abs, *cg, *last_x, *last_y, *cbf; uint8_t root_cbf; } rdoq_models;
static void rdoq_models_init(rdoq_models *m, const uint8_t *c, int type)
{
  m->sig = c + (type ? CTXO_SIG_CHROMA : CTXO_SIG_LUMA);
  m->one = c + (type ? CTXO_ONE_CHROMA : CTXO_ONE_LUMA);
  m->abs = c + (type ? CTXO_ABS_CHROMA : CTXO_ABS_LUMA);
  m->cg = c + CTXO_SIG_CG + type;
  m->last_x = c + (type ? CTXO_LAST_X_CHROMA : CTXO_LAST_X_LUMA);
  m->last_y = c + (type ? CTXO_LAST_Y_CHROMA : CTXO_LAST_Y_LUMA);
  m->cbf = c + (type ? CTXO_QT_CBF_CHROMA : CTXO_QT_CBF_LUMA);
  m->root_cbf = c[CTXO_ROOT_CBF];
}
typedef struct {
  double cost_coeff[1024], cost_sig[1024];
  int32_t inc[1024], dec[1024], sig_inc[1024], qdelta[1024];
  double cg_sig_cost[64];
  int32_t cg_flag[64];
  int32_t last_x_bits[12], last_y_bits[12];
  const uint32_t *blk;
} rdoq_local;
#define scaled_qp(type, qp, off) orc_get_scaled_qp((type), (qp), (off))
static int rdoq_level_rate(const rdoq_models *m, uint32_t abs_level, int ctx_one, int ctx_abs, int rice, uint32_t c1_idx, uint32_t c2_idx)
{
  int rate = 32768;                                               // the sign bin
  const uint32_t base_level = (c1_idx < 8) ? (2 + (c2_idx < 1)) : 1;
  if (abs_level >= base_level) {
    int symbol = (int)(abs_level - base_level);
    if (symbol < (3 << rice)) {
      rate += ((symbol >> rice) + 1 + rice) * 32768;
    } else {
      int length = rice;
      symbol -= 3 << rice;
      while (symbol >= (1 << length)) symbol -= 1 << (length++);
      rate += (3 + length + 1 - rice + length) * 32768;
    }
    if (c1_idx < 8) {
      rate += ebits(m->one[ctx_one], 1);
      if (c2_idx < 1) rate += ebits(m->abs[ctx_abs], 1);
    }
  } else if (abs_level == 1) {
    rate += ebits(m->one[ctx_one], 0);
  } else if (abs_level == 2) {
    rate += ebits(m->one[ctx_one], 1);
    rate += ebits(m->abs[ctx_abs], 0);
  }
  return rate;
}


static int rdoq_sig_ctx(int pattern, int scan_idx, int px, int py, int log2n, int type)
{
  if (px + py == 0) return 0;
  if (log2n == 2) { static const int map[16] = { 0, 1, 4, 5, 2, 3, 4, 5, 6, 6, 8, 8, 7, 7, 8, 8 }; return map[4 * py + px]; }
  const int offset = (log2n == 3) ? (scan_idx == 0 ? 9 : 15) : (type == 0 ? 21 : 12);
  const int sx = px & 3, sy = py & 3;
  int cnt;
  if (pattern == 0) cnt = (sx + sy <= 2) ? ((sx + sy == 0) ? 2 : 1) : 0;
  else if (pattern == 1) cnt = (sy <= 1) ? ((sy == 0) ? 2 : 1) : 0;
  else if (pattern == 2) cnt = (sx <= 1) ? ((sx == 0) ? 2 : 1) : 0;
  else cnt = 2;
  return ((type == 0 && ((px >> 2) + (py >> 2)) > 0) ? 3 : 0) + offset + cnt;
}


static void rdoq_sign_hiding(const rdoq_local *s, double lambda, int bitdepth, int qp_scaled, int scan_idx, int log2n, int last_pos,
                                 const int16_t *coef, int16_t *q)
{
  const int inv_quant = inv_quant_scales[qp_scaled % 6];
  const long long rd_factor = (long long)(inv_quant * inv_quant * (1 << (2 * (qp_scaled / 6))) / lambda / 16 / (1 << (2 * (bitdepth - 8))) + 0.5);
  const int last_cg = (last_pos - 1) >> 4;
  for (int cg = last_cg; cg >= 0; --cg) {
    const int base = cg << 4;
    const uint32_t *pos = s->blk + base;
    int last_nz = -1, first_nz = 16;
    for (int k = 15; k >= 0; --k) if (q[pos[k]]) { last_nz = k; break; }
    for (int k = 0; k <= last_nz; ++k) if (q[pos[k]]) { first_nz = k; break; }
    if (last_nz - first_nz < 4) continue;
    const int signbit = q[pos[first_nz]] <= 0;
    unsigned sum = 0;
    for (int k = first_nz; k <= last_nz; ++k) sum += (unsigned)(int)q[pos[k]];
    if (signbit == (int)(sum & 1)) continue;
    long long best_cost = 0x7FFFFFFFFFFFFFFFLL;
    int best_pos = 0, best_change = 0;
    const int start = (cg == last_cg) ? last_nz : 15;
    for (int k = start; k >= 0; --k) {
      const int p = pos[k];
      const long long quant_cost = rd_factor * s->qdelta[p];
      const int a = abs((int)q[p]);
      long long cost;
      int change;
      if (a != 0) {
        long long inc_bits = s->inc[p], dec_bits = s->dec[p];
        if (a == 1) dec_bits -= 32768 + s->sig_inc[p];
        if (cg == last_cg && last_nz == k && a == 1) dec_bits -= 4 * 32768;
        inc_bits = -quant_cost + inc_bits * 1;            // PRECISION_INC = 15 - CTX_FRAC_BITS = 0
        dec_bits = quant_cost + dec_bits * 1;
        if (inc_bits < dec_bits) { change = 1; cost = inc_bits; }
        else {
          change = -1; cost = dec_bits;
          if (k == first_nz && a == 1) cost = 0x7FFFFFFFFFFFFFFFLL;
        }
      } else {
        const int bits = 32768 + s->inc[p] + s->sig_inc[p];
        cost = -llabs(quant_cost) + (long long)bits;
        change = 1;
        if (k < first_nz && ((coef[p] >= 0) ? 0 : 1) != signbit) cost = 0x7FFFFFFFFFFFFFFFLL;
      }
      if (cost < best_cost) { best_cost = cost; best_pos = p; best_change = change; }
    }
    if (q[best_pos] == 32767 || q[best_pos] == -32768) best_change = -1;
    if (coef[best_pos] >= 0) q[best_pos] = (int16_t)(q[best_pos] + best_change);
    else q[best_pos] = (int16_t)(q[best_pos] - best_change);
  }
}


void orc_rdoq(const orc_rdoq_params *pp, const uint8_t *cabac, const int16_t *coef, int16_t *q, int width, int type, int scan_idx,
              int block_type, int tr_depth)
{
  const int log2n = width == 4 ? 2 : (width == 8 ? 3 : (width == 16 ? 4 : 5));
  const int n = 1 << log2n, nn = n * n;
  const int SH = pp->signhide_enable;
  static __thread rdoq_local sl;
  rdoq_local *s = &sl;
  s->blk = orc_scan_table(scan_idx, log2n);
  const int transform_shift = 15 - pp->bitdepth - log2n;
  const int qp_scaled = scaled_qp(type, pp->qp, (pp->bitdepth - 8) * 6);
  const int q_bits = 14 + qp_scaled / 6 + transform_shift;
  const int qc = quant_scales[qp_scaled % 6];
  const int half = 1 << (q_bits - 1);
  const double lambda = pp->lambda;
  const double err_scale = ldexp(32768.0, -2 * transform_shift) / qc / qc / (1 << (2 * (pp->bitdepth - 8)));
  rdoq_models mm; rdoq_models_init(&mm, cabac, type); const rdoq_models *m = &mm;
#define level_double(blk) ORC_MIN(abs((int)coef[blk]) * qc, 0x7FFFFFFF - half)

  // find_last_scanpos (quant-generic.c:376-399)
  int last_scanpos = -1;
  for (int sp = nn - 1; sp >= 0; --sp) {
    const int blk = s->blk[sp];
    if (((level_double(blk) + half) >> q_bits) > 0) { last_scanpos = sp; break; }
    q[blk] = 0;
  }
  if (last_scanpos < 0) return;
  for (int g = 0; g < nn / 16; ++g) { s->cg_flag[g] = 0; s->cg_sig_cost[g] = 0; }
  if (SH) s->sig_inc[s->blk[last_scanpos]] = 0;
  {
    const int cb = log2n - 2;
    const int off = type ? 0 : (cb * 3 + ((cb + 1) >> 2));
    const int sh = type ? cb : ((cb + 3) >> 2);
    int bx = 0, by = 0, ctx;
    const int groups = last_group(n - 1);
    for (ctx = 0; ctx < groups; ++ctx) {
      const int o = off + (ctx >> sh);
      s->last_x_bits[ctx] = bx + ebits(m->last_x[o], 0); bx += ebits(m->last_x[o], 1);
      s->last_y_bits[ctx] = by + ebits(m->last_y[o], 0); by += ebits(m->last_y[o], 1);
    }
    s->last_x_bits[ctx] = bx; s->last_y_bits[ctx] = by;
  }

  const int cg_last = last_scanpos >> 4;
  const int cgs_side = n >> 2;
  int ctx_set = (last_scanpos > 0 && type == 0) ? 2 : 0;
  int c1 = 1, c2 = 0, rice = 0;
  uint32_t c1_idx = 0, c2_idx = 0;
  double base_cost = 0, block_uncoded_cost = 0;

  for (int cg = cg_last; cg >= 0; --cg) {
    const int cg_first = s->blk[cg << 4];
    const int cgx = (cg_first & (n - 1)) >> 2, cgy = (cg_first >> log2n) >> 2;
    const int cg_blk = cgy * cgs_side + cgx;
    const int right = (cgx < cgs_side - 1) ? (s->cg_flag[cgy * cgs_side + cgx + 1] != 0) : 0;
    const int lower = (cgy < cgs_side - 1) ? (s->cg_flag[(cgy + 1) * cgs_side + cgx] != 0) : 0;
    const int pattern = (n == 4) ? -1 : right + (lower << 1);
    double st_coded = 0, st_uncoded = 0, st_sig = 0, st_sig0 = 0;
    int nnz_before_pos0 = 0;
    for (int k = 15; k >= 0; --k) {
      const int sp = (cg << 4) + k;
      if (sp > last_scanpos) continue;
      const int blk = s->blk[sp];
      const int ld = level_double(blk);
      const uint32_t max_abs = (uint32_t)((ld + half) >> q_bits);
      const double err0 = (double)ld;
      const double c0 = err0 * err0 * err_scale;
      block_uncoded_cost += c0;
      const int one_ctx = 4 * ctx_set + c1, abs_ctx = ctx_set + c2;
      const int last = sp == last_scanpos;
      int ctx_sig = 0;
      if (!last) {
        ctx_sig = rdoq_sig_ctx(pattern, scan_idx, blk & (n - 1), blk >> log2n, log2n, type);
        if (SH) s->sig_inc[blk] = ebits(m->sig[ctx_sig], 1) - ebits(m->sig[ctx_sig], 0);
      }
      uint32_t level = 0;
      double cc, cs = 0;
      if (!last && max_abs < 3) { cs = lambda * ebits(m->sig[ctx_sig], 0); cc = c0 + cs; }
      else cc = 1.7e+308;
      if (max_abs != 0) {
        const double sig_now = last ? 0.0 : lambda * ebits(m->sig[ctx_sig], 1);
        const int lo = max_abs > 1 ? (int)max_abs - 1 : 1;
        for (int lvl = (int)max_abs; lvl >= lo; --lvl) {
          const double err = (double)(ld - lvl * (1 << q_bits));
          double c = err * err * err_scale + lambda * rdoq_level_rate(m, (uint32_t)lvl, one_ctx, abs_ctx, rice, c1_idx, c2_idx);
          c += sig_now;
          if (c < cc) { level = (uint32_t)lvl; cc = c; cs = sig_now; }
        }
      }
      s->cost_coeff[sp] = cc;
      s->cost_sig[sp] = cs;
      if (SH) {
        s->qdelta[blk] = (ld - (int)level * (1 << q_bits)) >> (q_bits - 8);
        if (level > 0) {
          const int now = rdoq_level_rate(m, level, one_ctx, abs_ctx, rice, c1_idx, c2_idx);
          s->inc[blk] = rdoq_level_rate(m, level + 1, one_ctx, abs_ctx, rice, c1_idx, c2_idx) - now;
          s->dec[blk] = rdoq_level_rate(m, level - 1, one_ctx, abs_ctx, rice, c1_idx, c2_idx) - now;
        } else {
          s->inc[blk] = ebits(m->one[one_ctx], 0);
        }
      }
      q[blk] = (int16_t)level;
      base_cost += cc;
      const uint32_t base_level = (c1_idx < 8) ? (2 + (c2_idx < 1)) : 1;
      if (level >= base_level && level > (uint32_t)(3 * (1 << rice))) rice = ORC_MIN(rice + 1, 4);
      if (level >= 1) ++c1_idx;
      if (level > 1) { c1 = 0; c2 += (c2 < 2); ++c2_idx; }
      else if (c1 < 3 && c1 > 0 && level) ++c1;
      if (k == 0 && sp > 0) {
        c2 = 0; rice = 0; c1_idx = 0; c2_idx = 0;
        ctx_set = (sp == 16 || type != 0) ? 0 : 2;
        if (c1 == 0) ++ctx_set;
        c1 = 1;
      }
      st_sig += cs;
      if (k == 0) st_sig0 = cs;
      if (level) {
        s->cg_flag[cg_blk] = 1;
        st_coded += cc - cs;
        st_uncoded += c0;
        if (k != 0) ++nnz_before_pos0;
      }
    }
    if (cg) {
      const int ctx_cg = right || lower;
      if (s->cg_flag[cg_blk] == 0) {
        s->cg_sig_cost[cg] = lambda * ebits(m->cg[ctx_cg], 0);
        base_cost += s->cg_sig_cost[cg] - st_sig;
      } else if (cg < cg_last) {
        if (nnz_before_pos0 == 0) { base_cost -= st_sig0; st_sig -= st_sig0; }
        double cost_zero_cg = base_cost;
        s->cg_sig_cost[cg] = lambda * ebits(m->cg[ctx_cg], 1);
        base_cost += s->cg_sig_cost[cg];
        cost_zero_cg += lambda * ebits(m->cg[ctx_cg], 0);
        cost_zero_cg += st_uncoded;
        cost_zero_cg -= st_coded;
        cost_zero_cg -= st_sig;
        if (cost_zero_cg < base_cost) {
          s->cg_flag[cg_blk] = 0;
          base_cost = cost_zero_cg;
          s->cg_sig_cost[cg] = lambda * ebits(m->cg[ctx_cg], 0);
          for (int k = 15; k >= 0; --k) {
            const int sp = (cg << 4) + k, blk = s->blk[sp];
            if (q[blk]) { q[blk] = 0; const double e = (double)level_double(blk); s->cost_coeff[sp] = e * e * err_scale; s->cost_sig[sp] = 0; }
          }
        }
      }
    } else {
      s->cg_flag[cg_blk] = 1;
    }
  }

  // best last position (rdo.c:884-945)
  double best_cost;
  if (block_type != 1 && type == 0) {
    best_cost = block_uncoded_cost + lambda * ebits(m->root_cbf, 0);
    base_cost += lambda * ebits(m->root_cbf, 1);
  } else {
    const int ctx_cbf = type ? tr_depth : !tr_depth;
    best_cost = block_uncoded_cost + lambda * ebits(m->cbf[ctx_cbf], 0);
    base_cost += lambda * ebits(m->cbf[ctx_cbf], 1);
  }
  int best_last_p1 = 0;
  int found_last = 0;
  for (int cg = cg_last; cg >= 0 && !found_last; --cg) {
    const int cg_first = s->blk[cg << 4];
    const int cg_blk = ((cg_first >> log2n) >> 2) * cgs_side + ((cg_first & (n - 1)) >> 2);
    base_cost -= s->cg_sig_cost[cg];
    if (!s->cg_flag[cg_blk]) continue;
    for (int k = 15; k >= 0; --k) {
      const int sp = (cg << 4) + k;
      if (sp > last_scanpos) continue;
      const int blk = s->blk[sp];
      if (q[blk]) {
        const int py = blk >> log2n, px = blk & (n - 1);
        const int gx = last_group(scan_idx == 2 ? py : px), gy = last_group(scan_idx == 2 ? px : py);
        double bits = s->last_x_bits[gx] + s->last_y_bits[gy];
        if (gx > 3) bits += 32768 * ((gx - 2) >> 1);
        if (gy > 3) bits += 32768 * ((gy - 2) >> 1);
        const double total = base_cost + lambda * bits - s->cost_sig[sp];
        if (total < best_cost) { best_last_p1 = sp + 1; best_cost = total; }
        if (q[blk] > 1) { found_last = 1; break; }
        base_cost -= s->cost_coeff[sp];
        const double e = (double)level_double(blk);
        base_cost += e * e * err_scale;
      } else {
        base_cost -= s->cost_sig[sp];
      }
    }
  }
  unsigned abs_sum = 0;
  for (int sp = 0; sp < best_last_p1; ++sp) {
    const int blk = s->blk[sp];
    const int level = q[blk];
    abs_sum += (unsigned)level;
    q[blk] = (int16_t)(coef[blk] < 0 ? -level : level);
  }
  for (int sp = best_last_p1; sp <= last_scanpos; ++sp) q[s->blk[sp]] = 0;
  if (SH) {
    if (abs_sum >= 2) rdoq_sign_hiding(s, lambda, pp->bitdepth, qp_scaled, scan_idx, log2n, best_last_p1, coef, q);
  }
}
#undef level_double

/* ======================================================================================================
 * Intra mode signalling cost (groundwork for SURVEY §8f rank 2, the CTU search driver): most-probable-mode
 * derivation (kvz_intra_get_dir_luma_predictor, src/intra.c:84-127) and the bit estimates the rough search adds to
 * the SATD (kvz_luma_mode_bits / kvz_chroma_mode_bits, src/search_intra.c:641-698) in counting mode without
 * context adaptation.  Pinned by tests/test_rdoq.py::test_oracle_mode_bits_vs_reference.
 * ====================================================================================================== */
/* left / above: intra mode of the neighbouring PU, or -1 when it does not exist or is not intra; above is ignored on
 * the first row of a CTU (y % 64 == 0) */
void orc_intra_mpm(int left_mode, int above_mode, int y, int8_t preds[3])
{
  const int l = left_mode >= 0 ? left_mode : 1;                       /* DC when unavailable */
  const int a = (above_mode >= 0 && (y % 64) != 0) ? above_mode : 1;
  if (l == a) {
    if (l > 1) { preds[0] = (int8_t)l; preds[1] = (int8_t)(((l + 29) % 32) + 2); preds[2] = (int8_t)(((l - 1) % 32) + 2); }
    else { preds[0] = 0; preds[1] = 1; preds[2] = 26; }
  } else {
    preds[0] = (int8_t)l; preds[1] = (int8_t)a;
    preds[2] = (l && a) ? 0 : ((l + a) < 2 ? 26 : 1);
  }
}
/* cabac: image of cabac_data_t.ctx; intra_mode_model is byte 5, chroma_pred_model[0] byte 6 (src/cabac.h:67-75) */
double orc_luma_mode_bits(const uint8_t *cabac, int luma_mode, const int8_t preds[3])
{
  int in_preds = 0;
  for (int i = 0; i < 3; ++i) if (luma_mode == preds[i]) in_preds = 1;
  double bits = 0;
  bits += (double)ebits(cabac[5], in_preds) * (1.0 / 32768.0);      /* kvz_f_entropy_bits = kvz_entropy_bits / 2^15 */
  bits += in_preds ? (luma_mode == preds[0] ? 1 : 2) : 5;
  return bits;
}
double orc_chroma_mode_bits(const uint8_t *cabac, int chroma_mode, int luma_mode)
{
  double bits = 0;
  bits += (double)ebits(cabac[6], chroma_mode != luma_mode) * (1.0 / 32768.0);
  if (chroma_mode != luma_mode) bits += 2.0;
  return bits;
}
