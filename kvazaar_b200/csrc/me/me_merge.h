// me_merge.h -- merge analysis of one PU: which merge candidates are usable, their luma prediction (one or two lists)
// and Hadamard cost, sorted -- decision for decision as the reference's search_pu_inter does it for rdo < 3.
//
// Single source (device: one warp per PU; host test build: the 32 lane shares walked in turn).  As in me_frac.h every
// lane predicts the 8x8 / 4x4 sub-blocks of its share straight from the reference pictures in registers: the 14-bit
// intermediate sample of each list (the arithmetic of kvz_sample_14bit_quarterpel_luma; an integer MV gives
// sample << (14 - bitdepth), which is what the reference's pixel copy turns into inside kvz_bipred_average), then
//   one list :  (s + 2^(13-bitdepth)) >> (14 - bitdepth)        == kvz_sample_quarterpel_luma / the plain copy
//   two lists:  (s0 + s1 + 2^(14-bitdepth)) >> (15 - bitdepth)  == bipred_average_{px_px,px_im,im_im}
// clipped to the pixel range.
//
// What it follows in the reference (restated, nothing copied):
//   search_pu_inter, merge analysis      src/search_inter.c:1667-1730
//   merge_candidate_in_list              src/search_inter.c:1575-1594
//   kvz_inter_pred_pu / inter_recon_unipred / kvz_inter_recon_bipred   src/inter.c:374-668 (luma)
//   bipred_average_*                     src/strategies/generic/picture-generic.c:553-660
//   kvz_satd_any_size                    src/strategies/strategies-picture.h:76-112
//   kvz_sort_keys_by_cost                src/search.c:612-626
#pragma once
#include "me_frac.h"

namespace kvzme {

static_assert(sizeof(kvz_cuda_me_refs) == 256 && sizeof(kvz_cuda_me_merge_cost) == 96, "record layouts are part of the ABI");

// 14-bit intermediate samples of the N x N block whose top-left reference position is (bx, by) + fraction (fx, fy)
template <typename Pix, int N>
ME_FN void hi_block(const kvz_cuda_me_params &p, const Pix *ref, int ref_stride, int bx, int by, int fx, int fy, int32_t *out)
{
  const int xmax = p.width - 1, ymax = p.height - 1;
  const int shift1 = p.bitdepth - 8;
  if (fx == 0 && fy == 0) {
    for (int r = 0; r < N; ++r) {
      const int yy = by + r < 0 ? 0 : (by + r > ymax ? ymax : by + r);
      for (int c = 0; c < N; ++c) {
        const int xx = bx + c < 0 ? 0 : (bx + c > xmax ? xmax : bx + c);
        out[r * N + c] = (int32_t)ref[yy * ref_stride + xx] << (14 - p.bitdepth);
      }
    }
    return;
  }
  int32_t hor[(N + 7) * N];
  for (int r = 0; r < N + 7; ++r) {
    const int yy0 = by + r - 3;
    const int yy = yy0 < 0 ? 0 : (yy0 > ymax ? ymax : yy0);
    const Pix *row = ref + yy * ref_stride;
    for (int c = 0; c < N; ++c) {
      int32_t s = 0;
      if (fx == 0) {
        const int xx0 = bx + c;
        s = 64 * (int32_t)row[xx0 < 0 ? 0 : (xx0 > xmax ? xmax : xx0)];
      } else {
        for (int k = 0; k < 8; ++k) {
          const int xx0 = bx + c + k - 3;
          s += luma_tap(fx, k) * (int32_t)row[xx0 < 0 ? 0 : (xx0 > xmax ? xmax : xx0)];
        }
      }
      hor[r * N + c] = s >> shift1;
    }
  }
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < N; ++c) {
      int32_t s = 0;
      if (fy == 0) s = 64 * hor[(r + 3) * N + c];
      else
        for (int k = 0; k < 8; ++k) s += luma_tap(fy, k) * hor[(r + k) * N + c];
      out[r * N + c] = s >> 6;
    }
}

template <typename Pix>
struct RefSet {
  const Pix *plane[16];
  int stride[16];
};

// Hadamard cost of the N x N sub-block at (sx, sy) of the PU against the prediction of a merge candidate
template <typename Pix, int N>
ME_FN uint32_t merge_subblock_cost(const kvz_cuda_me_params &p, const kvz_cuda_me_refs &rf, const RefSet<Pix> &rs, const kvz_cuda_me_pu &pu,
                                   const Planes<Pix> &pl, const kvz_cuda_me_merge &cand, int sx, int sy)
{
  int32_t a[N * N], b[N * N];
  const int pix_max = (1 << p.bitdepth) - 1;
  const bool two = cand.dir == 3;
  const int l0 = (cand.dir & 1) ? 0 : 1;                       // the (first) list in use
  {
    const int pic = rf.ref_LX[l0][cand.ref[l0] & 15] & 15;
    const int qx = cand.mv[l0][0], qy = cand.mv[l0][1];
    hi_block<Pix, N>(p, rs.plane[pic], rs.stride[pic], pu.x + sx + (qx >> 2), pu.y + sy + (qy >> 2), qx & 3, qy & 3, a);
  }
  if (two) {
    const int pic = rf.ref_LX[1][cand.ref[1] & 15] & 15;
    const int qx = cand.mv[1][0], qy = cand.mv[1][1];
    hi_block<Pix, N>(p, rs.plane[pic], rs.stride[pic], pu.x + sx + (qx >> 2), pu.y + sy + (qy >> 2), qx & 3, qy & 3, b);
  }
  const int shift = two ? 15 - p.bitdepth : 14 - p.bitdepth;
  const int32_t offset = 1 << (shift - 1);
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < N; ++c) {
      int32_t v = ((two ? a[r * N + c] + b[r * N + c] : a[r * N + c]) + offset) >> shift;
      v = v < 0 ? 0 : (v > pix_max ? pix_max : v);
      a[r * N + c] = (int32_t)pl.cur[(pu.y + sy + r) * pl.cur_stride + pu.x + sx + c] - v;
    }
  const uint32_t s = wht_abs_sum<N>(a);
  return N == 4 ? (s + 1) >> 1 : (s + 2) >> 2;
}

// kvz_satd_any_size of the PU against a candidate's prediction: one lane's share, sub-blocks in the reference's order
template <typename Pix>
ME_FN uint32_t merge_satd_lane(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_refs &rf, const RefSet<Pix> &rs, const kvz_cuda_me_pu &pu,
                               const Planes<Pix> &pl, const kvz_cuda_me_merge &cand)
{
  int w = pu.w, h = pu.h;
  int x0 = 0, y0 = 0, k = 0;
  uint32_t s = 0;
  if (w % 8 != 0) {
    for (int y = 0; y < h; y += 4)
      if (k++ % ln.n == ln.lane) s += merge_subblock_cost<Pix, 4>(p, rf, rs, pu, pl, cand, 0, y);
    x0 = 4;
    w -= 4;
  }
  if (h % 8 != 0) {
    for (int x = 0; x < w; x += 4)
      if (k++ % ln.n == ln.lane) s += merge_subblock_cost<Pix, 4>(p, rf, rs, pu, pl, cand, x0 + x, 0);
    y0 = 4;
    h -= 4;
  }
  for (int y = 0; y < h; y += 8)
    for (int x = 0; x < w; x += 8)
      if (k++ % ln.n == ln.lane) s += merge_subblock_cost<Pix, 8>(p, rf, rs, pu, pl, cand, x0 + x, y0 + y);
  return s;
}

template <typename Pix>
ME_FN uint32_t merge_satd(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_refs &rf, const RefSet<Pix> &rs, const kvz_cuda_me_pu &pu,
                          const Planes<Pix> &pl, const kvz_cuda_me_merge &cand)
{
#if defined(__CUDA_ARCH__)
  return lane_sum(merge_satd_lane(ln, p, rf, rs, pu, pl, cand)) >> (p.bitdepth - 8);
#else
  uint32_t s = 0;
  for (int l = 0; l < ln.n; ++l) s += merge_satd_lane(Lanes{ l, ln.n }, p, rf, rs, pu, pl, cand);
  return s >> (p.bitdepth - 8);
#endif
}

ME_FN bool same_motion(const kvz_cuda_me_merge &a, const kvz_cuda_me_merge &b)
{
  // all fields, also those of a list the candidate does not use: as merge_candidate_in_list compares them
  return a.dir == b.dir && a.ref[0] == b.ref[0] && a.mv[0][0] == b.mv[0][0] && a.mv[0][1] == b.mv[0][1] && a.ref[1] == b.ref[1] &&
         a.mv[1][0] == b.mv[1][0] && a.mv[1][1] == b.mv[1][1];
}

template <typename Pix>
ME_FN void merge_cost_pu(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_refs &rf, const RefSet<Pix> &rs, const kvz_cuda_me_pu &pu,
                         const Planes<Pix> &pl, kvz_cuda_me_merge_cost *out)
{
  kvz_cuda_me_merge_cost m;
  m.size = 0;
  for (int i = 0; i < 5; ++i) { m.cost[i] = kMaxDouble; m.bits[i] = 0; m.keys[i] = -1; m.merge_idx[i] = 0; }
  m.pad[0] = m.pad[1] = 0;
  const int n_cand = pu_valid(p, pu) ? pu.num_merge : 0;
  for (int idx = 0; idx < n_cand; ++idx) {
    const kvz_cuda_me_merge &cand = pu.merge[idx];
    if (cand.dir < 1 || cand.dir > 3) continue;                                   // not a motion candidate: a bad record, not a reference case
    if (((cand.dir & 1) && rs.plane[rf.ref_LX[0][cand.ref[0] & 15] & 15] == nullptr) || ((cand.dir & 2) && rs.plane[rf.ref_LX[1][cand.ref[1] & 15] & 15] == nullptr)) continue;
    if (cand.dir == 3 && !rf.bipred) continue;
    if (cand.dir == 3 && !(pu.w + pu.h > 12)) continue;
    bool dup = false;
    for (int i = 0; i < m.size && !dup; ++i) dup = same_motion(cand, pu.merge[m.merge_idx[m.keys[i]]]);
    if (((cand.dir & 1) && !mv_allowed(p, pu, cand.mv[0][0], cand.mv[0][1])) || ((cand.dir & 2) && !mv_allowed(p, pu, cand.mv[1][0], cand.mv[1][1])) || dup)
      continue;
    const uint32_t satd = merge_satd(ln, p, rf, rs, pu, pl, cand);
    const double bits = rf.merge_flag_bits + idx + rf.merge_idx_bits[idx != 0];
    double cost = (double)satd;
    cost += bits * p.lambda_sqrt;
    m.merge_idx[m.size] = (int8_t)idx;
    m.cost[m.size] = cost;
    m.bits[m.size] = bits;
    m.keys[m.size] = (int8_t)m.size;
    m.size++;
  }
  for (int i = 1; i < m.size; ++i) {                         // kvz_sort_keys_by_cost: insertion sort, ties keep their order
    const int8_t cur = m.keys[i];
    const double cur_cost = m.cost[cur];
    int j = i;
    while (j > 0 && cur_cost < m.cost[m.keys[j - 1]]) { m.keys[j] = m.keys[j - 1]; --j; }
    m.keys[j] = cur;
  }
  if (ln.lane == 0) *out = m;
}

// ---- bi-prediction from the two best uni-predictions (search_pu_inter, search_inter.c:1937-2031)
static_assert(sizeof(kvz_cuda_me_bipred_pu) == 28 && sizeof(kvz_cuda_me_bipred_result) == 16, "record layouts are part of the ABI");

// select_mv_cand without a cost output (search_inter.c:351-391): 0 when both candidates are equal
ME_FN int pick_mv_cand(const int16_t mv_cand[2][2], int mvx, int mvy)
{
  if (mv_cand[0][0] == mv_cand[1][0] && mv_cand[0][1] == mv_cand[1][1]) return 0;
  const uint32_t c0 = mvd_bits(mvx - mv_cand[0][0], mvy - mv_cand[0][1]);
  const uint32_t c1 = mvd_bits(mvx - mv_cand[1][0], mvy - mv_cand[1][1]);
  return c1 < c0 ? 1 : 0;
}

template <typename Pix>
ME_FN void bipred_pu(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_refs &rf, const RefSet<Pix> &rs, const kvz_cuda_me_bipred_pu &bp,
                     const Planes<Pix> &pl, kvz_cuda_me_bipred_result *out)
{
  kvz_cuda_me_bipred_result r;
  r.cost = kMaxDouble; r.bits = 0; r.mv_cand_idx[0] = r.mv_cand_idx[1] = 0; r.valid = 0; r.pad = 0;
  kvz_cuda_me_pu pu;                                   // the geometry and AMVP candidates in the shape the shared helpers take
  pu.x = bp.x; pu.y = bp.y; pu.w = bp.w; pu.h = bp.h; pu.num_merge = 0; pu.pad = 0;
  pu.start_mv[0] = pu.start_mv[1] = 0;
  for (int c = 0; c < 2; ++c) { pu.mv_cand[c][0] = bp.mv_cand[c][0]; pu.mv_cand[c][1] = bp.mv_cand[c][1]; }
  const bool planes_ok = rs.plane[rf.ref_LX[0][bp.mv_ref[0] & 15] & 15] != nullptr && rs.plane[rf.ref_LX[1][bp.mv_ref[1] & 15] & 15] != nullptr;
  if (pu_valid(p, pu) && rf.bipred && bp.w + bp.h >= 16 && planes_ok) {
    kvz_cuda_me_merge cand;
    cand.dir = 3; cand.pad = 0;
    for (int l = 0; l < 2; ++l) { cand.mv[l][0] = bp.mv[l][0]; cand.mv[l][1] = bp.mv[l][1]; cand.ref[l] = bp.mv_ref[l]; }
    double cost = (double)merge_satd(ln, p, rf, rs, pu, pl, cand);
    const uint32_t b0 = qpel_mv_bits(pu, bp.mv[0][0], bp.mv[0][1]), b1 = qpel_mv_bits(pu, bp.mv[1][0], bp.mv[1][1]);
    cost += (double)b0 * p.lambda_sqrt;
    cost += (double)b1 * p.lambda_sqrt;
    const int extra_bits = bp.mv_ref[0] + bp.mv_ref[1] + 2;
    cost += p.lambda_sqrt * extra_bits;
    r.cost = cost;
    r.bits = (int32_t)(b0 + b1) + extra_bits;
    r.mv_cand_idx[0] = (uint8_t)pick_mv_cand(bp.mv_cand, bp.mv[0][0], bp.mv[0][1]);
    r.mv_cand_idx[1] = (uint8_t)pick_mv_cand(bp.mv_cand, bp.mv[1][0], bp.mv[1][1]);
    r.valid = 1;
  }
  if (ln.lane == 0) *out = r;
}

}  // namespace kvzme
