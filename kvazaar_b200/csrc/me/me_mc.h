// me_mc.h -- motion compensation of decided PUs, luma and chroma, one or two lists: the prediction picture the inter
// residual is taken against.  Single source (device: one warp per PU, each lane writes the 4x4 luma / 2x2 chroma blocks
// of its share -- disjoint pixels, nothing to reduce, no barrier; host test build: the lane shares walked in turn).
//
// What it follows in the reference (restated, nothing copied):
//   kvz_inter_pred_pu, inter_recon_unipred, kvz_inter_recon_bipred    src/inter.c:374-668
//   inter_recon_frac_luma(_hi) / inter_recon_frac_chroma(_hi)         src/inter.c:55-333 (block position, MV fraction)
//   kvz_sample_quarterpel_luma, kvz_sample_octpel_chroma and their 14-bit forms   src/strategies/generic/ipol-generic.c
//   kvz_bipred_average                                                 src/strategies/generic/picture-generic.c:553-700
// As in me_merge.h: 14-bit intermediate of each list, then (s + half) >> (14 - bitdepth) for one list and
// (s0 + s1 + half) >> (15 - bitdepth) for two, clipped -- what the copy / px / im combinations of the reference amount to.
#pragma once
#include "me_merge.h"

namespace kvzme {

static_assert(sizeof(kvz_cuda_me_mc_refs) == 416 && sizeof(kvz_cuda_me_mc_pu) == 20, "record layouts are part of the ABI");

// 14-bit intermediate samples of an N x N chroma block at (bx, by) + fraction (fx, fy) in 1/8 units; plane of cw x ch samples
template <typename Pix, int N>
ME_FN void hi_block_chroma(int bitdepth, const Pix *ref, int cw, int ch, int bx, int by, int fx, int fy, int32_t *out)
{
  const int xmax = cw - 1, ymax = ch - 1;
  const int shift1 = bitdepth - 8;
  int32_t hor[(N + 3) * N];
  for (int r = 0; r < N + 3; ++r) {
    const int yy0 = by + r - 1;
    const int yy = yy0 < 0 ? 0 : (yy0 > ymax ? ymax : yy0);
    const Pix *row = ref + yy * cw;
    for (int c = 0; c < N; ++c) {
      int32_t s = 0;
      for (int k = 0; k < 4; ++k) {
        const int xx0 = bx + c + k - 1;
        s += chroma_tap(fx, k) * (int32_t)row[xx0 < 0 ? 0 : (xx0 > xmax ? xmax : xx0)];
      }
      hor[r * N + c] = s >> shift1;
    }
  }
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < N; ++c) {
      int32_t s = 0;
      for (int k = 0; k < 4; ++k) s += chroma_tap(fy, k) * hor[(r + k) * N + c];
      out[r * N + c] = s >> 6;
    }
}

template <typename Pix>
struct McRefs {
  const Pix *y[16], *u[16], *v[16];
};

template <typename Pix>
ME_FN void predict_pu(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_mc_refs &rf, const McRefs<Pix> &rs, const kvz_cuda_me_mc_pu &pu,
                      Pix *out_y, Pix *out_u, Pix *out_v)
{
  const bool ok = pu.w >= 4 && pu.h >= 4 && pu.w <= 64 && pu.h <= 64 && (pu.w & 3) == 0 && (pu.h & 3) == 0 && pu.x >= 0 && pu.y >= 0 && (pu.x & 1) == 0 &&
                  (pu.y & 1) == 0 && pu.x + pu.w <= p.width && pu.y + pu.h <= p.height && pu.dir >= 1 && pu.dir <= 3;
  if (!ok) return;
  const bool two = pu.dir == 3;
  const int l0 = (pu.dir & 1) ? 0 : 1;
  const int pic0 = rf.ref_LX[l0][pu.mv_ref[l0] & 15] & 15, pic1 = rf.ref_LX[1][pu.mv_ref[1] & 15] & 15;
  if (rs.y[pic0] == nullptr || (two && rs.y[pic1] == nullptr)) return;
  const int pix_max = (1 << p.bitdepth) - 1;
  const int shift = two ? 15 - p.bitdepth : 14 - p.bitdepth;
  const int32_t offset = 1 << (shift - 1);
  const int cw = p.width / 2, ch = p.height / 2;
  int k = 0;
  // luma: 4x4 blocks
  for (int sy = 0; sy < pu.h; sy += 4)
    for (int sx = 0; sx < pu.w; sx += 4) {
      if (k++ % ln.n != ln.lane) continue;
      int32_t a[16], b[16];
      // hi_block reads the picture with stride = width
      hi_block<Pix, 4>(p, rs.y[pic0], p.width, pu.x + sx + (pu.mv[l0][0] >> 2), pu.y + sy + (pu.mv[l0][1] >> 2), pu.mv[l0][0] & 3, pu.mv[l0][1] & 3, a);
      if (two) hi_block<Pix, 4>(p, rs.y[pic1], p.width, pu.x + sx + (pu.mv[1][0] >> 2), pu.y + sy + (pu.mv[1][1] >> 2), pu.mv[1][0] & 3, pu.mv[1][1] & 3, b);
      for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
          int32_t v = ((two ? a[r * 4 + c] + b[r * 4 + c] : a[r * 4 + c]) + offset) >> shift;
          out_y[(pu.y + sy + r) * p.width + pu.x + sx + c] = (Pix)(v < 0 ? 0 : (v > pix_max ? pix_max : v));
        }
    }
  // chroma: 2x2 blocks of both planes
  for (int sy = 0; sy < pu.h / 2; sy += 2)
    for (int sx = 0; sx < pu.w / 2; sx += 2) {
      if (k++ % ln.n != ln.lane) continue;
      for (int plane = 0; plane < 2; ++plane) {
        const Pix *r0 = plane == 0 ? rs.u[pic0] : rs.v[pic0], *r1 = plane == 0 ? rs.u[pic1] : rs.v[pic1];
        Pix *o = plane == 0 ? out_u : out_v;
        int32_t a[4], b[4];
        hi_block_chroma<Pix, 2>(p.bitdepth, r0, cw, ch, pu.x / 2 + sx + (pu.mv[l0][0] >> 3), pu.y / 2 + sy + (pu.mv[l0][1] >> 3), pu.mv[l0][0] & 7, pu.mv[l0][1] & 7, a);
        if (two) hi_block_chroma<Pix, 2>(p.bitdepth, r1, cw, ch, pu.x / 2 + sx + (pu.mv[1][0] >> 3), pu.y / 2 + sy + (pu.mv[1][1] >> 3), pu.mv[1][0] & 7, pu.mv[1][1] & 7, b);
        for (int r = 0; r < 2; ++r)
          for (int c = 0; c < 2; ++c) {
            int32_t v = ((two ? a[r * 2 + c] + b[r * 2 + c] : a[r * 2 + c]) + offset) >> shift;
            o[(pu.y / 2 + sy + r) * cw + pu.x / 2 + sx + c] = (Pix)(v < 0 ? 0 : (v > pix_max ? pix_max : v));
          }
      }
    }
}

}  // namespace kvzme
