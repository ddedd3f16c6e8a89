// me_cand.h -- AMVP and merge candidates of one PU from a snapshot of the CU records, as the reference derives them.
//
// Single source: me_search.cu compiles it for the device (one thread per PU, integer logic only), tests/hostsim/
// me_hostsim.cpp for the host (TEST INFRASTRUCTURE).
//
// What it follows in the reference (restated, nothing copied):
//   is_a0_cand_coded / is_b0_cand_coded      src/inter.c:689-823
//   get_temporal_merge_candidates            src/inter.c:836-907
//   get_spatial_merge_candidates_cua         src/inter.c:1015-1076 (same availability rules as the lcu_t variant :922-1001)
//   get_scaled_mv, apply_mv_scaling_pocs     src/inter.c:1078-1103
//   add_temporal_candidate                   src/inter.c:1134-1184
//   add_mvp_candidate                        src/inter.c:1186-1220
//   get_mv_cand_from_candidates              src/inter.c:1225-1318
//   is_duplicate_candidate, add_merge_candidate, kvz_inter_get_merge_cand   src/inter.c:1385-1572
#pragma once
#include "me_search.h"

namespace kvzme {

static_assert(sizeof(kvz_cuda_me_cu) == 12 && sizeof(kvz_cuda_me_frame) == 260 && sizeof(kvz_cuda_me_cand_pu) == 12 &&
              sizeof(kvz_cuda_me_cand_out) == 80, "record layouts are part of the ABI");

constexpr int kCuInter = 2;

struct CuImage {
  const kvz_cuda_me_cu *cu;
  int stride;          // records per row
};
ME_FN const kvz_cuda_me_cu *cu_at(const CuImage &im, int x, int y) { return &im.cu[(y >> 2) * im.stride + (x >> 2)]; }

struct Neighbours {
  const kvz_cuda_me_cu *a[2], *b[3], *c3, *h;
};

ME_FN int low_bit(int v) { return v & ~(v - 1); }

// Is the CU holding A0 (below-left of the PU) decided before the PU?  Walks up the quadtree from the square block at
// the PU's lower-left corner: a first-column-first-row position has its A0 in the already coded left neighbour, the
// lower-left quadrant passes the question to its parent, the right column has A0 in a later CU.
ME_FN bool a0_coded(int x, int y, int w, int h)
{
  int size = low_bit(w) < low_bit(h) ? low_bit(w) : low_bit(h);
  if (h != size) y += h - size;
  for (; size < kLcuWidth; size *= 2) {
    const int parent = 2 * size;
    const bool right = x % parent != 0, lower = y % parent != 0;
    if (right) return false;
    if (!lower) return true;
    y -= size;
  }
  return false;                      // 64x64: A0 lies outside the LCU
}

// Same for B0 (above-right), from the square block at the PU's upper-right corner.
ME_FN bool b0_coded(int x, int y, int w, int h)
{
  int size = low_bit(w) < low_bit(h) ? low_bit(w) : low_bit(h);
  if (w != size) x += w - size;
  for (; size < kLcuWidth; size *= 2) {
    const int parent = 2 * size;
    const bool right = x % parent != 0, lower = y % parent != 0;
    if (!right) return true;         // upper-left: B0 above the parent; lower-left: B0 is the upper-right sibling
    if (lower) return false;         // lower-right: B0 in the CU to the right, coded later
    x -= size;                       // upper-right: ask the parent
  }
  return true;                       // the LCU above-right is coded already
}

ME_FN void spatial_neighbours(const CuImage &im, const kvz_cuda_me_frame &f, int x, int y, int w, int h, Neighbours &nb)
{
  const int xl = x & (kLcuWidth - 1), yl = y & (kLcuWidth - 1);
  if (x != 0) {
    const kvz_cuda_me_cu *a1 = cu_at(im, x - 1, y + h - 1);
    if (a1->type == kCuInter) nb.a[1] = a1;
    if (yl + h < kLcuWidth && y + h < f.height) {
      const kvz_cuda_me_cu *a0 = cu_at(im, x - 1, y + h);
      if (a0->type == kCuInter && a0_coded(x, y, w, h)) nb.a[0] = a0;
    }
  }
  if (y != 0) {
    if (x + w < f.width && (xl + w < kLcuWidth || yl == 0)) {
      const kvz_cuda_me_cu *b0 = cu_at(im, x + w, y - 1);
      if (b0->type == kCuInter && b0_coded(x, y, w, h)) nb.b[0] = b0;
    }
    const kvz_cuda_me_cu *b1 = cu_at(im, x + w - 1, y - 1);
    if (b1->type == kCuInter) nb.b[1] = b1;
    if (x != 0) {
      const kvz_cuda_me_cu *b2 = cu_at(im, x - 1, y - 1);
      if (b2->type == kCuInter) nb.b[2] = b2;
    }
  }
}

// colocated candidates: H (below-right, on the 16x16 grid, not across an LCU row) and C3 (centre)
ME_FN void temporal_neighbours(const CuImage &col, const kvz_cuda_me_frame &f, int x, int y, int w, int h, Neighbours &nb)
{
  nb.c3 = nb.h = nullptr;
  if (!f.used_size) return;
  if (!(f.ref_LX_size[0] > 0)) return;
  const unsigned xbr = (unsigned)(x + w), ybr = (unsigned)(y + h);
  if (xbr < (unsigned)f.width && ybr < (unsigned)f.height && ybr % kLcuWidth != 0) {
    const kvz_cuda_me_cu *c = &col.cu[((xbr >> 4) << 4) / 4 + (((ybr >> 4) << 4) / 4) * col.stride];
    if (c->type == kCuInter) nb.h = c;
  }
  const unsigned xc = (unsigned)(x + w / 2), yc = (unsigned)(y + h / 2);
  if (xc < (unsigned)f.width && yc < (unsigned)f.height) {
    const kvz_cuda_me_cu *c = &col.cu[((xc >> 4) << 4) / 4 + (((yc >> 4) << 4) / 4) * col.stride];
    if (c->type == kCuInter) nb.c3 = c;
  }
}

ME_FN int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

ME_FN int16_t scaled_mv(int16_t mv, int scale)
{
  const int32_t s = scale * mv;
  return (int16_t)clip3(-32768, 32767, (s + 127 + (s < 0)) >> 8);
}

ME_FN void scale_mv_pocs(int cur_poc, int cur_ref_poc, int nb_poc, int nb_ref_poc, int16_t mv[2])
{
  int dc = cur_poc - cur_ref_poc, dn = nb_poc - nb_ref_poc;
  if (dc == dn) return;
  dc = clip3(-128, 127, dc);
  dn = clip3(-128, 127, dn);
  const int an = dn < 0 ? -dn : dn;
  const int scale = clip3(-4096, 4095, (dc * ((0x4000 + (an >> 1)) / dn) + 32) >> 6);
  mv[0] = scaled_mv(mv[0], scale);
  mv[1] = scaled_mv(mv[1], scale);
}

ME_FN bool temporal_candidate(const kvz_cuda_me_frame &f, int current_ref, const kvz_cuda_me_cu *col, int reflist, int16_t mv_out[2])
{
  if (!col) return false;
  if (!(f.ref_LX_size[0] > 0)) return false;
  const int colocated_ref = f.ref_LX[0][0];
  int col_list = reflist;
  for (int i = 0; i < f.used_size; ++i)
    if (f.pocs[i] > f.poc) { col_list = 1; break; }
  if ((col->mv_dir & (col_list + 1)) == 0) col_list = 1 - col_list;
  mv_out[0] = col->mv[col_list][0];
  mv_out[1] = col->mv[col_list][1];
  scale_mv_pocs(f.poc, f.pocs[current_ref], f.pocs[colocated_ref], f.col_ref_pocs[col_list][col->mv_ref[col_list] & 15], mv_out);
  return true;
}

// one spatial neighbour as MVP of (reflist, cur_mv_ref): same reference picture (unscaled pass) or any, scaled
ME_FN bool mvp_candidate(const kvz_cuda_me_frame &f, int cur_mv_ref, const kvz_cuda_me_cu *cand, int reflist, bool scaling, int16_t out[2])
{
  if (!cand) return false;
  for (int i = 0; i < 2; ++i) {
    const int cl = i == 0 ? reflist : !reflist;
    if ((cand->mv_dir & (1 << cl)) == 0) continue;
    if (scaling) {
      out[0] = cand->mv[cl][0];
      out[1] = cand->mv[cl][1];
      scale_mv_pocs(f.poc, f.pocs[f.ref_LX[reflist][cur_mv_ref & 15]], f.poc, f.pocs[f.ref_LX[cl][cand->mv_ref[cl] & 15]], out);
      return true;
    }
    if (f.ref_LX[cl][cand->mv_ref[cl] & 15] == f.ref_LX[reflist][cur_mv_ref & 15]) {
      out[0] = cand->mv[cl][0];
      out[1] = cand->mv[cl][1];
      return true;
    }
  }
  return false;
}

ME_FN void amvp(const kvz_cuda_me_frame &f, const Neighbours &nb, int cur_mv_ref, int reflist, int16_t mv_cand[2][2])
{
  int n = 0, nb_b = 0;
  int16_t tmp[3][2] = { { 0, 0 }, { 0, 0 }, { 0, 0 } };      // candidate n is written before it is counted; never more than two are kept
  for (int i = 0; i < 2; ++i)
    if (mvp_candidate(f, cur_mv_ref, nb.a[i], reflist, false, tmp[n])) { ++n; break; }
  if (n == 0)
    for (int i = 0; i < 2; ++i)
      if (mvp_candidate(f, cur_mv_ref, nb.a[i], reflist, true, tmp[n])) { ++n; break; }
  for (int i = 0; i < 3; ++i)
    if (mvp_candidate(f, cur_mv_ref, nb.b[i], reflist, false, tmp[n])) { ++nb_b; break; }
  n += nb_b;
  if (nb.a[0] || nb.a[1]) nb_b = 1;
  else if (n != 2) nb_b = 0;
  if (!nb_b)
    for (int i = 0; i < 3; ++i)
      if (mvp_candidate(f, cur_mv_ref, nb.b[i], reflist, true, tmp[n])) { ++n; break; }
  if (n == 2 && tmp[0][0] == tmp[1][0] && tmp[0][1] == tmp[1][1]) n = 1;
  const bool tmvp = f.tmvp_enable && f.poc > 1 && f.used_size && n < 2 && (nb.h || nb.c3);
  if (tmvp && temporal_candidate(f, f.ref_LX[reflist][cur_mv_ref & 15], nb.h ? nb.h : nb.c3, reflist, tmp[n])) ++n;
  for (; n < 2; ++n) tmp[n][0] = tmp[n][1] = 0;
  for (int c = 0; c < 2; ++c) { mv_cand[c][0] = tmp[c][0]; mv_cand[c][1] = tmp[c][1]; }
}

ME_FN bool duplicate(const kvz_cuda_me_cu *a, const kvz_cuda_me_cu *b)
{
  if (!b) return false;
  if (a->mv_dir != b->mv_dir) return false;
  for (int l = 0; l < 2; ++l)
    if ((a->mv_dir & (1 << l)) && (a->mv[l][0] != b->mv[l][0] || a->mv[l][1] != b->mv[l][1] || a->mv_ref[l] != b->mv_ref[l])) return false;
  return true;
}

// a spatial merge candidate; the list the neighbour does not use reads as MV 0 / reference 255 (inter_clear_cu_unused, inter.c:669-678)
ME_FN bool merge_add(const kvz_cuda_me_cu *cand, const kvz_cuda_me_cu *dup1, const kvz_cuda_me_cu *dup2, kvz_cuda_me_merge *out, int n, int max_n)
{
  if (!cand || duplicate(cand, dup1) || duplicate(cand, dup2) || n >= max_n) return false;
  for (int l = 0; l < 2; ++l) {
    const bool used = (cand->mv_dir & (1 << l)) != 0;
    out->mv[l][0] = used ? cand->mv[l][0] : (int16_t)0;
    out->mv[l][1] = used ? cand->mv[l][1] : (int16_t)0;
    out->ref[l] = used ? cand->mv_ref[l] : (uint8_t)255;
  }
  out->dir = cand->mv_dir;
  return true;
}

ME_FN int merge_candidates(const kvz_cuda_me_frame &f, Neighbours nb, const kvz_cuda_me_cand_pu &pu, kvz_cuda_me_merge *mc)
{
  const int max_n = f.max_merge;
  int n = 0;
  if (!pu.use_a1) nb.a[1] = nullptr;
  if (!pu.use_b1) nb.b[1] = nullptr;
  if (merge_add(nb.a[1], nullptr, nullptr, &mc[n], n, max_n)) ++n;
  if (merge_add(nb.b[1], nb.a[1], nullptr, &mc[n], n, max_n)) ++n;
  if (merge_add(nb.b[0], nb.b[1], nullptr, &mc[n], n, max_n)) ++n;
  if (merge_add(nb.a[0], nb.a[1], nullptr, &mc[n], n, max_n)) ++n;
  if (n < 4 && merge_add(nb.b[2], nb.a[1], nb.b[1], &mc[n], n, max_n)) ++n;

  if (f.tmvp_enable && n < max_n && f.used_size) {
    mc[n].dir = 0;
    const kvz_cuda_me_cu *t = nb.h ? nb.h : nb.c3;
    for (int l = 0; l <= (f.slice_b ? 1 : 0); ++l)
      if (temporal_candidate(f, f.ref_LX[l][0], t, l, mc[n].mv[l])) {
        mc[n].ref[l] = 0;
        mc[n].dir |= (uint8_t)(1 << l);
      }
    if (mc[n].dir != 0) ++n;
  }

  if (n < max_n && f.slice_b) {
    // combined bi-predictive candidates: L0 of one, L1 of another, in the standard's pair order
    const uint8_t p0[12] = { 0, 1, 0, 2, 1, 2, 0, 3, 1, 3, 2, 3 }, p1[12] = { 1, 0, 2, 0, 2, 1, 3, 0, 3, 1, 3, 2 };
    const int cutoff = n;
    for (int idx = 0; idx < cutoff * (cutoff - 1) && n != max_n; ++idx) {
      const int i = p0[idx], j = p1[idx];
      if (i >= n || j >= n) break;
      if ((mc[i].dir & 1) && (mc[j].dir & 2)) {
        mc[n].dir = 3;
        mc[n].mv[0][0] = mc[i].mv[0][0]; mc[n].mv[0][1] = mc[i].mv[0][1];
        mc[n].mv[1][0] = mc[j].mv[1][0]; mc[n].mv[1][1] = mc[j].mv[1][1];
        mc[n].ref[0] = mc[i].ref[0];
        mc[n].ref[1] = mc[j].ref[1];
        const bool same = f.ref_LX[0][mc[i].ref[0] & 15] == f.ref_LX[1][mc[j].ref[1] & 15] && mc[i].mv[0][0] == mc[j].mv[1][0] &&
                          mc[i].mv[0][1] == mc[j].mv[1][1];
        if (!same) ++n;
      }
    }
  }

  int num_ref = f.used_size;
  if (n < max_n && f.slice_b) {
    int neg = 0, pos = 0;
    for (int j = 0; j < f.used_size; ++j) (f.pocs[j] < f.poc ? neg : pos)++;
    num_ref = neg < pos ? neg : pos;
  }
  for (int zero_idx = 0; n != max_n; ++zero_idx, ++n) {
    mc[n].mv[0][0] = mc[n].mv[0][1] = 0;
    mc[n].ref[0] = (uint8_t)(zero_idx >= num_ref - 1 ? 0 : zero_idx);
    mc[n].ref[1] = mc[n].ref[0];
    mc[n].dir = 1;
    if (f.slice_b) {
      mc[n].mv[1][0] = mc[n].mv[1][1] = 0;
      mc[n].dir = 3;
    }
  }
  return n;
}

ME_FN void candidates_of_pu(const kvz_cuda_me_frame &f, const CuImage &cur, const CuImage &col, const kvz_cuda_me_cand_pu &pu, kvz_cuda_me_cand_out *out)
{
  Neighbours nb = { { nullptr, nullptr }, { nullptr, nullptr, nullptr }, nullptr, nullptr };
  // a PU outside the picture (a bad record) gets no neighbours: zero AMVP candidates, zero-motion merge candidates
  const bool valid = pu.w >= 4 && pu.h >= 4 && pu.w <= 64 && pu.h <= 64 && pu.x >= 0 && pu.y >= 0 && pu.x + pu.w <= f.width && pu.y + pu.h <= f.height;
  if (valid) {
    spatial_neighbours(cur, f, pu.x, pu.y, pu.w, pu.h, nb);
    temporal_neighbours(col, f, pu.x, pu.y, pu.w, pu.h, nb);
  }
  kvz_cuda_me_cand_out o;
  for (int l = 0; l < 2; ++l)
    for (int c = 0; c < 2; ++c) o.mv_cand[l][c][0] = o.mv_cand[l][c][1] = 0;
  for (int m = 0; m < 5; ++m) {
    o.merge[m].mv[0][0] = o.merge[m].mv[0][1] = o.merge[m].mv[1][0] = o.merge[m].mv[1][1] = 0;
    o.merge[m].dir = 0; o.merge[m].ref[0] = o.merge[m].ref[1] = 0; o.merge[m].pad = 0;
  }
  for (int l = 0; l < 2; ++l)
    if (f.ref_LX_size[l] > 0) amvp(f, nb, pu.mv_ref[l], l, o.mv_cand[l]);
  o.num_merge = merge_candidates(f, nb, pu, o.merge);
  *out = o;
}

#if defined(__CUDACC__)
__host__
#endif
inline int frame_supported(const kvz_cuda_me_frame &f)
{
  if (f.width < 8 || f.height < 8 || f.width > 16384 || f.height > 16384) return -1;
  if (f.max_merge < 1 || f.max_merge > 5) return -1;
  if (f.used_size < 0 || f.used_size > 16) return -1;
  for (int l = 0; l < 2; ++l) {
    if (f.ref_LX_size[l] < 0 || f.ref_LX_size[l] > 16) return -1;
    for (int i = 0; i < 16; ++i)
      if (f.ref_LX[l][i] > 15) return -1;
  }
  return 0;
}

}  // namespace kvzme
