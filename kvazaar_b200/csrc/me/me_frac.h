// me_frac.h -- fractional motion search of one PU (half- then quarter-pel around the integer MV), decision for decision
// as the reference takes them.
//
// Single source (device: one warp per PU; host test build: the 32 lane shares walked in turn).  The reference filters
// four whole blocks per step with intermediates shared between its four filter stages
// (kvz_filter_hpel/qpel_blocks_*_luma); what those stages produce is, sample for sample, the standard's luma
// interpolation at the candidate MV (the same arithmetic as kvz_sample_quarterpel_luma, src/strategies/generic/
// ipol-generic.c), so here every lane interpolates the 8x8 / 4x4 sub-blocks of its share directly from the reference
// picture into registers, takes their Hadamard cost and the warp adds the shares up -- no staging buffers, no barrier.
//
// What it follows in the reference (restated, nothing copied):
//   search_frac                     src/search_inter.c:974-1168 (incl. the unsigned cost accumulator it adds the MV cost into)
//   kvz_get_extended_block          border samples = nearest picture sample
//   satd_any_size                   src/strategies/strategies-picture.h:76-112 (integer position)
//   satd_any_size_quad              src/strategies/generic/picture-generic.c:404-471 (fractional positions; with its
//                                   handling of heights that are 4 mod 8: rows 0-7 again instead of rows 4-11)
//   hadamard_4x4 / satd_8x8_subblock   picture-generic.c:117-199, 252-338 (sum of |2-D Walsh-Hadamard|, (s+1)>>1 and (s+2)>>2)
//   calc_mvd_cost                   src/search_inter.c:394-433, fracmv_within_tile :94-181
#pragma once
#include "me_search.h"

namespace kvzme {

// sum of absolute values of the 2-D Walsh-Hadamard transform of an n x n block (n = 4 or 8), d is overwritten
template <int N>
ME_FN uint32_t wht_abs_sum(int32_t *d)
{
  for (int r = 0; r < N; ++r) {               // rows
    int32_t *v = d + r * N;
    for (int half = 1; half < N; half *= 2)
      for (int i = 0; i < N; i += 2 * half)
        for (int j = i; j < i + half; ++j) {
          const int32_t a = v[j], b = v[j + half];
          v[j] = a + b;
          v[j + half] = a - b;
        }
  }
  for (int c = 0; c < N; ++c)                 // columns
    for (int half = 1; half < N; half *= 2)
      for (int i = 0; i < N; i += 2 * half)
        for (int j = i; j < i + half; ++j) {
          const int32_t a = d[j * N + c], b = d[(j + half) * N + c];
          d[j * N + c] = a + b;
          d[(j + half) * N + c] = a - b;
        }
  uint32_t s = 0;
  for (int i = 0; i < N * N; ++i) s += (uint32_t)(d[i] < 0 ? -d[i] : d[i]);
  return s;
}

// the standard's 8-tap luma filters for the 0, 1/4, 1/2, 3/4 positions and 4-tap chroma filters for the eight 1/8 positions
// (constant memory on the device: the fraction is the same for every lane of a warp, so a lookup is one broadcast)
#if defined(__CUDACC__)
#define ME_TABLE static __device__ __constant__ const
#else
#define ME_TABLE static const
#endif
ME_TABLE int8_t kLumaTaps[4][8] = { { 0, 0, 0, 64, 0, 0, 0, 0 }, { -1, 4, -10, 58, 17, -5, 1, 0 }, { -1, 4, -11, 40, 40, -11, 4, -1 }, { 0, 1, -5, 17, 58, -10, 4, -1 } };
ME_TABLE int8_t kChromaTaps[8][4] = { { 0, 64, 0, 0 }, { -2, 58, 10, -2 }, { -4, 54, 16, -2 }, { -6, 46, 28, -4 }, { -4, 36, 36, -4 }, { -4, 28, 46, -6 }, { -2, 16, 54, -4 }, { -2, 10, 58, -2 } };

ME_FN int luma_tap(int frac, int k) { return kLumaTaps[frac][k]; }
ME_FN int chroma_tap(int frac, int k) { return kChromaTaps[frac][k]; }

// Hadamard cost of the N x N sub-block at (sx, sy) of the PU against the prediction at quarter-pel MV (qx, qy)
template <typename Pix, int N>
ME_FN uint32_t subblock_cost(const kvz_cuda_me_params &p, const kvz_cuda_me_pu &pu, const Planes<Pix> &pl, int sx, int sy, int qx, int qy)
{
  const int bx = pu.x + sx + (qx >> 2), by = pu.y + sy + (qy >> 2);       // integer part (floor), fraction 0..3
  const int fx = qx & 3, fy = qy & 3;
  const int xmax = p.width - 1, ymax = p.height - 1;
  int32_t d[N * N];
  if (fx == 0 && fy == 0) {
    for (int r = 0; r < N; ++r) {
      const int yy = by + r < 0 ? 0 : (by + r > ymax ? ymax : by + r);
      for (int c = 0; c < N; ++c) {
        const int xx = bx + c < 0 ? 0 : (bx + c > xmax ? xmax : bx + c);
        d[r * N + c] = (int32_t)pl.cur[(pu.y + sy + r) * pl.cur_stride + pu.x + sx + c] - (int32_t)pl.ref[yy * pl.ref_stride + xx];
      }
    }
  } else {
    const int shift1 = p.bitdepth - 8, shift3 = 14 - p.bitdepth;
    const int32_t offset23 = 1 << (6 + shift3 - 1), pix_max = (1 << p.bitdepth) - 1;
    int32_t hor[(N + 7) * N];                  // horizontally filtered rows by-3 .. by+N+3
    for (int r = 0; r < N + 7; ++r) {
      const int yy0 = by + r - 3;
      const int yy = yy0 < 0 ? 0 : (yy0 > ymax ? ymax : yy0);
      const Pix *row = pl.ref + yy * pl.ref_stride;
      for (int c = 0; c < N; ++c) {
        int32_t s = 0;
        if (fx == 0) {                         // the integer-column filter is the single tap 64 at offset 0
          const int xx0 = bx + c;
          s = 64 * (int32_t)row[xx0 < 0 ? 0 : (xx0 > xmax ? xmax : xx0)];
        } else {
          for (int k = 0; k < 8; ++k) {
            const int xx0 = bx + c + k - 3;
            const int xx = xx0 < 0 ? 0 : (xx0 > xmax ? xmax : xx0);
            s += luma_tap(fx, k) * (int32_t)row[xx];
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
        int32_t v = ((s + offset23) >> 6) >> shift3;
        v = v < 0 ? 0 : (v > pix_max ? pix_max : v);
        d[r * N + c] = (int32_t)pl.cur[(pu.y + sy + r) * pl.cur_stride + pu.x + sx + c] - v;
      }
  }
  const uint32_t s = wht_abs_sum<N>(d);
  return N == 4 ? (s + 1) >> 1 : (s + 2) >> 2;
}

// One lane's share of the Hadamard cost of the whole PU at (qx, qy): the sub-blocks are numbered in the order the
// reference visits them and dealt out round-robin.  quad = the four-candidate variant used for the fractional positions.
template <typename Pix>
ME_FN uint32_t pu_satd_lane(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_pu &pu, const Planes<Pix> &pl, int qx, int qy, bool quad,
                            int &k)             // k: running number of the work item, continues over the positions of one step
{
  int w = pu.w, h = pu.h;
  const int wm = w % 8;
  int x0 = 0, y0 = 0;
  uint32_t s = 0;
  if (wm != 0) {                               // first column in 4x4 blocks
    for (int y = 0; y < h; y += 4)
      if (k++ % ln.n == ln.lane) s += subblock_cost<Pix, 4>(p, pu, pl, 0, y, qx, qy);
    x0 = 4;
    w -= 4;
  }
  if (h % 8 != 0) {                            // first row in 4x4 blocks
    // the single-candidate function continues right of the column strip; the four-candidate one starts at column 0 again
    const int xs = quad ? 0 : x0;
    for (int x = 0; x < w; x += 4)
      if (k++ % ln.n == ln.lane) s += subblock_cost<Pix, 4>(p, pu, pl, xs + x, 0, qx, qy);
    y0 = 4;
    h -= 4;
  }
  // the rest in 8x8 blocks; the four-candidate function restarts at row (h % 8) of the ORIGINAL block, i.e. row 0
  const int ys = quad ? h % 8 : y0;
  for (int y = 0; y < h - (quad ? h % 8 : 0); y += 8)
    for (int x = 0; x < w; x += 8)
      if (k++ % ln.n == ln.lane) s += subblock_cost<Pix, 8>(p, pu, pl, x0 + x, ys + y, qx, qy);
  return s;
}

template <typename Pix>
ME_FN uint32_t pu_satd(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_pu &pu, const Planes<Pix> &pl, int qx, int qy, bool quad)
{
#if defined(__CUDA_ARCH__)
  int k = 0;
  return lane_sum(pu_satd_lane(ln, p, pu, pl, qx, qy, quad, k)) >> (p.bitdepth - 8);
#else
  uint32_t s = 0;
  for (int l = 0; l < ln.n; ++l) {
    int k = 0;
    s += pu_satd_lane(Lanes{ l, ln.n }, p, pu, pl, qx, qy, quad, k);
  }
  return s >> (p.bitdepth - 8);
#endif
}

// The (up to) four candidate positions of one step at once: the sub-blocks of all of them are dealt out round-robin
// together, so a 16x16 PU keeps 16 lanes busy instead of 4; one butterfly per position adds the shares up.
template <typename Pix>
ME_FN void pu_satd_step_lane(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_pu &pu, const Planes<Pix> &pl, const int qx[4], const int qy[4],
                             const bool use[4], uint32_t s[4])
{
  int k = 0;
  for (int j = 0; j < 4; ++j) s[j] = use[j] ? pu_satd_lane(ln, p, pu, pl, qx[j], qy[j], true, k) : 0u;
}

template <typename Pix>
ME_FN void pu_satd_step(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_pu &pu, const Planes<Pix> &pl, const int qx[4], const int qy[4],
                        const bool use[4], uint32_t costs[4])
{
#if defined(__CUDA_ARCH__)
  uint32_t s[4];
  pu_satd_step_lane(ln, p, pu, pl, qx, qy, use, s);
  for (int j = 0; j < 4; ++j) costs[j] = lane_sum(s[j]) >> (p.bitdepth - 8);
#else
  for (int j = 0; j < 4; ++j) costs[j] = 0;
  for (int l = 0; l < ln.n; ++l) {
    uint32_t s[4];
    pu_satd_step_lane(Lanes{ l, ln.n }, p, pu, pl, qx, qy, use, s);
    for (int j = 0; j < 4; ++j) costs[j] += s[j];
  }
  for (int j = 0; j < 4; ++j) costs[j] >>= (p.bitdepth - 8);
#endif
}

// search_pu_inter_ref with cfg.fme_level == 0 (search_inter.c:1385-1397): no fractional search follows, so the integer
// winner's cost becomes its Hadamard cost (kvz_image_calc_satd, src/image.c:451-510) + bits * lambda_sqrt
template <typename Pix>
ME_FN void search_pu_satd_final(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_pu &pu, const Planes<Pix> &pl, kvz_cuda_me_result *out)
{
  Best best = search_pu_best<Pix>(ln, p, pu, pl);
  if (best.cost < kMaxDouble) {
    const uint32_t satd = pu_satd(ln, p, pu, pl, (best.mvx >> 2) * 4, (best.mvy >> 2) * 4, false);
    best.cost = (double)satd;
    best.cost += (double)best.bits * p.lambda_sqrt;
  }
  write_result(ln, best, out);
}

// calc_mvd_cost for a quarter-pel MV, no merge candidates
ME_FN uint32_t qpel_mv_bits(const kvz_cuda_me_pu &pu, int qx, int qy)
{
  const uint32_t c0 = mvd_bits(qx - pu.mv_cand[0][0], qy - pu.mv_cand[0][1]);
  const uint32_t c1 = mvd_bits(qx - pu.mv_cand[1][0], qy - pu.mv_cand[1][1]);
  return c0 < c1 ? c0 : c1;
}

// search_frac: pu.start_mv is the integer search's best MV (1/4 pel); levels = cfg.fme_level (1..4)
template <typename Pix>
ME_FN void frac_search_pu(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_pu &pu, const Planes<Pix> &pl, int levels,
                          kvz_cuda_me_result *out)
{
  const int sqx[9] = { 0, -1, 1, 0, 0, -1, 1, -1, 1 }, sqy[9] = { 0, 0, 0, -1, 1, -1, -1, 1, 1 };
  if (!pu_valid(p, pu)) {
    Best none;
    none.cost = kMaxDouble; none.bits = kMaxInt; none.mvx = none.mvy = 0; none.points = 0;
    write_result(ln, none, out);
    return;
  }
  int mx = pu.start_mv[0] >> 2, my = pu.start_mv[1] >> 2;
  int32_t points = 1;
  // integer position.  The reference keeps the costs in an unsigned and adds the (double) MV cost into it: truncation.
  uint32_t bits = qpel_mv_bits(pu, mx * 4, my * 4);
  uint32_t c0 = pu_satd(ln, p, pu, pl, mx * 4, my * 4, false);
  c0 = (uint32_t)((double)c0 + (double)bits * p.lambda_sqrt);
  double cost = (double)c0;
  uint32_t bitcost = bits;
  mx *= 2;
  my *= 2;
  int best_index = 0, i = 1;
  for (int step = 0; step < levels; ++step) {
    const int mv_shift = step < 2 ? 1 : 0;
    uint32_t costs[4], cbits[4];
    bool within[4];
    int qx[4], qy[4];
    for (int j = 0; j < 4; ++j) {
      qx[j] = (mx + sqx[i + j]) * (1 << mv_shift);
      qy[j] = (my + sqy[i + j]) * (1 << mv_shift);
      within[j] = mv_allowed(p, pu, qx[j], qy[j]);
    }
    pu_satd_step(ln, p, pu, pl, qx, qy, within, costs);      // the cost of a position that may not be used is never looked at
    for (int j = 0; j < 4; ++j) {
      cbits[j] = 0;
      if (within[j]) {
        cbits[j] = qpel_mv_bits(pu, qx[j], qy[j]);
        costs[j] = (uint32_t)((double)costs[j] + (double)cbits[j] * p.lambda_sqrt);
        ++points;
      }
    }
    for (int j = 0; j < 4; ++j)
      if (within[j] && (double)costs[j] < cost) {
        cost = (double)costs[j];
        bitcost = cbits[j];
        best_index = i + j;
      }
    i += 4;
    if (step == 1 || step == levels - 1) {
      mx += sqx[best_index];
      my += sqy[best_index];
      if (step == (levels - 1 < 1 ? levels - 1 : 1)) {       // last half-pel step: on to quarter-pel units
        mx *= 2;
        my *= 2;
        best_index = 0;
        i = 1;
      }
    }
  }
  if (ln.lane == 0) {
    out->cost = cost;
    out->bits = (int32_t)bitcost;
    out->mv[0] = (int16_t)mx;
    out->mv[1] = (int16_t)my;
    out->points = points;
    out->pad = 0;
  }
}

}  // namespace kvzme
