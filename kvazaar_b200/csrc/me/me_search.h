// me_search.h -- integer motion estimation of one PU, decision for decision as the reference takes them.
//
// Single source: me_search.cu compiles it for the device (one warp per PU: the 32 lanes share the pixels of every
// SAD, the decisions are warp-uniform), tests/hostsim/me_hostsim.cpp compiles it for the host with one "lane"
// (TEST INFRASTRUCTURE: checks the control flow without a GPU).
//
// What it follows in the reference (nothing is copied; the control flow is restated around a lane-parallel SAD):
//   check_mv_cost            src/search_inter.c:202-247
//   fracmv_within_tile       src/search_inter.c:94-181
//   kvz_image_calc_sad       src/image.c:407-447, image_interpolated_sad image.c:279-398
//   get_ep_ex_golomb_bitcost src/search_inter.c:250-270, get_mvd_coding_cost :333-348, select_mv_cand :351-391,
//   calc_mvd_cost            :394-433 (num_cand = 0 in the integer stage: no merge shortcut)
//   select_starting_point    :297-330, mv_in_merge :277-290
//   early_terminate          :436-485
//   hexagon_search           :712-792, diamond_search :812-888, kvz_tz_pattern_search :486-604 (diamond pattern),
//   tz_search                :623-697, search_mv_full :891-964
//   search_pu_inter_ref      :1349-1383 (the order of the three stages)
#pragma once
#include <stdint.h>

#include "../../../include/kvz_cuda.h"

#if defined(__CUDACC__)
#define ME_FN __device__ __forceinline__
#else
#define ME_FN static inline
#endif

namespace kvzme {

static_assert(sizeof(kvz_cuda_me_params) == 56 && sizeof(kvz_cuda_me_merge) == 12 && sizeof(kvz_cuda_me_pu) == 84 && sizeof(kvz_cuda_me_result) == 24,
              "record layouts are part of the ABI (tests/test_me_search.py mirrors them)");

struct Lanes {
  int lane;   // this thread's index among the PU's lanes
  int n;      // lanes per PU: 32 (a warp; the host build walks the 32 shares in turn)
};

ME_FN uint32_t lane_sum(uint32_t v)
{
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
  return v;
}

constexpr int kLcuWidth = 64;
constexpr double kMaxDouble = 1.7e+308;   // MAX_DOUBLE (global.h:293)
constexpr int32_t kMaxInt = 0x7FFFFFFF;   // MAX_INT (global.h:287)

template <typename Pix>
struct Planes {
  const Pix *cur;
  const Pix *ref;
  int cur_stride, ref_stride;
};

struct Best {
  double cost;
  int32_t bits;
  int mvx, mvy;      // 1/4 pel
  int32_t points;
};

// A PU record the device can work on: inside the picture, 4..64 samples in steps of 4, at most five merge candidates.
// (The reference asserts the same about its callers; a C ABI must not fault on a bad record.)
ME_FN bool pu_valid(const kvz_cuda_me_params &p, const kvz_cuda_me_pu &pu)
{
  return pu.w >= 4 && pu.h >= 4 && pu.w <= 64 && pu.h <= 64 && (pu.w & 3) == 0 && (pu.h & 3) == 0 && pu.x >= 0 && pu.y >= 0 &&
         pu.x + pu.w <= p.width && pu.y + pu.h <= p.height && pu.num_merge >= 0 && pu.num_merge <= 5;
}

// fracmv_within_tile: may the block at MV (x, y) (1/4 pel) be referenced?
ME_FN bool mv_allowed(const kvz_cuda_me_params &p, const kvz_cuda_me_pu &pu, int x, int y)
{
  const bool frac_luma = x % 4 != 0 || y % 4 != 0;
  const bool frac_chroma = x % 8 != 0 || y % 8 != 0;
  if (p.wpp_owf) {
    int margin = frac_luma ? 4 : (frac_chroma ? 2 : 0);
    margin += p.delay_px;
    const int lcu_x = pu.x / kLcuWidth, lcu_y = pu.y / kLcuWidth;
    const int dx = ((pu.x + pu.w + margin) * 4 + x) / (kLcuWidth << 2) - lcu_x;   // C division: truncates toward zero
    const int dy = ((pu.y + pu.h + margin) * 4 + y) / (kLcuWidth << 2) - lcu_y;
    if (dy > p.max_ref_lcu_down) return false;
    if (dx + dy > p.max_ref_lcu_down + p.max_ref_lcu_right) return false;
  }
  if (p.mv_constraint == 0) return true;
  int margin = 0;
  if (p.mv_constraint == 4) margin = frac_luma ? (4 << 2) : (frac_chroma ? (2 << 2) : 0);
  const int ax = pu.x * 4 + x, ay = pu.y * 4 + y;
  const int from_right = (p.width << 2) - (ax + (pu.w << 2));
  const int from_bottom = (p.height << 2) - (ay + (pu.h << 2));
  return ax >= margin && ay >= margin && from_right >= margin && from_bottom >= margin;
}

// One lane's share of the SAD of the PU against the reference block at full-pel offset (x, y): pixels lane, lane + n,
// lane + 2n, ... of the block in raster order.  Reference pixels outside the picture are the nearest edge pixels (what
// hor_sad / ver_sad / cor_sad of image_interpolated_sad add up to).
template <typename Pix>
ME_FN uint32_t pu_sad_lane(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_pu &pu, const Planes<Pix> &pl, int x, int y)
{
  const int w = pu.w, h = pu.h;
  const int rx = pu.x + x, ry = pu.y + y;
  const bool inside = rx >= 0 && rx <= p.width - w && ry >= 0 && ry <= p.height - h;
  uint32_t s = 0;
  int c = ln.lane % w, r = ln.lane / w;
  if (inside) {
    while (r < h) {
      const int a = pl.cur[(pu.y + r) * pl.cur_stride + pu.x + c];
      const int b = pl.ref[(ry + r) * pl.ref_stride + rx + c];
      s += (uint32_t)(a > b ? a - b : b - a);
      c += ln.n;
      while (c >= w) { c -= w; ++r; }
    }
  } else {
    const int xmax = p.width - 1, ymax = p.height - 1;
    while (r < h) {
      int xx = rx + c, yy = ry + r;
      xx = xx < 0 ? 0 : (xx > xmax ? xmax : xx);
      yy = yy < 0 ? 0 : (yy > ymax ? ymax : yy);
      const int a = pl.cur[(pu.y + r) * pl.cur_stride + pu.x + c];
      const int b = pl.ref[yy * pl.ref_stride + xx];
      s += (uint32_t)(a > b ? a - b : b - a);
      c += ln.n;
      while (c >= w) { c -= w; ++r; }
    }
  }
  return s;
}

// kvz_image_calc_sad.  Device: the warp's lanes each take their share, a shuffle butterfly leaves the total in every
// lane.  Host build: the same shares, walked one after the other.
template <typename Pix>
ME_FN uint32_t pu_sad(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_pu &pu, const Planes<Pix> &pl, int x, int y)
{
#if defined(__CUDA_ARCH__)
  return lane_sum(pu_sad_lane(ln, p, pu, pl, x, y)) >> (p.bitdepth - 8);
#else
  uint32_t s = 0;
  for (int l = 0; l < ln.n; ++l) s += pu_sad_lane(Lanes{ l, ln.n }, p, pu, pl, x, y);
  return s >> (p.bitdepth - 8);
#endif
}

ME_FN uint32_t golomb_bits(uint32_t symbol)
{
  uint32_t bins = 0;
  if (symbol >= 1u << 8) { bins += 16; symbol >>= 8; }
  if (symbol >= 1u << 4) { bins += 8; symbol >>= 4; }
  if (symbol >= 1u << 2) { bins += 4; symbol >>= 2; }
  if (symbol >= 1u << 1) { bins += 2; }
  return bins;
}

// get_mvd_coding_cost: whole bits (the reference accumulates them << CTX_FRAC_BITS and divides again: exact)
ME_FN uint32_t mvd_bits(int mvd_x, int mvd_y)
{
  const uint32_t ax = (uint32_t)(mvd_x < 0 ? -mvd_x : mvd_x), ay = (uint32_t)(mvd_y < 0 ? -mvd_y : mvd_y);
  return 4 + (ax == 1) + (ay == 1) + golomb_bits(ax) + golomb_bits(ay);
}

// calc_mvd_cost with mv_shift = 2 and no merge candidates: the cheaper of the two AMVP candidates
ME_FN uint32_t mv_bits(const kvz_cuda_me_pu &pu, int x, int y)
{
  const int qx = x * 4, qy = y * 4;
  const uint32_t c0 = mvd_bits(qx - pu.mv_cand[0][0], qy - pu.mv_cand[0][1]);
  const uint32_t c1 = mvd_bits(qx - pu.mv_cand[1][0], qy - pu.mv_cand[1][1]);
  return c0 < c1 ? c0 : c1;
}

// check_mv_cost: full-pel (x, y)
template <typename Pix>
ME_FN bool check_mv(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_pu &pu, const Planes<Pix> &pl, int x, int y, Best &best)
{
  if (!mv_allowed(p, pu, x * 4, y * 4)) return false;
  double cost = (double)pu_sad(ln, p, pu, pl, x, y);
  best.points += 1;
  if (cost + 0.001 >= best.cost) return false;
  const uint32_t bits = mv_bits(pu, x, y);
  cost += (double)bits * p.lambda_sqrt;
  if (cost + 0.001 >= best.cost) return false;
  best.mvx = x * 4;
  best.mvy = y * 4;
  best.cost = cost;
  best.bits = (int32_t)bits;
  return true;
}

ME_FN bool merge_mv(const kvz_cuda_me_merge &m, int &x, int &y)   // the candidate's MV rounded to full-pel; false for bi candidates
{
  if (m.dir == 3) return false;
  x = (m.mv[m.dir - 1][0] + 2) >> 2;
  y = (m.mv[m.dir - 1][1] + 2) >> 2;
  return true;
}

// select_starting_point: the 0-vector, the start MV (unless a merge candidate covers it), the merge candidates
template <typename Pix>
ME_FN void select_start(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_pu &pu, const Planes<Pix> &pl, int start_x, int start_y,
                        Best &best)
{
  check_mv(ln, p, pu, pl, 0, 0, best);
  const int ex = start_x >> 2, ey = start_y >> 2;       // the start MV by value: what (0,0) did to best_mv does not matter here
  if (ex != 0 || ey != 0) {
    bool in_merge = false;
    for (int i = 0; i < pu.num_merge; ++i) {
      int mx, my;
      if (merge_mv(pu.merge[i], mx, my) && mx == ex && my == ey) { in_merge = true; break; }
    }
    if (!in_merge) check_mv(ln, p, pu, pl, ex, ey, best);
  }
  for (int i = 0; i < pu.num_merge; ++i) {
    int mx, my;
    if (!merge_mv(pu.merge[i], mx, my)) continue;
    if (mx == 0 && my == 0) continue;
    check_mv(ln, p, pu, pl, mx, my, best);
  }
}

// early_terminate: two rounds of the small cross around the best MV; true = the pattern search is skipped
template <typename Pix>
ME_FN bool early_terminate(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_pu &pu, const Planes<Pix> &pl, Best &best)
{
  const int sx[7] = { 0, -1, 0, 1, 0, -1, 0 }, sy[7] = { -1, 0, 1, 0, -1, 0, 0 };
  int mx = best.mvx >> 2, my = best.mvy >> 2;
  int first = 0, last = 3;
  for (int k = 0; k < 2; ++k) {
    const double threshold = p.me_early_termination == 2 ? best.cost * 0.95 : best.cost;
    int best_index = 6;
    for (int i = first; i <= last; ++i)
      if (check_mv(ln, p, pu, pl, mx + sx[i], my + sy[i], best)) best_index = i;
    mx += sx[best_index];
    my += sy[best_index];
    if (best.cost >= threshold) return true;
    first = (best_index + 3) % 4;
    last = first + 2;
  }
  return false;
}

template <typename Pix>
ME_FN void hexagon_search(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_pu &pu, const Planes<Pix> &pl, Best &best)
{
  // large hexagon, its first two points repeated so that three consecutive entries are the new points of a move
  const int lx[9] = { 0, 1, 2, 1, -1, -2, -1, 1, 2 }, ly[9] = { 0, -2, 0, 2, 2, 0, -2, -2, 0 };
  const int qx[9] = { 0, 0, -1, 1, 0, -1, 1, -1, 1 }, qy[9] = { 0, -1, 0, 0, 1, -1, -1, 1, 1 };
  uint32_t steps = (uint32_t)p.me_max_steps;
  int mx = best.mvx >> 2, my = best.mvy >> 2;
  int best_index = 0;
  for (int i = 1; i < 7; ++i)
    if (check_mv(ln, p, pu, pl, mx + lx[i], my + ly[i], best)) best_index = i;
  while (best_index != 0 && steps != 0) {
    if (steps > 0) steps -= 1;
    const int start = best_index == 1 ? 6 : (best_index == 8 ? 1 : best_index - 1);
    mx += lx[best_index];
    my += ly[best_index];
    best_index = 0;
    for (int i = 0; i < 3; ++i)
      if (check_mv(ln, p, pu, pl, mx + lx[start + i], my + ly[start + i], best)) best_index = start + i;
  }
  // the centre is NOT moved to the last best point before the small pattern (the reference has that move commented out)
  for (int i = 1; i < 9; ++i) check_mv(ln, p, pu, pl, mx + qx[i], my + qy[i], best);
}

template <typename Pix>
ME_FN void diamond_search(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_pu &pu, const Planes<Pix> &pl, Best &best)
{
  const int dx[5] = { 0, 1, 0, -1, 0 }, dy[5] = { -1, 0, 1, 0, 0 };
  uint32_t steps = (uint32_t)p.me_max_steps;
  int mx = best.mvx >> 2, my = best.mvy >> 2;
  int best_index = 4;
  for (int i = 0; i < 5; ++i)
    if (check_mv(ln, p, pu, pl, mx + dx[i], my + dy[i], best)) best_index = i;
  if (best_index == 4) return;
  mx += dx[best_index];
  my += dy[best_index];
  int from_dir = 4;
  bool better;
  do {
    better = false;
    if (steps > 0) steps -= 1;
    for (int i = 0; i < 4; ++i) {
      if (i == from_dir) continue;
      if (check_mv(ln, p, pu, pl, mx + dx[i], my + dy[i], best)) { best_index = i; better = true; }
    }
    if (better) {
      mx += dx[best_index];
      my += dy[best_index];
      from_dir = best_index ^ 0x3;
    }
  } while (better && steps != 0);
}

// kvz_tz_pattern_search with the diamond pattern (type 0, the only one tz_search uses): 4 points at distance 1,
// else the 4 axis points and the 4 half-distance diagonal points
template <typename Pix>
ME_FN void tz_diamond(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_pu &pu, const Planes<Pix> &pl, int dist, int cx, int cy,
                      int &best_dist, Best &best)
{
  const int hd = dist / 2;
  const int ox[8] = { 0, dist, 0, -dist, hd, hd, -hd, -hd }, oy[8] = { dist, 0, -dist, 0, hd, -hd, -hd, hd };
  const int n_points = dist == 1 ? 4 : 8;
  bool improved = false;
  for (int i = 0; i < n_points; ++i)
    if (check_mv(ln, p, pu, pl, cx + ox[i], cy + oy[i], best)) improved = true;
  if (improved) best_dist = dist;
}

// tz_search (search_inter.c:623-697) with the reference's fixed parameters: range 96, diamond grid search from the
// start MV and again from the 0-vector, no raster step, star refinement until a round brings nothing
template <typename Pix>
ME_FN void tz_search(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_pu &pu, const Planes<Pix> &pl, Best &best)
{
  const int range = 96;
  int best_dist = 0;
  int sx = best.mvx >> 2, sy = best.mvy >> 2;
  int idle = 0;
  for (int dist = 1; dist <= range; dist *= 2) {
    tz_diamond(ln, p, pu, pl, dist, sx, sy, best_dist, best);
    if (best_dist != dist) ++idle;
    if (idle >= 3) break;
  }
  if (sx != 0 || sy != 0) {
    sx = sy = 0;
    idle = 0;
    for (int dist = 1; dist <= range / 2; dist *= 2) {
      tz_diamond(ln, p, pu, pl, dist, sx, sy, best_dist, best);
      if (best_dist != dist) ++idle;
      if (idle >= 3) break;
    }
  }
  while (best_dist > 0) {
    best_dist = 0;
    sx = best.mvx >> 2;
    sy = best.mvy >> 2;
    for (int dist = 1; dist <= range; dist *= 2) tz_diamond(ln, p, pu, pl, dist, sx, sy, best_dist, best);
  }
}

// search_mv_full (search_inter.c:891-964): the square around the 0-vector, around the MV found so far (unless a merge
// candidate rounds to it), and around every merge candidate, skipping what an earlier square of this last stage covered
template <typename Pix>
ME_FN void full_search(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_pu &pu, const Planes<Pix> &pl, int range, Best &best)
{
  const int ex = best.mvx >> 2, ey = best.mvy >> 2;       // extra_mv: best_mv by value at the call
  for (int y = -range; y <= range; ++y)
    for (int x = -range; x <= range; ++x) check_mv(ln, p, pu, pl, x, y, best);
  bool in_merge = false;
  for (int i = 0; i < pu.num_merge; ++i) {
    int mx, my;
    if (merge_mv(pu.merge[i], mx, my) && mx == ex && my == ey) { in_merge = true; break; }
  }
  if (!in_merge)
    for (int y = -range; y <= range; ++y)
      for (int x = -range; x <= range; ++x) check_mv(ln, p, pu, pl, ex + x, ey + y, best);
  for (int i = 0; i < pu.num_merge; ++i) {
    const kvz_cuda_me_merge &m = pu.merge[i];
    if (m.dir == 3) continue;
    const int mx = m.mv[m.dir - 1][0] >> 2, my = m.mv[m.dir - 1][1] >> 2;      // truncating here, rounding in select_start: as the reference
    if (mx == 0 && my == 0) continue;
    for (int y = my - range; y <= my + range; ++y)
      for (int x = mx - range; x <= mx + range; ++x) {
        if (!mv_allowed(p, pu, x * 4, y * 4)) continue;
        bool tested = false;
        for (int j = -1; j < i; ++j) {
          int xx = 0, yy = 0;
          if (j >= 0) {
            if (pu.merge[j].dir == 3) continue;
            xx = pu.merge[j].mv[pu.merge[j].dir - 1][0] >> 2;
            yy = pu.merge[j].mv[pu.merge[j].dir - 1][1] >> 2;
          }
          if (x >= xx - range && x <= xx + range && y >= yy - range && y <= yy + range) {
            tested = true;
            x = xx + range;                 // jump to the right edge of the covered square
            break;
          }
        }
        if (tested) continue;
        check_mv(ln, p, pu, pl, x, y, best);
      }
  }
}

// the integer stage of search_pu_inter_ref for one PU; every lane ends with the same result
template <typename Pix>
ME_FN Best search_pu_best(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_pu &pu, const Planes<Pix> &pl)
{
  Best best;
  best.cost = kMaxDouble;
  best.bits = kMaxInt;
  best.points = 0;
  best.mvx = best.mvy = 0;
  if (!pu_valid(p, pu)) return best;       // "no point was allowed"
  int sx = pu.start_mv[0], sy = pu.start_mv[1];
  if (!mv_allowed(p, pu, sx, sy)) { sx = 0; sy = 0; }      // search_inter.c:1334-1337
  best.mvx = sx;
  best.mvy = sy;
  select_start(ln, p, pu, pl, sx, sy, best);
  const bool skip = early_terminate(ln, p, pu, pl, best);
  if (!(p.me_early_termination && skip)) {
    switch (p.ime_algorithm) {               // enum kvz_ime_algorithm (kvazaar.h:110-119), search_inter.c:1360-1382
      case 1: tz_search(ln, p, pu, pl, best); break;
      case 2: case 5: full_search(ln, p, pu, pl, 32, best); break;
      case 3: full_search(ln, p, pu, pl, 8, best); break;
      case 4: full_search(ln, p, pu, pl, 16, best); break;
      case 6: full_search(ln, p, pu, pl, 64, best); break;
      case 7: diamond_search(ln, p, pu, pl, best); break;
      default: hexagon_search(ln, p, pu, pl, best); break;
    }
  }
  return best;
}

ME_FN void write_result(const Lanes &ln, const Best &best, kvz_cuda_me_result *out)
{
  if (ln.lane == 0) {
    out->cost = best.cost;
    out->bits = best.bits;
    out->mv[0] = (int16_t)best.mvx;
    out->mv[1] = (int16_t)best.mvy;
    out->points = best.points;
    out->pad = 0;
  }
}

template <typename Pix>
ME_FN void search_pu(const Lanes &ln, const kvz_cuda_me_params &p, const kvz_cuda_me_pu &pu, const Planes<Pix> &pl, kvz_cuda_me_result *out)
{
  write_result(ln, search_pu_best<Pix>(ln, p, pu, pl), out);
}

#if defined(__CUDACC__)
__host__
#endif
inline int params_supported(const kvz_cuda_me_params &p)
{
  if (p.width < 8 || p.height < 8 || p.width > 16384 || p.height > 16384) return -1;
  if (p.bitdepth != 8 && p.bitdepth != 10) return -1;
  if (p.ime_algorithm < 0 || p.ime_algorithm > 7) return -1;
  if (p.me_early_termination < 0 || p.me_early_termination > 2) return -1;
  if (p.mv_constraint < 0 || p.mv_constraint > 4) return -1;
  return 0;
}

}  // namespace kvzme
