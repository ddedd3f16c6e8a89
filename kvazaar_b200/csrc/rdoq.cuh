// rdoq.cuh -- rate-distortion optimised quantisation of one TU (ref: kvz_rdoq, src/rdo.c:661-977,
// kvz_get_coded_level :395-452, kvz_get_ic_rate :346-393, calc_last_bits/get_rate_last :465-508,
// kvz_rdoq_sign_hiding :518-653, find_last_scanpos_generic quant-generic.c:376-399, context derivation
// src/context.c:315-397).
//
// The algorithm is HM's serial trellis-free RDOQ: walking the scan backwards, every coefficient picks the level
// (of at most two candidates) with the least  distortion + lambda * CABAC-rate, where the rate of a level depends on
// the greater-than-1/2 context set and Rice parameter that the previously decided levels of the coefficient group
// left behind.  That chain is serial per TU; parallelism is across TUs (a warp per TU; lane 0 walks the chain,
// all lanes do the position-parallel parts).  Costs are doubles evaluated in the reference's operation order; the
// library is built with -fmad=false, and the reference's rdo.c is plain x86-64 code without FMA contraction.
#pragma once
#include "common.cuh"
#include "transform.cuh"

namespace kvzc {

// HM's fractional-bit estimates per CABAC state (15 fractional bits), split by symbol = MPS / LPS
// (kvz_entropy_bits, rdo.c:69-79: entry [2s] is the MPS cost of state s, [2s + 1] the LPS cost)
static __device__ const int32_t c_ebits_mps[64] = {
  32768, 30426, 28306, 26378, 24617, 23005, 21523, 20159, 18899, 17734, 16653, 15650, 14717, 13849, 13038, 12282,
  11575, 10914, 10294, 9714, 9169, 8658, 8178, 7727, 7303, 6903, 6527, 6173, 5840, 5525, 5228, 4948,
  4684, 4435, 4199, 3977, 3767, 3568, 3380, 3202, 3034, 2876, 2725, 2583, 2448, 2321, 2200, 2086,
  1978, 1875, 1778, 1686, 1599, 1517, 1439, 1364, 1294, 1228, 1165, 1105, 1048, 994, 943, 895 };
static __device__ const int32_t c_ebits_lps[64] = {
  32768, 35232, 37696, 40159, 42623, 45087, 47551, 50015, 52479, 54942, 57406, 59870, 62334, 64798, 67262, 69725,
  72189, 74653, 77117, 79581, 82044, 84508, 86972, 89436, 91900, 94363, 96827, 99291, 101755, 104219, 106683, 109146,
  111610, 114074, 116538, 119002, 121465, 123929, 126393, 128857, 131321, 133785, 136248, 138712, 141176, 143640, 146104, 148568,
  151031, 153495, 155959, 158423, 160887, 163351, 165814, 168278, 170742, 173207, 175669, 178134, 180598, 183061, 185525, 187989 };

// The 128-entry table in the reference's indexing (state byte ^ bin), staged in shared memory by the kernels: lanes
// look up different states at the same time, which constant memory would serialise.
__device__ __forceinline__ void rdoq_load_ebits(int32_t *table /* [128] */)
{
  for (int i = threadIdx.x; i < 128; i += blockDim.x) table[i] = (i & 1) ? c_ebits_lps[i >> 1] : c_ebits_mps[i >> 1];
}

constexpr int RDOQ_ONE_BIT = 1 << 15;

// position group of a last-significant coordinate (g_group_idx, rdo.c: 0,1,2,3,4,4,5,5,6,6,6,6,7,7,7,7,8..,9..)
__device__ __forceinline__ int last_group(int x) { if (x < 4) return x; const int l = 31 - __clz(x); return 2 * l + ((x >> (l - 1)) & 1); }

// scan position -> raster position of the 32x32 diagonal scan, shared by every 32x32 TU (filled once per process by
// rdoq_init_tables); smaller blocks build their table in shared memory
static __device__ uint16_t g_scan_diag32[1024];
static __global__ void rdoq_init_scan32_kernel() { g_scan_diag32[threadIdx.x] = (uint16_t)scan_pos(0, 5, threadIdx.x); }

// SH = sign hiding enabled: only then the per-position rate tables of kvz_sh_rates_t exist
// SH = sign hiding enabled: only then the per-position rate tables of kvz_sh_rates_t exist.
// cost_sig (rdo.c:683) is kept as one byte per position: the significance context (6 bits) and which of its two bin
// costs the position carries (0: flag = 0, 1: flag = 1, 2: none -- the last coefficient or a zeroed group); the double
// is lambda * entropy_bits again when the last-position search needs it.  That halves the table and lets all 32x32 TUs
// of a 1080p frame be resident at once.
template <int NN, bool SH>
struct RdoqScratch {
  double cost_coeff[NN];                                   // (the level-0 distortion cost_coeff0 is recomputed where needed)
  uint8_t sig_code[NN];
  int32_t inc[SH ? NN : 1], dec[SH ? NN : 1], sig_inc[SH ? NN : 1], qdelta[SH ? NN : 1];       // kvz_sh_rates_t (rdo.h:49-58)
  uint16_t blk[NN >= 1024 ? 1 : NN];                       // scan position -> raster position (32x32: g_scan_diag32)
  double cg_sig_cost[NN / 16];
  int32_t cg_flag[NN / 16];
  uint16_t cg_nz[NN / 16];                                 // per group: positions (bit k) whose level is non-zero
  int32_t last_x_bits[12], last_y_bits[12];
  // per coefficient group, filled by lanes 0..15 before lane 0 walks the group
  double prep_c0[16], prep_sig0[16], prep_sig1[16];
  int32_t prep_ld[16], prep_ctx_sig[16];
  int32_t best_last_p1;
};

struct RdoqModels {      // views into the kvz_cuda_cabac_ctx image for one texture type
  const uint8_t *sig, *one, *abs, *cg, *last_x, *last_y, *cbf;
  const int32_t *eb;     // entropy-bit table, see rdoq_load_ebits
  uint8_t root_cbf;
};
// cost in 1/32768 bits of coding `bin` with the context whose state byte is `st` (bit 0 = MPS value)
#define ebits(st, bin) (m.eb[(st) ^ (bin)])

__device__ __forceinline__ RdoqModels rdoq_models(const kvz_cuda_cabac_ctx *c, const int32_t *eb, int type)
{
  RdoqModels m;
  m.eb = eb;
  m.sig = type ? c->cu_sig_model_chroma : c->cu_sig_model_luma;
  m.one = type ? c->cu_one_model_chroma : c->cu_one_model_luma;
  m.abs = type ? c->cu_abs_model_chroma : c->cu_abs_model_luma;
  m.cg = c->cu_sig_coeff_group_model + type;               // rdo.c:720: indexed by type (0 luma, 2 chroma)
  m.last_x = type ? c->cu_ctx_last_x_chroma : c->cu_ctx_last_x_luma;
  m.last_y = type ? c->cu_ctx_last_y_chroma : c->cu_ctx_last_y_luma;
  m.cbf = type ? c->qt_cbf_model_chroma : c->qt_cbf_model_luma;
  m.root_cbf = c->cu_qt_root_cbf_model;
  return m;
}

// rate of the level bins of |level| given the greater1 / greater2 contexts and the Rice parameter (rdo.c:346-393)
__device__ __forceinline__ int rdoq_level_rate(const RdoqModels &m, uint32_t abs_level, int ctx_one, int ctx_abs, int rice, uint32_t c1_idx, uint32_t c2_idx)
{
  int rate = RDOQ_ONE_BIT;                                               // the sign bin
  const uint32_t base_level = (c1_idx < 8) ? (2 + (c2_idx < 1)) : 1;
  if (abs_level >= base_level) {
    int symbol = (int)(abs_level - base_level);
    if (symbol < (3 << rice)) {
      rate += ((symbol >> rice) + 1 + rice) * RDOQ_ONE_BIT;
    } else {
      int length = rice;
      symbol -= 3 << rice;
      while (symbol >= (1 << length)) symbol -= 1 << (length++);
      rate += (3 + length + 1 - rice + length) * RDOQ_ONE_BIT;
    }
    if (c1_idx < 8) {
      rate += ebits(m.one[ctx_one], 1);
      if (c2_idx < 1) rate += ebits(m.abs[ctx_abs], 1);
    }
  } else if (abs_level == 1) {
    rate += ebits(m.one[ctx_one], 0);
  } else if (abs_level == 2) {
    rate += ebits(m.one[ctx_one], 1);
    rate += ebits(m.abs[ctx_abs], 0);
  }
  return rate;
}

// best level of one coefficient (rdo.c:413-452); cost0 = distortion of level 0
__device__ __forceinline__ uint32_t rdoq_pick_level(const RdoqModels &m, double lambda, double &cost, double cost0, double &cost_sig, int level_double,
                                                    uint32_t max_abs_level, int ctx_sig, int ctx_one, int ctx_abs, int rice, uint32_t c1_idx,
                                                    uint32_t c2_idx, int q_bits, double err_scale, bool last)
{
  double sig_cost_now = 0;
  uint32_t best = 0;
  if (!last && max_abs_level < 3) {
    cost_sig = lambda * ebits(m.sig[ctx_sig], 0);
    cost = cost0 + cost_sig;
    if (max_abs_level == 0) return 0;
  } else {
    cost = 1.7e+308;
  }
  if (!last) sig_cost_now = lambda * ebits(m.sig[ctx_sig], 1);
  const int lo = max_abs_level > 1 ? (int)max_abs_level - 1 : 1;
  for (int lvl = (int)max_abs_level; lvl >= lo; --lvl) {
    const double err = (double)(level_double - lvl * (1 << q_bits));
    double c = err * err * err_scale + lambda * rdoq_level_rate(m, (uint32_t)lvl, ctx_one, ctx_abs, rice, c1_idx, c2_idx);
    c += sig_cost_now;
    if (c < cost) { best = (uint32_t)lvl; cost = c; cost_sig = sig_cost_now; }
  }
  return best;
}

// context increment of sig_coeff_flag (context.c:366-397)
__device__ __forceinline__ int rdoq_sig_ctx(int pattern, int scan_idx, int px, int py, int log2n, int type)
{
  if (px + py == 0) return 0;
  if (log2n == 2) { const int map[16] = { 0, 1, 4, 5, 2, 3, 4, 5, 6, 6, 8, 8, 7, 7, 8, 8 }; return map[4 * py + px]; }
  const int offset = (log2n == 3) ? (scan_idx == 0 ? 9 : 15) : (type == 0 ? 21 : 12);
  const int sx = px & 3, sy = py & 3;
  int cnt;
  if (pattern == 0) cnt = (sx + sy <= 2) ? ((sx + sy == 0) ? 2 : 1) : 0;
  else if (pattern == 1) cnt = (sy <= 1) ? ((sy == 0) ? 2 : 1) : 0;
  else if (pattern == 2) cnt = (sx <= 1) ? ((sx == 0) ? 2 : 1) : 0;
  else cnt = 2;
  return ((type == 0 && ((px >> 2) + (py >> 2)) > 0) ? 3 : 0) + offset + cnt;
}

// Sign-bit hiding on RDOQ output (rdo.c:518-653).  Serial; one thread.
template <class Scratch, class BlkT>
__device__ void rdoq_sign_hiding(const Scratch &s, const BlkT *blk, double lambda, int bitdepth, int qp_scaled, int scan_idx, int log2n, int last_pos,
                                 const int16_t *coef, int16_t *q)
{
  const int inv_quant = c_inv_quant_scales[qp_scaled % 6];
  const long long rd_factor = (long long)(inv_quant * inv_quant * (1 << (2 * (qp_scaled / 6))) / lambda / 16 / (1 << (2 * (bitdepth - 8))) + 0.5);
  const int last_cg = (last_pos - 1) >> 4;
  for (int cg = last_cg; cg >= 0; --cg) {
    const int base = cg << 4;
    const BlkT *pos = blk + base;
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
      const long long quant_cost = rd_factor * s.qdelta[p];
      const int a = abs((int)q[p]);
      long long cost;
      int change;
      if (a != 0) {
        long long inc_bits = s.inc[p], dec_bits = s.dec[p];
        if (a == 1) dec_bits -= RDOQ_ONE_BIT + s.sig_inc[p];
        if (cg == last_cg && last_nz == k && a == 1) dec_bits -= 4 * RDOQ_ONE_BIT;
        inc_bits = -quant_cost + inc_bits * 1;            // PRECISION_INC = 15 - CTX_FRAC_BITS = 0
        dec_bits = quant_cost + dec_bits * 1;
        if (inc_bits < dec_bits) { change = 1; cost = inc_bits; }
        else {
          change = -1; cost = dec_bits;
          if (k == first_nz && a == 1) cost = 0x7FFFFFFFFFFFFFFFLL;
        }
      } else {
        const int bits = RDOQ_ONE_BIT + s.inc[p] + s.sig_inc[p];
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

// kvz_rdoq for one TU by one warp (all 32 lanes must call; `lane` = lane id).  coef / q: n x n row-major in shared
// memory.  type: 0 luma, 2 chroma (the reference passes 2 for U and V, quant-generic.c:239).  block_type: 1 intra,
// 2 inter.  tr_depth: depth below the CU (+1 for NxN).
//
// Work split: everything that does not depend on the serial context state is done by the lanes in parallel -- the scan
// table, the last significant position, and per coefficient group (once its neighbour pattern is known) the
// distortion of level 0 and the significance costs of its 16 positions.  Lane 0 then walks the group: level choice,
// context-set / Rice state, and the cost sums in the reference's order (double additions are not associative).
template <int NN, bool SH>
__device__ void rdoq_tu(const kvz_cuda_rdoq_params &p, const kvz_cuda_cabac_ctx *cabac, const int32_t *ebits_table, const int16_t *coef, int16_t *q, int log2n, int type,
                        int scan_idx, int block_type, int tr_depth, RdoqScratch<NN, SH> &s, int lane)
{
  const int n = 1 << log2n, nn = n * n;
  const int transform_shift = 15 - p.bitdepth - log2n;
  const int qp_scaled = scaled_qp(type, p.qp, (p.bitdepth - 8) * 6);
  const int q_bits = 14 + qp_scaled / 6 + transform_shift;
  const int qc = c_quant_scales[qp_scaled % 6];                          // flat scaling list
  const int half = 1 << (q_bits - 1);
  const double lambda = p.lambda;
  // error scale (scalinglist.c:351-368): 2^15 * 2^(-2 * transform_shift) / q / q / 2^(2 * (bitdepth - 8))
  const double err_scale = ldexp(32768.0, -2 * transform_shift) / qc / qc / (1 << (2 * (p.bitdepth - 8)));
  const RdoqModels m = rdoq_models(cabac, ebits_table, type);
  // distortion of quantising the coefficient at raster position blk to 0 (cost_coeff0, rdo.c:770-771)
  // cost_sig of a position from its code byte
  auto sig_cost_of = [&](uint8_t code) { return (code >> 6) == 2 ? 0.0 : lambda * ebits(m.sig[code & 63], code >> 6); };
  auto level0_cost = [&](int blk) { const double e = (double)min(abs((int)coef[blk]) * qc, 0x7FFFFFFF - half); return e * e * err_scale; };

  // ---- scan table and last significant scan position (find_last_scanpos)
  constexpr bool GLOBAL_SCAN = NN >= 1024;
  const uint16_t *blk_of = GLOBAL_SCAN ? g_scan_diag32 : s.blk;
  int my_last = -1;
  for (int sp = lane; sp < nn; sp += 32) {
    int blk;
    if constexpr (GLOBAL_SCAN) blk = blk_of[sp];
    else { blk = scan_pos(scan_idx, log2n, sp); s.blk[sp] = (uint16_t)blk; }
    const int ld = min(abs((int)coef[blk]) * qc, 0x7FFFFFFF - half);
    if (((ld + half) >> q_bits) > 0) my_last = sp;                       // increasing sp: the last assignment is the largest
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) my_last = max(my_last, __shfl_xor_sync(0xffffffffu, my_last, o));
  const int last_scanpos = my_last;
  __syncwarp();
  for (int sp = lane; sp < nn; sp += 32) if (sp > last_scanpos) q[blk_of[sp]] = 0;
  if (last_scanpos < 0) { __syncwarp(); return; }
  for (int g = lane; g < nn / 16; g += 32) { s.cg_flag[g] = 0; s.cg_sig_cost[g] = 0; }
  if (lane == 0) {
    if (SH) s.sig_inc[blk_of[last_scanpos]] = 0;
    // last-position bin costs (calc_last_bits, rdo.c:479-508)
    const int cb = log2n - 2;
    const int off = type ? 0 : (cb * 3 + ((cb + 1) >> 2));
    const int sh = type ? cb : ((cb + 3) >> 2);
    int bx = 0, by = 0, ctx;
    const int groups = last_group(n - 1);
    for (ctx = 0; ctx < groups; ++ctx) {
      const int o = off + (ctx >> sh);
      s.last_x_bits[ctx] = bx + ebits(m.last_x[o], 0); bx += ebits(m.last_x[o], 1);
      s.last_y_bits[ctx] = by + ebits(m.last_y[o], 0); by += ebits(m.last_y[o], 1);
    }
    s.last_x_bits[ctx] = bx; s.last_y_bits[ctx] = by;
  }
  __syncwarp();

  const int cg_last = last_scanpos >> 4;
  const int cgs_side = n >> 2;
  // serial state (meaningful in lane 0 only)
  int ctx_set = (last_scanpos > 0 && type == 0) ? 2 : 0;
  int c1 = 1, c2 = 0, rice = 0;
  uint32_t c1_idx = 0, c2_idx = 0;
  double base_cost = 0, block_uncoded_cost = 0;

  for (int cg = cg_last; cg >= 0; --cg) {
    const int cg_first = blk_of[cg << 4];                                  // raster position of the group's first coefficient
    const int cgx = (cg_first & (n - 1)) >> 2, cgy = (cg_first >> log2n) >> 2;
    const int cg_blk = cgy * cgs_side + cgx;
    // neighbouring coded groups: right and below (context.c:315-351); both were decided earlier in this walk
    const int right = (cgx < cgs_side - 1) ? (s.cg_flag[cgy * cgs_side + cgx + 1] != 0) : 0;
    const int lower = (cgy < cgs_side - 1) ? (s.cg_flag[(cgy + 1) * cgs_side + cgx] != 0) : 0;
    const int pattern = (n == 4) ? -1 : right + (lower << 1);
    // position-parallel part of the group: distortion of level 0, significance costs; positions that can only be
    // zero (max_abs_level == 0, rdo.c:428-432) are finished here except for their place in the ordered cost sums
    bool valid = false, cand = false;
    if (lane < 16) {
      const int sp = (cg << 4) + lane;
      if (sp <= last_scanpos) {
        valid = true;
        const int blk = blk_of[sp];
        const int ld = min(abs((int)coef[blk]) * qc, 0x7FFFFFFF - half);
        const double err = (double)ld;
        const double c0 = err * err * err_scale;
        s.prep_ld[lane] = ld;
        s.prep_c0[lane] = c0;
        cand = sp == last_scanpos || ((ld + half) >> q_bits) != 0;
        if (sp != last_scanpos) {
          const int ctx_sig = rdoq_sig_ctx(pattern, scan_idx, blk & (n - 1), blk >> log2n, log2n, type);
          const double sig0 = lambda * ebits(m.sig[ctx_sig], 0);
          s.prep_sig0[lane] = sig0;
          s.prep_sig1[lane] = lambda * ebits(m.sig[ctx_sig], 1);
          if (SH) s.sig_inc[blk] = ebits(m.sig[ctx_sig], 1) - ebits(m.sig[ctx_sig], 0);
          s.prep_ctx_sig[lane] = ctx_sig;
          if (!cand) {
            s.sig_code[sp] = (uint8_t)ctx_sig; s.cost_coeff[sp] = c0 + sig0;
            q[blk] = 0;
            if (SH) s.qdelta[blk] = ld >> (q_bits - 8);
          }
        }
      }
    }
    const unsigned valid_mask = __ballot_sync(0xffffffffu, valid), cand_mask = __ballot_sync(0xffffffffu, cand);
    __syncwarp();

    if (lane == 0) {
      double st_coded = 0, st_uncoded = 0, st_sig = 0, st_sig0 = 0;
      int nnz_before_pos0 = 0;
      unsigned nz_mask = 0;
      // the group's 32 inputs of the ordered sums come into registers first (independent loads), so that the walk
      // below is bound by the three addition chains only
      // (large TUs only: small ones rarely reach past their first groups, and the unrolled walk costs registers)
      constexpr bool BIG = NN >= 1024;
      double c0r[BIG ? 16 : 1], s0r[BIG ? 16 : 1];
      if constexpr (BIG) {
#pragma unroll
        for (int k = 0; k < 16; ++k) { c0r[k] = s.prep_c0[k]; s0r[k] = s.prep_sig0[k]; }
      }
#pragma unroll (BIG ? 16 : 1)
      for (int k = 15; k >= 0; --k) {
        if (!((valid_mask >> k) & 1)) continue;
        const int sp = (cg << 4) + k;
        double c0, sig0k;
        if constexpr (BIG) { c0 = c0r[k]; sig0k = s0r[k]; } else { c0 = s.prep_c0[k]; sig0k = s.prep_sig0[k]; }
        block_uncoded_cost += c0;
        if (!((cand_mask >> k) & 1)) {
          // level 0 is the only candidate: coded cost = c0 + cost of a zero significance flag
          const double cs = sig0k;
          base_cost += c0 + cs;
          st_sig += cs;
          if (k == 0) st_sig0 = cs;
          if (SH) s.inc[blk_of[sp]] = ebits(m.one[4 * ctx_set + c1], 0);
          if (k == 0 && sp > 0) {
            c2 = 0; rice = 0; c1_idx = 0; c2_idx = 0;
            ctx_set = (sp == 16 || type != 0) ? 0 : 2;
            if (c1 == 0) ++ctx_set;
            c1 = 1;
          }
          continue;
        }
        const int blk = blk_of[sp];
        const int ld = s.prep_ld[k];
        const uint32_t max_abs = (uint32_t)((ld + half) >> q_bits);
        const int one_ctx = 4 * ctx_set + c1, abs_ctx = ctx_set + c2;
        const bool last = sp == last_scanpos;
        // kvz_get_coded_level (rdo.c:413-452)
        uint32_t level = 0;
        double cc, cs = 0;
        int cs_kind = 2;
        if (!last && max_abs < 3) { cs = sig0k; cc = c0 + cs; cs_kind = 0; }
        else cc = 1.7e+308;
        if (max_abs != 0) {
          const double sig_now = last ? 0.0 : s.prep_sig1[k];
          const int lo = max_abs > 1 ? (int)max_abs - 1 : 1;
          for (int lvl = (int)max_abs; lvl >= lo; --lvl) {
            const double err = (double)(ld - lvl * (1 << q_bits));
            double c = err * err * err_scale + lambda * rdoq_level_rate(m, (uint32_t)lvl, one_ctx, abs_ctx, rice, c1_idx, c2_idx);
            c += sig_now;
            if (c < cc) { level = (uint32_t)lvl; cc = c; cs = sig_now; cs_kind = last ? 2 : 1; }
          }
        }
        s.cost_coeff[sp] = cc;
        s.sig_code[sp] = (uint8_t)((last ? 0 : s.prep_ctx_sig[k]) | (cs_kind << 6));
        if (SH) {
          s.qdelta[blk] = (ld - (int)level * (1 << q_bits)) >> (q_bits - 8);
          if (level > 0) {
            const int now = rdoq_level_rate(m, level, one_ctx, abs_ctx, rice, c1_idx, c2_idx);
            s.inc[blk] = rdoq_level_rate(m, level + 1, one_ctx, abs_ctx, rice, c1_idx, c2_idx) - now;
            s.dec[blk] = rdoq_level_rate(m, level - 1, one_ctx, abs_ctx, rice, c1_idx, c2_idx) - now;
          } else {
            s.inc[blk] = ebits(m.one[one_ctx], 0);
          }
        }
        q[blk] = (int16_t)level;
        base_cost += cc;

        const uint32_t base_level = (c1_idx < 8) ? (2 + (c2_idx < 1)) : 1;
        if (level >= base_level && level > (uint32_t)(3 * (1 << rice))) rice = min(rice + 1, 4);
        if (level >= 1) ++c1_idx;
        if (level > 1) { c1 = 0; c2 += (c2 < 2); ++c2_idx; }
        else if (c1 < 3 && c1 > 0 && level) ++c1;
        if (k == 0 && sp > 0) {                                            // context set for the next group down the scan
          c2 = 0; rice = 0; c1_idx = 0; c2_idx = 0;
          ctx_set = (sp == 16 || type != 0) ? 0 : 2;
          if (c1 == 0) ++ctx_set;
          c1 = 1;
        }
        st_sig += cs;
        if (k == 0) st_sig0 = cs;
        if (level) {
          nz_mask |= 1u << k;
          s.cg_flag[cg_blk] = 1;
          st_coded += cc - cs;
          st_uncoded += c0;
          if (k != 0) ++nnz_before_pos0;
        }
      }

      if (cg) {
        const int ctx_cg = right || lower;
        if (s.cg_flag[cg_blk] == 0) {
          s.cg_sig_cost[cg] = lambda * ebits(m.cg[ctx_cg], 0);
          base_cost += s.cg_sig_cost[cg] - st_sig;
        } else if (cg < cg_last) {
          if (nnz_before_pos0 == 0) { base_cost -= st_sig0; st_sig -= st_sig0; }
          double cost_zero_cg = base_cost;
          s.cg_sig_cost[cg] = lambda * ebits(m.cg[ctx_cg], 1);
          base_cost += s.cg_sig_cost[cg];
          cost_zero_cg += lambda * ebits(m.cg[ctx_cg], 0);
          cost_zero_cg += st_uncoded;
          cost_zero_cg -= st_coded;
          cost_zero_cg -= st_sig;
          if (cost_zero_cg < base_cost) {
            nz_mask = 0;
            s.cg_flag[cg_blk] = 0;
            base_cost = cost_zero_cg;
            s.cg_sig_cost[cg] = lambda * ebits(m.cg[ctx_cg], 0);
            for (int k = 15; k >= 0; --k) {
              const int sp = (cg << 4) + k, blk = blk_of[sp];
              if (q[blk]) { q[blk] = 0; s.cost_coeff[sp] = level0_cost(blk); s.sig_code[sp] = 2 << 6; }
            }
          }
        }
      } else {
        s.cg_flag[cg_blk] = 1;
      }
      s.cg_nz[cg] = (uint16_t)nz_mask;
    }
    __syncwarp();
  }

  if (lane == 0) {
    // ---- best last position (rdo.c:884-945)
    double best_cost;
    if (block_type != 1 && type == 0) {
      best_cost = block_uncoded_cost + lambda * ebits(m.root_cbf, 0);
      base_cost += lambda * ebits(m.root_cbf, 1);
    } else {
      const int ctx_cbf = type ? tr_depth : !tr_depth;
      best_cost = block_uncoded_cost + lambda * ebits(m.cbf[ctx_cbf], 0);
      base_cost += lambda * ebits(m.cbf[ctx_cbf], 1);
    }
    int best_last_p1 = 0;
    bool found_last = false;
    for (int cg = cg_last; cg >= 0 && !found_last; --cg) {
      const int cg_first = blk_of[cg << 4];
      const int cg_blk = ((cg_first >> log2n) >> 2) * cgs_side + ((cg_first & (n - 1)) >> 2);
      base_cost -= s.cg_sig_cost[cg];
      if (!s.cg_flag[cg_blk]) continue;
      const unsigned nz = s.cg_nz[cg];
      const int top = cg == cg_last ? (last_scanpos & 15) : 15;
      constexpr bool BIG = NN >= 1024;
      double csr[BIG ? 16 : 1];
      if constexpr (BIG) {
#pragma unroll
        for (int k = 0; k < 16; ++k) csr[k] = sig_cost_of(s.sig_code[(cg << 4) + k]);
      }
#pragma unroll (BIG ? 16 : 1)
      for (int k = 15; k >= 0; --k) {
        if (k > top) continue;
        const int sp = (cg << 4) + k;
        double csk;
        if constexpr (BIG) csk = csr[k]; else csk = sig_cost_of(s.sig_code[sp]);
        if (!((nz >> k) & 1)) { base_cost -= csk; continue; }
        const int blk = blk_of[sp];
        const int py = blk >> log2n, px = blk & (n - 1);
        const int gx = last_group(scan_idx == 2 ? py : px), gy = last_group(scan_idx == 2 ? px : py);
        double bits = s.last_x_bits[gx] + s.last_y_bits[gy];
        if (gx > 3) bits += RDOQ_ONE_BIT * ((gx - 2) >> 1);
        if (gy > 3) bits += RDOQ_ONE_BIT * ((gy - 2) >> 1);
        const double total = base_cost + lambda * bits - csk;
        if (total < best_cost) { best_last_p1 = sp + 1; best_cost = total; }
        if (q[blk] > 1) { found_last = true; break; }
        base_cost -= s.cost_coeff[sp];
        base_cost += level0_cost(blk);
      }
    }
    s.best_last_p1 = best_last_p1;
  }
  __syncwarp();

  // ---- signs and clean-up in parallel, then sign hiding
  const int best_last_p1 = s.best_last_p1;
  int abs_sum = 0;
  for (int sp = lane; sp <= last_scanpos; sp += 32) {
    const int blk = blk_of[sp];
    if (sp < best_last_p1) {
      const int level = q[blk];
      abs_sum += level;
      q[blk] = (int16_t)(coef[blk] < 0 ? -level : level);
    } else {
      q[blk] = 0;
    }
  }
  if constexpr (SH) {
    abs_sum = warp_sum(abs_sum);
    __syncwarp();
    if (lane == 0 && abs_sum >= 2) rdoq_sign_hiding(s, blk_of, lambda, p.bitdepth, qp_scaled, scan_idx, log2n, best_last_p1, coef, q);
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------------------------
// One THREAD per TU, for 4x4 and 8x8 blocks: there are 10^4..10^5 of them per frame and most are empty, so running
// HM's chain once per lane beats one chain per warp.  Same arithmetic, same order; the cost tables live in the
// thread's local memory, the scan table and the entropy table in shared memory.
template <int NN, bool SH>
struct RdoqLocal {
  double cost_coeff[NN], cost_sig[NN];
  int32_t inc[SH ? NN : 1], dec[SH ? NN : 1], sig_inc[SH ? NN : 1], qdelta[SH ? NN : 1];
  double cg_sig_cost[NN / 16];
  int32_t cg_flag[NN / 16];
  int32_t last_x_bits[8], last_y_bits[8];
  const uint8_t *blk;                                      // scan position -> raster position (shared memory)
};

template <int NN, bool SH>
__device__ void rdoq_tu_thread(const kvz_cuda_rdoq_params &p, const kvz_cuda_cabac_ctx *cabac, const int32_t *ebits_table, const int16_t *coef,
                               int16_t *q, int log2n, int type, int scan_idx, int block_type, int tr_depth, RdoqLocal<NN, SH> &s)
{
  const int n = 1 << log2n;
  const int transform_shift = 15 - p.bitdepth - log2n;
  const int qp_scaled = scaled_qp(type, p.qp, (p.bitdepth - 8) * 6);
  const int q_bits = 14 + qp_scaled / 6 + transform_shift;
  const int qc = c_quant_scales[qp_scaled % 6];
  const int half = 1 << (q_bits - 1);
  const double lambda = p.lambda;
  const double err_scale = ldexp(32768.0, -2 * transform_shift) / qc / qc / (1 << (2 * (p.bitdepth - 8)));
  const RdoqModels m = rdoq_models(cabac, ebits_table, type);
  auto level_double = [&](int blk) { return min(abs((int)coef[blk]) * qc, 0x7FFFFFFF - half); };

  // find_last_scanpos (quant-generic.c:376-399)
  int last_scanpos = -1;
  for (int sp = NN - 1; sp >= 0; --sp) {
    const int blk = s.blk[sp];
    if (((level_double(blk) + half) >> q_bits) > 0) { last_scanpos = sp; break; }
    q[blk] = 0;
  }
  if (last_scanpos < 0) return;
  for (int g = 0; g < NN / 16; ++g) { s.cg_flag[g] = 0; s.cg_sig_cost[g] = 0; }
  if (SH) s.sig_inc[s.blk[last_scanpos]] = 0;
  {
    const int cb = log2n - 2;
    const int off = type ? 0 : (cb * 3 + ((cb + 1) >> 2));
    const int sh = type ? cb : ((cb + 3) >> 2);
    int bx = 0, by = 0, ctx;
    const int groups = last_group(n - 1);
    for (ctx = 0; ctx < groups; ++ctx) {
      const int o = off + (ctx >> sh);
      s.last_x_bits[ctx] = bx + ebits(m.last_x[o], 0); bx += ebits(m.last_x[o], 1);
      s.last_y_bits[ctx] = by + ebits(m.last_y[o], 0); by += ebits(m.last_y[o], 1);
    }
    s.last_x_bits[ctx] = bx; s.last_y_bits[ctx] = by;
  }

  const int cg_last = last_scanpos >> 4;
  const int cgs_side = n >> 2;
  int ctx_set = (last_scanpos > 0 && type == 0) ? 2 : 0;
  int c1 = 1, c2 = 0, rice = 0;
  uint32_t c1_idx = 0, c2_idx = 0;
  double base_cost = 0, block_uncoded_cost = 0;

  for (int cg = cg_last; cg >= 0; --cg) {
    const int cg_first = s.blk[cg << 4];
    const int cgx = (cg_first & (n - 1)) >> 2, cgy = (cg_first >> log2n) >> 2;
    const int cg_blk = cgy * cgs_side + cgx;
    const int right = (cgx < cgs_side - 1) ? (s.cg_flag[cgy * cgs_side + cgx + 1] != 0) : 0;
    const int lower = (cgy < cgs_side - 1) ? (s.cg_flag[(cgy + 1) * cgs_side + cgx] != 0) : 0;
    const int pattern = (n == 4) ? -1 : right + (lower << 1);
    double st_coded = 0, st_uncoded = 0, st_sig = 0, st_sig0 = 0;
    int nnz_before_pos0 = 0;
    for (int k = 15; k >= 0; --k) {
      const int sp = (cg << 4) + k;
      if (sp > last_scanpos) continue;
      const int blk = s.blk[sp];
      const int ld = level_double(blk);
      const uint32_t max_abs = (uint32_t)((ld + half) >> q_bits);
      const double err0 = (double)ld;
      const double c0 = err0 * err0 * err_scale;
      block_uncoded_cost += c0;
      const int one_ctx = 4 * ctx_set + c1, abs_ctx = ctx_set + c2;
      const bool last = sp == last_scanpos;
      int ctx_sig = 0;
      if (!last) {
        ctx_sig = rdoq_sig_ctx(pattern, scan_idx, blk & (n - 1), blk >> log2n, log2n, type);
        if (SH) s.sig_inc[blk] = ebits(m.sig[ctx_sig], 1) - ebits(m.sig[ctx_sig], 0);
      }
      uint32_t level = 0;
      double cc, cs = 0;
      if (!last && max_abs < 3) { cs = lambda * ebits(m.sig[ctx_sig], 0); cc = c0 + cs; }
      else cc = 1.7e+308;
      if (max_abs != 0) {
        const double sig_now = last ? 0.0 : lambda * ebits(m.sig[ctx_sig], 1);
        const int lo = max_abs > 1 ? (int)max_abs - 1 : 1;
        for (int lvl = (int)max_abs; lvl >= lo; --lvl) {
          const double err = (double)(ld - lvl * (1 << q_bits));
          double c = err * err * err_scale + lambda * rdoq_level_rate(m, (uint32_t)lvl, one_ctx, abs_ctx, rice, c1_idx, c2_idx);
          c += sig_now;
          if (c < cc) { level = (uint32_t)lvl; cc = c; cs = sig_now; }
        }
      }
      s.cost_coeff[sp] = cc;
      s.cost_sig[sp] = cs;
      if (SH) {
        s.qdelta[blk] = (ld - (int)level * (1 << q_bits)) >> (q_bits - 8);
        if (level > 0) {
          const int now = rdoq_level_rate(m, level, one_ctx, abs_ctx, rice, c1_idx, c2_idx);
          s.inc[blk] = rdoq_level_rate(m, level + 1, one_ctx, abs_ctx, rice, c1_idx, c2_idx) - now;
          s.dec[blk] = rdoq_level_rate(m, level - 1, one_ctx, abs_ctx, rice, c1_idx, c2_idx) - now;
        } else {
          s.inc[blk] = ebits(m.one[one_ctx], 0);
        }
      }
      q[blk] = (int16_t)level;
      base_cost += cc;
      const uint32_t base_level = (c1_idx < 8) ? (2 + (c2_idx < 1)) : 1;
      if (level >= base_level && level > (uint32_t)(3 * (1 << rice))) rice = min(rice + 1, 4);
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
        s.cg_flag[cg_blk] = 1;
        st_coded += cc - cs;
        st_uncoded += c0;
        if (k != 0) ++nnz_before_pos0;
      }
    }
    if (cg) {
      const int ctx_cg = right || lower;
      if (s.cg_flag[cg_blk] == 0) {
        s.cg_sig_cost[cg] = lambda * ebits(m.cg[ctx_cg], 0);
        base_cost += s.cg_sig_cost[cg] - st_sig;
      } else if (cg < cg_last) {
        if (nnz_before_pos0 == 0) { base_cost -= st_sig0; st_sig -= st_sig0; }
        double cost_zero_cg = base_cost;
        s.cg_sig_cost[cg] = lambda * ebits(m.cg[ctx_cg], 1);
        base_cost += s.cg_sig_cost[cg];
        cost_zero_cg += lambda * ebits(m.cg[ctx_cg], 0);
        cost_zero_cg += st_uncoded;
        cost_zero_cg -= st_coded;
        cost_zero_cg -= st_sig;
        if (cost_zero_cg < base_cost) {
          s.cg_flag[cg_blk] = 0;
          base_cost = cost_zero_cg;
          s.cg_sig_cost[cg] = lambda * ebits(m.cg[ctx_cg], 0);
          for (int k = 15; k >= 0; --k) {
            const int sp = (cg << 4) + k, blk = s.blk[sp];
            if (q[blk]) { q[blk] = 0; const double e = (double)level_double(blk); s.cost_coeff[sp] = e * e * err_scale; s.cost_sig[sp] = 0; }
          }
        }
      }
    } else {
      s.cg_flag[cg_blk] = 1;
    }
  }

  // best last position (rdo.c:884-945)
  double best_cost;
  if (block_type != 1 && type == 0) {
    best_cost = block_uncoded_cost + lambda * ebits(m.root_cbf, 0);
    base_cost += lambda * ebits(m.root_cbf, 1);
  } else {
    const int ctx_cbf = type ? tr_depth : !tr_depth;
    best_cost = block_uncoded_cost + lambda * ebits(m.cbf[ctx_cbf], 0);
    base_cost += lambda * ebits(m.cbf[ctx_cbf], 1);
  }
  int best_last_p1 = 0;
  bool found_last = false;
  for (int cg = cg_last; cg >= 0 && !found_last; --cg) {
    const int cg_first = s.blk[cg << 4];
    const int cg_blk = ((cg_first >> log2n) >> 2) * cgs_side + ((cg_first & (n - 1)) >> 2);
    base_cost -= s.cg_sig_cost[cg];
    if (!s.cg_flag[cg_blk]) continue;
    for (int k = 15; k >= 0; --k) {
      const int sp = (cg << 4) + k;
      if (sp > last_scanpos) continue;
      const int blk = s.blk[sp];
      if (q[blk]) {
        const int py = blk >> log2n, px = blk & (n - 1);
        const int gx = last_group(scan_idx == 2 ? py : px), gy = last_group(scan_idx == 2 ? px : py);
        double bits = s.last_x_bits[gx] + s.last_y_bits[gy];
        if (gx > 3) bits += RDOQ_ONE_BIT * ((gx - 2) >> 1);
        if (gy > 3) bits += RDOQ_ONE_BIT * ((gy - 2) >> 1);
        const double total = base_cost + lambda * bits - s.cost_sig[sp];
        if (total < best_cost) { best_last_p1 = sp + 1; best_cost = total; }
        if (q[blk] > 1) { found_last = true; break; }
        base_cost -= s.cost_coeff[sp];
        const double e = (double)level_double(blk);
        base_cost += e * e * err_scale;
      } else {
        base_cost -= s.cost_sig[sp];
      }
    }
  }
  unsigned abs_sum = 0;
  for (int sp = 0; sp < best_last_p1; ++sp) {
    const int blk = s.blk[sp];
    const int level = q[blk];
    abs_sum += (unsigned)level;
    q[blk] = (int16_t)(coef[blk] < 0 ? -level : level);
  }
  for (int sp = best_last_p1; sp <= last_scanpos; ++sp) q[s.blk[sp]] = 0;
  if constexpr (SH) {
    if (abs_sum >= 2) rdoq_sign_hiding(s, s.blk, lambda, p.bitdepth, qp_scaled, scan_idx, log2n, best_last_p1, coef, q);
  }
}

}  // namespace kvzc
