/*
 * ref_interpass.c -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The frame-level INTER pass of kvazaar_b200/csrc/interpass.cu executed on the CPU through the unmodified
 * reference's own selected (AVX2) strategy pointers: kvz_reg_sad, the four kvz_filter_*_blocks_*_luma stages,
 * kvz_satd_any_size(_quad), kvz_sample_quarterpel_luma, kvz_sample_octpel_chroma, kvz_quantize_residual,
 * kvz_pixels_calc_ssd.  The control flow between them is search_frac's (ref: search_inter.c:974-1168) without
 * the MV bit cost, plus a +-R raster full search.  Parity checker for the CUDA pass; pthread pool over PUs.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "global.h"
#include "kvazaar.h"
#include "kvazaar_internal.h"
#include "encoder.h"
#include "encoderstate.h"
#include "strategyselector.h"
#include "cu.h"

#include "../include/kvz_cuda.h"

typedef struct { const kvz_api *api; kvz_config *cfg; kvz_encoder *enc; } kvzref_ctx;

typedef struct {
  kvzref_ctx *ctx;
  const uint8_t *cur, *ref;
  int W, H, R;
  const kvz_cuda_ip_layout *L;
  uint8_t *blob, *pred;
  volatile int next;
} ip_job;

static void ip_do_pu(ip_job *j, int i, kvz_pixel (*filtered)[LCU_LUMA_SIZE], int16_t (*im)[KVZ_IPOL_MAX_IM_SIZE_LUMA_SIMD],
                     int16_t (*cols)[KVZ_EXT_BLOCK_W_LUMA + 1])
{
  static const int sqx[9] = { 0, -1, 1, 0, 0, -1, 1, -1, 1 }, sqy[9] = { 0, 0, 0, -1, 1, -1, -1, 1, 1 };
  const kvz_cuda_ip_layout *L = j->L;
  const int W = j->W, H = j->H, R = j->R, Wc = W / 2;
  const int x = (1 + i % L->pus_x) * 16, y = (1 + i / L->pus_x) * 16;
  const encoder_control_t *enc = j->ctx->enc->control;
  encoder_state_t *state = &j->ctx->enc->states[0];
  kvz_pixel *cur = (kvz_pixel *)j->cur + y * W + x;
  kvz_pixel *refp = (kvz_pixel *)j->ref;

  /* 1. integer full search */
  unsigned best = 0; int bx = 0, by = 0, first = 1;
  for (int dy = -R; dy <= R; ++dy)
    for (int dx = -R; dx <= R; ++dx) {
      const unsigned s = kvz_reg_sad(cur, refp + (y + dy) * W + x + dx, 16, 16, W, W);
      if (first || s < best) { best = s; bx = dx; by = dy; first = 0; }
    }
  ((int16_t *)(j->blob + L->mv_int))[2 * i] = (int16_t)bx; ((int16_t *)(j->blob + L->mv_int))[2 * i + 1] = (int16_t)by;
  ((uint32_t *)(j->blob + L->sad_int))[i] = best;

  /* 2. fractional search (search_frac) */
  kvz_pixel *ext_origin = refp + (y + by - 1) * W + (x + bx - 1);
  unsigned cost = kvz_satd_any_size(16, 16, cur, W, ext_origin + W + 1, W);
  int mvx = bx * 2, mvy = by * 2, best_index = 0, off_x = 0, off_y = 0, idx = 1;
  ipol_blocks_func *steps[4] = { kvz_filter_hpel_blocks_hor_ver_luma, kvz_filter_hpel_blocks_diag_luma,
                                 kvz_filter_qpel_blocks_hor_ver_luma, kvz_filter_qpel_blocks_diag_luma };
  for (int step = 0; step < 4; ++step) {
    steps[step](enc, ext_origin, W, 16, 16, filtered, im, 4, cols, off_x, off_y);
    const kvz_pixel *pos[4] = { filtered[0], filtered[1], filtered[2], filtered[3] };
    unsigned costs[4]; int8_t valid[4] = { 1, 1, 1, 1 };
    kvz_satd_any_size_quad(16, 16, pos, LCU_WIDTH, cur, W, 4, costs, valid);
    for (int k = 0; k < 4; ++k) if (costs[k] < cost) { cost = costs[k]; best_index = idx + k; }
    idx += 4;
    if (step == 1 || step == 3) {
      mvx += sqx[best_index]; mvy += sqy[best_index];
      if (step == 1) { mvx *= 2; mvy *= 2; off_x = sqx[best_index]; off_y = sqy[best_index]; best_index = 0; idx = 1; }
    }
  }
  ((int16_t *)(j->blob + L->mv_final))[2 * i] = (int16_t)mvx; ((int16_t *)(j->blob + L->mv_final))[2 * i + 1] = (int16_t)mvy;
  ((uint32_t *)(j->blob + L->satd_best))[i] = cost;

  /* 3. motion compensation */
  const int16_t mv[2] = { (int16_t)mvx, (int16_t)mvy };
  kvz_pixel *pred = j->pred, *rec = j->blob + L->rec;
  kvz_sample_quarterpel_luma(enc, refp + (y + (mvy >> 2)) * W + x + (mvx >> 2), W, 16, 16, pred + y * W + x, W, 0, 0, mv);
  const size_t po[3] = { 0, (size_t)W * H, (size_t)W * H * 5 / 4 };
  for (int c = 1; c <= 2; ++c)
    kvz_sample_octpel_chroma(enc, refp + po[c] + (y / 2 + (mvy >> 3)) * Wc + x / 2 + (mvx >> 3), Wc, 8, 8,
                             pred + po[c] + (y / 2) * Wc + x / 2, Wc, 0, 0, mv);

  /* 4. inter residual coding + SSD */
  cu_info_t cu; memset(&cu, 0, sizeof(cu));
  cu.type = CU_INTER; cu.part_size = SIZE_2Nx2N;
  int has = kvz_quantize_residual(state, &cu, 16, COLOR_Y, SCAN_DIAG, 0, W, W, cur, pred + y * W + x, rec + y * W + x,
                                  (coeff_t *)(j->blob + L->coeff_y) + (size_t)i * 256, false);
  ((int32_t *)(j->blob + L->has_y))[i] = has;
  ((uint32_t *)(j->blob + L->ssd_y))[i] = kvz_pixels_calc_ssd(cur, rec + y * W + x, W, W, 16);
  for (int c = 1; c <= 2; ++c) {
    const size_t o = po[c] + (y / 2) * Wc + x / 2;
    has = kvz_quantize_residual(state, &cu, 8, (color_t)c, SCAN_DIAG, 0, Wc, Wc, (kvz_pixel *)j->cur + o, pred + o, rec + o,
                                (coeff_t *)(j->blob + (c == 1 ? L->coeff_u : L->coeff_v)) + (size_t)i * 64, false);
    ((int32_t *)(j->blob + (c == 1 ? L->has_u : L->has_v)))[i] = has;
    ((uint32_t *)(j->blob + (c == 1 ? L->ssd_u : L->ssd_v)))[i] = kvz_pixels_calc_ssd((kvz_pixel *)j->cur + o, rec + o, Wc, Wc, 8);
  }
}

static void *ip_worker(void *arg)
{
  ip_job *j = (ip_job *)arg;
  void *mem = NULL;
  const size_t sz = sizeof(kvz_pixel) * 4 * LCU_LUMA_SIZE + sizeof(int16_t) * 5 * KVZ_IPOL_MAX_IM_SIZE_LUMA_SIMD +
                    sizeof(int16_t) * 5 * (KVZ_EXT_BLOCK_W_LUMA + 1) + 256;
  if (posix_memalign(&mem, 64, sz)) abort();
  memset(mem, 0, sz);
  kvz_pixel (*filtered)[LCU_LUMA_SIZE] = mem;
  int16_t (*im)[KVZ_IPOL_MAX_IM_SIZE_LUMA_SIMD] = (void *)((uint8_t *)mem + sizeof(kvz_pixel) * 4 * LCU_LUMA_SIZE);
  int16_t (*cols)[KVZ_EXT_BLOCK_W_LUMA + 1] = (void *)((uint8_t *)im + ((sizeof(int16_t) * 5 * KVZ_IPOL_MAX_IM_SIZE_LUMA_SIMD + 63) & ~(size_t)63));
  for (;;) {
    const int i0 = __sync_fetch_and_add(&j->next, 8);
    if (i0 >= j->L->npu) break;
    for (int i = i0; i < i0 + 8 && i < j->L->npu; ++i) ip_do_pu(j, i, filtered, im, cols);
  }
  free(mem);
  return NULL;
}

int kvzref_inter_pass(kvzref_ctx *ctx, const uint8_t *cur, const uint8_t *ref, int W, int H, int qp, int R,
                      const kvz_cuda_ip_layout *L, uint8_t *blob, int nthreads)
{
  if (KVZ_BIT_DEPTH != 8) return -1;
  ip_job j; memset(&j, 0, sizeof(j));
  j.ctx = ctx; j.cur = cur; j.ref = ref; j.W = W; j.H = H; j.R = R; j.L = L; j.blob = blob;
  encoder_state_t *st = &ctx->enc->states[0];
  st->qp = (int8_t)qp; st->frame->slicetype = KVZ_SLICE_P;
  void *pred = NULL;
  if (posix_memalign(&pred, 64, (size_t)W * H * 3 / 2 + 128)) abort();
  memset(pred, 0, (size_t)W * H * 3 / 2 + 128);
  j.pred = pred;
  memset(blob + L->rec, 0, (size_t)W * H * 3 / 2);
  pthread_t th[256];
  if (nthreads > 256) nthreads = 256;
  for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, ip_worker, &j);
  ip_worker(&j);
  for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
  free(pred);
  return 0;
}
