/*
 * ref_framepass.c -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The frame-level hot-path pass (see kvazaar_b200/csrc/framepass.cu for the definition) executed on the CPU through
 * the UNMODIFIED reference's own strategy function pointers -- i.e. whatever kvz_strategyselector_init selected on
 * this host (AVX2 where available), exactly like the reference encoder would call them -- with a pthread pool
 * over blocks.  It is (1) the parity checker for the CUDA frame pass (byte-identical result blob) and (2) the
 * `--impl reference` / cpu_baseline arm of bench.py.  The only non-reference arithmetic is the frame-plane
 * restatement of kvz_intra_build_reference from oracle/kvz_oracle.c (the reference's version needs an lcu_t)
 * and the three integer decision rules of the pass (argmin, SAO offset = clip(sum / count), SAO class argmin).
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
#include "intra.h"
#include <math.h>
#include "context.h"
#include "rdo.h"
#include "transform.h"
#include "sao.h"
#include "filter.h"
#include "cu.h"
#include "videoframe.h"
#include "nal.h"

#include "../include/kvz_cuda.h"   /* kvz_cuda_fp_layout only: the blob layout both arms fill */
#include "kvz_oracle.h"

typedef struct { const kvz_api *api; kvz_config *cfg; kvz_encoder *enc; } kvzref_ctx;   /* as in ref_shim.c */

typedef struct {
  kvzref_ctx *ctx;
  const kvz_pixel *src;
  int W, H, qp;
  const kvz_cuda_fp_layout *L;
  uint8_t *blob;
  kvz_pixel *rec[3][4];     /* per colour, per depth reconstruction planes (scratch, not part of the blob) */
  int stage;
  volatile int next;
  int total;
  int wl[4];
  int *dbk_done;            /* per LCU row: number of LCUs deblocked (wavefront progress) */
} job_t;

static void *xaligned(size_t bytes) { void *p = NULL; if (posix_memalign(&p, 64, bytes + 64)) abort(); memset(p, 0, bytes + 64); return p; }

static int scan_for(int is_c, int w, int mode)
{
  if ((!is_c && w <= 8) || (is_c && w == 4)) return (mode >= 6 && mode <= 14) ? 2 : ((mode >= 22 && mode <= 30) ? 1 : 0);
  return 0;
}

/* one block of one depth: rough search + selection + luma recon (+ chroma recon) */
static void do_block(job_t *j, int d, int b, kvz_pixel *buf /* scratch: 6 * 1024 px, aligned */)
{
  const int W = j->W, H = j->H, w = j->wl[d], log2w = 5 - d;
  const int bx = b % (W / w), by = b / (W / w);
  const kvz_cuda_fp_layout *L = j->L;
  encoder_state_t *state = &j->ctx->enc->states[0];
  kvz_pixel *orig = buf, *pred = buf + 1024, *recb = buf + 2048, *pred2 = buf + 3072;
  kvz_intra_references refs;
  cu_info_t cu; memset(&cu, 0, sizeof(cu));
  cu.type = CU_INTRA; cu.part_size = SIZE_2Nx2N;
  /* transform depth below the CU as the pass defines it per depth index (csrc/framepass.cu k_fp_tr_depth); only
   * kvz_rdoq's cbf context reads it (quant-generic.c:237, rdo.c:895-899) */
  { static const int trd[4] = { 1, 0, 0, 1 }; cu.depth = 0; cu.tr_depth = trd[d]; }

  /* --- luma rough search: 35 modes, SATD against the source block */
  memset(&refs, 0, sizeof(refs));
  orc_intra_build_reference(log2w, 0, bx * w, by * w, W, H, j->src, W, refs.ref.top, refs.ref.left);
  for (int y = 0; y < w; ++y) memcpy(orig + y * w, j->src + (size_t)(by * w + y) * W + bx * w, w * sizeof(kvz_pixel));
  cost_pixel_nxn_func *satd = kvz_pixels_get_satd_func(w);
  unsigned best = 0; int best_mode = 0;
  for (int mode = 0; mode < 35; ++mode) {
    kvz_intra_predict(&refs, log2w, mode, COLOR_Y, pred, true);
    const unsigned c = satd(pred, orig);
    if (mode == 0 || c < best) { best = c; best_mode = mode; }
  }
  ((int8_t *)(j->blob + L->mode_y[d]))[b] = (int8_t)best_mode;
  ((uint32_t *)(j->blob + L->cost_y[d]))[b] = best;

  /* --- luma reconstruction of the chosen mode */
  kvz_intra_predict(&refs, log2w, best_mode, COLOR_Y, pred, true);
  coeff_t *coeff = (coeff_t *)(j->blob + L->coeff_y[d]) + (size_t)b * w * w;
  int has;
  if (w == 4 && state->encoder_control->cfg.trskip_enable) {
    /* 4x4 luma: DST or transform skip, whichever has the smaller SSD + bits * lambda (transform.c:241-288, as quantize_tr_residual calls it) */
    int8_t tr_skip = 0;
    has = kvz_quantize_residual_trskip(state, &cu, w, COLOR_Y, scan_for(0, w, best_mode), &tr_skip, w, w, orig, pred, recb, coeff);
    (j->blob + L->trskip_y)[b] = (uint8_t)tr_skip;
  } else {
    has = kvz_quantize_residual(state, &cu, w, COLOR_Y, scan_for(0, w, best_mode), 0, w, w, orig, pred, recb, coeff, false);
    if (w == 4) (j->blob + L->trskip_y)[b] = 0;
  }
  (j->blob + L->has_y[d])[b] = (uint8_t)has;
  ((uint32_t *)(j->blob + L->ssd_y[d]))[b] = kvz_pixels_calc_ssd(orig, recb, w, w, w);
  ((double *)(j->blob + L->bits_y[d]))[b] = kvz_get_coeff_cost(state, coeff, w, 0, (int8_t)scan_for(0, w, best_mode));
  for (int y = 0; y < w; ++y) memcpy(j->rec[0][d] + (size_t)(by * w + y) * W + bx * w, recb + y * w, w * sizeof(kvz_pixel));

  /* --- chroma, co-located luma mode */
  if (d < 3) {
    const int wc = w / 2, Wc = W / 2;
    for (int color = 1; color <= 2; ++color) {
      const kvz_pixel *plane = j->src + (color == 1 ? (size_t)W * H : (size_t)W * H * 5 / 4);
      memset(&refs, 0, sizeof(refs));
      orc_intra_build_reference(log2w - 1, color, bx * w, by * w, W, H, plane, Wc, refs.ref.top, refs.ref.left);
      for (int y = 0; y < wc; ++y) memcpy(orig + y * wc, plane + (size_t)(by * wc + y) * Wc + bx * wc, wc * sizeof(kvz_pixel));
      kvz_intra_predict(&refs, log2w - 1, best_mode, (color_t)color, pred2, true);
      coeff_t *cc = (coeff_t *)(j->blob + (color == 1 ? L->coeff_u[d] : L->coeff_v[d])) + (size_t)b * wc * wc;
      has = kvz_quantize_residual(state, &cu, wc, (color_t)color, scan_for(1, wc, best_mode), 0, wc, wc, orig, pred2, recb, cc, false);
      (j->blob + (color == 1 ? L->has_u[d] : L->has_v[d]))[b] = (uint8_t)has;
      ((double *)(j->blob + (color == 1 ? L->bits_u[d] : L->bits_v[d])))[b] = kvz_get_coeff_cost(state, cc, wc, 2, (int8_t)scan_for(1, wc, best_mode));
      ((uint32_t *)(j->blob + (color == 1 ? L->ssd_u[d] : L->ssd_v[d])))[b] = kvz_pixels_calc_ssd(orig, recb, wc, wc, wc);
      for (int y = 0; y < wc; ++y) memcpy(j->rec[color][d] + (size_t)(by * wc + y) * Wc + bx * wc, recb + y * wc, wc * sizeof(kvz_pixel));
    }
  }
}

/* SAO for one (plane, CTU): statistics, offsets, delta-distortions, decision, reconstruction */
static void do_sao(job_t *j, int i, kvz_pixel *buf /* 2 * 4096 px */)
{
  const int W = j->W, H = j->H;
  const kvz_cuda_fp_layout *L = j->L;
  const int nctu = L->nctu, nctu3 = 3 * nctu;
  const int color = i / nctu, ctu = i % nctu, cx = (W + 63) / 64;
  const int Wp = color ? W / 2 : W, Hp = color ? H / 2 : H, lw = color ? 32 : 64;
  const int x0 = (ctu % cx) * lw, y0 = (ctu / cx) * lw;
  const int bw = MIN(lw, Wp - x0), bh = MIN(lw, Hp - y0);
  const kvz_pixel *splane = j->src + (color == 0 ? 0 : (color == 1 ? (size_t)W * H : (size_t)W * H * 5 / 4));
  const kvz_pixel *rplane = j->rec[color][2];
  kvz_pixel *sao_plane = (kvz_pixel *)(j->blob + L->sao_rec) + (color == 0 ? 0 : (color == 1 ? (size_t)W * H : (size_t)W * H * 5 / 4));
  const encoder_control_t *enc = j->ctx->enc->control;
  encoder_state_t *state = &j->ctx->enc->states[0];
  kvz_pixel *orig = buf, *rec = buf + 4096;
  /* contiguous copies, as sao_search_luma/chroma hand them to the strategies (ref: sao.c:605-669) */
  for (int y = 0; y < bh; ++y) { memcpy(orig + y * bw, splane + (size_t)(y0 + y) * Wp + x0, bw * sizeof(kvz_pixel)); memcpy(rec + y * bw, rplane + (size_t)(y0 + y) * Wp + x0, bw * sizeof(kvz_pixel)); }
  int32_t *stats = (int32_t *)(j->blob + L->sao_stats) + (size_t)i * 40;
  int32_t *dd = (int32_t *)(j->blob + L->sao_dd);
  int offsets[4][NUM_SAO_EDGE_CATEGORIES];
  int best_eo = 0, best_dd = 0;
  for (int eo = 0; eo < 4; ++eo) {
    int csc[2][NUM_SAO_EDGE_CATEGORIES]; memset(csc, 0, sizeof(csc));
    kvz_calc_sao_edge_dir(enc, orig, rec, eo, bw, bh, csc);
    memcpy(stats + eo * 10, csc, sizeof(csc));
    offsets[eo][0] = 0;
    for (int k = 1; k < 5; ++k) offsets[eo][k] = csc[1][k] ? CLIP(-7, 7, csc[0][k] / csc[1][k]) : 0;
    const int v = kvz_sao_edge_ddistortion(enc, orig, rec, bw, bh, eo, offsets[eo]);
    dd[(size_t)eo * nctu3 + i] = v;
    if (eo == 0 || v < best_dd) { best_dd = v; best_eo = eo; }
  }
  const int bands[4] = { 1, -1, 2, -2 };
  ((int32_t *)(j->blob + L->sao_band_dd))[i] = kvz_sao_band_ddistortion(state, orig, rec, bw, bh, (i * 7) % 29, bands);
  ((int8_t *)(j->blob + L->sao_best))[i] = (int8_t)(best_dd < 0 ? best_eo : -1);
  /* reconstruction rectangle: CTU area minus the picture's one-pixel border */
  const int rx0 = MAX(x0, 1), ry0 = MAX(y0, 1), rx1 = MIN(x0 + bw, Wp - 1), ry1 = MIN(y0 + bh, Hp - 1);
  if (best_dd < 0 && rx1 > rx0 && ry1 > ry0) {
    sao_info_t sao; memset(&sao, 0, sizeof(sao));
    sao.type = SAO_TYPE_EDGE; sao.eo_class = (sao_eo_class)best_eo;
    for (int k = 0; k < 5; ++k) sao.offsets[k + (color == 2 ? 5 : 0)] = offsets[best_eo][k];
    kvz_sao_reconstruct_color(enc, rplane + (size_t)ry0 * Wp + rx0, sao_plane + (size_t)ry0 * Wp + rx0, &sao, Wp, Wp,
                              rx1 - rx0, ry1 - ry0, (color_t)color);
  }
}

/* deblocking of one LCU row with the reference's own wavefront rule: LCU (x, y) runs after (x + 1, y - 1)
 * (encoderstate.c:1156-1190 dependencies); rows are claimed in order, progress is published per row */
static void do_deblock_row(job_t *j, int row)
{
  encoder_state_t *st = &j->ctx->enc->states[0];
  const int lcus_x = (j->W + 63) / 64;
  for (int x = 0; x < lcus_x; ++x) {
    if (row > 0) {
      const int need = x + 2 < lcus_x ? x + 2 : lcus_x;
      while (__atomic_load_n(&j->dbk_done[row - 1], __ATOMIC_ACQUIRE) < need) { }
    }
    kvz_filter_deblock_lcu(st, x * 64, row * 64);
    __atomic_store_n(&j->dbk_done[row], x + 1, __ATOMIC_RELEASE);
  }
}

static void *worker(void *arg)
{
  job_t *j = (job_t *)arg;
  kvz_pixel *buf = (kvz_pixel *)xaligned(8192 * sizeof(kvz_pixel));
  /* work is claimed in chunks so that 100+ threads do not serialise on the shared counter (and neighbouring
   * blocks, whose outputs share cache lines, stay on one thread) */
  const int chunk = j->stage == 0 ? 256 : (j->stage == 2 ? 1 : 4);
  for (;;) {
    const int i0 = __sync_fetch_and_add(&j->next, chunk);
    if (i0 >= j->total) break;
    const int i1 = i0 + chunk < j->total ? i0 + chunk : j->total;
    for (int i = i0; i < i1; ++i) {
      if (j->stage == 0) {
        int d = 0, b = i;
        while (b >= j->L->nblk[d]) { b -= j->L->nblk[d]; ++d; }
        do_block(j, d, b, buf);
      } else if (j->stage == 2) {
        do_deblock_row(j, i);
      } else {
        do_sao(j, i, buf);
      }
    }
  }
  free(buf);
  return NULL;
}

static void run_stage(job_t *j, int stage, int total, int nthreads)
{
  pthread_t th[256];
  j->stage = stage; j->next = 0; j->total = total;
  if (nthreads > 256) nthreads = 256;
  for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, worker, j);
  worker(j);
  for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
}

/* ctx must have been opened with rdoq = 0, signhide as wanted, qp = the pass QP (kvzref_ctx_open in ref_shim.c) */
int kvzref_frame_pass(kvzref_ctx *ctx, const kvz_pixel *src, int W, int H, int qp, const kvz_cuda_fp_layout *L,
                      uint8_t *blob, int nthreads)
{
  job_t j; memset(&j, 0, sizeof(j));
  j.ctx = ctx; j.src = src; j.W = W; j.H = H; j.qp = qp; j.L = L; j.blob = blob;
  encoder_state_t *st = &ctx->enc->states[0];
  st->qp = (int8_t)qp; st->frame->slicetype = KVZ_SLICE_I;
  /* RDOQ inputs (used when the ctx was opened with rdoq = 1): slice-initial context models, constant-QP lambda
   * (qp_to_lambda, rate_control.c:678-691) */
  kvz_init_contexts(st, (int8_t)qp, KVZ_SLICE_I);
  st->lambda = 0.57 * pow(2.0, (qp - 12) / 3.0);
  /* coefficient bit cost (kvz_get_coeff_cost, rdo.c:291-330): CABAC counting on the same models, no adaptation */
  memcpy(&st->search_cabac.ctx, &st->cabac.ctx, sizeof(st->cabac.ctx));
  st->search_cabac.update = 0;
  for (int d = 0; d < 4; ++d) {
    j.wl[d] = 32 >> d;
    j.rec[0][d] = (kvz_pixel *)xaligned((size_t)W * H * sizeof(kvz_pixel));
    if (d < 3) { j.rec[1][d] = (kvz_pixel *)xaligned((size_t)W * H / 4 * sizeof(kvz_pixel)); j.rec[2][d] = (kvz_pixel *)xaligned((size_t)W * H / 4 * sizeof(kvz_pixel)); }
  }
  run_stage(&j, 0, L->nblk[0] + L->nblk[1] + L->nblk[2] + L->nblk[3], nthreads);
  /* deblocking of the 8x8-level reconstruction (every CU: intra 8x8, 2Nx2N, one TU) through kvz_filter_deblock_lcu */
  {
    videoframe_t *frame = st->tile->frame;
    if (!frame->cu_array) frame->cu_array = kvz_cu_array_alloc(W, H);
    cu_info_t cu; memset(&cu, 0, sizeof(cu));
    cu.type = CU_INTRA; cu.depth = 3; cu.part_size = SIZE_2Nx2N; cu.tr_depth = 3; cu.qp = (uint8_t)qp;
    const int n_scu = (frame->cu_array->stride / 4) * (frame->cu_array->height / 4);
    for (int i = 0; i < n_scu; ++i) frame->cu_array->data[i] = cu;
    kvz_picture *saved = frame->rec, pic;
    memset(&pic, 0, sizeof(pic));
    pic.y = pic.data[0] = j.rec[0][2]; pic.u = pic.data[1] = j.rec[1][2]; pic.v = pic.data[2] = j.rec[2][2];
    pic.width = W; pic.height = H; pic.stride = W; pic.chroma_format = KVZ_CSP_420;
    frame->rec = &pic;
    st->frame->max_qp_delta_depth = -1;
    const int rows = (H + 63) / 64;
    j.dbk_done = (int *)calloc((size_t)rows, sizeof(int));
    run_stage(&j, 2, rows, nthreads < rows ? nthreads : rows);
    free(j.dbk_done);
    frame->rec = saved;
  }
  /* SAO works on the deblocked 8x8-level reconstruction; unfiltered pixels are copied first */
  kvz_pixel *sao_out = (kvz_pixel *)(blob + L->sao_rec);
  memcpy(sao_out, j.rec[0][2], (size_t)W * H * sizeof(kvz_pixel));
  memcpy(sao_out + (size_t)W * H, j.rec[1][2], (size_t)W * H / 4 * sizeof(kvz_pixel));
  memcpy(sao_out + (size_t)W * H * 5 / 4, j.rec[2][2], (size_t)W * H / 4 * sizeof(kvz_pixel));
  run_stage(&j, 1, 3 * L->nctu, nthreads);
  for (int color = 0; color < 3; ++color) {
    const int Wp = color ? W / 2 : W, Hp = color ? H / 2 : H;
    unsigned char ck[SEI_HASH_MAX_LENGTH] = { 0 };
    kvz_array_checksum(sao_out + (color == 0 ? 0 : (color == 1 ? (size_t)W * H : (size_t)W * H * 5 / 4)), Hp, Wp, Wp, ck, KVZ_BIT_DEPTH);
    memcpy(blob + L->checksum + 4 * color, ck, 4);
  }
  for (int d = 0; d < 4; ++d) for (int c = 0; c < 3; ++c) free(j.rec[c][d]);
  return 0;
}
