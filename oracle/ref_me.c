/*
 * ref_me.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The integer motion search of the UNMODIFIED reference, callable per PU: this translation unit includes the
 * reference's src/search_inter.c from where it lies (nothing is copied into the repository) so that its static
 * functions -- select_starting_point, early_terminate, hexagon_search, diamond_search, check_mv_cost, calc_mvd_cost,
 * fracmv_within_tile -- run exactly as compiled from the reference, against a real encoder_state_t of an opened
 * encoder.  Only the three calls of search_pu_inter_ref (search_inter.c:1349-1383) that sequence them are restated
 * below.  Parity checker for kvz_cuda_me_search_batch (tests/test_me_search.py).
 */
#include "search_inter.c"      /* -I/root/reference/src: the reference source, compiled in place */

#include "kvazaar_internal.h"

#include "../include/kvz_cuda.h"

typedef struct { const kvz_api *api; kvz_config *cfg; kvz_encoder *enc; } kvzref_ctx;

int kvzref_me_search(kvzref_ctx *ctx, const kvz_cuda_me_params *p, const kvz_pixel *cur, int cur_stride, const kvz_pixel *ref, int ref_stride,
                     const kvz_cuda_me_pu *pus, int count, kvz_cuda_me_result *out)
{
  encoder_state_t *state = &ctx->enc->states[0];
  encoder_control_t *ctrl = (encoder_control_t *)state->encoder_control;
  if (p->bitdepth != KVZ_BIT_DEPTH) return -1;
  if (state->tile->frame->width != p->width || state->tile->frame->height != p->height) return -2;
  if (state->tile->offset_x != 0 || state->tile->offset_y != 0) return -3;

  /* the configuration fields the search reads; restored below */
  const kvz_config saved_cfg = ctrl->cfg;
  const int saved_right = ctrl->max_inter_ref_lcu.right, saved_down = ctrl->max_inter_ref_lcu.down;
  const double saved_lambda_sqrt = state->lambda_sqrt;
  ctrl->cfg.owf = p->wpp_owf ? 1 : 0;
  ctrl->cfg.wpp = p->wpp_owf ? 1 : 0;
  ctrl->cfg.sao_type = p->delay_px == SAO_DELAY_PX ? KVZ_SAO_FULL : KVZ_SAO_OFF;
  ctrl->cfg.deblock_enable = p->delay_px == DEBLOCK_DELAY_PX ? 1 : 0;
  if (p->delay_px != 0 && p->delay_px != SAO_DELAY_PX && p->delay_px != DEBLOCK_DELAY_PX) return -4;
  ctrl->cfg.mv_constraint = (enum kvz_mv_constraint)p->mv_constraint;
  ctrl->cfg.me_early_termination = (enum kvz_me_early_termination)p->me_early_termination;
  ctrl->cfg.me_max_steps = (uint32_t)p->me_max_steps;
  ctrl->cfg.ime_algorithm = (enum kvz_ime_algorithm)p->ime_algorithm;
  ctrl->cfg.mv_rdo = 0;
  ctrl->max_inter_ref_lcu.right = p->max_ref_lcu_right;
  ctrl->max_inter_ref_lcu.down = p->max_ref_lcu_down;
  state->lambda_sqrt = p->lambda_sqrt;

  kvz_picture pic, rpic;
  memset(&pic, 0, sizeof(pic));
  memset(&rpic, 0, sizeof(rpic));
  pic.y = (kvz_pixel *)cur;  pic.width = p->width;  pic.height = p->height;  pic.stride = cur_stride;
  rpic.y = (kvz_pixel *)ref; rpic.width = p->width; rpic.height = p->height; rpic.stride = ref_stride;

  for (int i = 0; i < count; ++i) {
    const kvz_cuda_me_pu *u = &pus[i];
    inter_search_info_t info;
    memset(&info, 0, sizeof(info));
    info.state = state;
    info.pic = &pic;
    info.ref = &rpic;
    info.ref_idx = 0;
    info.origin.x = u->x; info.origin.y = u->y;
    info.width = u->w;    info.height = u->h;
    for (int c = 0; c < 2; ++c) { info.mv_cand[c][0] = u->mv_cand[c][0]; info.mv_cand[c][1] = u->mv_cand[c][1]; }
    info.num_merge_cand = u->num_merge;
    for (int m = 0; m < u->num_merge; ++m) {
      info.merge_cand[m].dir = u->merge[m].dir;
      for (int l = 0; l < 2; ++l) { info.merge_cand[m].mv[l][0] = u->merge[m].mv[l][0]; info.merge_cand[m].mv[l][1] = u->merge[m].mv[l][1]; }
    }
    info.mvd_cost_func = calc_mvd_cost;
    info.optimized_sad = kvz_get_optimized_sad(info.width);

    /* search_pu_inter_ref, search_inter.c:1281, 1334-1383 */
    vector2d_t best_mv = { 0, 0 };
    const vector2d_t mv_previous = { u->start_mv[0], u->start_mv[1] };
    if (fracmv_within_tile(&info, mv_previous.x, mv_previous.y)) best_mv = mv_previous;
    double best_cost = MAX_DOUBLE;
    double best_bits = MAX_INT;
    select_starting_point(&info, best_mv, &best_cost, &best_bits, &best_mv);
    bool skip_me = early_terminate(&info, &best_cost, &best_bits, &best_mv);
    if (!(ctrl->cfg.me_early_termination && skip_me)) {
      if (ctrl->cfg.ime_algorithm == KVZ_IME_DIA) diamond_search(&info, best_mv, ctrl->cfg.me_max_steps, &best_cost, &best_bits, &best_mv);
      else hexagon_search(&info, best_mv, ctrl->cfg.me_max_steps, &best_cost, &best_bits, &best_mv);
    }
    out[i].cost = best_cost;
    out[i].bits = (int32_t)best_bits;
    out[i].mv[0] = (int16_t)best_mv.x;
    out[i].mv[1] = (int16_t)best_mv.y;
    out[i].points = 0;
    out[i].pad = 0;
  }

  ctrl->cfg = saved_cfg;
  ctrl->max_inter_ref_lcu.right = saved_right;
  ctrl->max_inter_ref_lcu.down = saved_down;
  state->lambda_sqrt = saved_lambda_sqrt;
  return 0;
}
