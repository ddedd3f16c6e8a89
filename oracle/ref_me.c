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
      int search_range = 32;
      switch (ctrl->cfg.ime_algorithm) {
        case KVZ_IME_FULL64: search_range = 64; break;
        case KVZ_IME_FULL32: search_range = 32; break;
        case KVZ_IME_FULL16: search_range = 16; break;
        case KVZ_IME_FULL8: search_range = 8; break;
        default: break;
      }
      switch (ctrl->cfg.ime_algorithm) {
        case KVZ_IME_TZ: tz_search(&info, best_mv, &best_cost, &best_bits, &best_mv); break;
        case KVZ_IME_FULL64: case KVZ_IME_FULL32: case KVZ_IME_FULL16: case KVZ_IME_FULL8: case KVZ_IME_FULL:
          search_mv_full(&info, search_range, best_mv, &best_cost, &best_bits, &best_mv); break;
        case KVZ_IME_DIA: diamond_search(&info, best_mv, ctrl->cfg.me_max_steps, &best_cost, &best_bits, &best_mv); break;
        default: hexagon_search(&info, best_mv, ctrl->cfg.me_max_steps, &best_cost, &best_bits, &best_mv); break;
      }
    }
    if (p->satd_final && best_cost < MAX_DOUBLE) {        /* cfg->fme_level == 0, search_inter.c:1385-1397 */
      best_cost = kvz_image_calc_satd(&pic, info.ref, info.origin.x, info.origin.y, state->tile->offset_x + info.origin.x + (best_mv.x >> 2),
                                      state->tile->offset_y + info.origin.y + (best_mv.y >> 2), info.width, info.height);
      best_cost += best_bits * state->lambda_sqrt;
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

/* The reference's own search_frac (search_inter.c:974-1168) per PU, starting from pus[i].start_mv. */
int kvzref_me_frac_search(kvzref_ctx *ctx, const kvz_cuda_me_params *p, int fme_level, const kvz_pixel *cur, int cur_stride, const kvz_pixel *ref,
                          int ref_stride, const kvz_cuda_me_pu *pus, int count, kvz_cuda_me_result *out)
{
  encoder_state_t *state = &ctx->enc->states[0];
  encoder_control_t *ctrl = (encoder_control_t *)state->encoder_control;
  if (p->bitdepth != KVZ_BIT_DEPTH) return -1;
  if (state->tile->frame->width != p->width || state->tile->frame->height != p->height) return -2;
  if (p->delay_px != 0 && p->delay_px != SAO_DELAY_PX && p->delay_px != DEBLOCK_DELAY_PX) return -4;
  const kvz_config saved_cfg = ctrl->cfg;
  const int saved_right = ctrl->max_inter_ref_lcu.right, saved_down = ctrl->max_inter_ref_lcu.down;
  const double saved_lambda_sqrt = state->lambda_sqrt;
  ctrl->cfg.owf = p->wpp_owf ? 1 : 0;
  ctrl->cfg.wpp = p->wpp_owf ? 1 : 0;
  ctrl->cfg.sao_type = p->delay_px == SAO_DELAY_PX ? KVZ_SAO_FULL : KVZ_SAO_OFF;
  ctrl->cfg.deblock_enable = p->delay_px == DEBLOCK_DELAY_PX ? 1 : 0;
  ctrl->cfg.mv_constraint = (enum kvz_mv_constraint)p->mv_constraint;
  ctrl->cfg.mv_rdo = 0;
  ctrl->cfg.fme_level = fme_level;
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
    info.origin.x = u->x; info.origin.y = u->y;
    info.width = u->w;    info.height = u->h;
    for (int c = 0; c < 2; ++c) { info.mv_cand[c][0] = u->mv_cand[c][0]; info.mv_cand[c][1] = u->mv_cand[c][1]; }
    info.mvd_cost_func = calc_mvd_cost;
    info.optimized_sad = kvz_get_optimized_sad(info.width);
    vector2d_t mv = { u->start_mv[0], u->start_mv[1] };
    double cost = MAX_DOUBLE, bits = MAX_INT;
    search_frac(&info, &cost, &bits, &mv);
    out[i].cost = cost;
    out[i].bits = (int32_t)bits;
    out[i].mv[0] = (int16_t)mv.x;
    out[i].mv[1] = (int16_t)mv.y;
    out[i].points = 0;
    out[i].pad = 0;
  }
  ctrl->cfg = saved_cfg;
  ctrl->max_inter_ref_lcu.right = saved_right;
  ctrl->max_inter_ref_lcu.down = saved_down;
  state->lambda_sqrt = saved_lambda_sqrt;
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * AMVP and merge candidates through the reference's own kvz_inter_get_mv_cand_cua / kvz_inter_get_merge_cand
 * (src/inter.c, linked from the unmodified library): the CU records of the test become a cu_array_t (and, for the merge
 * candidates, the lcu_t work copy the search uses), the reference lists an image_list_t.  Parity checker for
 * kvz_cuda_me_candidates_batch. */
static cu_array_t *cua_from_records(const kvz_cuda_me_cu *recs, int stride, int rows, int width, int height)
{
  cu_array_t *cua = kvz_cu_array_alloc(width, height);
  const int w4 = cua->width / 4, h4 = cua->height / 4;
  for (int y = 0; y < h4 && y < rows; ++y)
    for (int x = 0; x < w4 && x < stride; ++x) {
      const kvz_cuda_me_cu *r = &recs[y * stride + x];
      cu_info_t *c = &cua->data[y * w4 + x];
      c->type = r->type;
      if (r->type == CU_INTER) {
        c->inter.mv_dir = r->mv_dir;
        for (int l = 0; l < 2; ++l) { c->inter.mv[l][0] = r->mv[l][0]; c->inter.mv[l][1] = r->mv[l][1]; c->inter.mv_ref[l] = r->mv_ref[l]; }
      }
    }
  return cua;
}

int kvzref_me_candidates(kvzref_ctx *ctx, const kvz_cuda_me_frame *f, const int32_t *col_pic_ref_pocs, const uint8_t *col_ref_LXs /* [2][16] */,
                         const kvz_cuda_me_cu *cus, int cu_stride, const kvz_cuda_me_cu *col_cus, int col_stride, int cu_rows,
                         const kvz_cuda_me_cand_pu *pus, int count, kvz_cuda_me_cand_out *out)
{
  encoder_state_t *state = &ctx->enc->states[0];
  encoder_control_t *ctrl = (encoder_control_t *)state->encoder_control;
  if (ctrl->in.width != f->width || ctrl->in.height != f->height) return -2;
  if (state->tile->frame->width != f->width || state->tile->frame->height != f->height) return -2;

  cu_array_t *cua = cua_from_records(cus, cu_stride, cu_rows, f->width, f->height);
  cu_array_t *col = cua_from_records(col_cus, col_stride, cu_rows, f->width, f->height);
  image_list_t *list = kvz_image_list_alloc(16);
  kvz_picture *pics[16];
  for (int i = 0; i < 16; ++i) {
    pics[i] = calloc(1, sizeof(kvz_picture));
    for (int r = 0; r < 16; ++r) pics[i]->ref_pocs[r] = col_pic_ref_pocs[r];
    list->images[i] = pics[i];
    list->cu_arrays[i] = col;
    list->pocs[i] = f->pocs[i];
    memcpy(list->ref_LXs[i], col_ref_LXs, 32);
  }
  list->used_size = (uint32_t)f->used_size;

  /* swap the test's frame description in; everything is restored below */
  image_list_t *saved_ref = state->frame->ref;
  cu_array_t *saved_cua = state->tile->frame->cu_array;
  const int32_t saved_poc = state->frame->poc;
  const int saved_slicetype = state->frame->slicetype;
  uint8_t saved_LX[2][16]; memcpy(saved_LX, state->frame->ref_LX, 32);
  uint8_t saved_LX_size[2] = { state->frame->ref_LX_size[0], state->frame->ref_LX_size[1] };
  const kvz_config saved_cfg = ctrl->cfg;
  state->frame->ref = list;
  state->tile->frame->cu_array = cua;
  state->frame->poc = f->poc;
  state->frame->slicetype = f->slice_b ? KVZ_SLICE_B : KVZ_SLICE_P;
  memcpy(state->frame->ref_LX, f->ref_LX, 32);
  state->frame->ref_LX_size[0] = (uint8_t)f->ref_LX_size[0];
  state->frame->ref_LX_size[1] = (uint8_t)f->ref_LX_size[1];
  ctrl->cfg.tmvp_enable = f->tmvp_enable;
  ctrl->cfg.max_merge = (uint8_t)f->max_merge;

  lcu_t *lcu = calloc(1, sizeof(lcu_t));
  const int w4 = cua->width / 4;
  for (int i = 0; i < count; ++i) {
    const kvz_cuda_me_cand_pu *u = &pus[i];
    memset(&out[i], 0, sizeof(out[i]));
    cu_info_t cur; memset(&cur, 0, sizeof(cur));
    cur.type = CU_INTER;
    cur.inter.mv_ref[0] = u->mv_ref[0];
    cur.inter.mv_ref[1] = u->mv_ref[1];
    for (int l = 0; l < 2; ++l) {
      if (f->ref_LX_size[l] <= 0) continue;
      int16_t mvc[2][2] = { { 0, 0 }, { 0, 0 } };
      kvz_inter_get_mv_cand_cua(state, u->x, u->y, u->w, u->h, mvc, &cur, (int8_t)l);
      memcpy(out[i].mv_cand[l], mvc, sizeof(mvc));
    }
    /* the lcu_t work copy: the CUs of this LCU plus the column to the left, the row above and the top-right CU */
    memset(lcu->cu, 0, sizeof(lcu->cu));
    const int lx = (u->x / LCU_WIDTH) * LCU_WIDTH, ly = (u->y / LCU_WIDTH) * LCU_WIDTH;
    for (int yl = -4; yl < LCU_WIDTH; yl += 4)
      for (int xl = -4; xl < LCU_WIDTH; xl += 4) {
        const int X = lx + xl, Y = ly + yl;
        if (X < 0 || Y < 0 || X >= cua->width || Y >= cua->height) continue;
        *LCU_GET_CU_AT_PX(lcu, xl, yl) = cua->data[(Y / 4) * w4 + X / 4];
      }
    if (ly > 0 && lx + LCU_WIDTH < cua->width) *LCU_GET_TOP_RIGHT_CU(lcu) = cua->data[((ly - 1) / 4) * w4 + (lx + LCU_WIDTH) / 4];
    inter_merge_cand_t mc[MRG_MAX_NUM_CANDS];
    memset(mc, 0, sizeof(mc));
    out[i].num_merge = kvz_inter_get_merge_cand(state, u->x, u->y, u->w, u->h, u->use_a1 != 0, u->use_b1 != 0, mc, lcu);
    for (int m = 0; m < MRG_MAX_NUM_CANDS; ++m) {
      out[i].merge[m].dir = mc[m].dir;
      for (int l = 0; l < 2; ++l) { out[i].merge[m].ref[l] = mc[m].ref[l]; out[i].merge[m].mv[l][0] = mc[m].mv[l][0]; out[i].merge[m].mv[l][1] = mc[m].mv[l][1]; }
    }
  }
  free(lcu);

  state->frame->ref = saved_ref;
  state->tile->frame->cu_array = saved_cua;
  state->frame->poc = saved_poc;
  state->frame->slicetype = saved_slicetype;
  memcpy(state->frame->ref_LX, saved_LX, 32);
  state->frame->ref_LX_size[0] = saved_LX_size[0];
  state->frame->ref_LX_size[1] = saved_LX_size[1];
  ctrl->cfg = saved_cfg;
  for (int i = 0; i < 16; ++i) { list->images[i] = NULL; list->cu_arrays[i] = NULL; free(pics[i]); }
  list->used_size = 0;
  kvz_image_list_destroy(list);
  kvz_cu_array_free(&cua);
  kvz_cu_array_free(&col);
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * The merge analysis of search_pu_inter (search_inter.c:1667-1730, rdo < 3) with the reference's own primitives:
 * merge_candidate_in_list, fracmv_within_tile (static, from the included file), kvz_inter_pred_pu, kvz_satd_any_size,
 * kvz_sort_keys_by_cost and the CTX_ENTROPY_FBITS macro on real context models.  Only the loop that sequences them is
 * restated.  cu[i] = { x_cu, y_cu, width_cu, part_mode, i_pu } of the PU.  Parity checker for kvz_cuda_me_merge_cost_batch. */
int kvzref_me_merge_cost(kvzref_ctx *ctx, const kvz_cuda_me_params *p, int n_pics, const kvz_pixel *const *pic_planes, const uint8_t *ref_LX /* [2][16] */,
                         const int32_t *ref_LX_size, int bipred, int merge_flag_state, int merge_idx_state, double *bits_out /* [3] */,
                         const kvz_pixel *cur, int cur_stride, const kvz_cuda_me_pu *pus, const int32_t *cu /* [count][5] */, int count,
                         kvz_cuda_me_merge_cost *out)
{
  encoder_state_t *state = &ctx->enc->states[0];
  encoder_control_t *ctrl = (encoder_control_t *)state->encoder_control;
  if (p->bitdepth != KVZ_BIT_DEPTH) return -1;
  if (state->tile->frame->width != p->width || state->tile->frame->height != p->height) return -2;
  if (p->delay_px != 0 && p->delay_px != SAO_DELAY_PX && p->delay_px != DEBLOCK_DELAY_PX) return -4;

  const kvz_config saved_cfg = ctrl->cfg;
  const int saved_right = ctrl->max_inter_ref_lcu.right, saved_down = ctrl->max_inter_ref_lcu.down;
  const double saved_lambda_sqrt = state->lambda_sqrt;
  image_list_t *saved_ref = state->frame->ref;
  uint8_t saved_LX[2][16]; memcpy(saved_LX, state->frame->ref_LX, 32);
  uint8_t saved_LX_size[2] = { state->frame->ref_LX_size[0], state->frame->ref_LX_size[1] };
  ctrl->cfg.owf = p->wpp_owf ? 1 : 0;
  ctrl->cfg.wpp = p->wpp_owf ? 1 : 0;
  ctrl->cfg.sao_type = p->delay_px == SAO_DELAY_PX ? KVZ_SAO_FULL : KVZ_SAO_OFF;
  ctrl->cfg.deblock_enable = p->delay_px == DEBLOCK_DELAY_PX ? 1 : 0;
  ctrl->cfg.mv_constraint = (enum kvz_mv_constraint)p->mv_constraint;
  ctrl->cfg.bipred = bipred;
  ctrl->max_inter_ref_lcu.right = p->max_ref_lcu_right;
  ctrl->max_inter_ref_lcu.down = p->max_ref_lcu_down;
  state->lambda_sqrt = p->lambda_sqrt;

  image_list_t *list = kvz_image_list_alloc(16);
  kvz_picture *pics[16];
  for (int i = 0; i < 16; ++i) {
    pics[i] = calloc(1, sizeof(kvz_picture));
    pics[i]->y = (kvz_pixel *)pic_planes[i < n_pics ? i : 0];
    pics[i]->width = p->width; pics[i]->height = p->height; pics[i]->stride = p->width;
    list->images[i] = pics[i];
  }
  list->used_size = (uint32_t)n_pics;
  state->frame->ref = list;
  memcpy(state->frame->ref_LX, ref_LX, 32);
  state->frame->ref_LX_size[0] = (uint8_t)ref_LX_size[0];
  state->frame->ref_LX_size[1] = (uint8_t)ref_LX_size[1];

  /* the two context models the merge bits read */
  cabac_ctx_t saved_flag = state->search_cabac.ctx.cu_merge_flag_ext_model, saved_idx = state->search_cabac.ctx.cu_merge_idx_ext_model;
  state->search_cabac.ctx.cu_merge_flag_ext_model.uc_state = (uint8_t)merge_flag_state;
  state->search_cabac.ctx.cu_merge_idx_ext_model.uc_state = (uint8_t)merge_idx_state;
  const double merge_flag_cost = CTX_ENTROPY_FBITS(&state->search_cabac.ctx.cu_merge_flag_ext_model, 1);
  bits_out[0] = merge_flag_cost;
  bits_out[1] = CTX_ENTROPY_FBITS(&state->search_cabac.ctx.cu_merge_idx_ext_model, 0);
  bits_out[2] = CTX_ENTROPY_FBITS(&state->search_cabac.ctx.cu_merge_idx_ext_model, 1);

  kvz_picture pic;
  memset(&pic, 0, sizeof(pic));
  pic.y = (kvz_pixel *)cur; pic.width = p->width; pic.height = p->height; pic.stride = cur_stride;
  lcu_t *lcu = calloc(1, sizeof(lcu_t));
  unit_stats_map_t *merge = calloc(1, sizeof(unit_stats_map_t));
  for (int i = 0; i < count; ++i) {
    const kvz_cuda_me_pu *u = &pus[i];
    const int x_cu = cu[5 * i], y_cu = cu[5 * i + 1], width_cu = cu[5 * i + 2], part_mode = cu[5 * i + 3], i_pu = cu[5 * i + 4];
    const int x = u->x, y = u->y, width = u->w, height = u->h;
    const int x_local = SUB_SCU(x), y_local = SUB_SCU(y);
    inter_search_info_t info;
    memset(&info, 0, sizeof(info));
    info.state = state; info.pic = &pic; info.origin.x = x; info.origin.y = y; info.width = width; info.height = height;
    info.num_merge_cand = u->num_merge;
    for (int m = 0; m < u->num_merge; ++m) {
      info.merge_cand[m].dir = u->merge[m].dir;
      for (int l = 0; l < 2; ++l) {
        info.merge_cand[m].ref[l] = u->merge[m].ref[l];
        info.merge_cand[m].mv[l][0] = u->merge[m].mv[l][0]; info.merge_cand[m].mv[l][1] = u->merge[m].mv[l][1];
      }
    }
    /* the LCU work copy: source pixels of the PU, the CU's partitioning */
    memset(lcu->cu, 0, sizeof(lcu->cu));
    for (int r = 0; r < height; ++r) memcpy(&lcu->ref.y[(y_local + r) * LCU_WIDTH + x_local], &cur[(size_t)(y + r) * cur_stride + x], (size_t)width * sizeof(kvz_pixel));
    cu_info_t *cu_rec = LCU_GET_CU_AT_PX(lcu, SUB_SCU(x_cu), SUB_SCU(y_cu));
    cu_rec->type = CU_INTER; cu_rec->part_size = part_mode;
    cu_info_t *cur_pu = LCU_GET_CU_AT_PX(lcu, x_local, y_local);
    cur_pu->type = CU_INTER; cur_pu->part_size = part_mode;

    merge->size = 0;
    for (int k = 0; k < MRG_MAX_NUM_CANDS; ++k) { merge->keys[k] = -1; merge->cost[k] = MAX_DOUBLE; merge->bits[k] = 0; }
    for (int merge_idx = 0; merge_idx < info.num_merge_cand; ++merge_idx) {
      inter_merge_cand_t *cur_cand = &info.merge_cand[merge_idx];
      cur_pu->inter.mv_dir = cur_cand->dir;
      cur_pu->inter.mv_ref[0] = cur_cand->ref[0]; cur_pu->inter.mv_ref[1] = cur_cand->ref[1];
      cur_pu->inter.mv[0][0] = cur_cand->mv[0][0]; cur_pu->inter.mv[0][1] = cur_cand->mv[0][1];
      cur_pu->inter.mv[1][0] = cur_cand->mv[1][0]; cur_pu->inter.mv[1][1] = cur_cand->mv[1][1];
      if (cur_pu->inter.mv_dir == 3 && !ctrl->cfg.bipred) continue;
      if (cur_pu->inter.mv_dir == 3 && !(width + height > 12)) continue;
      bool is_duplicate = merge_candidate_in_list(info.merge_cand, cur_cand, merge);
      bool active_L0 = cur_pu->inter.mv_dir & 1, active_L1 = cur_pu->inter.mv_dir & 2;
      if ((active_L0 && !fracmv_within_tile(&info, cur_pu->inter.mv[0][0], cur_pu->inter.mv[0][1])) ||
          (active_L1 && !fracmv_within_tile(&info, cur_pu->inter.mv[1][0], cur_pu->inter.mv[1][1])) || is_duplicate) continue;
      kvz_inter_pred_pu(state, lcu, x_cu, y_cu, width_cu, true, false, i_pu);
      merge->unit[merge->size] = *cur_pu;
      merge->unit[merge->size].merge_idx = merge_idx;
      double bits = merge_flag_cost + merge_idx + CTX_ENTROPY_FBITS(&(state->search_cabac.ctx.cu_merge_idx_ext_model), merge_idx != 0);
      merge->cost[merge->size] = kvz_satd_any_size(width, height, lcu->rec.y + y_local * LCU_WIDTH + x_local, LCU_WIDTH,
                                                   lcu->ref.y + y_local * LCU_WIDTH + x_local, LCU_WIDTH);
      merge->cost[merge->size] += bits * info.state->lambda_sqrt;
      merge->bits[merge->size] = bits;
      merge->keys[merge->size] = merge->size;
      merge->size++;
    }
    kvz_sort_keys_by_cost(merge);
    memset(&out[i], 0, sizeof(out[i]));
    out[i].size = merge->size;
    for (int k = 0; k < 5; ++k) {
      out[i].cost[k] = merge->cost[k];
      out[i].bits[k] = k < merge->size ? merge->bits[k] : 0;
      out[i].keys[k] = merge->keys[k];
      out[i].merge_idx[k] = k < merge->size ? (int8_t)merge->unit[k].merge_idx : 0;
    }
  }
  free(merge);
  free(lcu);

  state->search_cabac.ctx.cu_merge_flag_ext_model = saved_flag;
  state->search_cabac.ctx.cu_merge_idx_ext_model = saved_idx;
  state->frame->ref = saved_ref;
  memcpy(state->frame->ref_LX, saved_LX, 32);
  state->frame->ref_LX_size[0] = saved_LX_size[0];
  state->frame->ref_LX_size[1] = saved_LX_size[1];
  for (int i = 0; i < 16; ++i) { list->images[i] = NULL; free(pics[i]); }
  list->used_size = 0;
  kvz_image_list_destroy(list);
  ctrl->cfg = saved_cfg;
  ctrl->max_inter_ref_lcu.right = saved_right;
  ctrl->max_inter_ref_lcu.down = saved_down;
  state->lambda_sqrt = saved_lambda_sqrt;
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Bi-prediction from two uni-predictions (search_pu_inter, search_inter.c:1937-2031) with the reference's own
 * kvz_inter_recon_bipred, kvz_satd_any_size, calc_mvd_cost and select_mv_cand; only their sequence is restated (the
 * AMVP candidates info->mv_cand holds at that point are an input).  Parity checker for kvz_cuda_me_bipred_batch. */
int kvzref_me_bipred(kvzref_ctx *ctx, const kvz_cuda_me_params *p, int n_pics, const kvz_pixel *const *pic_planes, const uint8_t *ref_LX /* [2][16] */,
                     int bipred, const kvz_pixel *cur, int cur_stride, const kvz_cuda_me_bipred_pu *pus, int count, kvz_cuda_me_bipred_result *out)
{
  encoder_state_t *state = &ctx->enc->states[0];
  encoder_control_t *ctrl = (encoder_control_t *)state->encoder_control;
  if (p->bitdepth != KVZ_BIT_DEPTH) return -1;
  if (state->tile->frame->width != p->width || state->tile->frame->height != p->height) return -2;
  const kvz_config saved_cfg = ctrl->cfg;
  const double saved_lambda_sqrt = state->lambda_sqrt;
  image_list_t *saved_ref = state->frame->ref;
  uint8_t saved_LX[2][16]; memcpy(saved_LX, state->frame->ref_LX, 32);
  ctrl->cfg.bipred = bipred;
  ctrl->cfg.mv_rdo = 0;
  state->lambda_sqrt = p->lambda_sqrt;
  image_list_t *list = kvz_image_list_alloc(16);
  kvz_picture *pics[16];
  for (int i = 0; i < 16; ++i) {
    pics[i] = calloc(1, sizeof(kvz_picture));
    pics[i]->y = (kvz_pixel *)pic_planes[i < n_pics ? i : 0];
    pics[i]->width = p->width; pics[i]->height = p->height; pics[i]->stride = p->width;
    list->images[i] = pics[i];
  }
  list->used_size = (uint32_t)n_pics;
  state->frame->ref = list;
  memcpy(state->frame->ref_LX, ref_LX, 32);

  lcu_t *lcu = calloc(1, sizeof(lcu_t));
  for (int i = 0; i < count; ++i) {
    const kvz_cuda_me_bipred_pu *u = &pus[i];
    const int x = u->x, y = u->y, width = u->w, height = u->h;
    memset(&out[i], 0, sizeof(out[i]));
    out[i].cost = MAX_DOUBLE;
    const bool can_use_bipred = ctrl->cfg.bipred && width + height >= 16;      /* the slice-type condition is the caller's */
    if (!can_use_bipred) continue;
    for (int r = 0; r < height; ++r)
      memcpy(&lcu->ref.y[(SUB_SCU(y) + r) * LCU_WIDTH + SUB_SCU(x)], &cur[(size_t)(y + r) * cur_stride + x], (size_t)width * sizeof(kvz_pixel));
    int16_t mv[2][2] = { { u->mv[0][0], u->mv[0][1] }, { u->mv[1][0], u->mv[1][1] } };
    int16_t mv_cand[2][2] = { { u->mv_cand[0][0], u->mv_cand[0][1] }, { u->mv_cand[1][0], u->mv_cand[1][1] } };
    kvz_inter_recon_bipred(state, list->images[state->frame->ref_LX[0][u->mv_ref[0]]], list->images[state->frame->ref_LX[1][u->mv_ref[1]]],
                           x, y, width, height, mv, lcu, true, false);
    const kvz_pixel *rec = &lcu->rec.y[SUB_SCU(y) * LCU_WIDTH + SUB_SCU(x)];
    const kvz_pixel *src = &lcu->ref.y[SUB_SCU(y) * LCU_WIDTH + SUB_SCU(x)];
    double best_bipred_cost = kvz_satd_any_size(width, height, rec, LCU_WIDTH, src, LCU_WIDTH);
    double bitcost[2] = { 0, 0 };
    best_bipred_cost += calc_mvd_cost(state, mv[0][0], mv[0][1], 0, mv_cand, NULL, 0, 0, &bitcost[0]);
    best_bipred_cost += calc_mvd_cost(state, mv[1][0], mv[1][1], 0, mv_cand, NULL, 0, 0, &bitcost[1]);
    const uint8_t mv_ref_coded[2] = { u->mv_ref[0], u->mv_ref[1] };
    const int extra_bits = mv_ref_coded[0] + mv_ref_coded[1] + 2 /* mv dir cost */;
    best_bipred_cost += state->lambda_sqrt * extra_bits;
    out[i].cost = best_bipred_cost;
    out[i].bits = (int32_t)(bitcost[0] + bitcost[1] + extra_bits);
    for (int reflist = 0; reflist < 2; reflist++) out[i].mv_cand_idx[reflist] = (uint8_t)select_mv_cand(state, mv_cand, mv[reflist][0], mv[reflist][1], NULL);
    out[i].valid = 1;
  }
  free(lcu);
  state->frame->ref = saved_ref;
  memcpy(state->frame->ref_LX, saved_LX, 32);
  for (int i = 0; i < 16; ++i) { list->images[i] = NULL; free(pics[i]); }
  list->used_size = 0;
  kvz_image_list_destroy(list);
  ctrl->cfg = saved_cfg;
  state->lambda_sqrt = saved_lambda_sqrt;
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Motion compensation through the reference's own kvz_inter_pred_pu (luma + chroma, one or two lists): the PU's prediction
 * is taken from the lcu_t work copy into an I420 picture.  cu[i] = { x_cu, y_cu, width_cu, part_mode, i_pu }.
 * Parity checker for kvz_cuda_me_predict_batch. */
int kvzref_me_predict(kvzref_ctx *ctx, const kvz_cuda_me_params *p, int n_pics, const kvz_pixel *const *ys, const kvz_pixel *const *us,
                      const kvz_pixel *const *vs, const uint8_t *ref_LX /* [2][16] */, const kvz_cuda_me_mc_pu *pus, const int32_t *cu, int count,
                      kvz_pixel *out_y, kvz_pixel *out_u, kvz_pixel *out_v)
{
  encoder_state_t *state = &ctx->enc->states[0];
  encoder_control_t *ctrl = (encoder_control_t *)state->encoder_control;
  if (p->bitdepth != KVZ_BIT_DEPTH) return -1;
  if (state->tile->frame->width != p->width || state->tile->frame->height != p->height) return -2;
  const kvz_config saved_cfg = ctrl->cfg;
  image_list_t *saved_ref = state->frame->ref;
  uint8_t saved_LX[2][16]; memcpy(saved_LX, state->frame->ref_LX, 32);
  ctrl->cfg.bipred = 1;
  image_list_t *list = kvz_image_list_alloc(16);
  kvz_picture *pics[16];
  for (int i = 0; i < 16; ++i) {
    const int k = i < n_pics ? i : 0;
    pics[i] = calloc(1, sizeof(kvz_picture));
    pics[i]->y = (kvz_pixel *)ys[k]; pics[i]->u = (kvz_pixel *)us[k]; pics[i]->v = (kvz_pixel *)vs[k];
    pics[i]->width = p->width; pics[i]->height = p->height; pics[i]->stride = p->width;
    list->images[i] = pics[i];
  }
  list->used_size = (uint32_t)n_pics;
  state->frame->ref = list;
  memcpy(state->frame->ref_LX, ref_LX, 32);

  lcu_t *lcu = calloc(1, sizeof(lcu_t));
  const int W = p->width, Wc = p->width / 2;
  for (int i = 0; i < count; ++i) {
    const kvz_cuda_me_mc_pu *u = &pus[i];
    const int x_cu = cu[5 * i], y_cu = cu[5 * i + 1], width_cu = cu[5 * i + 2], part_mode = cu[5 * i + 3], i_pu = cu[5 * i + 4];
    const int xl = SUB_SCU(u->x), yl = SUB_SCU(u->y);
    memset(lcu->cu, 0, sizeof(lcu->cu));
    cu_info_t *cu_rec = LCU_GET_CU_AT_PX(lcu, SUB_SCU(x_cu), SUB_SCU(y_cu));
    cu_rec->type = CU_INTER; cu_rec->part_size = part_mode;
    cu_info_t *pu = LCU_GET_CU_AT_PX(lcu, xl, yl);
    pu->type = CU_INTER; pu->part_size = part_mode;
    pu->inter.mv_dir = u->dir;
    for (int l = 0; l < 2; ++l) { pu->inter.mv_ref[l] = u->mv_ref[l]; pu->inter.mv[l][0] = u->mv[l][0]; pu->inter.mv[l][1] = u->mv[l][1]; }
    kvz_inter_pred_pu(state, lcu, x_cu, y_cu, width_cu, true, true, i_pu);
    for (int r = 0; r < u->h; ++r) memcpy(&out_y[(size_t)(u->y + r) * W + u->x], &lcu->rec.y[(yl + r) * LCU_WIDTH + xl], (size_t)u->w * sizeof(kvz_pixel));
    for (int r = 0; r < u->h / 2; ++r) {
      memcpy(&out_u[(size_t)(u->y / 2 + r) * Wc + u->x / 2], &lcu->rec.u[(yl / 2 + r) * LCU_WIDTH_C + xl / 2], (size_t)(u->w / 2) * sizeof(kvz_pixel));
      memcpy(&out_v[(size_t)(u->y / 2 + r) * Wc + u->x / 2], &lcu->rec.v[(yl / 2 + r) * LCU_WIDTH_C + xl / 2], (size_t)(u->w / 2) * sizeof(kvz_pixel));
    }
  }
  free(lcu);
  state->frame->ref = saved_ref;
  memcpy(state->frame->ref_LX, saved_LX, 32);
  for (int i = 0; i < 16; ++i) { list->images[i] = NULL; free(pics[i]); }
  list->used_size = 0;
  kvz_image_list_destroy(list);
  ctrl->cfg = saved_cfg;
  return 0;
}
