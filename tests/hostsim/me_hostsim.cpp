// me_hostsim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Host build of the integer motion search's single-source algorithm (kvazaar_b200/csrc/me/me_search.h) with the 32
// lane shares of every SAD walked in turn: checks the control flow of the device code against the reference without a GPU (tests/test_me_search.py).
// Exports the same entry points as libkvzcuda.so, with host pointers.
#include "../../kvazaar_b200/csrc/me/me_search.h"
#include "../../kvazaar_b200/csrc/me/me_cand.h"
#include "../../kvazaar_b200/csrc/me/me_frac.h"
#include "../../kvazaar_b200/csrc/me/me_merge.h"
#include "../../kvazaar_b200/csrc/me/me_mc.h"

extern "C" int kvz_cuda_me_params_supported(const kvz_cuda_me_params *p) { return p ? kvzme::params_supported(*p) : -1; }

template <typename Pix>
static void run(const kvz_cuda_me_params *p, const void *cur, int cur_stride, const void *ref, int ref_stride, const kvz_cuda_me_pu *pus, int count,
                kvz_cuda_me_result *out)
{
  const kvzme::Lanes ln = { 0, 32 };      // lane 0 writes the result; pu_sad walks all 32 shares
  const kvzme::Planes<Pix> pl = { (const Pix *)cur, (const Pix *)ref, cur_stride, ref_stride };
  for (int i = 0; i < count; ++i) {
    if (p->satd_final) kvzme::search_pu_satd_final<Pix>(ln, *p, pus[i], pl, &out[i]);
    else kvzme::search_pu<Pix>(ln, *p, pus[i], pl, &out[i]);
  }
}

extern "C" int kvz_cuda_call_me_search(const kvz_cuda_me_params *p, const void *cur, int cur_stride, const void *ref, int ref_stride,
                                       const kvz_cuda_me_pu *pus, int count, kvz_cuda_me_result *out)
{
  if (!p || kvzme::params_supported(*p) != 0) return -2;
  if (p->bitdepth == 8) run<uint8_t>(p, cur, cur_stride, ref, ref_stride, pus, count, out);
  else run<uint16_t>(p, cur, cur_stride, ref, ref_stride, pus, count, out);
  return 0;
}

extern "C" int kvz_cuda_me_search_batch(const kvz_cuda_me_params *p, const void *cur, int cur_stride, const void *ref, int ref_stride,
                                        const kvz_cuda_me_pu *pus, int count, kvz_cuda_me_result *out, void *)
{
  return kvz_cuda_call_me_search(p, cur, cur_stride, ref, ref_stride, pus, count, out);
}

extern "C" int kvz_cuda_call_me_candidates(const kvz_cuda_me_frame *f, const kvz_cuda_me_cu *cus, int cu_stride, const kvz_cuda_me_cu *col_cus,
                                           int col_stride, int, const kvz_cuda_me_cand_pu *pus, int count, kvz_cuda_me_cand_out *out)
{
  if (!f || kvzme::frame_supported(*f) != 0) return -2;
  const kvzme::CuImage cur = { cus, cu_stride }, col = { col_cus, col_stride };
  for (int i = 0; i < count; ++i) kvzme::candidates_of_pu(*f, cur, col, pus[i], &out[i]);
  return 0;
}

extern "C" int kvz_cuda_me_candidates_batch(const kvz_cuda_me_frame *f, const kvz_cuda_me_cu *cus, int cu_stride, const kvz_cuda_me_cu *col_cus,
                                            int col_stride, const kvz_cuda_me_cand_pu *pus, int count, kvz_cuda_me_cand_out *out, void *)
{
  return kvz_cuda_call_me_candidates(f, cus, cu_stride, col_cus, col_stride, 0, pus, count, out);
}

template <typename Pix>
static void run_frac(const kvz_cuda_me_params *p, int levels, const void *cur, int cur_stride, const void *ref, int ref_stride, const kvz_cuda_me_pu *pus,
                     int count, kvz_cuda_me_result *out)
{
  const kvzme::Lanes ln = { 0, 32 };
  const kvzme::Planes<Pix> pl = { (const Pix *)cur, (const Pix *)ref, cur_stride, ref_stride };
  for (int i = 0; i < count; ++i) kvzme::frac_search_pu<Pix>(ln, *p, pus[i], pl, levels, &out[i]);
}

extern "C" int kvz_cuda_call_me_frac_search(const kvz_cuda_me_params *p, int fme_level, const void *cur, int cur_stride, const void *ref,
                                            int ref_stride, const kvz_cuda_me_pu *pus, int count, kvz_cuda_me_result *out)
{
  if (!p || kvzme::params_supported(*p) != 0 || fme_level < 1 || fme_level > 4) return -2;
  if (p->bitdepth == 8) run_frac<uint8_t>(p, fme_level, cur, cur_stride, ref, ref_stride, pus, count, out);
  else run_frac<uint16_t>(p, fme_level, cur, cur_stride, ref, ref_stride, pus, count, out);
  return 0;
}

extern "C" int kvz_cuda_me_frac_search_batch(const kvz_cuda_me_params *p, int fme_level, const void *cur, int cur_stride, const void *ref,
                                             int ref_stride, const kvz_cuda_me_pu *pus, int count, kvz_cuda_me_result *out, void *)
{
  return kvz_cuda_call_me_frac_search(p, fme_level, cur, cur_stride, ref, ref_stride, pus, count, out);
}

template <typename Pix>
static void run_merge(const kvz_cuda_me_params *p, const kvz_cuda_me_refs *rf, const void *cur, int cur_stride, const kvz_cuda_me_pu *pus, int count,
                      kvz_cuda_me_merge_cost *out)
{
  const kvzme::Lanes ln = { 0, 32 };
  const kvzme::Planes<Pix> pl = { (const Pix *)cur, nullptr, cur_stride, 0 };
  kvzme::RefSet<Pix> rs;
  for (int i = 0; i < 16; ++i) { rs.plane[i] = (const Pix *)rf->plane[i]; rs.stride[i] = rf->stride[i]; }
  for (int i = 0; i < count; ++i) kvzme::merge_cost_pu<Pix>(ln, *p, *rf, rs, pus[i], pl, &out[i]);
}

extern "C" int kvz_cuda_me_merge_cost_batch(const kvz_cuda_me_params *p, const kvz_cuda_me_refs *refs, const void *cur, int cur_stride,
                                            const kvz_cuda_me_pu *pus, int count, kvz_cuda_me_merge_cost *out, void *)
{
  if (!p || !refs || kvzme::params_supported(*p) != 0) return -2;
  if (p->bitdepth == 8) run_merge<uint8_t>(p, refs, cur, cur_stride, pus, count, out);
  else run_merge<uint16_t>(p, refs, cur, cur_stride, pus, count, out);
  return 0;
}

template <typename Pix>
static void run_bipred(const kvz_cuda_me_params *p, const kvz_cuda_me_refs *rf, const void *cur, int cur_stride, const kvz_cuda_me_bipred_pu *pus, int count,
                       kvz_cuda_me_bipred_result *out)
{
  const kvzme::Lanes ln = { 0, 32 };
  const kvzme::Planes<Pix> pl = { (const Pix *)cur, nullptr, cur_stride, 0 };
  kvzme::RefSet<Pix> rs;
  for (int i = 0; i < 16; ++i) { rs.plane[i] = (const Pix *)rf->plane[i]; rs.stride[i] = rf->stride[i]; }
  for (int i = 0; i < count; ++i) kvzme::bipred_pu<Pix>(ln, *p, *rf, rs, pus[i], pl, &out[i]);
}

extern "C" int kvz_cuda_me_bipred_batch(const kvz_cuda_me_params *p, const kvz_cuda_me_refs *refs, const void *cur, int cur_stride,
                                        const kvz_cuda_me_bipred_pu *pus, int count, kvz_cuda_me_bipred_result *out, void *)
{
  if (!p || !refs || kvzme::params_supported(*p) != 0) return -2;
  if (p->bitdepth == 8) run_bipred<uint8_t>(p, refs, cur, cur_stride, pus, count, out);
  else run_bipred<uint16_t>(p, refs, cur, cur_stride, pus, count, out);
  return 0;
}

template <typename Pix>
static void run_predict(const kvz_cuda_me_params *p, const kvz_cuda_me_mc_refs *rf, const kvz_cuda_me_mc_pu *pus, int count, void *y, void *u, void *v)
{
  kvzme::McRefs<Pix> rs;
  for (int i = 0; i < 16; ++i) { rs.y[i] = (const Pix *)rf->y[i]; rs.u[i] = (const Pix *)rf->u[i]; rs.v[i] = (const Pix *)rf->v[i]; }
  for (int i = 0; i < count; ++i)
    for (int l = 0; l < 32; ++l) kvzme::predict_pu<Pix>(kvzme::Lanes{ l, 32 }, *p, *rf, rs, pus[i], (Pix *)y, (Pix *)u, (Pix *)v);
}

extern "C" int kvz_cuda_me_predict_batch(const kvz_cuda_me_params *p, const kvz_cuda_me_mc_refs *refs, const kvz_cuda_me_mc_pu *pus, int count,
                                         void *pred_y, void *pred_u, void *pred_v, void *)
{
  if (!p || !refs || kvzme::params_supported(*p) != 0) return -2;
  if (p->bitdepth == 8) run_predict<uint8_t>(p, refs, pus, count, pred_y, pred_u, pred_v);
  else run_predict<uint16_t>(p, refs, pus, count, pred_y, pred_u, pred_v);
  return 0;
}
