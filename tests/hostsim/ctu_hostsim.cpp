// ctu_hostsim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Compiles the single-source CTU search driver (kvazaar_b200/csrc/ctu/*.h) for the host with a "CTA" of one thread
// (see ctu_common.h) behind the same C ABI as the CUDA library (include/kvz_cuda_ctu.h), so that the driver's
// control flow can be checked against the compiled reference (oracle/_ref) on a machine without a GPU
// (tests/test_ctu_hostsim.py).  Only tests/ may load the resulting library; libkvzcuda.so never does.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <mutex>
#include <condition_variable>

#include "../../include/kvz_cuda_ctu.h"
#include "../../kvazaar_b200/csrc/ctu/ctu_frame.h"

using namespace kvzctu;

static_assert(sizeof(kvz_cuda_ctu_config) == sizeof(CtuConfig), "config layout");
static_assert(sizeof(kvz_cuda_ctu_cu) == sizeof(CuRec), "cu layout");
static_assert(sizeof(kvz_cuda_ctu_sao) == sizeof(SaoRec), "sao layout");

struct Slot {
  bool busy = false;
  std::vector<uint8_t> src[3], rec[3], out[3], hor[3], ver[3], dbg[3];
  std::vector<CuRec> cu;
  std::vector<int16_t> coeff;
  std::vector<SaoRec> sao;
  std::vector<CabacState> row_ctx;
  std::vector<uint8_t> dbg_ctx;
  FrameDev F;
};

struct kvz_cuda_ctu_enc {
  CtuConfig cfg;
  CtuTables *T;
  CtuWork *W;
  CtuS *S;
  SaoStats *st;
  std::vector<Slot> slots;
  std::mutex mtx;                // one scratch set: pictures are searched one at a time
  std::condition_variable cv;
};

extern "C" {

int kvz_cuda_ctu_config_supported(const kvz_cuda_ctu_config *c)
{
  if (!c) return -1;
  if (c->width < 8 || c->height < 8 || (c->width & 7) || (c->height & 7)) return -1;
  if (c->rdo < 0 || c->rdo > 3) return -1;
  if (c->pu_depth_intra_min < 1 || c->pu_depth_intra_max > 4 || c->pu_depth_intra_min > c->pu_depth_intra_max) return -1;
  if (c->qp < 0 || c->qp > 51) return -1;
  return 0;
}

kvz_cuda_ctu_enc *kvz_cuda_ctu_open(const kvz_cuda_ctu_config *cfg, int slots)
{
  if (kvz_cuda_ctu_config_supported(cfg)) return NULL;
  kvz_cuda_ctu_enc *e = new kvz_cuda_ctu_enc;
  memcpy(&e->cfg, cfg, sizeof(CtuConfig));
  e->T = new CtuTables;
  ctu_tables_init(e->T);
  e->W = (CtuWork *)calloc(1, sizeof(CtuWork));
  e->S = (CtuS *)calloc(1, sizeof(CtuS));
  e->st = (SaoStats *)calloc(1, sizeof(SaoStats));
  e->slots.resize(slots > 0 ? slots : 1);
  const int W = cfg->width, H = cfg->height, wl = (W + 63) / 64, hl = (H + 63) / 64;
  for (Slot &s : e->slots) {
    for (int p = 0; p < 3; ++p) {
      const int pw = p ? W / 2 : W, ph = p ? H / 2 : H;
      s.src[p].assign((size_t)pw * ph, 0); s.rec[p].assign((size_t)pw * ph, 0); s.out[p].assign((size_t)pw * ph, 0); s.dbg[p].assign((size_t)pw * ph, 0);
      s.hor[p].assign((size_t)pw * hl, 0); s.ver[p].assign((size_t)ph * wl, 0);
    }
    s.cu.assign((size_t)(wl * 16) * (hl * 16), CuRec());
    s.coeff.assign((size_t)wl * hl * 6144, 0);
    s.sao.assign((size_t)wl * hl * 2, SaoRec());
    s.row_ctx.assign(hl, CabacState());
    s.dbg_ctx.assign((size_t)wl * hl * CTX_COUNT, 0);
    FrameDev &F = s.F;
    F.src_y = s.src[0].data(); F.src_u = s.src[1].data(); F.src_v = s.src[2].data();
    F.rec_y = s.rec[0].data(); F.rec_u = s.rec[1].data(); F.rec_v = s.rec[2].data();
    F.out_y = s.out[0].data(); F.out_u = s.out[1].data(); F.out_v = s.out[2].data();
    F.dbg_y = s.dbg[0].data(); F.dbg_u = s.dbg[1].data(); F.dbg_v = s.dbg[2].data();
    F.hor_y = s.hor[0].data(); F.hor_u = s.hor[1].data(); F.hor_v = s.hor[2].data();
    F.ver_y = s.ver[0].data(); F.ver_u = s.ver[1].data(); F.ver_v = s.ver[2].data();
    F.cu = s.cu.data(); F.coeff = s.coeff.data(); F.sao = s.sao.data(); F.row_ctx = s.row_ctx.data();
    F.cu_stride = wl * 16; F.wlcu = wl; F.hlcu = hl;
  }
  return e;
}

void kvz_cuda_ctu_close(kvz_cuda_ctu_enc *e)
{
  if (!e) return;
  delete e->T; free(e->W); free(e->S); free(e->st);
  delete e;
}

int kvz_cuda_ctu_submit(kvz_cuda_ctu_enc *e, const uint8_t *y, const uint8_t *u, const uint8_t *v, int stride_y, int stride_c,
                        const uint8_t *ctx_init, double lambda, double lambda_sqrt, int qp)
{
  // like the CUDA library: blocks while every slot is busy
  std::unique_lock<std::mutex> lock(e->mtx);
  int id = -1;
  e->cv.wait(lock, [&] { for (size_t i = 0; i < e->slots.size(); ++i) if (!e->slots[i].busy) { id = (int)i; return true; } return false; });
  Slot &s = e->slots[id];
  s.busy = true;
  e->cfg.lambda = lambda; e->cfg.lambda_sqrt = lambda_sqrt; e->cfg.qp = qp;
  if (getenv("KVZ_CTU_DEBUG")) fprintf(stderr, "hostsim: qp %d lambda %.17g sqrt %.17g rdo %d pu %d-%d rdoq %d/%d sh %d ts %d sao %d dbk %d\n", qp, lambda, lambda_sqrt, e->cfg.rdo, e->cfg.pu_depth_intra_min, e->cfg.pu_depth_intra_max, e->cfg.rdoq_enable, e->cfg.rdoq_skip, e->cfg.signhide_enable, e->cfg.trskip_enable, e->cfg.sao_type, e->cfg.deblock_enable);
  const int W = e->cfg.width, H = e->cfg.height;
  for (int r = 0; r < H; ++r) memcpy(&s.src[0][(size_t)r * W], y + (size_t)r * stride_y, W);
  for (int r = 0; r < H / 2; ++r) { memcpy(&s.src[1][(size_t)r * (W / 2)], u + (size_t)r * stride_c, W / 2); memcpy(&s.src[2][(size_t)r * (W / 2)], v + (size_t)r * stride_c, W / 2); }
  memset(s.cu.data(), 0, s.cu.size() * sizeof(CuRec));
  for (CabacState &c : s.row_ctx) { memcpy(c.ctx, ctx_init, CTX_COUNT); c.update = 0; }
  Ctx c = { e->T, &e->cfg, e->W, e->S };
  for (int cy = 0; cy < s.F.hlcu; ++cy)
    for (int cx = 0; cx < s.F.wlcu; ++cx) {
      memcpy(&s.dbg_ctx[(size_t)(cy * s.F.wlcu + cx) * CTX_COUNT], s.row_ctx[cy].ctx, CTX_COUNT);
      ctu_job(c, &s.F, e->st, cx, cy);
    }
  for (int cy = 0; cy < s.F.hlcu; ++cy)
    for (int cx = 0; cx < s.F.wlcu; ++cx) ctu_sao_apply(&e->cfg, &s.F, cx, cy);
  return id;
}

int kvz_cuda_ctu_wait(kvz_cuda_ctu_enc *e, int slot, kvz_cuda_ctu_result *out)
{
  if (slot < 0 || slot >= (int)e->slots.size() || !e->slots[slot].busy) return -1;
  Slot &s = e->slots[slot];
  out->cu = (const kvz_cuda_ctu_cu *)s.cu.data();
  out->cu_stride = s.F.cu_stride;
  out->width_in_lcu = s.F.wlcu; out->height_in_lcu = s.F.hlcu;
  out->coeff = s.coeff.data();
  out->sao = (const kvz_cuda_ctu_sao *)s.sao.data();
  out->rec_y = s.out[0].data(); out->rec_u = s.out[1].data(); out->rec_v = s.out[2].data();
  out->dbg_ctx = s.dbg_ctx.data();
  out->dbg_y = s.dbg[0].data(); out->dbg_u = s.dbg[1].data(); out->dbg_v = s.dbg[2].data();
  return 0;
}

// "device" memory of the host build is host memory
int kvz_cuda_ctu_submit_device(kvz_cuda_ctu_enc *e, const uint8_t *y, const uint8_t *u, const uint8_t *v, int stride_y, int stride_c,
                               const uint8_t *ctx_init, double lambda, double lambda_sqrt, int qp)
{
  return kvz_cuda_ctu_submit(e, y, u, v, stride_y, stride_c, ctx_init, lambda, lambda_sqrt, qp);
}
int kvz_cuda_ctu_wait_device(kvz_cuda_ctu_enc *e, int slot, kvz_cuda_ctu_device_result *out)
{
  if (slot < 0 || slot >= (int)e->slots.size() || !e->slots[slot].busy) return -1;
  Slot &s = e->slots[slot];
  memset(out, 0, sizeof(*out));
  out->cu = (const kvz_cuda_ctu_cu *)s.cu.data(); out->coeff = s.coeff.data(); out->sao = (const kvz_cuda_ctu_sao *)s.sao.data();
  out->rec = s.out[0].data();
  out->cu_stride = s.F.cu_stride; out->width_in_lcu = s.F.wlcu; out->height_in_lcu = s.F.hlcu;
  return 0;
}

void kvz_cuda_ctu_release(kvz_cuda_ctu_enc *e, int slot)
{
  {
    std::lock_guard<std::mutex> lock(e->mtx);
    if (slot >= 0 && slot < (int)e->slots.size()) e->slots[slot].busy = false;
  }
  e->cv.notify_all();
}

uint64_t kvz_cuda_ctu_launches(const kvz_cuda_ctu_enc *) { return 0; }

}  // extern "C"
