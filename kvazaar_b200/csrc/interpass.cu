// interpass.cu -- frame-level INTER hot-path pass (BASELINE config 4 kernels: SAD search, fractional-ME filters,
// SATD, luma/chroma interpolation, inter residual coding), composed from the batched device API.
//
// For every 16x16 luma PU of the frame that is at least one PU away from the picture border (so that every
// candidate and every filter tap stays inside the reference picture):
//   1. integer full search, +-R samples: reg_sad of the PU against the reference frame; first minimum in raster
//      order of (dy, dx)                                            (kvz_reg_sad, ref: picture-generic.c:98-111)
//   2. fractional search exactly like search_frac (ref: search_inter.c:974-1168) without the MV bit cost:
//      SATD of the integer position, then the four filter stages (hpel hor/ver, hpel diag, qpel hor/ver, qpel diag,
//      ref: ipol-generic.c:213-679) each followed by satd_any_size_quad of its four candidates (ref:
//      picture-generic.c:404-471); strict "<" keeps the earlier candidate on ties
//   3. motion compensation with the final quarter-pel MV: sample_quarterpel_luma 16x16, sample_octpel_chroma 8x8
//      for U and V                                                  (ref: ipol-generic.c:134-211, 681-758)
//   4. kvz_quantize_residual (inter CU, RDOQ-off branch) of the luma 16x16 and the two chroma 8x8 TUs + SSD
// The decisions between the stages are three tiny integer kernels here; everything else is the batched API.
#include <vector>

#include "common.cuh"

namespace kvzc {

constexpr int PU = 16;

struct IpGeom { int W, H, ax, ay, n, R; };

int launch_recon_inter(const kvz_cuda_quant_params &qp, const uint8_t *src, const uint8_t *pred, int stride, int color, int log2w,
                       int blocks_x, int nblk, uint8_t *rec, int16_t *coeff, int32_t *has, uint32_t *ssd, cudaStream_t st);   // framepass.cu       // ax x ay active PUs starting at PU (1,1)

__device__ __forceinline__ void pu_xy(const IpGeom &g, int i, int &x, int &y) { x = (1 + i % g.ax) * PU; y = (1 + i / g.ax) * PU; }

// 1. integer full search: one CTA per PU, one thread per candidate (ceil((2R+1)^2 / 32) warps).  The PU and the
//    search window live in shared memory as 32-bit words; a candidate row is 16 bytes at an arbitrary byte offset:
//    five aligned words funnel-shifted into four, then __vabsdiffu4 + __dp4a (4 absolute differences per pair).
__global__ void __launch_bounds__(320) me_full_search_kernel(IpGeom g, const uint8_t *__restrict__ cur, const uint8_t *__restrict__ ref,
                                                             int16_t *__restrict__ mv_int, uint32_t *__restrict__ sad_int)
{
  constexpr int WSP = 40;                                         // padded window row: (16 + 2*8) bytes + slack, 10 words
  __shared__ __align__(16) uint32_t s_cur[PU * PU / 4];
  __shared__ __align__(16) uint32_t s_ref[(PU + 16) * WSP / 4];
  __shared__ unsigned long long s_best;
  const int R = g.R, D = 2 * R + 1, WS = PU + 2 * R;
  int x0, y0;
  pu_xy(g, blockIdx.x, x0, y0);
  uint8_t *cb = reinterpret_cast<uint8_t *>(s_cur), *rb = reinterpret_cast<uint8_t *>(s_ref);
  for (int i = threadIdx.x; i < PU * PU; i += blockDim.x) cb[i] = cur[(long)(y0 + i / PU) * g.W + x0 + i % PU];
  for (int i = threadIdx.x; i < WS * WSP; i += blockDim.x) {
    const int yy = i / WSP, xx = i - yy * WSP;
    rb[i] = xx < WS ? ref[(long)(y0 - R + yy) * g.W + x0 - R + xx] : 0;
  }
  if (threadIdx.x == 0) s_best = ~0ull;
  __syncthreads();
  for (int c = threadIdx.x; c < D * D; c += blockDim.x) {
    const int dy = c / D, dx = c - dy * D;
    const uint32_t sh = (uint32_t)(dx & 3) * 8;
    uint32_t sad = 0;
#pragma unroll 4
    for (int y = 0; y < PU; ++y) {
      const uint32_t *r = s_ref + (dy + y) * (WSP / 4) + (dx >> 2);
      const uint32_t w0 = r[0], w1 = r[1], w2 = r[2], w3 = r[3], w4 = r[4];
      const uint4 cw = *reinterpret_cast<const uint4 *>(s_cur + y * 4);
      sad = __dp4a(__vabsdiffu4(cw.x, __funnelshift_r(w0, w1, sh)), 0x01010101u, sad);
      sad = __dp4a(__vabsdiffu4(cw.y, __funnelshift_r(w1, w2, sh)), 0x01010101u, sad);
      sad = __dp4a(__vabsdiffu4(cw.z, __funnelshift_r(w2, w3, sh)), 0x01010101u, sad);
      sad = __dp4a(__vabsdiffu4(cw.w, __funnelshift_r(w3, w4, sh)), 0x01010101u, sad);
    }
    atomicMin(&s_best, ((unsigned long long)sad << 32) | (unsigned)c);     // first minimum in raster order
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned c = (unsigned)(s_best & 0xffffffffu);
    mv_int[2 * blockIdx.x] = (int16_t)((int)(c % D) - R);
    mv_int[2 * blockIdx.x + 1] = (int16_t)((int)(c / D) - R);
    sad_int[blockIdx.x] = (uint32_t)(s_best >> 32);
  }
}

struct IpState {        // per-PU running state of the fractional search
  uint32_t cost; int8_t best_index, off_x, off_y, pad; int16_t mvx, mvy;
};

// 2a. after the integer search: FME source offsets, SATD descriptors of the integer position
__global__ void ip_prepare_kernel(IpGeom g, const int16_t *__restrict__ mv_int, int32_t *__restrict__ src_off,
                                  kvz_cuda_blk *__restrict__ int_desc, int8_t *__restrict__ hpel_off)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  int x, y;
  pu_xy(g, i, x, y);
  const int mx = mv_int[2 * i], my = mv_int[2 * i + 1];
  src_off[i] = (y + my - 1) * g.W + (x + mx - 1);            // ext_origin: one sample up-left of the integer position
  int_desc[i].off_a = y * g.W + x;
  int_desc[i].off_b = (y + my) * g.W + x + mx;
  int_desc[i].w = PU; int_desc[i].h = PU; int_desc[i].left = 0; int_desc[i].right = 0;
  hpel_off[2 * i] = 0; hpel_off[2 * i + 1] = 0;
}

// 2b. decision after each filter stage: search_frac's bookkeeping (ref: search_inter.c:1133-1163)
__global__ void ip_decide_kernel(IpGeom g, int step, const int16_t *__restrict__ mv_int, const uint32_t *__restrict__ int_cost,
                                 const uint32_t *__restrict__ quad_cost, IpState *__restrict__ st, int8_t *__restrict__ hpel_off,
                                 int16_t *__restrict__ mv_final, uint32_t *__restrict__ satd_best,
                                 kvz_cuda_ipol *__restrict__ desc_y, kvz_cuda_ipol *__restrict__ desc_c)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  const int sqx[9] = { 0, -1, 1, 0, 0, -1, 1, -1, 1 }, sqy[9] = { 0, 0, 0, -1, 1, -1, -1, 1, 1 };
  IpState s = st[i];
  if (step == 0) {
    s.cost = int_cost[i]; s.best_index = 0; s.off_x = 0; s.off_y = 0;
    s.mvx = (int16_t)(mv_int[2 * i] * 2); s.mvy = (int16_t)(mv_int[2 * i + 1] * 2);       // half-pel units
  }
  const int base = (step & 1) ? 5 : 1;
  for (int j = 0; j < 4; ++j) {
    const uint32_t c = quad_cost[4 * i + j];
    if (c < s.cost) { s.cost = c; s.best_index = (int8_t)(base + j); }
  }
  if (step == 1 || step == 3) {
    s.mvx = (int16_t)(s.mvx + sqx[s.best_index]); s.mvy = (int16_t)(s.mvy + sqy[s.best_index]);
    if (step == 1) {
      s.mvx = (int16_t)(s.mvx * 2); s.mvy = (int16_t)(s.mvy * 2);                         // quarter-pel units
      s.off_x = (int8_t)sqx[s.best_index]; s.off_y = (int8_t)sqy[s.best_index];
      hpel_off[2 * i] = s.off_x; hpel_off[2 * i + 1] = s.off_y;
      s.best_index = 0;
    }
  }
  st[i] = s;
  if (step == 3) {
    int x, y;
    pu_xy(g, i, x, y);
    mv_final[2 * i] = s.mvx; mv_final[2 * i + 1] = s.mvy;
    satd_best[i] = s.cost;
    // motion-compensation descriptors (ref: inter.c:98-160 luma, :218-309 chroma: integer part mv >> 2 / >> 3)
    kvz_cuda_ipol dy;
    dy.off_src = (y + (s.mvy >> 2)) * g.W + x + (s.mvx >> 2); dy.off_dst = y * g.W + x;
    dy.w = PU; dy.h = PU; dy.mvx = s.mvx; dy.mvy = s.mvy;
    desc_y[i] = dy;
    kvz_cuda_ipol dc;
    const int Wc = g.W / 2;
    dc.off_src = (y / 2 + (s.mvy >> 3)) * Wc + x / 2 + (s.mvx >> 3); dc.off_dst = (y / 2) * Wc + x / 2;
    dc.w = PU / 2; dc.h = PU / 2; dc.mvx = s.mvx; dc.mvy = s.mvy;
    desc_c[i] = dc;
  }
}

}  // namespace kvzc

using namespace kvzc;

struct kvz_cuda_inter_pass {
  kvz_cuda_ip_params prm;
  IpGeom g;
  kvz_cuda_ip_layout lay;
  size_t host_bytes, total_bytes;
  uint8_t *blob = nullptr;
  // device-only sections
  size_t o_src_off, o_int_desc, o_hpel_off, o_int_cost, o_quad_cost, o_state, o_quad_desc, o_desc_y, o_desc_c;
  size_t o_tu_y, o_tu_c, o_ssd_y, o_ssd_c, o_pred, o_filtered, o_im, o_cols, o_cur, o_ref;
  size_t init_off = 0, init_bytes = 0;
  std::vector<uint8_t> host_init;
};

static size_t ip_align(size_t v) { return (v + 255) & ~size_t(255); }

static kvz_cuda_inter_pass *ip_build(const kvz_cuda_ip_params *p, bool alloc)
{
  if (alloc && g_device < 0 && kvz_cuda_init(-1) != 0) return nullptr;
  if (!p || p->bitdepth != 8 || p->width % 16 || p->height % 8 || p->width < 48 || p->height < 48 || p->search_range < 1 ||
      p->search_range > 8) {
    set_error("inter pass: need 8-bit, width multiple of 16, height multiple of 8, both >= 48, search_range 1..8");
    return nullptr;
  }
  kvz_cuda_inter_pass *ip = new kvz_cuda_inter_pass();
  ip->prm = *p;
  IpGeom &g = ip->g;
  g.W = p->width; g.H = p->height; g.R = p->search_range;
  g.ax = g.W / PU - 2; g.ay = g.H / PU - 2; g.n = g.ax * g.ay;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = ip_align(off + bytes); return o; };
  kvz_cuda_ip_layout &L = ip->lay;
  memset(&L, 0, sizeof(L));
  const size_t n = (size_t)g.n, fb = (size_t)g.W * g.H * 3 / 2;
  L.npu = g.n; L.pus_x = g.ax; L.pus_y = g.ay;
  L.mv_int = take(4 * n); L.sad_int = take(4 * n); L.mv_final = take(4 * n); L.satd_best = take(4 * n);
  L.has_y = take(4 * n); L.ssd_y = take(4 * n); L.coeff_y = take(2 * n * 256);
  L.has_u = take(4 * n); L.has_v = take(4 * n); L.ssd_u = take(4 * n); L.ssd_v = take(4 * n);
  L.coeff_u = take(2 * n * 64); L.coeff_v = take(2 * n * 64);
  L.rec = take(fb);
  L.host_bytes = ip->host_bytes = off;
  ip->o_src_off = take(4 * n); ip->o_int_desc = take(sizeof(kvz_cuda_blk) * n); ip->o_hpel_off = take(2 * n);
  ip->o_int_cost = take(4 * n); ip->o_quad_cost = take(16 * n); ip->o_state = take(sizeof(IpState) * n);
  ip->o_desc_y = take(sizeof(kvz_cuda_ipol) * n); ip->o_desc_c = take(sizeof(kvz_cuda_ipol) * n);
  ip->o_pred = take(fb);
  ip->o_filtered = take(n * 4 * 4096);
  ip->o_im = take(n * 5 * KVZ_CUDA_IPOL_IM_SIZE * 2);
  ip->o_cols = take(n * 5 * KVZ_CUDA_IPOL_FIRST_COLS * 2);
  ip->o_cur = take(fb); ip->o_ref = take(fb);
  // constant descriptors, initialised from the host once
  ip->init_off = off;
  ip->o_quad_desc = take(sizeof(kvz_cuda_quad) * n);
  ip->o_tu_y = take(sizeof(kvz_cuda_tu) * n); ip->o_tu_c = take(sizeof(kvz_cuda_tu) * n);
  ip->o_ssd_y = take(sizeof(kvz_cuda_blk) * n); ip->o_ssd_c = take(sizeof(kvz_cuda_blk) * n);
  ip->init_bytes = off - ip->init_off;
  ip->total_bytes = off;
  if (!alloc) return ip;
  ip->host_init.assign(ip->init_bytes, 0);
  uint8_t *base = ip->host_init.data() - ip->init_off;
  kvz_cuda_quad *qd = (kvz_cuda_quad *)(base + ip->o_quad_desc);
  kvz_cuda_tu *ty = (kvz_cuda_tu *)(base + ip->o_tu_y), *tc = (kvz_cuda_tu *)(base + ip->o_tu_c);
  kvz_cuda_blk *sy = (kvz_cuda_blk *)(base + ip->o_ssd_y), *sc = (kvz_cuda_blk *)(base + ip->o_ssd_c);
  const int Wc = g.W / 2;
  for (int i = 0; i < g.n; ++i) {
    const int x = (1 + i % g.ax) * PU, y = (1 + i / g.ax) * PU;
    for (int k = 0; k < 4; ++k) qd[i].off_pred[k] = (int32_t)(((size_t)i * 4 + k) * 4096);
    qd[i].off_orig = y * g.W + x; qd[i].w = PU; qd[i].h = PU;
    memset(&ty[i], 0, sizeof(kvz_cuda_tu));
    ty[i].off_ref = ty[i].off_pred = ty[i].off_rec = y * g.W + x; ty[i].off_coeff = i * 256; ty[i].width = 16; ty[i].color = 0;
    memset(&tc[i], 0, sizeof(kvz_cuda_tu));
    tc[i].off_ref = tc[i].off_pred = tc[i].off_rec = (y / 2) * Wc + x / 2; tc[i].off_coeff = i * 64; tc[i].width = 8; tc[i].color = 1;
    sy[i].off_a = sy[i].off_b = y * g.W + x; sy[i].w = 16; sy[i].h = 16; sy[i].left = sy[i].right = 0;
    sc[i].off_a = sc[i].off_b = (y / 2) * Wc + x / 2; sc[i].w = 8; sc[i].h = 8; sc[i].left = sc[i].right = 0;
  }
  if (cudaMalloc((void **)&ip->blob, ip->total_bytes) != cudaSuccess) { set_error("inter pass: cudaMalloc(%zu) failed", ip->total_bytes); delete ip; return nullptr; }
  cudaMemset(ip->blob, 0, ip->total_bytes);
  cudaMemcpy(ip->blob + ip->init_off, ip->host_init.data(), ip->init_bytes, cudaMemcpyHostToDevice);
  return ip;
}

extern "C" {

kvz_cuda_inter_pass *kvz_cuda_ip_create(const kvz_cuda_ip_params *p) { return ip_build(p, true); }
void kvz_cuda_ip_destroy(kvz_cuda_inter_pass *ip) { if (ip) { cudaFree(ip->blob); delete ip; } }
int kvz_cuda_ip_layout_for(const kvz_cuda_ip_params *p, kvz_cuda_ip_layout *out)
{
  KVZC_ARG(out != nullptr);
  kvz_cuda_inter_pass *ip = ip_build(p, false);
  if (!ip) return KVZ_CUDA_E_ARG;
  *out = ip->lay;
  delete ip;
  return 0;
}
void *kvz_cuda_ip_result_dev(kvz_cuda_inter_pass *ip) { return ip ? ip->blob : nullptr; }

int kvz_cuda_ip_run_dev(kvz_cuda_inter_pass *ip, const void *cur_dev, const void *ref_dev, void *stream)
{
  KVZC_REQUIRE_DEVICE();
  KVZC_ARG(ip && cur_dev && ref_dev);
  cudaStream_t st = as_stream(stream);
  const IpGeom g = ip->g;
  if (g.n <= 0) return 0;
  uint8_t *B = ip->blob;
  const kvz_cuda_ip_layout &L = ip->lay;
  const uint8_t *cur = (const uint8_t *)cur_dev, *ref = (const uint8_t *)ref_dev;
  const int W = g.W, H = g.H, n = g.n, Wc = W / 2;
  const size_t poff[3] = { 0, (size_t)W * H, (size_t)W * H * 5 / 4 };
  int16_t *mv_int = (int16_t *)(B + L.mv_int);
  int32_t *src_off = (int32_t *)(B + ip->o_src_off);
  int8_t *hpel_off = (int8_t *)(B + ip->o_hpel_off);
  uint32_t *int_cost = (uint32_t *)(B + ip->o_int_cost), *quad_cost = (uint32_t *)(B + ip->o_quad_cost);
  kvz_cuda_ipol *desc_y = (kvz_cuda_ipol *)(B + ip->o_desc_y), *desc_c = (kvz_cuda_ipol *)(B + ip->o_desc_c);
  uint8_t *filtered = B + ip->o_filtered;
  int16_t *im = (int16_t *)(B + ip->o_im), *cols = (int16_t *)(B + ip->o_cols);
  uint8_t *pred = B + ip->o_pred, *rec = B + L.rec;

  const int nthr = ((2 * g.R + 1) * (2 * g.R + 1) + 31) / 32 * 32;
  me_full_search_kernel<<<n, nthr > 320 ? 320 : nthr, 0, st>>>(g, cur, ref, mv_int, (uint32_t *)(B + L.sad_int));
  KVZC_LAUNCHED();
  ip_prepare_kernel<<<(n + 255) / 256, 256, 0, st>>>(g, mv_int, src_off, (kvz_cuda_blk *)(B + ip->o_int_desc), hpel_off);
  KVZC_LAUNCHED();
  if (int r = kvz_cuda_block_cost_batch(KVZ_CUDA_OP_SATD_ANY, 8, cur, W, ref, W, (const kvz_cuda_blk *)(B + ip->o_int_desc), n, int_cost, st)) return r;
  for (int step = 0; step < 4; ++step) {
    if (int r = kvz_cuda_filter_fme_batch(step, 8, ref, W, src_off, PU, PU, filtered, im, 4, cols, hpel_off, n, st)) return r;
    if (int r = kvz_cuda_satd_any_size_quad_batch(8, filtered, 64, cur, W, (const kvz_cuda_quad *)(B + ip->o_quad_desc), n, quad_cost, st)) return r;
    ip_decide_kernel<<<(n + 255) / 256, 256, 0, st>>>(g, step, mv_int, int_cost, quad_cost, (IpState *)(B + ip->o_state), hpel_off,
                                                      (int16_t *)(B + L.mv_final), (uint32_t *)(B + L.satd_best), desc_y, desc_c);
    KVZC_LAUNCHED();
  }
  // motion compensation into the prediction frame, then residual coding into the reconstruction frame
  if (int r = kvz_cuda_sample_batch(KVZ_CUDA_IPOL_LUMA, 8, ref, W, pred, W, desc_y, n, st)) return r;
  if (int r = kvz_cuda_sample_batch(KVZ_CUDA_IPOL_CHROMA, 8, ref + poff[1], Wc, pred + poff[1], Wc, desc_c, n, st)) return r;
  if (int r = kvz_cuda_sample_batch(KVZ_CUDA_IPOL_CHROMA, 8, ref + poff[2], Wc, pred + poff[2], Wc, desc_c, n, st)) return r;
  // residual coding of the active PU grid (origin PU (1,1)): one grouped transform/quant/recon/SSD launch per plane
  kvz_cuda_quant_params qp = { ip->prm.qp, 8, 0 /* P slice */, 0, 0 };
  const size_t oy = (size_t)PU * W + PU, oc = (size_t)(PU / 2) * Wc + PU / 2;
  if (int r = launch_recon_inter(qp, cur + oy, pred + oy, W, 0, 4, g.ax, n, rec + oy, (int16_t *)(B + L.coeff_y), (int32_t *)(B + L.has_y),
                                 (uint32_t *)(B + L.ssd_y), st)) return r;
  for (int c = 1; c <= 2; ++c)
    if (int r = launch_recon_inter(qp, cur + poff[c] + oc, pred + poff[c] + oc, Wc, c, 3, g.ax, n, rec + poff[c] + oc,
                                   (int16_t *)(B + (c == 1 ? L.coeff_u : L.coeff_v)), (int32_t *)(B + (c == 1 ? L.has_u : L.has_v)),
                                   (uint32_t *)(B + (c == 1 ? L.ssd_u : L.ssd_v)), st)) return r;
  return 0;
}

int kvz_cuda_ip_run_host(kvz_cuda_inter_pass *ip, const void *cur_host, const void *ref_host, void *result_host, void *stream)
{
  KVZC_REQUIRE_DEVICE();
  KVZC_ARG(ip && cur_host && ref_host && result_host);
  cudaStream_t st = as_stream(stream);
  const size_t fb = (size_t)ip->g.W * ip->g.H * 3 / 2;
  KVZC_CHECK(cudaMemcpyAsync(ip->blob + ip->o_cur, cur_host, fb, cudaMemcpyHostToDevice, st));
  KVZC_CHECK(cudaMemcpyAsync(ip->blob + ip->o_ref, ref_host, fb, cudaMemcpyHostToDevice, st));
  if (int r = kvz_cuda_ip_run_dev(ip, ip->blob + ip->o_cur, ip->blob + ip->o_ref, st)) return r;
  KVZC_CHECK(cudaMemcpyAsync(result_host, ip->blob, ip->host_bytes, cudaMemcpyDeviceToHost, st));
  return 0;
}

}  // extern "C"
