// strategies.cu -- layer 2 (per-call strategy functions with the reference's exact typedefs) and layer 3
// (registrars) of include/kvz_cuda.h.
//
// A per-call function stages its host operands through the calling thread's pinned buffer, launches the batched
// kernel with count == 1 on the thread's stream and waits for the result: the reference's function pointers are
// synchronous, context-free and called concurrently from all worker threads (ref: strategyselector.h:99,
// threadqueue.c:275).  They exist to honour the dispatch contract and to run the reference's own unit tests
// against the CUDA kernels; throughput comes from the batched layer.
// Kernels have no error channel: a CUDA failure aborts loudly instead of returning garbage.
#include <dlfcn.h>
#include <stdlib.h>

#include "common.cuh"

using namespace kvzc;

static void die(const char *what)
{
  fprintf(stderr, "libkvzcuda: %s failed: %s\n", what, kvz_cuda_last_error());
  abort();
}
#define MUST(expr) do { if ((expr) != 0) die(#expr); } while (0)

namespace {

template <class T> struct Bits { static constexpr int v = sizeof(T) == 1 ? 8 : 10; };

// ---- picture ------------------------------------------------------------------------------------
template <class T, int N, bool SATD> unsigned strat_cost_nxn(const T *a, const T *b)
{
  Call c(2 * N * N * sizeof(T) + 64);
  if (!c.ok) die("staging");
  const T *da = c.in(a, N * N), *db = c.in(b, N * N);
  uint32_t *out = c.out<uint32_t>(1);
  MUST(c.upload());
  MUST((SATD ? kvz_cuda_satd_nxn_batch : kvz_cuda_sad_nxn_batch)(N, Bits<T>::v, da, db, 1, out, c.s.stream));
  MUST(c.download());
  return *c.host_ptr(out);
}

// cost_pixel_nxn_multi_func (ref: strategies-picture.h:124): preds = kvz_pixel[2][32*32]
template <class T, int N, bool SATD> void strat_cost_nxn_dual(const T (*preds)[32 * 32], const T *orig, unsigned num_modes, unsigned *costs)
{
  (void)num_modes;
  constexpr int SPAN = 1024 + N * N;          // mode 1 starts 1024 pixels in, runs N*N (overlapping for N == 64)
  Call c((SPAN + N * N) * sizeof(T) + 64);
  if (!c.ok) die("staging");
  const T *dp = c.in(&preds[0][0], SPAN), *dorig = c.in(orig, N * N);
  uint32_t *out = c.out<uint32_t>(2);
  MUST(c.upload());
  MUST(kvz_cuda_cost_nxn_multi_batch(SATD, N, Bits<T>::v, dp, 0, 1024, 2, dorig, 1, out, c.s.stream));
  MUST(c.download());
  const uint32_t *h = c.host_ptr(out);
  costs[0] = h[0]; costs[1] = h[1];
}

template <class T> unsigned strat_block_cost(int op, const T *a, const T *b, int w, int h, unsigned s1, unsigned s2,
                                             int b_rows, int left, int right)
{
  Call c(((size_t)w * h + (size_t)w * b_rows) * sizeof(T) + 256);
  if (!c.ok) die("staging");
  const T *da = c.in2d(a, w, h, s1), *db = c.in2d(b, w, b_rows, s2);
  kvz_cuda_blk d = { 0, 0, (int16_t)w, (int16_t)h, (int16_t)left, (int16_t)right };
  const kvz_cuda_blk *dd = c.in(&d, 1);
  uint32_t *out = c.out<uint32_t>(1);
  MUST(c.upload());
  MUST(kvz_cuda_block_cost_batch(op, Bits<T>::v, da, w, db, w, dd, 1, out, c.s.stream));
  MUST(c.download());
  return *c.host_ptr(out);
}
template <class T> unsigned strat_reg_sad(const T *a, const T *b, int w, int h, unsigned s1, unsigned s2)
{ return strat_block_cost<T>(KVZ_CUDA_OP_REG_SAD, a, b, w, h, s1, s2, h, 0, 0); }
template <class T> unsigned strat_satd_any_size(int w, int h, const T *a, int s1, const T *b, int s2)
{ return strat_block_cost<T>(KVZ_CUDA_OP_SATD_ANY, a, b, w, h, s1, s2, h, 0, 0); }
template <class T> unsigned strat_ssd(const T *ref, const T *rec, int rs, int cs, int width)
{ return strat_block_cost<T>(KVZ_CUDA_OP_SSD, ref, rec, width, width, rs, cs, width, 0, 0); }
template <class T> uint32_t strat_ver_sad(const T *pic, const T *ref, int32_t w, int32_t h, uint32_t ps)
{ return strat_block_cost<T>(KVZ_CUDA_OP_VER_SAD, pic, ref, w, h, ps, w, 1, 0, 0); }
template <class T> uint32_t strat_hor_sad(const T *pic, const T *ref, int32_t w, int32_t h, uint32_t ps, uint32_t rs,
                                          uint32_t left, uint32_t right)
{ return strat_block_cost<T>(KVZ_CUDA_OP_HOR_SAD, pic, ref, w, h, ps, rs, h, (int)left, (int)right); }

template <class T> void strat_satd_quad(int w, int h, const T **preds, int stride, const T *orig, int orig_stride,
                                        unsigned num_modes, unsigned *costs, int8_t *valid)
{
  (void)num_modes; (void)valid;
  Call c((size_t)5 * w * h * sizeof(T) + 2048);
  if (!c.ok) die("staging");
  const T *dp[4];
  for (int k = 0; k < 4; ++k) dp[k] = c.in2d(preds[k], w, h, stride);
  const T *dorig = c.in2d(orig, w, h, orig_stride);
  kvz_cuda_quad q;
  for (int k = 0; k < 4; ++k) q.off_pred[k] = (int32_t)(dp[k] - dp[0]);
  q.off_orig = 0; q.w = (int16_t)w; q.h = (int16_t)h;
  const kvz_cuda_quad *dq = c.in(&q, 1);
  uint32_t *out = c.out<uint32_t>(4);
  MUST(c.upload());
  MUST(kvz_cuda_satd_any_size_quad_batch(Bits<T>::v, dp[0], w, dorig, w, dq, 1, out, c.s.stream));
  MUST(c.download());
  const uint32_t *hres = c.host_ptr(out);
  for (int k = 0; k < 4; ++k) costs[k] = hres[k];
}

template <class T> double strat_pixel_var(const T *buf, uint32_t len)
{
  Call c((size_t)len * sizeof(T) + 64);
  if (!c.ok) die("staging");
  const T *d = c.in(buf, len);
  double *out = c.out<double>(1);
  MUST(c.upload());
  MUST(kvz_cuda_pixel_var_batch(Bits<T>::v, d, len, 1, out, c.s.stream));
  MUST(c.download());
  return *c.host_ptr(out);
}

// get_optimized_sad: like the generic strategy, no width-specialised variants -> callers use reg_sad
// (ref: picture-generic.c:671-674, image.c:230-237)
typedef uint32_t (*optimized_sad_ptr)(const void *, const void *, int32_t, uint32_t, uint32_t);
optimized_sad_ptr strat_get_optimized_sad(int32_t) { return nullptr; }

// ---- dct ----------------------------------------------------------------------------------------
template <int KIND, int N> void strat_transform(int8_t bitdepth, const int16_t *in, int16_t *out)
{
  Call c(2 * N * N * sizeof(int16_t) + 64);
  if (!c.ok) die("staging");
  const int16_t *din = c.in(in, N * N);
  int16_t *dout = c.out<int16_t>(N * N);
  MUST(c.upload());
  MUST(kvz_cuda_transform_batch(KIND, N, bitdepth, din, dout, 1, c.s.stream));
  MUST(c.download());
  memcpy(out, c.host_ptr(dout), N * N * sizeof(int16_t));
}

// ---- quant (plain-typed members) ------------------------------------------------------------------
uint32_t strat_coeff_abs_sum(const int16_t *coeffs, size_t length)
{
  Call c(length * sizeof(int16_t) + 64);
  if (!c.ok) die("staging");
  const int16_t *d = c.in(coeffs, length);
  uint32_t *out = c.out<uint32_t>(1);
  MUST(c.upload());
  MUST(kvz_cuda_coeff_abs_sum_batch(d, length, 1, out, c.s.stream));
  MUST(c.download());
  return *c.host_ptr(out);
}
double strat_fast_coeff_cost(const int16_t *coeff, int32_t width, uint64_t weights)
{
  Call c((size_t)width * width * sizeof(int16_t) + 64);
  if (!c.ok) die("staging");
  const int16_t *d = c.in(coeff, (size_t)width * width);
  uint32_t *out = c.out<uint32_t>(1);
  MUST(c.upload());
  MUST(kvz_cuda_fast_coeff_cost_batch(d, width, weights, 1, out, c.s.stream));
  MUST(c.download());
  return (double)*c.host_ptr(out) / 256.0;       // ref: quant-generic.c:374
}

// ---- intra --------------------------------------------------------------------------------------
template <class T> void strat_intra(int log2w, int mode, const T *top, const T *left, T *dst)
{
  const int n = 2 * (1 << log2w) + 1, ww = 1 << (2 * log2w);
  Call c((2 * n + ww) * sizeof(T) + 1024);
  if (!c.ok) die("staging");
  const T *dt = c.in(top, n), *dl = c.in(left, n);
  const int8_t m = (int8_t)mode;
  const int8_t *dm = c.in(&m, 1);
  T *dd = c.out<T>(ww);
  MUST(c.upload());
  MUST(kvz_cuda_intra_predict_batch(0, log2w, 0, 0, Bits<T>::v, dt, dl, dm, 1, dd, c.s.stream));
  MUST(c.download());
  memcpy(dst, c.host_ptr(dd), ww * sizeof(T));
}
// int_fast8_t is signed char on the reference's targets (glibc x86-64 / aarch64)
template <class T> void strat_angular(const int_fast8_t log2w, const int_fast8_t mode, const T *top, const T *left, T *dst)
{ strat_intra<T>(log2w, mode, top, left, dst); }
template <class T> void strat_planar(const int_fast8_t log2w, const T *top, const T *left, T *dst)
{ strat_intra<T>(log2w, 0, top, left, dst); }
template <class T> void strat_filtered_dc(const int_fast8_t log2w, const T *top, const T *left, T *dst)
{ strat_intra<T>(log2w, 1, top, left, dst); }

// ---- nal ----------------------------------------------------------------------------------------
template <class T> void strat_checksum(const T *data, const int height, const int width, const int stride,
                                       unsigned char checksum_out[16], const uint8_t bitdepth)
{
  (void)bitdepth;
  Call c((size_t)width * height * sizeof(T) + 64);
  if (!c.ok) die("staging");
  const T *d = c.in2d(data, width, height, stride);
  uint8_t *out = c.out<uint8_t>(4);
  MUST(c.upload());
  MUST(kvz_cuda_array_checksum(Bits<T>::v, d, height, width, width, out, c.s.stream));
  MUST(c.download());
  memcpy(checksum_out, c.host_ptr(out), 4);
}

struct Entry { const char *type; void *f8; void *f16; const char *group; };

#define NXN(kind, SATD, n) { #kind "_" #n "x" #n, (void *)&strat_cost_nxn<uint8_t, n, SATD>, (void *)&strat_cost_nxn<uint16_t, n, SATD>, "picture" }
#define DUAL(kind, SATD, n) { #kind "_" #n "x" #n "_dual", (void *)&strat_cost_nxn_dual<uint8_t, n, SATD>, (void *)&strat_cost_nxn_dual<uint16_t, n, SATD>, "picture" }
#define TR(name, KIND, n) { name, (void *)&strat_transform<KIND, n>, (void *)&strat_transform<KIND, n>, "dct" }

const Entry g_entries[] = {
  { "reg_sad", (void *)&strat_reg_sad<uint8_t>, (void *)&strat_reg_sad<uint16_t>, "picture" },
  NXN(sad, false, 4), NXN(sad, false, 8), NXN(sad, false, 16), NXN(sad, false, 32), NXN(sad, false, 64),
  NXN(satd, true, 4), NXN(satd, true, 8), NXN(satd, true, 16), NXN(satd, true, 32), NXN(satd, true, 64),
  { "satd_any_size", (void *)&strat_satd_any_size<uint8_t>, (void *)&strat_satd_any_size<uint16_t>, "picture" },
  DUAL(sad, false, 4), DUAL(sad, false, 8), DUAL(sad, false, 16), DUAL(sad, false, 32), DUAL(sad, false, 64),
  DUAL(satd, true, 4), DUAL(satd, true, 8), DUAL(satd, true, 16), DUAL(satd, true, 32), DUAL(satd, true, 64),
  { "satd_any_size_quad", (void *)&strat_satd_quad<uint8_t>, (void *)&strat_satd_quad<uint16_t>, "picture" },
  { "pixels_calc_ssd", (void *)&strat_ssd<uint8_t>, (void *)&strat_ssd<uint16_t>, "picture" },
  { "get_optimized_sad", (void *)&strat_get_optimized_sad, (void *)&strat_get_optimized_sad, "picture" },
  { "ver_sad", (void *)&strat_ver_sad<uint8_t>, (void *)&strat_ver_sad<uint16_t>, "picture" },
  { "hor_sad", (void *)&strat_hor_sad<uint8_t>, (void *)&strat_hor_sad<uint16_t>, "picture" },
  { "pixel_var", (void *)&strat_pixel_var<uint8_t>, (void *)&strat_pixel_var<uint16_t>, "picture" },
  TR("fast_forward_dst_4x4", KVZ_CUDA_TR_DST, 4),
  TR("dct_4x4", KVZ_CUDA_TR_DCT, 4), TR("dct_8x8", KVZ_CUDA_TR_DCT, 8), TR("dct_16x16", KVZ_CUDA_TR_DCT, 16), TR("dct_32x32", KVZ_CUDA_TR_DCT, 32),
  TR("fast_inverse_dst_4x4", KVZ_CUDA_TR_IDST, 4),
  TR("idct_4x4", KVZ_CUDA_TR_IDCT, 4), TR("idct_8x8", KVZ_CUDA_TR_IDCT, 8), TR("idct_16x16", KVZ_CUDA_TR_IDCT, 16), TR("idct_32x32", KVZ_CUDA_TR_IDCT, 32),
  { "coeff_abs_sum", (void *)&strat_coeff_abs_sum, (void *)&strat_coeff_abs_sum, "quant" },
  { "fast_coeff_cost", (void *)&strat_fast_coeff_cost, (void *)&strat_fast_coeff_cost, "quant" },
  { "angular_pred", (void *)&strat_angular<uint8_t>, (void *)&strat_angular<uint16_t>, "intra" },
  { "intra_pred_planar", (void *)&strat_planar<uint8_t>, (void *)&strat_planar<uint16_t>, "intra" },
  { "intra_pred_filtered_dc", (void *)&strat_filtered_dc<uint8_t>, (void *)&strat_filtered_dc<uint16_t>, "intra" },
  { "array_checksum", (void *)&strat_checksum<uint8_t>, (void *)&strat_checksum<uint16_t>, "nal" },
};

kvz_cuda_register_fn g_register = nullptr;

int register_group(void *opaque, uint8_t bitdepth, const char *group)
{
  // A missing/unsupported device must degrade to "don't register" (SURVEY.md 8b): report success with
  // nothing registered so that encoder_open still works on the host's own strategies.
  if (!kvz_cuda_available()) return 1;
  kvz_cuda_register_fn reg = g_register;
  if (!reg) reg = (kvz_cuda_register_fn)dlsym(RTLD_DEFAULT, "kvz_strategyselector_register");
  if (!reg) { set_error("kvz_strategyselector_register not found: call kvz_cuda_set_register_fn first"); return 0; }
  int ok = 1;
  for (const Entry &e : g_entries)
    if (strcmp(e.group, group) == 0)
      ok &= reg(opaque, e.type, "cuda", KVZ_CUDA_PRIORITY, bitdepth == 8 ? e.f8 : e.f16);
  return ok;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// Plain-parameter per-call entries for the strategy types whose reference typedefs take encoder structs
// (quant, sao, ipol, bipred_average).  integration/strategies-cuda-glue.c unpacks the structs and calls these.
// ---------------------------------------------------------------------------------------------------------
extern "C" {

void kvz_cuda_call_quant(const kvz_cuda_quant_params *p, const int16_t *coef, int16_t *q_coef, int n, int type, int scan_idx)
{
  Call c(2 * (size_t)n * n * sizeof(int16_t) + 512);
  if (!c.ok) die("staging");
  const int16_t *dc = c.in(coef, (size_t)n * n);
  const int8_t sc = (int8_t)scan_idx;
  const int8_t *ds = c.in(&sc, 1);
  int16_t *dq = c.out<int16_t>((size_t)n * n);
  MUST(c.upload());
  MUST(kvz_cuda_quant_batch(p, dc, dq, n, type, ds, 1, c.s.stream));
  MUST(c.download());
  memcpy(q_coef, c.host_ptr(dq), (size_t)n * n * sizeof(int16_t));
}

void kvz_cuda_call_dequant(const kvz_cuda_quant_params *p, const int16_t *q_coef, int16_t *coef, int n, int type)
{
  Call c(2 * (size_t)n * n * sizeof(int16_t) + 512);
  if (!c.ok) die("staging");
  const int16_t *dq = c.in(q_coef, (size_t)n * n);
  int16_t *dc = c.out<int16_t>((size_t)n * n);
  MUST(c.upload());
  MUST(kvz_cuda_dequant_batch(p, dq, dc, n, type, 1, c.s.stream));
  MUST(c.download());
  memcpy(coef, c.host_ptr(dc), (size_t)n * n * sizeof(int16_t));
}

// kvz_quantize_residual with plain parameters.  phase as in kvz_cuda_tu.  Returns has_coeffs.
int kvz_cuda_call_quantize_residual(const kvz_cuda_quant_params *p, int width, int color, int scan_idx, int use_trskip,
                                    int cu_is_intra, int early_skip, int phase, int in_stride, int out_stride,
                                    const void *ref_in, const void *pred_in, void *rec_out, int16_t *coeff_out)
{
  const size_t px = p->bitdepth == 8 ? 1 : 2;
  const int n = width;
  Call c(3 * (size_t)n * n * px + (size_t)n * n * 2 + 1024);
  if (!c.ok) die("staging");
  // inputs compacted to stride n; the coefficient buffer is an input for phase 2 and an output otherwise
  const uint8_t *dref = c.in2d((const uint8_t *)ref_in, (int)(n * px), n, (long)in_stride * px);
  const uint8_t *dpred = c.in2d((const uint8_t *)pred_in, (int)(n * px), n, (long)in_stride * px);
  int16_t *dcoef_in = phase == 2 ? c.in(coeff_out, (size_t)n * n) : nullptr;
  kvz_cuda_tu tu; memset(&tu, 0, sizeof(tu));
  tu.width = (uint8_t)n; tu.color = (uint8_t)color; tu.scan_idx = (uint8_t)scan_idx; tu.use_trskip = (uint8_t)use_trskip;
  tu.cu_is_intra = (uint8_t)cu_is_intra; tu.early_skip = (uint8_t)early_skip; tu.phase = (uint8_t)phase;
  const kvz_cuda_tu *dtu = c.in(&tu, 1);
  uint8_t *drec = c.out<uint8_t>((size_t)n * n * px);
  int16_t *dcoef = phase == 2 ? dcoef_in : c.out<int16_t>((size_t)n * n);
  int32_t *dhas = c.out<int32_t>(1);
  MUST(c.upload());
  MUST(kvz_cuda_quantize_residual_batch(p, dref, dpred, n, drec, n, dcoef, dtu, 1, dhas, c.s.stream));
  MUST(c.download());
  const int has = *c.host_ptr(dhas);
  if (phase != 2) memcpy(coeff_out, c.host_ptr(dcoef), (size_t)n * n * sizeof(int16_t));
  if (phase != 1) {
    const uint8_t *hrec = c.host_ptr(drec);
    for (int y = 0; y < n; ++y) memcpy((uint8_t *)rec_out + (size_t)y * out_stride * px, hrec + (size_t)y * n * px, n * px);
  }
  return has;
}

// kvz_quantize_residual with RDOQ, whole function on the device (one upload, three launches, one download)
int kvz_cuda_call_quantize_residual_rdoq(const kvz_cuda_quant_params *p, const kvz_cuda_rdoq_params *rp, const kvz_cuda_cabac_ctx *cabac,
                                         int width, int color, int scan_idx, int use_trskip, int cu_is_intra, int early_skip, int tr_depth,
                                         int in_stride, int out_stride, const void *ref_in, const void *pred_in, void *rec_out, int16_t *coeff_out)
{
  const size_t px = p->bitdepth == 8 ? 1 : 2;
  const int n = width;
  Call c(3 * (size_t)n * n * px + (size_t)n * n * 2 + 2048);
  if (!c.ok) die("staging");
  const uint8_t *dref = c.in2d((const uint8_t *)ref_in, (int)(n * px), n, (long)in_stride * px);
  const uint8_t *dpred = c.in2d((const uint8_t *)pred_in, (int)(n * px), n, (long)in_stride * px);
  kvz_cuda_tu tu; memset(&tu, 0, sizeof(tu));
  tu.width = (uint8_t)n; tu.color = (uint8_t)color; tu.scan_idx = (uint8_t)scan_idx; tu.use_trskip = (uint8_t)use_trskip;
  tu.cu_is_intra = (uint8_t)cu_is_intra; tu.early_skip = (uint8_t)early_skip; tu.tr_depth = (uint8_t)tr_depth;
  const kvz_cuda_tu *dtu = c.in(&tu, 1);
  const kvz_cuda_cabac_ctx *dctx = c.in(cabac, 1);
  uint8_t *drec = c.out<uint8_t>((size_t)n * n * px);
  int16_t *dcoef = c.out<int16_t>((size_t)n * n);
  int32_t *dhas = c.out<int32_t>(1);
  MUST(c.upload());
  MUST(kvz_cuda_quantize_residual_rdoq_batch(p, rp, dctx, dref, dpred, n, drec, n, dcoef, dtu, 1, n, dhas, c.s.stream));
  MUST(c.download());
  memcpy(coeff_out, c.host_ptr(dcoef), (size_t)n * n * sizeof(int16_t));
  const uint8_t *hrec = c.host_ptr(drec);
  for (int y = 0; y < n; ++y) memcpy((uint8_t *)rec_out + (size_t)y * out_stride * px, hrec + (size_t)y * n * px, n * px);
  return *c.host_ptr(dhas);
}

void kvz_cuda_call_sao_edge_stats(int bitdepth, const void *orig, const void *rec, int eo_class, int bw, int bh, int *cat_sum_cnt)
{
  const size_t px = bitdepth == 8 ? 1 : 2;
  Call c(2 * (size_t)bw * bh * px + 1024);
  if (!c.ok) die("staging");
  const uint8_t *dorig = c.in((const uint8_t *)orig, (size_t)bw * bh * px), *drec = c.in((const uint8_t *)rec, (size_t)bw * bh * px);
  kvz_cuda_sao_blk b = { 0, 0, (int16_t)bw, (int16_t)bh, 0, 0 };
  const kvz_cuda_sao_blk *db = c.in(&b, 1);
  int32_t *out = c.out<int32_t>(40);
  MUST(c.upload());
  MUST(kvz_cuda_sao_edge_stats_batch(bitdepth, dorig, drec, db, 1, out, c.s.stream));
  MUST(c.download());
  const int32_t *h = c.host_ptr(out) + eo_class * 10;
  for (int k = 0; k < 10; ++k) cat_sum_cnt[k] += h[k];          // the reference accumulates (sao-generic.c:76-77)
}

int kvz_cuda_call_sao_edge_ddistortion(int bitdepth, const void *orig, const void *rec, int bw, int bh, int eo_class, const int *offsets)
{
  const size_t px = bitdepth == 8 ? 1 : 2;
  Call c(2 * (size_t)bw * bh * px + 1024);
  if (!c.ok) die("staging");
  const uint8_t *dorig = c.in((const uint8_t *)orig, (size_t)bw * bh * px), *drec = c.in((const uint8_t *)rec, (size_t)bw * bh * px);
  kvz_cuda_sao_blk b = { 0, 0, (int16_t)bw, (int16_t)bh, 0, 0 };
  const kvz_cuda_sao_blk *db = c.in(&b, 1);
  const int8_t eo = (int8_t)eo_class;
  const int8_t *deo = c.in(&eo, 1);
  const int32_t *doff = c.in((const int32_t *)offsets, 5);
  int32_t *out = c.out<int32_t>(1);
  MUST(c.upload());
  MUST(kvz_cuda_sao_edge_ddistortion_batch(bitdepth, dorig, drec, db, deo, doff, 1, out, c.s.stream));
  MUST(c.download());
  return *c.host_ptr(out);
}

int kvz_cuda_call_sao_band_ddistortion(int bitdepth, const void *orig, const void *rec, int bw, int bh, int band_pos, const int *bands)
{
  const size_t px = bitdepth == 8 ? 1 : 2;
  Call c(2 * (size_t)bw * bh * px + 1024);
  if (!c.ok) die("staging");
  const uint8_t *dorig = c.in((const uint8_t *)orig, (size_t)bw * bh * px), *drec = c.in((const uint8_t *)rec, (size_t)bw * bh * px);
  kvz_cuda_sao_blk b = { 0, 0, (int16_t)bw, (int16_t)bh, 0, 0 };
  const kvz_cuda_sao_blk *db = c.in(&b, 1);
  const int32_t bp = band_pos;
  const int32_t *dbp = c.in(&bp, 1), *dbands = c.in((const int32_t *)bands, 4);
  int32_t *out = c.out<int32_t>(1);
  MUST(c.upload());
  MUST(kvz_cuda_sao_band_ddistortion_batch(bitdepth, dorig, drec, db, dbp, dbands, 1, out, c.s.stream));
  MUST(c.download());
  return *c.host_ptr(out);
}

// sao_reconstruct_color: rec_data points at the block; edge types read one sample around it (the caller
// guarantees that halo exists, exactly as for the reference function).
void kvz_cuda_call_sao_reconstruct(int bitdepth, const void *rec_data, void *new_rec_data, int sao_type, int eo_class,
                                   const int *band_position, const int *offsets, int stride, int new_stride, int bw, int bh, int color)
{
  const size_t px = bitdepth == 8 ? 1 : 2;
  const int halo = sao_type == 2 ? 1 : 0;
  const int ww = bw + 2 * halo, hh = bh + 2 * halo;
  Call c((size_t)ww * hh * px + (size_t)bw * bh * px + 1024);
  if (!c.ok) die("staging");
  const uint8_t *drec = c.in2d((const uint8_t *)rec_data - ((size_t)halo * stride + halo) * px, (int)(ww * px), hh, (long)stride * px);
  kvz_cuda_sao_rec d; memset(&d, 0, sizeof(d));
  d.off_rec = halo * ww + halo; d.off_new = 0; d.bw = (int16_t)bw; d.bh = (int16_t)bh;
  d.type = (int8_t)sao_type; d.eo_class = (int8_t)eo_class; d.color = (int8_t)color;
  d.band_position[0] = band_position[0]; d.band_position[1] = band_position[1];
  for (int k = 0; k < 10; ++k) d.offsets[k] = offsets[k];
  const kvz_cuda_sao_rec *dd = c.in(&d, 1);
  uint8_t *dnew = c.out<uint8_t>((size_t)bw * bh * px);
  MUST(c.upload());
  MUST(kvz_cuda_sao_reconstruct_batch(bitdepth, drec, ww, dnew, bw, dd, 1, c.s.stream));
  MUST(c.download());
  const uint8_t *h = c.host_ptr(dnew);
  for (int y = 0; y < bh; ++y) memcpy((uint8_t *)new_rec_data + (size_t)y * new_stride * px, h + (size_t)y * bw * px, bw * px);
}

// sample_quarterpel_luma(_hi) / sample_octpel_chroma(_hi); kind = KVZ_CUDA_IPOL_*
void kvz_cuda_call_sample(int kind, int bitdepth, const void *src, int src_stride, int w, int h, void *dst, int dst_stride, int mvx, int mvy)
{
  const size_t px = bitdepth == 8 ? 1 : 2;
  const int taps = kind >= KVZ_CUDA_IPOL_CHROMA ? 4 : 8, off = taps / 2 - 1;
  const int ww = w + taps - 1, hh = h + taps - 1;
  const size_t opx = (kind & 1) ? 2 : px;
  Call c((size_t)ww * hh * px + (size_t)w * h * opx + 1024);
  if (!c.ok) die("staging");
  const uint8_t *dsrc = c.in2d((const uint8_t *)src - ((size_t)off * src_stride + off) * px, (int)(ww * px), hh, (long)src_stride * px);
  kvz_cuda_ipol d = { off * ww + off, 0, (int16_t)w, (int16_t)h, (int16_t)mvx, (int16_t)mvy };
  const kvz_cuda_ipol *dd = c.in(&d, 1);
  uint8_t *ddst = c.out<uint8_t>((size_t)w * h * opx);
  MUST(c.upload());
  MUST(kvz_cuda_sample_batch(kind, bitdepth, dsrc, ww, ddst, w, dd, 1, c.s.stream));
  MUST(c.download());
  const uint8_t *hres = c.host_ptr(ddst);
  for (int y = 0; y < h; ++y) memcpy((uint8_t *)dst + (size_t)y * dst_stride * opx, hres + (size_t)y * w * opx, w * opx);
}

// The four FME stages; filtered [4][64*64] pixels, hor_intermediate [5][KVZ_CUDA_IPOL_IM_SIZE], hor_first_cols [5][KVZ_CUDA_IPOL_FIRST_COLS].
void kvz_cuda_call_filter_fme(int stage, int bitdepth, const void *src, int src_stride, int w, int h, void *filtered,
                              int16_t *hor_intermediate, int fme_level, int16_t *hor_first_cols, int hpel_off_x, int hpel_off_y)
{
  const size_t px = bitdepth == 8 ? 1 : 2;
  // source window: rows -3 .. h+4 (h + 8 rows), columns -3 .. w+4+1
  const int ww = w + 1 + 7 + 1, hh = h + 1 + 7;
  const size_t n_im = (size_t)5 * KVZ_CUDA_IPOL_IM_SIZE, n_col = (size_t)5 * KVZ_CUDA_IPOL_FIRST_COLS, n_f = (size_t)4 * 4096;
  Call c((size_t)ww * hh * px + n_im * 2 + n_col * 2 + n_f * px + 2048);
  if (!c.ok) die("staging");
  // state arrays are both input and output: place them in the input region, read them back from the same place
  int16_t *dim = c.in(hor_intermediate, n_im);
  int16_t *dcol = c.in(hor_first_cols, n_col);
  uint8_t *dflt = c.in((const uint8_t *)filtered, n_f * px);
  const uint8_t *dsrc = c.in2d((const uint8_t *)src - ((size_t)3 * src_stride + 3) * px, (int)(ww * px), hh, (long)src_stride * px);
  const int32_t soff = 3 * ww + 3;
  const int32_t *dsoff = c.in(&soff, 1);
  const int8_t ho[2] = { (int8_t)hpel_off_x, (int8_t)hpel_off_y };
  const int8_t *dho = c.in(ho, 2);
  MUST(c.upload());
  MUST(kvz_cuda_filter_fme_batch(stage, bitdepth, dsrc, ww, dsoff, w, h, dflt, dim, fme_level, dcol, dho, 1, c.s.stream));
  // download the three state arrays (they sit at the start of the staging buffer)
  const size_t span = (size_t)((uint8_t *)dflt + n_f * px - c.s.d);
  if (cudaMemcpyAsync(c.s.h, c.s.d, span, cudaMemcpyDeviceToHost, c.s.stream) != cudaSuccess || cudaStreamSynchronize(c.s.stream) != cudaSuccess) die("fme download");
  memcpy(hor_intermediate, c.host_ptr(dim), n_im * 2);
  memcpy(hor_first_cols, c.host_ptr(dcol), n_col * 2);
  memcpy(filtered, c.host_ptr(dflt), n_f * px);
}

void kvz_cuda_call_extend_block(int bitdepth, const void *src, int src_w, int src_h, int src_s, int blk_x, int blk_y, int blk_w,
                                int blk_h, int pad_l, int pad_r, int pad_t, int pad_b, int pad_b_simd, void *buf)
{
  const size_t px = bitdepth == 8 ? 1 : 2;
  // only the rows/columns the block can touch are staged: the clipped source window
  const int x0 = blk_x - pad_l < 0 ? 0 : (blk_x - pad_l > src_w - 1 ? src_w - 1 : blk_x - pad_l);
  const int x1 = blk_x + blk_w + pad_r - 1 < 0 ? 0 : (blk_x + blk_w + pad_r - 1 > src_w - 1 ? src_w - 1 : blk_x + blk_w + pad_r - 1);
  const int y0 = blk_y - pad_t < 0 ? 0 : (blk_y - pad_t > src_h - 1 ? src_h - 1 : blk_y - pad_t);
  const int y1 = blk_y + blk_h + pad_b - 1 < 0 ? 0 : (blk_y + blk_h + pad_b - 1 > src_h - 1 ? src_h - 1 : blk_y + blk_h + pad_b - 1);
  const int ww = x1 - x0 + 1, hh = y1 - y0 + 1;
  const size_t total = (size_t)(pad_l + blk_w + pad_r) * (pad_t + blk_h + pad_b + pad_b_simd) + 1;
  Call c((size_t)ww * hh * px + total * px + 1024);
  if (!c.ok) die("staging");
  const uint8_t *dsrc = c.in2d((const uint8_t *)src + ((size_t)y0 * src_s + x0) * px, (int)(ww * px), hh, (long)src_s * px);
  uint8_t *dbuf = c.out<uint8_t>(total * px);
  MUST(c.upload());
  MUST(kvz_cuda_extend_block(bitdepth, dsrc, ww, hh, ww, blk_x - x0, blk_y - y0, blk_w, blk_h, pad_l, pad_r, pad_t, pad_b, pad_b_simd, dbuf, c.s.stream));
  MUST(c.download());
  memcpy(buf, c.host_ptr(dbuf), total * px);
}

void kvz_cuda_call_bipred_plane(int bitdepth, void *dst, int dst_stride, const void *l0, const void *l1, int l0_is_im, int l1_is_im, int w, int h)
{
  const size_t px = bitdepth == 8 ? 1 : 2;
  Call c((size_t)w * h * (2 + 2 + px) + 1024);
  if (!c.ok) die("staging");
  const uint8_t *d0 = c.in((const uint8_t *)l0, (size_t)w * h * (l0_is_im ? 2 : px));
  const uint8_t *d1 = c.in((const uint8_t *)l1, (size_t)w * h * (l1_is_im ? 2 : px));
  uint8_t *dd = c.out<uint8_t>((size_t)w * h * px);
  MUST(c.upload());
  MUST(kvz_cuda_bipred_average_plane(bitdepth, dd, w, d0, d1, l0_is_im, l1_is_im, w, h, c.s.stream));
  MUST(c.download());
  const uint8_t *hres = c.host_ptr(dd);
  for (int y = 0; y < h; ++y) memcpy((uint8_t *)dst + (size_t)y * dst_stride * px, hres + (size_t)y * w * px, w * px);
}

}  // extern "C"

extern "C" {

void kvz_cuda_set_register_fn(kvz_cuda_register_fn fn) { g_register = fn; }

void *kvz_cuda_strategy_fptr(const char *type, uint8_t bitdepth)
{
  for (const Entry &e : g_entries)
    if (strcmp(e.type, type) == 0) return bitdepth == 8 ? e.f8 : e.f16;
  return nullptr;
}

int kvz_strategy_register_picture_cuda(void *opaque, uint8_t bitdepth) { return register_group(opaque, bitdepth, "picture"); }
int kvz_strategy_register_dct_cuda(void *opaque, uint8_t bitdepth) { return register_group(opaque, bitdepth, "dct"); }
int kvz_strategy_register_intra_cuda(void *opaque, uint8_t bitdepth) { return register_group(opaque, bitdepth, "intra"); }
int kvz_strategy_register_nal_cuda(void *opaque, uint8_t bitdepth) { return register_group(opaque, bitdepth, "nal"); }
int kvz_strategy_register_quant_plain_cuda(void *opaque, uint8_t bitdepth) { return register_group(opaque, bitdepth, "quant"); }

}  // extern "C"
