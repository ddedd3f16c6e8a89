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
