// rdoq.cu -- batched kvz_rdoq: one warp per TU (see rdoq.cuh).
#include "rdoq.cuh"

#include <mutex>

namespace kvzc {

// one-time fill of g_scan_diag32 (the copy of this translation unit: every 32x32 RDOQ kernel lives here)
static int rdoq_init_tables(cudaStream_t st)
{
  static std::once_flag once;
  static int rc = 0;
  std::call_once(once, [&] {
    rdoq_init_scan32_kernel<<<1, 1024, 0, st>>>();
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) rc = KVZ_CUDA_E_RUNTIME;
  });
  return rc;
}

template <int LOG2N, int WARPS, bool SH>
__global__ void __launch_bounds__(WARPS * 32) rdoq_kernel(kvz_cuda_rdoq_params p, const kvz_cuda_cabac_ctx *__restrict__ cabac,
                                                          const int16_t *__restrict__ coef, int16_t *__restrict__ dest,
                                                          const kvz_cuda_rdoq_tu *__restrict__ tus, int count)
{
  constexpr int NN = 1 << (2 * LOG2N);
  __shared__ RdoqScratch<NN, SH> scratch[WARPS];
  __shared__ kvz_cuda_cabac_ctx s_ctx;
  __shared__ int32_t s_ebits[128];
  rdoq_load_ebits(s_ebits);
  __shared__ __align__(4) int16_t s_q[WARPS][NN];
  for (int i = threadIdx.x; i < (int)sizeof(kvz_cuda_cabac_ctx); i += blockDim.x) ((uint8_t *)&s_ctx)[i] = ((const uint8_t *)cabac)[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * WARPS + warp;
  const bool active = t < count;
  kvz_cuda_rdoq_tu tu = {};
  if (active) tu = tus[t];
  __syncthreads();
  if (!active) return;
  rdoq_tu<NN, SH>(p, &s_ctx, s_ebits, coef + tu.off_coef, s_q[warp], LOG2N, tu.type, tu.scan_idx, tu.block_type, tu.tr_depth, scratch[warp], lane);
  for (int e = lane; e < NN; e += 32) dest[tu.off_dest + e] = s_q[warp][e];
}

// The same for kvz_cuda_tu descriptors (quantize_residual's RDOQ branch): in place on coeff; TUs of other widths
// are skipped (one launch per width present in the batch).
template <int LOG2N, int WARPS, bool SH>
__global__ void __launch_bounds__(WARPS * 32) rdoq_tu_kernel(kvz_cuda_rdoq_params p, const kvz_cuda_cabac_ctx *__restrict__ cabac,
                                                             int16_t *__restrict__ coeff, const kvz_cuda_tu *__restrict__ tus, int count)
{
  constexpr int NN = 1 << (2 * LOG2N);
  __shared__ RdoqScratch<NN, SH> scratch[WARPS];
  __shared__ kvz_cuda_cabac_ctx s_ctx;
  __shared__ int32_t s_ebits[128];
  rdoq_load_ebits(s_ebits);
  __shared__ __align__(4) int16_t s_q[WARPS][NN];
  for (int i = threadIdx.x; i < (int)sizeof(kvz_cuda_cabac_ctx); i += blockDim.x) ((uint8_t *)&s_ctx)[i] = ((const uint8_t *)cabac)[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * WARPS + warp;
  kvz_cuda_tu tu = {};
  bool active = t < count;
  if (active) { tu = tus[t]; active = tu.width == (1 << LOG2N); }
  __syncthreads();
  if (!active) return;
  rdoq_tu<NN, SH>(p, &s_ctx, s_ebits, coeff + tu.off_coeff, s_q[warp], LOG2N, tu.color == 0 ? 0 : 2, tu.scan_idx, tu.cu_is_intra ? 1 : 2, tu.tr_depth, scratch[warp], lane);
  for (int e = lane; e < NN; e += 32) coeff[tu.off_coeff + e] = s_q[warp][e];
}

// Uniform TU grid of the frame-level pass: TU t occupies coeff[t * NN ..), every TU intra; the scan follows the
// intra mode exactly as in the reconstruction kernel (kvz_get_scan_order, search_intra.c / intra.c call sites).
template <int LOG2N, int WARPS, bool SH>
__global__ void __launch_bounds__(WARPS * 32) rdoq_grid_kernel(kvz_cuda_rdoq_params p, const kvz_cuda_cabac_ctx *__restrict__ cabac,
                                                               int16_t *__restrict__ coeff, int16_t *__restrict__ coeff2, int count,
                                                               const int8_t *__restrict__ modes, int is_chroma, int tr_depth)
{
  constexpr int NN = 1 << (2 * LOG2N), W = 1 << LOG2N;
  __shared__ RdoqScratch<NN, SH> scratch[WARPS];
  __shared__ kvz_cuda_cabac_ctx s_ctx;
  __shared__ int32_t s_ebits[128];
  rdoq_load_ebits(s_ebits);
  __shared__ __align__(4) int16_t s_q[WARPS][NN];
  for (int i = threadIdx.x; i < (int)sizeof(kvz_cuda_cabac_ctx); i += blockDim.x) ((uint8_t *)&s_ctx)[i] = ((const uint8_t *)cabac)[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // two planes (U and V) can share one launch: TU indices [count, 2 * count) address coeff2
  int t = blockIdx.x * WARPS + warp;
  const bool active = t < (coeff2 ? 2 * count : count);
  int16_t *base = coeff;
  if (t >= count) { t -= count; base = coeff2; }
  __syncthreads();
  if (!active) return;
  int scan = 0;
  if ((!is_chroma && W <= 8) || (is_chroma && W == 4)) { const int m = modes[t]; scan = (m >= 6 && m <= 14) ? 2 : ((m >= 22 && m <= 30) ? 1 : 0); }
  // (in place: the coefficients are read from global memory, the levels are written back after the walk)
  rdoq_tu<NN, SH>(p, &s_ctx, s_ebits, base + (size_t)t * NN, s_q[warp], LOG2N, is_chroma ? 2 : 0, scan, 1, tr_depth, scratch[warp], lane);
  for (int e = lane; e < NN; e += 32) base[(size_t)t * NN + e] = s_q[warp][e];
}

// Thread-per-TU form of the uniform grid for 4x4 and 8x8 TUs (rdoq_tu_thread).
template <int LOG2N, bool SH>
__global__ void __launch_bounds__(128) rdoq_grid_thread_kernel(kvz_cuda_rdoq_params p, const kvz_cuda_cabac_ctx *__restrict__ cabac,
                                                               int16_t *__restrict__ coeff, int16_t *__restrict__ coeff2, int count,
                                                               const int8_t *__restrict__ modes, int is_chroma, int tr_depth)
{
  constexpr int NN = 1 << (2 * LOG2N), W = 1 << LOG2N;
  __shared__ kvz_cuda_cabac_ctx s_ctx;
  __shared__ int32_t s_ebits[128];
  __shared__ uint8_t s_scan[3][NN];
  rdoq_load_ebits(s_ebits);
  for (int i = threadIdx.x; i < (int)sizeof(kvz_cuda_cabac_ctx); i += blockDim.x) ((uint8_t *)&s_ctx)[i] = ((const uint8_t *)cabac)[i];
  for (int i = threadIdx.x; i < 3 * NN; i += blockDim.x) s_scan[i / NN][i % NN] = (uint8_t)scan_pos(i / NN, LOG2N, i % NN);
  __syncthreads();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (coeff2 ? 2 * count : count)) return;
  if (t >= count) { t -= count; coeff = coeff2; }
  int16_t coef[NN], q[NN];
  const uint4 *src = reinterpret_cast<const uint4 *>(coeff + (size_t)t * NN);
  bool any = false;
#pragma unroll
  for (int i = 0; i < NN / 8; ++i) { const uint4 v = src[i]; reinterpret_cast<uint4 *>(coef)[i] = v; any |= (v.x | v.y | v.z | v.w) != 0; }
  if (!any) return;                                     // all-zero coefficients quantise to all-zero levels (already in place)
  int scan = 0;
  if ((!is_chroma && W <= 8) || (is_chroma && W == 4)) { const int m = modes[t]; scan = (m >= 6 && m <= 14) ? 2 : ((m >= 22 && m <= 30) ? 1 : 0); }
  RdoqLocal<NN, SH> loc;
  loc.blk = s_scan[scan];
  rdoq_tu_thread<NN, SH>(p, &s_ctx, s_ebits, coef, q, LOG2N, is_chroma ? 2 : 0, scan, 1, tr_depth, loc);
  uint4 *dst = reinterpret_cast<uint4 *>(coeff + (size_t)t * NN);
#pragma unroll
  for (int i = 0; i < NN / 8; ++i) dst[i] = reinterpret_cast<const uint4 *>(q)[i];
}

int rdoq_launch_grid(const kvz_cuda_rdoq_params &p, const kvz_cuda_cabac_ctx *ctx_dev, int16_t *coeff, int16_t *coeff2, int count, int log2n,
                     const int8_t *modes, int is_chroma, int tr_depth, cudaStream_t st)
{
  const int total = coeff2 ? 2 * count : count;
  const bool SHV = p.signhide_enable != 0;
  switch (log2n) {
    case 2: if (SHV) rdoq_grid_thread_kernel<2, true><<<(total + 127) / 128, 128, 0, st>>>(p, ctx_dev, coeff, coeff2, count, modes, is_chroma, tr_depth); else rdoq_grid_thread_kernel<2, false><<<(total + 127) / 128, 128, 0, st>>>(p, ctx_dev, coeff, coeff2, count, modes, is_chroma, tr_depth); break;
    case 3: if (SHV) rdoq_grid_kernel<3, 8, true><<<(total + 7) / 8, 256, 0, st>>>(p, ctx_dev, coeff, coeff2, count, modes, is_chroma, tr_depth); else rdoq_grid_kernel<3, 8, false><<<(total + 7) / 8, 256, 0, st>>>(p, ctx_dev, coeff, coeff2, count, modes, is_chroma, tr_depth); break;
    case 4: if (SHV) rdoq_grid_kernel<4, 2, true><<<(total + 1) / 2, 64, 0, st>>>(p, ctx_dev, coeff, coeff2, count, modes, is_chroma, tr_depth); else rdoq_grid_kernel<4, 2, false><<<(total + 1) / 2, 64, 0, st>>>(p, ctx_dev, coeff, coeff2, count, modes, is_chroma, tr_depth); break;
    default: if (int r = rdoq_init_tables(st)) return r;
      if (SHV) rdoq_grid_kernel<5, 1, true><<<total, 32, 0, st>>>(p, ctx_dev, coeff, coeff2, count, modes, is_chroma, tr_depth); else rdoq_grid_kernel<5, 1, false><<<total, 32, 0, st>>>(p, ctx_dev, coeff, coeff2, count, modes, is_chroma, tr_depth); break;
  }
  KVZC_LAUNCHED();
  return 0;
}

int rdoq_launch_tus(const kvz_cuda_rdoq_params &p, const kvz_cuda_cabac_ctx *ctx_dev, int16_t *coeff, const kvz_cuda_tu *tus, int count, int n, cudaStream_t st)
{
  const bool SHV = p.signhide_enable != 0;
  switch (n) {
    case 4: if (SHV) rdoq_tu_kernel<2, 8, true><<<(count + 7) / 8, 256, 0, st>>>(p, ctx_dev, coeff, tus, count); else rdoq_tu_kernel<2, 8, false><<<(count + 7) / 8, 256, 0, st>>>(p, ctx_dev, coeff, tus, count); break;
    case 8: if (SHV) rdoq_tu_kernel<3, 8, true><<<(count + 7) / 8, 256, 0, st>>>(p, ctx_dev, coeff, tus, count); else rdoq_tu_kernel<3, 8, false><<<(count + 7) / 8, 256, 0, st>>>(p, ctx_dev, coeff, tus, count); break;
    case 16: if (SHV) rdoq_tu_kernel<4, 2, true><<<(count + 1) / 2, 64, 0, st>>>(p, ctx_dev, coeff, tus, count); else rdoq_tu_kernel<4, 2, false><<<(count + 1) / 2, 64, 0, st>>>(p, ctx_dev, coeff, tus, count); break;
    default: if (int r = rdoq_init_tables(st)) return r;
      if (SHV) rdoq_tu_kernel<5, 1, true><<<count, 32, 0, st>>>(p, ctx_dev, coeff, tus, count); else rdoq_tu_kernel<5, 1, false><<<count, 32, 0, st>>>(p, ctx_dev, coeff, tus, count); break;
  }
  KVZC_LAUNCHED();
  return 0;
}

}  // namespace kvzc

using namespace kvzc;

extern "C" int kvz_cuda_rdoq_batch(const kvz_cuda_rdoq_params *p, const kvz_cuda_cabac_ctx *ctx_dev, const int16_t *coef, int16_t *dest,
                                   int n, const kvz_cuda_rdoq_tu *tus, int count, void *stream)
{
  KVZC_REQUIRE_DEVICE();
  KVZC_ARG(p && ctx_dev && coef && dest && tus && count >= 0);
  KVZC_ARG(n == 4 || n == 8 || n == 16 || n == 32);
  KVZC_ARG(p->bitdepth == 8 || p->bitdepth == 10);
  KVZC_ARG(p->lambda > 0);
  if (count == 0) return 0;
  cudaStream_t st = as_stream(stream);
  const bool SHV = p->signhide_enable != 0;
  switch (n) {
    case 4: if (SHV) rdoq_kernel<2, 8, true><<<(count + 7) / 8, 256, 0, st>>>(*p, ctx_dev, coef, dest, tus, count); else rdoq_kernel<2, 8, false><<<(count + 7) / 8, 256, 0, st>>>(*p, ctx_dev, coef, dest, tus, count); break;
    case 8: if (SHV) rdoq_kernel<3, 8, true><<<(count + 7) / 8, 256, 0, st>>>(*p, ctx_dev, coef, dest, tus, count); else rdoq_kernel<3, 8, false><<<(count + 7) / 8, 256, 0, st>>>(*p, ctx_dev, coef, dest, tus, count); break;
    case 16: if (SHV) rdoq_kernel<4, 2, true><<<(count + 1) / 2, 64, 0, st>>>(*p, ctx_dev, coef, dest, tus, count); else rdoq_kernel<4, 2, false><<<(count + 1) / 2, 64, 0, st>>>(*p, ctx_dev, coef, dest, tus, count); break;
    default: if (int r = rdoq_init_tables(st)) return r;
      if (SHV) rdoq_kernel<5, 1, true><<<count, 32, 0, st>>>(*p, ctx_dev, coef, dest, tus, count); else rdoq_kernel<5, 1, false><<<count, 32, 0, st>>>(*p, ctx_dev, coef, dest, tus, count); break;
  }
  KVZC_LAUNCHED();
  return 0;
}
