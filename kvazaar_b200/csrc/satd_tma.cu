// satd_tma.cu -- persistent, TMA-fed batched SATD 8x8 for 8-bit pixels (the HBM-streaming roofline kernel).
//
// The plain kernel (picture.cu) lets every thread pull its own 64-byte block with four 128-bit loads whose
// addresses are 64 bytes apart across the warp.  Here the block pairs are streamed by the TMA engine instead:
// lane 0 of each warp issues 1-D bulk copies (cp.async.bulk, SASS UBLKCP) of 32 pairs (2 KiB of `a`, 2 KiB of `b`)
// into shared-memory rings guarded by mbarriers -- one private 3-deep ring per warp, so no CTA-wide barrier ever
// stalls the arithmetic -- the grid is persistent (a multiple of the SM count) and threads only read shared memory.  Bank conflicts of the "thread t owns bytes 64t..64t+63" pattern
// are avoided without any data swizzle: thread t reads its four 16-byte chunks in the order j ^ ((t >> 1) & 3);
// that XOR-permutes the rows of the 8x8 block, which leaves sum |H d H^T| unchanged (Walsh-Hadamard rows are
// closed under XOR of the index), so no un-permute is needed.
#include "common.cuh"
#include "satd.cuh"

namespace kvzc {

constexpr int kWarps = 8;                  // warps per CTA; every warp owns a private ring (no CTA-wide barrier)
constexpr int kStages = 3;
constexpr int kTilePairs = 32;             // one pair per lane
constexpr int kTileBytes = kTilePairs * 64;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(kWarps * 32) satd8_tma_kernel(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, int count,
                                                                uint32_t *__restrict__ out)
{
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[kWarps][kStages];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t *sa = smem + (size_t)warp * 2 * kStages * kTileBytes;      // [kStages][kTileBytes] for a, then for b
  uint8_t *sb = sa + kStages * kTileBytes;
  const int ntiles = (count + kTilePairs - 1) / kTilePairs;
  const int gwarp = blockIdx.x * kWarps + warp, nwarps = gridDim.x * kWarps;
  if (lane == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[warp][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto issue = [&](int tile, int slot) {
    const int pairs = min(kTilePairs, count - tile * kTilePairs);
    const uint32_t bytes = (uint32_t)pairs * 64u;
    mbar_expect_tx(&full[warp][slot], 2 * bytes);
    bulk_load(sa + slot * kTileBytes, a + (size_t)tile * kTileBytes, bytes, &full[warp][slot]);
    bulk_load(sb + slot * kTileBytes, b + (size_t)tile * kTileBytes, bytes, &full[warp][slot]);
  };
  if (lane == 0)
    for (int s = 0; s < kStages; ++s) {
      const int tile = gwarp + s * nwarps;
      if (tile < ntiles) issue(tile, s);
    }
  const int m = (lane >> 1) & 3;
  int it = 0;
  for (int tile = gwarp; tile < ntiles; tile += nwarps, ++it) {
    const int slot = it % kStages;
    mbar_wait(&full[warp][slot], (uint32_t)(it / kStages) & 1u);
    const int pair = tile * kTilePairs + lane;
    uint32_t cost = 0;
    if (pair < count) {
      const uint4 *pa = reinterpret_cast<const uint4 *>(sa + slot * kTileBytes + lane * 64);
      const uint4 *pb = reinterpret_cast<const uint4 *>(sb + slot * kTileBytes + lane * 64);
      uint2 ra[8], rb[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 x = pa[j ^ m], y = pb[j ^ m];
        ra[2 * j] = make_uint2(x.x, x.y); ra[2 * j + 1] = make_uint2(x.z, x.w);
        rb[2 * j] = make_uint2(y.x, y.y); rb[2 * j + 1] = make_uint2(y.z, y.w);
      }
      cost = (hadamard8x8_u8(ra, rb) + 2) >> 2;
    }
    __syncwarp();                                        // the whole warp has consumed this slot
    if (lane == 0) {
      const int next = tile + kStages * nwarps;
      if (next < ntiles) issue(next, slot);
    }
    if (pair < count) out[pair] = cost;
  }
}

int satd8_tma(const uint8_t *a, const uint8_t *b, int count, uint32_t *out, cudaStream_t st)
{
  static bool attr_set = false;
  const int smem = kWarps * 2 * kStages * kTileBytes;   // 96 KiB: two CTAs (16 warps) per SM
  if (!attr_set) {
    KVZC_CHECK(cudaFuncSetAttribute(satd8_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int ntiles = (count + kTilePairs - 1) / kTilePairs;
  int grid = g_sm_count * 2;                            // persistent: a multiple of the SM count
  if (grid * kWarps > ntiles) grid = (ntiles + kWarps - 1) / kWarps;
  satd8_tma_kernel<<<grid, kWarps * 32, smem, st>>>(a, b, count, out);
  KVZC_LAUNCHED();
  return 0;
}

}  // namespace kvzc
