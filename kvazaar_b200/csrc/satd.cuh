// satd.cuh -- Hadamard (SATD) device primitives.
//
// Arithmetic notes (bit-exactness vs ref: picture-generic.c:117-340):
//  * SATD only sums |coefficients| of the 2-D Walsh-Hadamard transform of the difference block, so any
//    butterfly ordering (and any XOR-permutation of rows/columns) gives the identical result.
//  * 8-bit pixels: |diff| <= 255, an 8x8 2-D Hadamard coefficient is <= 255*64 = 16320, so two values are
//    carried per 32-bit register as  lo + hi*65536  ("arithmetic packing": plain IADD/ISUB act lane-wise,
//    no carries leak because each lane stays inside int16).  The last butterfly stage is folded into the
//    absolute sum with |x+y| + |x-y| = 2*max(|x|,|y|).
//  * 10-bit pixels: a full 8x8 coefficient does not fit (1023*64 > 32767), but the packed path only materialises five of
//    the six butterfly stages (1023*32 = 32736 fits) -- the fused rough search and the 16-bit strided sub-block SATD
//    (satd_sub_strided<uint16_t, N>) use it; inputs must be <= 10-bit samples.
#pragma once
#include "common.cuh"

namespace kvzc {

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }

// lanes (lo,hi) of an arithmetic-packed register -> 2*max(|lo|,|hi|) accumulated as max only
__device__ __forceinline__ int packed_absmax(int x)
{
  const int lo = (int)(short)x;                        // sign-extend low half (PRMT/SGXT)
  const int hi = (x - lo) >> 16;
  return max(abs(lo), abs(hi));
}

#define KVZC_BFLY(p, q) { const int t__ = (p) - (q); (p) = (p) + (q); (q) = t__; }

// Raw sum of |H d H^T| from packed difference lanes d[row][k]: k=0 cols (0,2), k=1 cols (1,3), k=2 cols (4,6),
// k=3 cols (5,7), each register = lo + hi*65536.  Returns sum (caller applies (s + 2) >> 2).
__device__ __forceinline__ uint32_t hadamard8x8_lanes(int (&d)[8][4])
{
#pragma unroll
  for (int r = 0; r < 8; ++r) {           // horizontal, distances 1 and 4 (distance 2 is in-register, done last)
    KVZC_BFLY(d[r][0], d[r][1]); KVZC_BFLY(d[r][2], d[r][3]);
    KVZC_BFLY(d[r][0], d[r][2]); KVZC_BFLY(d[r][1], d[r][3]);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {           // vertical, distances 1, 2, 4
#pragma unroll
    for (int r = 0; r < 8; r += 2) KVZC_BFLY(d[r][k], d[r + 1][k]);
#pragma unroll
    for (int r = 0; r < 8; r += 4) { KVZC_BFLY(d[r][k], d[r + 2][k]); KVZC_BFLY(d[r + 1][k], d[r + 3][k]); }
#pragma unroll
    for (int r = 0; r < 4; ++r) KVZC_BFLY(d[r][k], d[r + 4][k]);
  }
  int s = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k) s += packed_absmax(d[r][k]);
  return (uint32_t)(2 * s);
}

// Raw sum of |H d H^T| for one 8x8 block of 8-bit pixels; a[r], b[r] = the 8 bytes of row r.
__device__ __forceinline__ uint32_t hadamard8x8_u8(const uint2 (&a)[8], const uint2 (&b)[8])
{
  int d[8][4];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    d[r][0] = (int)prmt(a[r].x, 0u, 0x4240) - (int)prmt(b[r].x, 0u, 0x4240);
    d[r][1] = (int)prmt(a[r].x, 0u, 0x4341) - (int)prmt(b[r].x, 0u, 0x4341);
    d[r][2] = (int)prmt(a[r].y, 0u, 0x4240) - (int)prmt(b[r].y, 0u, 0x4240);
    d[r][3] = (int)prmt(a[r].y, 0u, 0x4341) - (int)prmt(b[r].y, 0u, 0x4341);
  }
  return hadamard8x8_lanes(d);
}

// 4x4 from packed difference lanes d[row][k]: k=0 cols (0,2), k=1 cols (1,3). Returns raw sum (caller: (s+1)>>1).
__device__ __forceinline__ uint32_t hadamard4x4_lanes(int (&d)[4][2])
{
#pragma unroll
  for (int r = 0; r < 4; ++r) KVZC_BFLY(d[r][0], d[r][1]);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    KVZC_BFLY(d[0][k], d[1][k]); KVZC_BFLY(d[2][k], d[3][k]);
    KVZC_BFLY(d[0][k], d[2][k]); KVZC_BFLY(d[1][k], d[3][k]);
  }
  int s = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r) { s += packed_absmax(d[r][0]); s += packed_absmax(d[r][1]); }
  return (uint32_t)(2 * s);
}

// 4x4 block of 8-bit pixels, a = 16 contiguous bytes (rows in x,y,z,w). Returns raw sum (caller: (s+1)>>1).
__device__ __forceinline__ uint32_t hadamard4x4_u8(const uint32_t (&a)[4], const uint32_t (&b)[4])
{
  int d[4][2];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    d[r][0] = (int)prmt(a[r], 0u, 0x4240) - (int)prmt(b[r], 0u, 0x4240);
    d[r][1] = (int)prmt(a[r], 0u, 0x4341) - (int)prmt(b[r], 0u, 0x4341);
    KVZC_BFLY(d[r][0], d[r][1]);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    KVZC_BFLY(d[0][k], d[1][k]); KVZC_BFLY(d[2][k], d[3][k]);
    KVZC_BFLY(d[0][k], d[2][k]); KVZC_BFLY(d[1][k], d[3][k]);
  }
  int s = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r) { s += packed_absmax(d[r][0]); s += packed_absmax(d[r][1]); }
  return (uint32_t)(2 * s);
}

// ---------------------------------------------------------------- generic int32 path (any pixel type, any stride)
template <int N> __device__ __forceinline__ uint32_t hadamard_abs_sum_i32(int (&d)[N][N])
{
#pragma unroll
  for (int r = 0; r < N; ++r)
#pragma unroll
    for (int half = 1; half < N; half <<= 1)
#pragma unroll
      for (int i = 0; i < N; ++i)
        if ((i & half) == 0) KVZC_BFLY(d[r][i], d[r][i + half]);
#pragma unroll
  for (int c = 0; c < N; ++c)
#pragma unroll
    for (int half = 1; half < N; half <<= 1)
#pragma unroll
      for (int i = 0; i < N; ++i)
        if ((i & half) == 0) KVZC_BFLY(d[i][c], d[i + half][c]);
  uint32_t s = 0;
#pragma unroll
  for (int r = 0; r < N; ++r)
#pragma unroll
    for (int c = 0; c < N; ++c) s += (uint32_t)abs(d[r][c]);
  return s;
}

// strided NxN sub-block SATD with element loads (used by any_size paths and by the 10-bit build)
template <class T, int N> __device__ __forceinline__ uint32_t satd_sub_strided(const T *a, int sa, const T *b, int sb)
{
  int d[N][N];
#pragma unroll
  for (int y = 0; y < N; ++y)
#pragma unroll
    for (int x = 0; x < N; ++x) d[y][x] = (int)a[y * sa + x] - (int)b[y * sb + x];
  const uint32_t s = hadamard_abs_sum_i32<N>(d);
  return N == 4 ? (s + 1) >> 1 : (s + 2) >> 2;
}

// 16-bit samples of up to 10 significant bits: the packed path is exact too (five materialised butterfly stages of
// |d| <= 1023 stay below 2^15, see the notes on top), with (c, c+2) lanes assembled from element loads
__device__ __forceinline__ uint32_t pair16(const uint16_t *p, int i) { return (uint32_t)p[i] | ((uint32_t)p[i + 2] << 16); }
template <> __device__ __forceinline__ uint32_t satd_sub_strided<uint16_t, 8>(const uint16_t *a, int sa, const uint16_t *b, int sb)
{
  int d[8][4];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const uint16_t *pa = a + r * sa, *pb = b + r * sb;
    d[r][0] = (int)pair16(pa, 0) - (int)pair16(pb, 0);
    d[r][1] = (int)pair16(pa, 1) - (int)pair16(pb, 1);
    d[r][2] = (int)pair16(pa, 4) - (int)pair16(pb, 4);
    d[r][3] = (int)pair16(pa, 5) - (int)pair16(pb, 5);
  }
  return (hadamard8x8_lanes(d) + 2) >> 2;
}
template <> __device__ __forceinline__ uint32_t satd_sub_strided<uint16_t, 4>(const uint16_t *a, int sa, const uint16_t *b, int sb)
{
  int d[4][2];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const uint16_t *pa = a + r * sa, *pb = b + r * sb;
    d[r][0] = (int)pair16(pa, 0) - (int)pair16(pb, 0);
    d[r][1] = (int)pair16(pa, 1) - (int)pair16(pb, 1);
  }
  return (hadamard4x4_lanes(d) + 1) >> 1;
}

// 8-bit strided 8x8 via packed path with unaligned-safe row loads
__device__ __forceinline__ uint2 load_row8_u8(const uint8_t *p)
{
  if ((((uintptr_t)p) & 7) == 0) return *reinterpret_cast<const uint2 *>(p);
  if ((((uintptr_t)p) & 3) == 0) { uint2 v; v.x = *reinterpret_cast<const uint32_t *>(p); v.y = *reinterpret_cast<const uint32_t *>(p + 4); return v; }
  uint2 v;
  v.x = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
  v.y = (uint32_t)p[4] | ((uint32_t)p[5] << 8) | ((uint32_t)p[6] << 16) | ((uint32_t)p[7] << 24);
  return v;
}
__device__ __forceinline__ uint32_t load_row4_u8(const uint8_t *p)
{
  if ((((uintptr_t)p) & 3) == 0) return *reinterpret_cast<const uint32_t *>(p);
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

template <class T> __device__ __forceinline__ uint32_t satd8_sub(const T *a, int sa, const T *b, int sb);
template <> __device__ __forceinline__ uint32_t satd8_sub<uint8_t>(const uint8_t *a, int sa, const uint8_t *b, int sb)
{
  uint2 ra[8], rb[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) { ra[r] = load_row8_u8(a + r * sa); rb[r] = load_row8_u8(b + r * sb); }
  return (hadamard8x8_u8(ra, rb) + 2) >> 2;
}
template <> __device__ __forceinline__ uint32_t satd8_sub<uint16_t>(const uint16_t *a, int sa, const uint16_t *b, int sb)
{
  return satd_sub_strided<uint16_t, 8>(a, sa, b, sb);
}
template <class T> __device__ __forceinline__ uint32_t satd4_sub(const T *a, int sa, const T *b, int sb);
template <> __device__ __forceinline__ uint32_t satd4_sub<uint8_t>(const uint8_t *a, int sa, const uint8_t *b, int sb)
{
  uint32_t ra[4], rb[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) { ra[r] = load_row4_u8(a + r * sa); rb[r] = load_row4_u8(b + r * sb); }
  return (hadamard4x4_u8(ra, rb) + 1) >> 1;
}
template <> __device__ __forceinline__ uint32_t satd4_sub<uint16_t>(const uint16_t *a, int sa, const uint16_t *b, int sb)
{
  return satd_sub_strided<uint16_t, 4>(a, sa, b, sb);
}

}  // namespace kvzc
