// Device helpers shared by the SAT-build kernels (sat_encode.cu, sat_onepass.cu).
#pragma once
#include <stdint.h>

namespace fov {

constexpr int kStripPx = 128;             // pixels per warp strip (32 lanes x 4 px)
#ifndef FOV360_STAGE_BUFS
#define FOV360_STAGE_BUFS 2
#endif
constexpr int kStageBufs = FOV360_STAGE_BUFS;             // TMA-store ring depth per warp
constexpr int kRowBytes = kStripPx * 12;  // one SAT row segment of a warp strip

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// Loads the 4 pixels [x0, x0+4) of a row as four RGB0-packed words (r | g<<8 | b<<16).
// FAST: RGB0 input with 16-byte aligned rows -> one 128-bit load; otherwise byte loads with
// the reference's pixel stride linesize / W (sat_encoder_encode_kernels.cl:9).
template <bool FAST>
__device__ __forceinline__ uint4 load_px4(const uint8_t *row, int x0, int W, int bpp) {
  uint4 v = make_uint4(0, 0, 0, 0);
  if (FAST) {
    if (x0 < W) v = ldg_stream(reinterpret_cast<const uint4 *>(row + (size_t)x0 * 4));
  } else {
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = x0 + k;
      if (x < W) {
        const uint8_t *q = row + (size_t)x * bpp;
        w[k] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16);
      }
    }
    v = make_uint4(w[0], w[1], w[2], w[3]);
  }
  return v;
}

__device__ __forceinline__ void unpack_px4(const uint4 v, uint32_t (&p)[12]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    p[3 * k + 0] = w[k] & 0xffu;
    p[3 * k + 1] = (w[k] >> 8) & 0xffu;
    p[3 * k + 2] = (w[k] >> 16) & 0xffu;
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// Inclusive warp scan of three independent u32 values (Kogge-Stone on shuffles).
__device__ __forceinline__ void warp_scan3(uint32_t &a, uint32_t &b, uint32_t &c, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t ta = __shfl_up_sync(0xffffffffu, a, d);
    const uint32_t tb = __shfl_up_sync(0xffffffffu, b, d);
    const uint32_t tc = __shfl_up_sync(0xffffffffu, c, d);
    if (lane >= d) {
      a += ta;
      b += tb;
      c += tc;
    }
  }
}

// Inclusive warp scan of two independent u32 values.
__device__ __forceinline__ void warp_scan2(uint32_t &a, uint32_t &b, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t ta = __shfl_up_sync(0xffffffffu, a, d);
    const uint32_t tb = __shfl_up_sync(0xffffffffu, b, d);
    if (lane >= d) {
      a += ta;
      b += tb;
    }
  }
}

}  // namespace fov
