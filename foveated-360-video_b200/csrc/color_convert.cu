// Colour conversions either side of the foveation path, on the device.
//
// RGB0 -> YUV420P / NV12:
//
// Replaces the colour conversion inside VideoEncoder::EncodeFrame (video_encoder.cc:380-398):
// the reference copies the foveated RGB0 buffer to the host, runs
// sws_getContext(w, h, RGB0, w, h, YUV420P, SWS_BILINEAR) + sws_scale on the CPU and uploads the
// planes again with av_hwframe_transfer_data.  Here the planes are produced where the reduced
// buffer already lives, straight into the device surface the hardware encoder reads (planar
// YUV420P = the reference's sw_format, video_encoder.cc:549; NV12 = NVENC's native layout, the
// same samples with U and V interleaved).
//
// The arithmetic is libswscale's C path, integer and exact (FFmpeg 4.2 libswscale: input.c:252-345,
// swscale.c:96-122, output.c:380-403, utils.c:403-418,807-816; the parity tests pin it to golden
// vectors produced by a real libswscale):
//   luma    Y  = clip8((min(2 * ((RY r + GY g + BY b + (32 << 14) + 256) >> 9), 32767) + 64) >> 7)
//   chroma  c15(row, pair) = min(2 * ((RU (r0+r1) + GU (g0+g1) + BU (b0+b1) + (256 << 15) + 512) >> 10), 32767)
//           U  = clip8(((64 << 12) + 512 c15(2j-1) + 1536 c15(2j) + 1536 c15(2j+1) + 512 c15(2j+2)) >> 19)
// with rows outside the frame replicating the edge row.  HBM traffic: 4 B/px read, 1.5 B/px
// written - 6 % of the foveation pipeline's bytes at the same frame size, so it stays a separate
// kernel behind sample_rect (whose boxes do not line up with the 2x4 chroma footprint).
#include "fov360_internal.h"

namespace fov {
namespace {

constexpr int kRY = 8414, kGY = 16519, kBY = 3208;
constexpr int kRU = -4865, kGU = -9528, kBU = 14392;
constexpr int kRV = 14392, kGV = -12061, kBV = -2332;

struct YuvArgs {
  uint8_t *y, *u, *v;  // v unused for NV12 (u = the interleaved plane)
  const uint8_t *src;
  size_t y_stride, c_stride, src_stride;  // per-frame strides in bytes
  int y_ls, c_ls, v_ls, src_ls;           // linesizes in bytes (c_ls: U or interleaved plane)
  int W, H;
};

__device__ __forceinline__ int clip8(int v) { return min(max(v, 0), 255); }

// c + k0 * byte0(p) + k1 * byte1(p) (lo) / c + k0 * byte2(p) + k1 * byte3(p) (hi): IDP.2A, the
// 16-bit coefficients packed as {k0, k1}.  Exact 32-bit integer arithmetic.
__device__ __forceinline__ int dp2a_lo(int k01, uint32_t p, int c) {
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(k01), "r"(p), "r"(c));
  return d;
}
__device__ __forceinline__ int dp2a_hi(int k01, uint32_t p, int c) {
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(k01), "r"(p), "r"(c));
  return d;
}
__host__ __device__ constexpr int pack16(int lo, int hi) { return (int)(((uint32_t)hi << 16) | ((uint32_t)lo & 0xffffu)); }

// The saturations and clips of the reference arithmetic cannot trigger for 8-bit input, which
// collapses the per-sample work (the parity oracle keeps the literal form):
//  * luma: t = RY r + GY g + BY b + (32 << 14) + 256 lies in [524544, 7700499], so y14 = t >> 9 is
//    at most 15040, 2 * y14 never reaches 32767, and (2 * y14 + 64) >> 7 = (t + (32 << 9)) >> 15 lies
//    in [16, 235];
//  * chroma: c14 = (dot + (256 << 15) + 512) >> 10 lies in [1024, 15360] (again no saturation of
//    2 * c14), and with the taps 512 * (1, 3, 3, 1)
//    ((64 << 12) + sum tap_k * 2 c14_k) >> 19 = (c14_0 + 3 (c14_1 + c14_2) + c14_3 + 256) >> 9,
//    which lies in [16, 240].
__device__ __forceinline__ int luma8(uint32_t p) {
  // R, G from bytes 0, 1; B from byte 2 (byte 3, the padding, is multiplied by 0)
  return dp2a_hi(pack16(kBY, 0), p,
                 dp2a_lo(pack16(kRY, kGY), p, (32 << 14) + 256 + (32 << 9))) >> 15;
}

// 14-bit chroma of one horizontal pixel pair: the matrix is linear, so the pair sum of the
// reference is accumulated pixel by pixel
__device__ __forceinline__ void chroma14(uint32_t p0, uint32_t p1, int &u14, int &v14) {
  constexpr int rnd = (256 << 15) + 512;
  int u = dp2a_hi(pack16(kBU, 0), p0, dp2a_lo(pack16(kRU, kGU), p0, rnd));
  u = dp2a_hi(pack16(kBU, 0), p1, dp2a_lo(pack16(kRU, kGU), p1, u));
  int v = dp2a_hi(pack16(kBV, 0), p0, dp2a_lo(pack16(kRV, kGV), p0, rnd));
  v = dp2a_hi(pack16(kBV, 0), p1, dp2a_lo(pack16(kRV, kGV), p1, v));
  u14 = u >> 10;
  v14 = v >> 10;
}

// One thread: 4 consecutive pixels x kYuvRows chroma rows (2 * kYuvRows frame rows).  All
// 2 * kYuvRows + 2 frame rows the thread needs (its own plus one above and one below for the
// vertical chroma filter) are requested before any is consumed - the kernel is bound by load
// latency, not by arithmetic - and every row goes through the chroma matrix once per thread.
constexpr int kYuvRows = 4;

template <bool kNV12>
__global__ void __launch_bounds__(256) rgb0_to_yuv_kernel(const YuvArgs a) {
  const int x0 = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int j0 = (blockIdx.y * 8 + threadIdx.y) * kYuvRows;  // first chroma row
  const int W = a.W, H = a.H;
  if (x0 >= W || 2 * j0 >= H) return;
  const int f = blockIdx.z;
  const uint8_t *src = a.src + (size_t)f * a.src_stride;
  const bool full = x0 + 4 <= W;  // false only for the last thread of a row when W % 4 == 2
  const bool vec = full && ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)a.src_ls) & 15) == 0;
  constexpr int kRows = 2 * kYuvRows + 2;

  uint32_t px[kRows][4];  // frame rows 2*j0 - 1 .. 2*j0 + 2*kYuvRows, clamped to the frame
#pragma unroll
  for (int k = 0; k < kRows; ++k) {
    const int yy = min(max(2 * j0 - 1 + k, 0), H - 1);
    const uint8_t *row = src + (size_t)yy * a.src_ls + (size_t)x0 * 4;
    if (vec) {
      const uint4 q = __ldg(reinterpret_cast<const uint4 *>(row));
      px[k][0] = q.x, px[k][1] = q.y, px[k][2] = q.z, px[k][3] = q.w;
    } else {
      const uint32_t *r32 = reinterpret_cast<const uint32_t *>(row);
      px[k][0] = __ldg(r32), px[k][1] = __ldg(r32 + 1);
      px[k][2] = full ? __ldg(r32 + 2) : 0u, px[k][3] = full ? __ldg(r32 + 3) : 0u;
    }
  }

  // 14-bit chroma of every row, both pairs
  int cu[kRows][2], cv[kRows][2];
#pragma unroll
  for (int k = 0; k < kRows; ++k)
#pragma unroll
    for (int p = 0; p < 2; ++p) chroma14(px[k][2 * p], px[k][2 * p + 1], cu[k][p], cv[k][p]);

#pragma unroll
  for (int t = 0; t < kYuvRows; ++t) {
    const int j = j0 + t;
    if (2 * j >= H) break;
    // luma: frame rows 2j and 2j+1 are px[2t+1] and px[2t+2]
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const uint32_t *q = px[2 * t + 1 + r];
      uint8_t *yrow = a.y + (size_t)f * a.y_stride + (size_t)(2 * j + r) * a.y_ls + x0;
      const int l0 = luma8(q[0]), l1 = luma8(q[1]), l2 = luma8(q[2]), l3 = luma8(q[3]);
      if (full && (reinterpret_cast<uintptr_t>(yrow) & 3) == 0) {
        *reinterpret_cast<uint32_t *>(yrow) = (uint32_t)l0 | (l1 << 8) | (l2 << 16) | (l3 << 24);
      } else {
        yrow[0] = (uint8_t)l0, yrow[1] = (uint8_t)l1;
        if (full) yrow[2] = (uint8_t)l2, yrow[3] = (uint8_t)l3;
      }
    }
    // chroma: the vertical taps (1, 3, 3, 1) / 8 over frame rows 2j-1 .. 2j+2 = rows 2t .. 2t+3
    int uu[2], vv[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      uu[p] = (cu[2 * t][p] + 3 * (cu[2 * t + 1][p] + cu[2 * t + 2][p]) + cu[2 * t + 3][p] + 256) >> 9;
      vv[p] = (cv[2 * t][p] + 3 * (cv[2 * t + 1][p] + cv[2 * t + 2][p]) + cv[2 * t + 3][p] + 256) >> 9;
    }
    const int u0 = uu[0], u1 = uu[1], v0 = vv[0], v1 = vv[1];
    const int cx = x0 / 2;
    if (kNV12) {
      uint8_t *c = a.u + (size_t)f * a.c_stride + (size_t)j * a.c_ls + (size_t)cx * 2;
      if (full && (reinterpret_cast<uintptr_t>(c) & 3) == 0) {
        *reinterpret_cast<uint32_t *>(c) = (uint32_t)u0 | (v0 << 8) | (u1 << 16) | (v1 << 24);
      } else {
        c[0] = (uint8_t)u0, c[1] = (uint8_t)v0;
        if (full) c[2] = (uint8_t)u1, c[3] = (uint8_t)v1;
      }
    } else {
      uint8_t *up = a.u + (size_t)f * a.c_stride + (size_t)j * a.c_ls + cx;
      uint8_t *vp = a.v + (size_t)f * a.c_stride + (size_t)j * a.v_ls + cx;
      if (full && ((reinterpret_cast<uintptr_t>(up) | reinterpret_cast<uintptr_t>(vp)) & 1) == 0) {
        *reinterpret_cast<uint16_t *>(up) = (uint16_t)(u0 | (u1 << 8));
        *reinterpret_cast<uint16_t *>(vp) = (uint16_t)(v0 | (v1 << 8));
      } else {
        up[0] = (uint8_t)u0, vp[0] = (uint8_t)v0;
        if (full) up[1] = (uint8_t)u1, vp[1] = (uint8_t)v1;
      }
    }
  }
}

// ---- YUV420P / NV12 -> RGB0 --------------------------------------------
//
// Replaces the sws_scale inside VideoDecoder::GetFrame (video_decoder.cc:165-170, :222) so that a
// frame decoded on the device (NVDEC hands out NV12 surfaces) becomes the RGB0 frame
// EncodeFrameGPU reads without a host round trip.  libswscale converts same-size YUV420P to packed
// RGB with its yuv2rgb converter: chroma is replicated over each 2x2 block and the arithmetic is
// 16-bit fixed point (x86/yuv2rgb_template.c:95-98, coefficients yuv2rgb.c:830-837):
//   Y' = (((Y << 3) - 128) * 9539) >> 16,  U' = (U << 3) - 1024,  V' = (V << 3) - 1024
//   R = clip8(Y' + (V' * 13075 >> 16)),  G = clip8(Y' + (U' * -3209 >> 16) + (V' * -6660 >> 16)),
//   B = clip8(Y' + (U' * 16525 >> 16)),  4th byte = 255.
struct RgbArgs {
  uint8_t *dst;
  const uint8_t *y, *u, *v;  // v unused for NV12 (u = the interleaved plane)
  size_t dst_stride, y_stride, c_stride;
  int dst_ls, y_ls, c_ls, v_ls;
  int W, H;
};

__device__ __forceinline__ uint32_t yuv_px(int Y, int rv, int guv, int bu) {
  const int yt = (((Y << 3) - 128) * 9539) >> 16;
  return (uint32_t)clip8(yt + rv) | ((uint32_t)clip8(yt + guv) << 8) |
         ((uint32_t)clip8(yt + bu) << 16) | 0xff000000u;
}

// One thread: 4 consecutive pixels x 2 rows (two chroma samples).
template <bool kNV12>
__global__ void __launch_bounds__(256) yuv_to_rgb0_kernel(const RgbArgs a) {
  const int x0 = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int j = blockIdx.y * 8 + threadIdx.y;  // chroma row
  const int W = a.W, H = a.H;
  if (x0 >= W || 2 * j >= H) return;
  const int f = blockIdx.z;
  const bool full = x0 + 4 <= W;  // false only for the last thread of a row when W % 4 == 2
  int U[2], V[2];
  if (kNV12) {
    const uint8_t *c = a.u + (size_t)f * a.c_stride + (size_t)j * a.c_ls + x0;
    U[0] = c[0], V[0] = c[1];
    U[1] = full ? c[2] : 0, V[1] = full ? c[3] : 0;
  } else {
    const uint8_t *up = a.u + (size_t)f * a.c_stride + (size_t)j * a.c_ls + x0 / 2;
    const uint8_t *vp = a.v + (size_t)f * a.c_stride + (size_t)j * a.v_ls + x0 / 2;
    U[0] = up[0], V[0] = vp[0];
    U[1] = full ? up[1] : 0, V[1] = full ? vp[1] : 0;
  }
  int rv[2], guv[2], bu[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int us = (U[p] << 3) - 1024, vs = (V[p] << 3) - 1024;
    rv[p] = (vs * 13075) >> 16;
    guv[p] = ((us * -3209) >> 16) + ((vs * -6660) >> 16);
    bu[p] = (us * 16525) >> 16;
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const uint8_t *yrow = a.y + (size_t)f * a.y_stride + (size_t)(2 * j + r) * a.y_ls + x0;
    uint32_t px[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      px[k] = (k < 2 || full) ? yuv_px(yrow[k], rv[k / 2], guv[k / 2], bu[k / 2]) : 0u;
    uint8_t *o = a.dst + (size_t)f * a.dst_stride + (size_t)(2 * j + r) * a.dst_ls + (size_t)x0 * 4;
    if (full && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
      *reinterpret_cast<uint4 *>(o) = make_uint4(px[0], px[1], px[2], px[3]);
    } else {
      uint32_t *o32 = reinterpret_cast<uint32_t *>(o);
      o32[0] = px[0], o32[1] = px[1];
      if (full) o32[2] = px[2], o32[3] = px[3];
    }
  }
}

}  // namespace

cudaError_t launch_yuv_to_rgb0(const LaunchCtx &lc, bool nv12, int n, uint8_t *dst,
                               size_t dst_stride, int dst_ls, const uint8_t *y, size_t y_stride,
                               int y_ls, const uint8_t *u, int u_ls, const uint8_t *v, int v_ls,
                               size_t c_stride, int W, int H) {
  RgbArgs a;
  a.dst = dst, a.y = y, a.u = u, a.v = v;
  a.dst_stride = dst_stride, a.y_stride = y_stride, a.c_stride = c_stride;
  a.dst_ls = dst_ls, a.y_ls = y_ls, a.c_ls = u_ls, a.v_ls = v_ls;
  a.W = W, a.H = H;
  const dim3 grid((W + 127) / 128, (H / 2 + 7) / 8, n), block(32, 8);
  KernelScope ks(lc, nv12 ? "nv12_to_rgb0" : "yuv420p_to_rgb0");
  if (nv12)
    yuv_to_rgb0_kernel<true><<<grid, block, 0, lc.stream>>>(a);
  else
    yuv_to_rgb0_kernel<false><<<grid, block, 0, lc.stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_rgb0_to_yuv(const LaunchCtx &lc, bool nv12, int n, uint8_t *y, size_t y_stride,
                               int y_ls, uint8_t *u, int u_ls, uint8_t *v, int v_ls, size_t c_stride,
                               const uint8_t *src, size_t src_stride, int src_ls, int W, int H) {
  YuvArgs a;
  a.y = y, a.u = u, a.v = v, a.src = src;
  a.y_stride = y_stride, a.c_stride = c_stride, a.src_stride = src_stride;
  a.y_ls = y_ls, a.c_ls = u_ls, a.v_ls = v_ls, a.src_ls = src_ls;
  a.W = W, a.H = H;
  const dim3 grid((W + 127) / 128, (H / 2 + 8 * kYuvRows - 1) / (8 * kYuvRows), n), block(32, 8);
  KernelScope ks(lc, nv12 ? "rgb0_to_nv12" : "rgb0_to_yuv420p");
  if (nv12)
    rgb0_to_yuv_kernel<true><<<grid, block, 0, lc.stream>>>(a);
  else
    rgb0_to_yuv_kernel<false><<<grid, block, 0, lc.stream>>>(a);
  return cudaGetLastError();
}

}  // namespace fov
