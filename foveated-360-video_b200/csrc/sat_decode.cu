// SATDecoder kernels for sm_100a: gaze-dependent box sampling from the SAT, the inverse
// log-rectilinear warp back to full resolution, and the exact 1x1 SAT->image decode.
//
// Replaces sample_rect_kernel (sat_decoder_sample_rect_kernel.cl:138-241),
// interpolate_rect_kernel (sat_decoder_interpolate_kernel.cl:1-152) and decode_kernel
// (sat_decoder_decode_kernel.cl:1-58).  None of the three evaluates a transcendental on the
// device: the separable, gaze-independent parts of the transform come from host-built 1-D
// tables (luts.cc), so the kernels are integer/gather work bounded by HBM/L2 bandwidth.
#include "fov360_internal.h"

namespace fov {
namespace {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// (int)(c * dim): float multiply, truncation toward zero
// (sat_decoder_sample_rect_kernel.cl:176, sat_decoder_interpolate_kernel.cl:24-25).
__device__ __forceinline__ int gaze_px(float c, int dim) {
  return __float2int_rz(__fmul_rn(c, (float)dim));
}

// Exact unsigned division num / den given r ~= 1/den.  Box sums are tiny next to 2^32 on
// real frames, so the float estimate + one correction step is the common path.
__device__ __forceinline__ uint32_t udiv_exact(uint32_t num, uint32_t den, float rden) {
  if (num >= (1u << 22)) return num / den;
  uint32_t q = (uint32_t)__float2int_rz(__fmul_rn((float)num, rden));
  const int32_t r = (int32_t)(num - q * den);
  if (r < 0)
    --q;
  else if ((uint32_t)r >= den)
    ++q;
  return q;
}

// ---------------------------------------------------------------------------------------
// sample_rect: one thread per reduced pixel; 4 SAT corners (12 B each) -> box average.
// ---------------------------------------------------------------------------------------
struct SampleArgs {
  uint8_t *out;
  const uint32_t *sat;
  const int16_t *xedge, *yedge;
  size_t out_stride, sat_stride;
  int ow, oh, o_linesize_px, W, H;
};

__global__ void __launch_bounds__(256) sat_sample_rect_kernel(const SampleArgs a,
                                                              const GazeBatch g) {
  const int i = blockIdx.x * 32 + threadIdx.x;
  const int j = blockIdx.y * 8 + threadIdx.y;
  const int f = blockIdx.z;
  if (i >= a.ow || j >= a.oh) return;
  const int W = a.W, H = a.H;
  const int cxp = gaze_px(g.xy[2 * f], W);
  const int cyp = gaze_px(g.xy[2 * f + 1], H);

  int px = cxp + a.xedge[i + 1];  // grid[j+1][i+1].x  (:168-169)
  int mx = cxp + a.xedge[i];      // grid[j+1][i].x    (:170-171)
  int py = cyp + a.yedge[j + 1];  // grid[j+1][i+1].y  (:172-173)
  int my = cyp + a.yedge[j];      // grid[j][i+1].y    (:174-175)
  if (px >= W && mx >= W) {       // :181-187
    px -= W;
    mx -= W;
  } else if (px < 0 && mx < 0) {
    px += W;
    mx += W;
  }
  const bool x_in = (px >= 0 && px < W) || (mx >= 0 && mx < W);
  const bool y_in = (py >= 0 && py < H) || (my >= 0 && my < H);
  if (!(x_in && y_in)) return;  // :197-200: leave the pixel untouched
  px = clampi(px, 1, W - 1);    // :201-204
  py = clampi(py, 1, H - 1);
  mx = clampi(mx, 0, px - 1);
  my = clampi(my, 0, py - 1);

  const uint32_t *sat =
      reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(a.sat) +
                                         (size_t)f * a.sat_stride);
  const uint32_t *tl = sat + ((size_t)my * W + mx) * 3;
  const uint32_t *tr = sat + ((size_t)my * W + px) * 3;
  const uint32_t *bl = sat + ((size_t)py * W + mx) * 3;
  const uint32_t *br = sat + ((size_t)py * W + px) * 3;
  uint32_t s0 = __ldg(br + 0) - __ldg(tr + 0) + __ldg(tl + 0) - __ldg(bl + 0);  // :212-217
  uint32_t s1 = __ldg(br + 1) - __ldg(tr + 1) + __ldg(tl + 1) - __ldg(bl + 1);
  uint32_t s2 = __ldg(br + 2) - __ldg(tr + 2) + __ldg(tl + 2) - __ldg(bl + 2);
  const uint32_t area = (uint32_t)((px - mx) * (py - my));  // :211
  if (area != 1u) {
    const float r = __frcp_rn((float)area);
    s0 = udiv_exact(s0, area, r);
    s1 = udiv_exact(s1, area, r);
    s2 = udiv_exact(s2, area, r);
  }
  // `.xyz =` store: byte 3 of the uchar4 keeps its previous value (:212).
  uint32_t *o = reinterpret_cast<uint32_t *>(a.out + (size_t)f * a.out_stride) +
                (size_t)j * a.o_linesize_px + i;
  const uint32_t old = *o;
  *o = (old & 0xff000000u) | (s0 & 0xffu) | ((s1 & 0xffu) << 8) | ((s2 & 0xffu) << 16);
}

// ---------------------------------------------------------------------------------------
// interpolate_rect: inverse warp; one thread per 4 consecutive full-resolution pixels.
// ---------------------------------------------------------------------------------------
struct InterpArgs {
  uint8_t *out;
  const uint8_t *red;
  const InterpEntry *lx, *ly;
  size_t out_stride, red_stride;
  int W, H, ow, oh;
};

struct AxisSel {
  int lo, hi, exact_idx;
  float ratio;
  bool exact;
};

// Applies the gaze-dependent border fix-ups (sat_decoder_interpolate_kernel.cl:105-116) to a
// table entry and clamps the reduced-buffer indices (:118-133).
__device__ __forceinline__ AxisSel resolve_axis(const InterpEntry e, int centre, int n_full,
                                                int n_red, bool wrapped) {
  int min_u = e.min_u, max_u = e.max_u;
  const int lo = centre + e.rel_lo, hi = centre + e.rel_hi;
  if (lo < 0 && !wrapped) min_u = max_u;
  if (hi >= n_full && !wrapped) max_u = min_u;
  AxisSel s;
  s.lo = clampi(min_u + n_red / 2, 0, n_red - 1);
  s.hi = clampi(max_u + n_red / 2, 0, n_red - 1);
  s.exact_idx = e.idx_exact;
  s.ratio = e.ratio;
  s.exact = e.exact != 0;
  return s;
}

__device__ __forceinline__ InterpEntry load_entry(const InterpEntry *p) {
  const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
  InterpEntry e;
  e.idx_exact = (int16_t)(v.x & 0xffffu);
  e.min_u = (int16_t)(v.x >> 16);
  e.max_u = (int16_t)(v.y & 0xffffu);
  e.exact = (int16_t)(v.y >> 16);
  e.rel_lo = (int16_t)(v.z & 0xffffu);
  e.rel_hi = (int16_t)(v.z >> 16);
  e.ratio = __uint_as_float(v.w);
  return e;
}

// mix(a, b, t) = a + (b - a) * t with every operation rounded separately (no FMA), matching
// the oracle's scalar float arithmetic (:143-150).
__device__ __forceinline__ float mix_rn(float a, float b, float t) {
  return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), t));
}

__device__ __forceinline__ uint32_t lerp_pixel(uint32_t tl, uint32_t tr, uint32_t bl, uint32_t br,
                                               float tx, float ty) {
  uint32_t outp = 0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float ftl = (float)((tl >> (8 * c)) & 0xffu);
    const float ftr = (float)((tr >> (8 * c)) & 0xffu);
    const float fbl = (float)((bl >> (8 * c)) & 0xffu);
    const float fbr = (float)((br >> (8 * c)) & 0xffu);
    const float l = mix_rn(ftl, fbl, ty);
    const float r = mix_rn(ftr, fbr, ty);
    const int v = __float2int_rz(mix_rn(l, r, tx));
    outp |= ((uint32_t)v & 0xffu) << (8 * c);
  }
  return outp;  // byte 3 = 0 (convert_uchar3 result)
}

constexpr int kInterpPx = 4;

__global__ void __launch_bounds__(256) sat_interpolate_rect_kernel(const InterpArgs a,
                                                                   const GazeBatch g) {
  const int x4 = (blockIdx.x * 32 + threadIdx.x) * kInterpPx;
  const int y = blockIdx.y * 8 + threadIdx.y;
  const int f = blockIdx.z;
  if (x4 >= a.W || y >= a.H) return;
  const int W = a.W, H = a.H, ow = a.ow, oh = a.oh;
  const int cxp = gaze_px(g.xy[2 * f], W);
  const int cyp = gaze_px(g.xy[2 * f + 1], H);
  const uint32_t *red =
      reinterpret_cast<const uint32_t *>(a.red + (size_t)f * a.red_stride);

  const int dy = clampi(y - cyp, -H, H);
  const AxisSel sy = resolve_axis(load_entry(a.ly + (dy + H)), cyp, H, oh, false);
  const uint32_t *row_lo = red + (size_t)sy.lo * ow;
  const uint32_t *row_hi = red + (size_t)sy.hi * ow;
  const uint32_t *row_ex = red + (size_t)sy.exact_idx * ow;

  uint32_t px[kInterpPx];
#pragma unroll
  for (int k = 0; k < kInterpPx; ++k) {
    int x = x4 + k;
    px[k] = 0;
    if (x < W) {
      bool wrapped = false;  // :26-33
      if (x - cxp > W / 2) {
        x -= W;
        wrapped = true;
      } else if (x - cxp < -(W / 2)) {
        x += W;
        wrapped = true;
      }
      const int dx = clampi(x - cxp, -W, W);
      const AxisSel sx = resolve_axis(load_entry(a.lx + (dx + W)), cxp, W, ow, wrapped);
      if (sx.exact && sy.exact) {  // :67-72
        px[k] = __ldg(row_ex + sx.exact_idx);
      } else {
        px[k] = lerp_pixel(__ldg(row_lo + sx.lo), __ldg(row_lo + sx.hi), __ldg(row_hi + sx.lo),
                           __ldg(row_hi + sx.hi), sx.ratio, sy.ratio);
      }
    }
  }
  uint32_t *o = reinterpret_cast<uint32_t *>(a.out + (size_t)f * a.out_stride) + (size_t)y * W + x4;
  if (x4 + kInterpPx <= W && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
    __stcs(reinterpret_cast<uint4 *>(o), make_uint4(px[0], px[1], px[2], px[3]));
  } else {
#pragma unroll
    for (int k = 0; k < kInterpPx; ++k)
      if (x4 + k < W) o[k] = px[k];
  }
}

// ---------------------------------------------------------------------------------------
// decode: 1x1 boxes, exact inverse of the SAT.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sat_decode_kernel(uint8_t *out, int out_linesize, int bpp,
                                                         const uint32_t *sat, int W, int H) {
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x >= W || y >= H) return;
  const size_t row = (size_t)3 * W;
  const uint32_t *br = sat + (size_t)y * row + (size_t)3 * x;
  uint8_t *o = out + (size_t)y * out_linesize + (size_t)x * bpp;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    uint32_t v;
    if (x > 0 && y > 0)
      v = br[c] - br[c - (ptrdiff_t)row] + br[c - (ptrdiff_t)row - 3] - br[c - 3];  // :21-33
    else if (x > 0)
      v = br[c] - br[c - 3];  // :34-42
    else if (y > 0)
      v = br[c] - br[c - (ptrdiff_t)row];  // :43-51
    else
      v = sat[c];  // :52-57
    o[c] = (uint8_t)min(v, 255u);
  }
}

}  // namespace

cudaError_t launch_sat_sample_rect(const LaunchCtx &lc, int n, uint8_t *out, size_t out_stride, int ow,
                                   int oh, int out_linesize, const uint32_t *sat,
                                   size_t sat_stride, int W, int H, const int16_t *xedge,
                                   const int16_t *yedge, const GazeBatch &gaze) {
  SampleArgs a;
  a.out = out;
  a.sat = sat;
  a.xedge = xedge;
  a.yedge = yedge;
  a.out_stride = out_stride;
  a.sat_stride = sat_stride;
  a.ow = ow;
  a.oh = oh;
  a.o_linesize_px = out_linesize / 4;  // :153
  a.W = W;
  a.H = H;
  const dim3 grid((ow + 31) / 32, (oh + 7) / 8, n), block(32, 8);
  KernelScope ks(lc, "sat_sample_rect");
  sat_sample_rect_kernel<<<grid, block, 0, lc.stream>>>(a, gaze);
  return cudaGetLastError();
}

cudaError_t launch_sat_interpolate_rect(const LaunchCtx &lc, int n, uint8_t *out, size_t out_stride,
                                        int W, int H, const uint8_t *red, size_t red_stride,
                                        int ow, int oh, const InterpEntry *lx,
                                        const InterpEntry *ly, const GazeBatch &gaze) {
  InterpArgs a;
  a.out = out;
  a.red = red;
  a.lx = lx;
  a.ly = ly;
  a.out_stride = out_stride;
  a.red_stride = red_stride;
  a.W = W;
  a.H = H;
  a.ow = ow;
  a.oh = oh;
  const dim3 grid((W + 32 * kInterpPx - 1) / (32 * kInterpPx), (H + 7) / 8, n), block(32, 8);
  KernelScope ks(lc, "sat_interpolate_rect");
  sat_interpolate_rect_kernel<<<grid, block, 0, lc.stream>>>(a, gaze);
  return cudaGetLastError();
}

cudaError_t launch_sat_decode(const LaunchCtx &lc, uint8_t *out, int out_linesize, const uint32_t *sat,
                              int W, int H) {
  const dim3 grid((W + 31) / 32, (H + 7) / 8), block(32, 8);
  KernelScope ks(lc, "sat_decode");
  sat_decode_kernel<<<grid, block, 0, lc.stream>>>(out, out_linesize, out_linesize / W, sat, W, H);
  return cudaGetLastError();
}

}  // namespace fov
