// SATDecoder kernels for sm_100a: gaze-dependent box sampling from the SAT, the inverse
// log-rectilinear warp back to full resolution, and the exact 1x1 SAT->image decode.
//
// Replaces sample_rect_kernel (sat_decoder_sample_rect_kernel.cl:138-241),
// interpolate_rect_kernel (sat_decoder_interpolate_kernel.cl:1-152) and decode_kernel
// (sat_decoder_decode_kernel.cl:1-58).  None of the three evaluates a transcendental on the
// device: the separable, gaze-independent parts of the transform come from host-built 1-D
// tables (luts.cc), so the kernels are integer/gather work bounded by HBM/L2 bandwidth.
#include <cstdlib>

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
// sample_rect: box average from 4 SAT corners per reduced pixel.
//
// Neighbouring boxes share corners: the right edge of pixel i is the left edge of pixel i+1 and
// the bottom edge of row j is the top edge of row j+1 (both come from the same edge-table entry),
// except where the reference's clamps / seam wrap move one of them (frame borders).  A warp
// covers 31 pixel columns x kSampleRows rows: every lane gathers only its LEFT corner of the new
// bottom edge (12 B), takes the right corner from lane+1 by shuffle and reuses the previous
// bottom edge as its top edge - 3 gathered words per pixel instead of 12.  Lane 31 only supplies
// the right edge of lane 30.  Wherever a shared coordinate differs from the one the reference
// would use, the lane falls back to gathering it itself, so results stay bit-exact.
// ---------------------------------------------------------------------------------------
struct SampleArgs {
  uint8_t *out;
  const uint32_t *sat;
  const int16_t *xedge, *yedge;
  size_t out_stride, sat_stride;
  int ow, oh, o_linesize_px, W, H;
};

constexpr int kSampleCols = 31;

struct Rgb32 {
  uint32_t r, g, b;
};

__device__ __forceinline__ Rgb32 ld_sat(const uint32_t *p) {
  Rgb32 v;
  v.r = __ldg(p);
  v.g = __ldg(p + 1);
  v.b = __ldg(p + 2);
  return v;
}

__device__ __forceinline__ Rgb32 shfl_down1(const Rgb32 v) {
  Rgb32 o;
  o.r = __shfl_down_sync(0xffffffffu, v.r, 1);
  o.g = __shfl_down_sync(0xffffffffu, v.g, 1);
  o.b = __shfl_down_sync(0xffffffffu, v.b, 1);
  return o;
}

template <int kSampleRows>
__global__ void __launch_bounds__(256) sat_sample_rect_kernel(const SampleArgs a,
                                                              const GazeBatch g) {
  const int lane = threadIdx.x;
  const int i = blockIdx.x * kSampleCols + lane;
  const int j0 = (blockIdx.y * 8 + threadIdx.y) * kSampleRows;
  const int f = blockIdx.z;
  if (j0 >= a.oh) return;  // warp-uniform
  const int W = a.W, H = a.H, ow = a.ow, oh = a.oh;
  const int cxp = gaze_px(g.xy[2 * f], W);
  const int cyp = gaze_px(g.xy[2 * f + 1], H);
  const bool has_px = lane < kSampleCols && i < ow;

  // x edges of this lane's pixel (lanes past the last pixel compute harmless in-range values)
  const int ic = min(i, ow - 1);
  int px = cxp + a.xedge[ic + 1];  // grid[j+1][i+1].x  (:168-169)
  int mx = cxp + a.xedge[ic];      // grid[j+1][i].x    (:170-171)
  if (px >= W && mx >= W) {        // :181-187
    px -= W;
    mx -= W;
  } else if (px < 0 && mx < 0) {
    px += W;
    mx += W;
  }
  const bool x_in = (px >= 0 && px < W) || (mx >= 0 && mx < W);  // :197-198
  px = clampi(px, 1, W - 1);                                      // :201, :203
  mx = clampi(mx, 0, px - 1);
  // lane+1 gathers its own left corner at mx(lane+1); usable as this lane's right corner iff equal
  const int mx_next = __shfl_down_sync(0xffffffffu, mx, 1);
  const bool share = lane < 31 && i + 1 < ow && px == mx_next;
  const bool own_right = has_px && !share;

  const uint32_t *sat =
      reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(a.sat) +
                                         (size_t)f * a.sat_stride);
  const uint32_t *colL = sat + (size_t)mx * 3;
  const uint32_t *colR = sat + (size_t)px * 3;
  const size_t row_words = (size_t)W * 3;

  // Gather the left corner of all kSampleRows+1 horizontal edges up front (27 loads in flight
  // per lane).  Edge e is the bottom edge of row e-1 and, away from the frame border, the top
  // edge of row e; it is fetched at the bottom-edge form of its coordinate (:202).
  int yraw[kSampleRows + 1];
  Rgb32 L[kSampleRows + 1];
#pragma unroll
  for (int e = 0; e <= kSampleRows; ++e) {
    yraw[e] = cyp + a.yedge[min(j0 + e, oh)];
    const int ye = (e == 0) ? clampi(yraw[0], 0, H - 2) : clampi(yraw[e], 1, H - 1);
    L[e] = ld_sat(colL + (size_t)ye * row_words);
  }

  // The store keeps byte 3 of the target pixel (`.xyz =`, :212): fetch the old pixels now too.
  uint32_t *orow = reinterpret_cast<uint32_t *>(a.out + (size_t)f * a.out_stride) +
                   (size_t)j0 * a.o_linesize_px + i;
  uint32_t old[kSampleRows];
#pragma unroll
  for (int r = 0; r < kSampleRows; ++r)
    old[r] = (has_px && x_in && j0 + r < oh) ? orow[(size_t)r * a.o_linesize_px] : 0u;

#pragma unroll
  for (int r = 0; r < kSampleRows; ++r, orow += a.o_linesize_px) {
    if (j0 + r >= oh) break;  // warp-uniform
    int py = yraw[r + 1];     // grid[j+1][i+1].y  (:172-173)
    int my = yraw[r];         // grid[j][i+1].y    (:174-175)
    const bool y_in = (py >= 0 && py < H) || (my >= 0 && my < H);  // :199-200
    if (!y_in) continue;        // warp-uniform: the whole row keeps its contents
    py = clampi(py, 1, H - 1);  // :202, :204  (== the coordinate L[r+1] was fetched at)
    my = clampi(my, 0, py - 1);
    const int fetched_top = (r == 0) ? clampi(yraw[0], 0, H - 2) : clampi(yraw[r], 1, H - 1);
    Rgb32 tl = L[r];
    if (my != fetched_top) tl = ld_sat(colL + (size_t)my * row_words);  // warp-uniform, borders
    const Rgb32 bl = L[r + 1];
    Rgb32 tr = shfl_down1(tl);
    Rgb32 br = shfl_down1(bl);
    if (own_right) {  // seam / border lanes and the last column gather their own right corners
      tr = ld_sat(colR + (size_t)my * row_words);
      br = ld_sat(colR + (size_t)py * row_words);
    }
    if (has_px && x_in) {
      uint32_t s0 = br.r - tr.r + tl.r - bl.r;  // :212-217
      uint32_t s1 = br.g - tr.g + tl.g - bl.g;
      uint32_t s2 = br.b - tr.b + tl.b - bl.b;
      const uint32_t area = (uint32_t)((px - mx) * (py - my));  // :211
      if (area != 1u) {
        const float rcp = __frcp_rn((float)area);
        s0 = udiv_exact(s0, area, rcp);
        s1 = udiv_exact(s1, area, rcp);
        s2 = udiv_exact(s2, area, rcp);
      }
      // `.xyz =` store: byte 3 of the uchar4 keeps its previous value (:212).
      *orow = (old[r] & 0xff000000u) | (s0 & 0xffu) | ((s1 & 0xffu) << 8) | ((s2 & 0xffu) << 16);
    }
  }
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

// u8 -> float without the conversion pipe: splice the byte into the mantissa of 2^23 (one PRMT)
// and subtract 2^23 (one FADD); exact for 0..255.
template <int C>
__device__ __forceinline__ float byte_to_float(uint32_t v) {
  return __fsub_rn(__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7440u | C)), 8388608.0f);
}

// mix(a, b, t) = a + (b - a) * t with every operation rounded separately (no FMA), matching
// the oracle's scalar float arithmetic (:143-150).
__device__ __forceinline__ float mix_rn(float a, float b, float t) {
#ifdef FOV360_FUSED_LERP
  return __fmaf_rn(__fsub_rn(b, a), t, a);  // <= 1 LSB after truncation, not bit-exact
#else
  return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), t));
#endif
}

// trunc(v) for v in [0, 256): the low mantissa bits of v + 2^23 rounded toward zero
// (convert_uchar3, :150).  Byte 1 of the result is 0, which pack_rgb0 uses as the padding byte.
__device__ __forceinline__ uint32_t trunc_bits(float v) {
  return __float_as_uint(__fadd_rz(v, 8388608.0f));
}

__device__ __forceinline__ uint32_t pack_rgb0(uint32_t c0, uint32_t c1, uint32_t c2) {
  return __byte_perm(__byte_perm(c0, c1, 0x0040u), c2, 0x5410u);  // c0.b0, c1.b0, c2.b0, 0
}

constexpr int kInterpPx = 4;      // consecutive pixels per lane: one 16-byte store per row
constexpr int kInterpRows = 32;   // consecutive rows per warp: the x-axis work is done once
constexpr int kInterpMaxCols = 136;  // widest reduced-column window a warp stages per row

// Row descriptor, resolved by lane r for row y0 + r and broadcast by shuffle.
struct RowSel {
  int rows;   // first reduced row | second reduced row << 16 (equal when ty is 0 or 1)
  float ty;   // vertical ratio
  int yex;    // exact-hit reduced row, or -1 when this row is not an exact hit
};

template <int C>
__device__ __forceinline__ float vmix_channel(uint32_t a, uint32_t b, float ty) {
  return mix_rn(byte_to_float<C>(a), byte_to_float<C>(b), ty);
}

// interpolate_rect, one warp = 128 columns x kInterpRows rows.
//
// Everything that depends on x only (table entry, wrap, border fix-ups, clamped reduced columns,
// ratio) is resolved once per lane; the kInterpRows y entries are resolved by one lane each.
// The bilinear tap is separable exactly as the reference evaluates it - mix vertically at the two
// columns, then mix horizontally - so per row a warp first forms the vertical mixes V[c] for the
// window of reduced columns its 128 pixels touch (one per column, not two per pixel; in the
// periphery a column serves ~4.6 pixels) in shared memory as floats, then every pixel is one
// horizontal mix of two staged values.  A ratio of exactly 0 or 1 makes mix() return one operand
// unchanged, so 1:1 (foveal) columns/rows are handled by selecting that operand: bit-identical,
// no arithmetic.  Warps that are entirely inside the 1:1 column band skip the staging.
__global__ void __launch_bounds__(256, 4) sat_interpolate_rect_kernel(const InterpArgs a,
                                                                      const GazeBatch g) {
  __shared__ float4 vstage[8][2][kInterpMaxCols];
  const int lane = threadIdx.x, warp = threadIdx.y;
  const int x4 = (blockIdx.x * 32 + lane) * kInterpPx;
  const int y0 = (blockIdx.y * 8 + warp) * kInterpRows;
  const int f = blockIdx.z;
  if (y0 >= a.H) return;  // warp-uniform
  const int W = a.W, H = a.H, ow = a.ow, oh = a.oh;
  const int cxp = gaze_px(g.xy[2 * f], W);
  const int cyp = gaze_px(g.xy[2 * f + 1], H);
  const uint32_t *red = reinterpret_cast<const uint32_t *>(a.red + (size_t)f * a.red_stride);

  // ---- x axis: once per lane ------------------------------------------------------------
  int xlo[kInterpPx], xhi[kInterpPx], xex[kInterpPx];
  float xr[kInterpPx];
  bool all_deg = true;
  int cmin = 0x7fffffff, cmax = -1;
#pragma unroll
  for (int k = 0; k < kInterpPx; ++k) {
    int x = min(x4 + k, W - 1);  // lanes past the right edge repeat the last pixel (never stored)
    bool wrapped = false;        // :26-33
    if (x - cxp > W / 2) {
      x -= W;
      wrapped = true;
    } else if (x - cxp < -(W / 2)) {
      x += W;
      wrapped = true;
    }
    const int dx = clampi(x - cxp, -W, W);
    const AxisSel sx = resolve_axis(load_entry(a.lx + (dx + W)), cxp, W, ow, wrapped);
    const bool deg = sx.ratio == 0.0f || sx.ratio == 1.0f;
    const int sel = sx.ratio == 1.0f ? sx.hi : sx.lo;
    xlo[k] = deg ? sel : sx.lo;  // degenerate: both taps are the selected column
    xhi[k] = deg ? sel : sx.hi;
    xr[k] = sx.ratio;
    xex[k] = sx.exact ? sx.exact_idx : -1;
    all_deg = all_deg && deg;
    cmin = min(cmin, min(xlo[k], xhi[k]));
    cmax = max(cmax, max(xlo[k], xhi[k]));
  }
  all_deg = __all_sync(0xffffffffu, all_deg);
  cmin = __reduce_min_sync(0xffffffffu, cmin);
  cmax = __reduce_max_sync(0xffffffffu, cmax);
  const int ncols = cmax - cmin + 1;

  // ---- y axis: lane r resolves row y0 + r --------------------------------------------------
  RowSel mine;
  {
    const int y = min(y0 + (lane & (kInterpRows - 1)), H - 1);
    const int dy = clampi(y - cyp, -H, H);
    const AxisSel sy = resolve_axis(load_entry(a.ly + (dy + H)), cyp, H, oh, false);
    const bool deg = sy.ratio == 0.0f || sy.ratio == 1.0f;
    const int sel = sy.ratio == 1.0f ? sy.hi : sy.lo;
    mine.rows = deg ? (sel | (sel << 16)) : (sy.lo | (sy.hi << 16));
    mine.ty = sy.ratio;
    mine.yex = sy.exact ? sy.exact_idx : -1;
  }

  uint32_t *orow = reinterpret_cast<uint32_t *>(a.out + (size_t)f * a.out_stride) +
                   (size_t)y0 * W + x4;
  const bool in_x = x4 < W;
  const bool vec_ok = (x4 + kInterpPx <= W) && ((reinterpret_cast<uintptr_t>(orow) & 15) == 0) &&
                      ((W & 3) == 0);
  const int nrows = min(kInterpRows, H - y0);

  auto store_row = [&](const uint32_t (&px)[kInterpPx]) {
    if (vec_ok) {
      __stcs(reinterpret_cast<uint4 *>(orow), make_uint4(px[0], px[1], px[2], px[3]));
    } else if (in_x) {
#pragma unroll
      for (int k = 0; k < kInterpPx; ++k)
        if (x4 + k < W) orow[k] = px[k];
    }
  };

  if (all_deg) {
    // Every pixel of this warp maps onto a single reduced column: vertical mix (or copy) only.
    for (int r = 0; r < nrows; ++r, orow += W) {
      const int rows = __shfl_sync(0xffffffffu, mine.rows, r);
      const float ty = __shfl_sync(0xffffffffu, mine.ty, r);
      const int yex = __shfl_sync(0xffffffffu, mine.yex, r);
      const uint32_t *ra = red + (size_t)(rows & 0xffff) * ow;
      const uint32_t *rb = red + (size_t)(rows >> 16) * ow;
      uint32_t px[kInterpPx];
      if ((rows & 0xffff) == (rows >> 16)) {  // warp-uniform: the row is a copy of a reduced row
        const uint32_t *rex = red + (size_t)max(yex, 0) * ow;
#pragma unroll
        for (int k = 0; k < kInterpPx; ++k)
          px[k] = (yex >= 0 && xex[k] >= 0) ? __ldg(rex + xex[k])  // :67-72: all 4 bytes
                                            : (__ldg(ra + xlo[k]) & 0x00ffffffu);
      } else {
#pragma unroll
        for (int k = 0; k < kInterpPx; ++k) {
          const uint32_t p = __ldg(ra + xlo[k]), q = __ldg(rb + xlo[k]);
          px[k] = pack_rgb0(trunc_bits(vmix_channel<0>(p, q, ty)),
                            trunc_bits(vmix_channel<1>(p, q, ty)),
                            trunc_bits(vmix_channel<2>(p, q, ty)));
        }
      }
      store_row(px);
    }
    return;
  }

  if (ncols <= kInterpMaxCols) {
    // Stage the vertical mixes of the column window, then one horizontal mix per pixel.
    int olo[kInterpPx], ohi[kInterpPx];
#pragma unroll
    for (int k = 0; k < kInterpPx; ++k) {
      olo[k] = xlo[k] - cmin;
      ohi[k] = xhi[k] - cmin;
    }
    for (int r = 0; r < nrows; ++r, orow += W) {
      const int rows = __shfl_sync(0xffffffffu, mine.rows, r);
      const float ty = __shfl_sync(0xffffffffu, mine.ty, r);
      const int yex = __shfl_sync(0xffffffffu, mine.yex, r);
      const uint32_t *ra = red + (size_t)(rows & 0xffff) * ow + cmin;
      const uint32_t *rb = red + (size_t)(rows >> 16) * ow + cmin;
      float4 *vs = vstage[warp][r & 1];
      if ((rows & 0xffff) == (rows >> 16)) {  // warp-uniform
        for (int c = lane; c < ncols; c += 32) {
          const uint32_t p = __ldg(ra + c);
          vs[c] = make_float4(byte_to_float<0>(p), byte_to_float<1>(p), byte_to_float<2>(p), 0.f);
        }
      } else {
        for (int c = lane; c < ncols; c += 32) {
          const uint32_t p = __ldg(ra + c), q = __ldg(rb + c);
          vs[c] = make_float4(vmix_channel<0>(p, q, ty), vmix_channel<1>(p, q, ty),
                              vmix_channel<2>(p, q, ty), 0.f);
        }
      }
      __syncwarp();
      uint32_t px[kInterpPx];
#pragma unroll
      for (int k = 0; k < kInterpPx; ++k) {
        const float4 l = vs[olo[k]], rr = vs[ohi[k]];
        px[k] = pack_rgb0(trunc_bits(mix_rn(l.x, rr.x, xr[k])), trunc_bits(mix_rn(l.y, rr.y, xr[k])),
                          trunc_bits(mix_rn(l.z, rr.z, xr[k])));
      }
      if (yex >= 0) {  // warp-uniform: pixels that hit a sample on both axes copy all 4 bytes
        const uint32_t *rex = red + (size_t)yex * ow;
#pragma unroll
        for (int k = 0; k < kInterpPx; ++k)
          if (xex[k] >= 0) px[k] = __ldg(rex + xex[k]);
      }
      store_row(px);
    }
    return;
  }

  // Seam warps (the window spans the whole reduced width): gather every tap directly.
  for (int r = 0; r < nrows; ++r, orow += W) {
    const int rows = __shfl_sync(0xffffffffu, mine.rows, r);
    const float ty = __shfl_sync(0xffffffffu, mine.ty, r);
    const int yex = __shfl_sync(0xffffffffu, mine.yex, r);
    const uint32_t *ra = red + (size_t)(rows & 0xffff) * ow;
    const uint32_t *rb = red + (size_t)(rows >> 16) * ow;
    const uint32_t *rex = red + (size_t)max(yex, 0) * ow;
    uint32_t px[kInterpPx];
#pragma unroll
    for (int k = 0; k < kInterpPx; ++k) {
      if (yex >= 0 && xex[k] >= 0) {
        px[k] = __ldg(rex + xex[k]);
      } else {
        const uint32_t tl = __ldg(ra + xlo[k]), tr = __ldg(ra + xhi[k]);
        const uint32_t bl = __ldg(rb + xlo[k]), br = __ldg(rb + xhi[k]);
        px[k] = pack_rgb0(
            trunc_bits(mix_rn(vmix_channel<0>(tl, bl, ty), vmix_channel<0>(tr, br, ty), xr[k])),
            trunc_bits(mix_rn(vmix_channel<1>(tl, bl, ty), vmix_channel<1>(tr, br, ty), xr[k])),
            trunc_bits(mix_rn(vmix_channel<2>(tl, bl, ty), vmix_channel<2>(tr, br, ty), xr[k])));
      }
    }
    store_row(px);
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
  static const int rows_per_warp = [] {
    const char *e = getenv("FOV360_SAMPLE_ROWS");
    return e ? atoi(e) : 4;
  }();
  const int rpw = rows_per_warp == 8 ? 8 : 4;
  const dim3 grid((ow + kSampleCols - 1) / kSampleCols, (oh + 8 * rpw - 1) / (8 * rpw), n),
      block(32, 8);
  KernelScope ks(lc, "sat_sample_rect");
  if (rpw == 8)
    sat_sample_rect_kernel<8><<<grid, block, 0, lc.stream>>>(a, gaze);
  else
    sat_sample_rect_kernel<4><<<grid, block, 0, lc.stream>>>(a, gaze);
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
  const dim3 grid((W + 32 * kInterpPx - 1) / (32 * kInterpPx),
                  (H + 8 * kInterpRows - 1) / (8 * kInterpRows), n),
      block(32, 8);
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
