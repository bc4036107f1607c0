// SATDecoder kernels for sm_100a: gaze-dependent box sampling from the SAT, the inverse
// log-rectilinear warp back to full resolution, and the exact 1x1 SAT->image decode.
//
// Replaces sample_rect_kernel (sat_decoder_sample_rect_kernel.cl:138-241),
// interpolate_rect_kernel (sat_decoder_interpolate_kernel.cl:1-152) and decode_kernel
// (sat_decoder_decode_kernel.cl:1-58).  None of the three evaluates a transcendental on the
// device: the separable, gaze-independent parts of the transform come from host-built 1-D
// tables (luts.cc), so the kernels are integer/gather work bounded by HBM/L2 bandwidth.
#include <cstdlib>

#include "bounds_check.cuh"
#include "fov360_internal.h"
#include "pixel_math.cuh"
#include "projection_common.cuh"

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
  // Optional (fused encode+sample only): the RGB0 frames the SATs were just built from.  A 1x1 box
  // sums to the source pixel itself, so the 1:1 region around the gaze reads 4 B per pixel from
  // the frame instead of ~15 B of SAT corners - same bits by construction.
  const uint32_t *src;
  size_t src_stride;
  int src_linesize_px;
};

constexpr int kSampleCols = 31;
constexpr int kSampleWarps = 8;

struct Rgb32 {
  uint32_t r, g, b;
};

// One 12-byte SAT entry at a 32-bit word offset (a frame's SAT has fewer than 2^32 words).
__device__ __forceinline__ Rgb32 ld_sat(const uint32_t *sat, uint32_t off) {
  const uint32_t *p = sat + off;
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

// Everything a reduced row contributes to its boxes, resolved once per CTA: the clamped SAT rows
// of its top and bottom edge as word offsets, their distance, and two flags.
struct __align__(16) SampleRow {
  uint32_t top_off, bot_off;  // my * W * 3, py * W * 3
  int py;                     // bottom SAT row
  int dyf;                    // (py - my) << 2 | flags
                              // 1: the row is sampled at all (:199-200, and j < oh)
                              // 2: its top edge is the bottom edge of the row above
};

#ifndef FOV360_SAMPLE_MIN_CTAS
#define FOV360_SAMPLE_MIN_CTAS 6
#endif
// kWholePx (fov_ctx_set_option FOV_OPT_REDUCED_PAD_ZERO): the caller promised that byte 3 of every
// reduced pixel is 0 and may stay 0, so a sampled pixel is ONE 32-bit store (r, g, b, 0) instead of
// the reference's 3-byte .xyz store - same buffer contents under the promise, and L2 no longer has
// to read-fill the sectors the partial stores touch (-7 % kernel time at 8K x 16).
template <int kSampleRows, bool kDevGaze, bool kWholePx>
__global__ void __launch_bounds__(32 * kSampleWarps, FOV360_SAMPLE_MIN_CTAS)
    sat_sample_rect_kernel(const SampleArgs a, const GazeBatch g) {
  __shared__ SampleRow srow[kSampleWarps * kSampleRows];
  pdl_trigger();
  const int lane = threadIdx.x, warp = threadIdx.y;
  const int i = blockIdx.x * kSampleCols + lane;
  const int jb = blockIdx.y * kSampleWarps * kSampleRows;
  const int f = blockIdx.z;
  const int W = a.W, H = a.H, ow = a.ow, oh = a.oh;
  // kDevGaze is a template parameter: the run-time form of this choice cost the by-value path 1 %.
  // A device-side gaze array may itself have been produced on this stream (a copy, or the caller's
  // own kernel), so those instantiations wait before they read it and overlap the launch only.
  if (kDevGaze) pdl_wait();
  const int cxp = gaze_px(kDevGaze ? __ldg(g.dev + 2 * f) : g.xy[2 * f], W);
  const int cyp = gaze_px(kDevGaze ? __ldg(g.dev + 2 * f + 1) : g.xy[2 * f + 1], H);
  const uint32_t row_words = (uint32_t)W * 3u;

  // ---- y edges of the CTA's rows: one thread per row ------------------------------------------
  {
    const int t = warp * 32 + lane;
    if (t < kSampleWarps * kSampleRows) {
      const int j = jb + t;
      const int y0 = cyp + a.yedge[min(j, oh)];      // grid[j][i+1].y    (:174-175)
      const int y1 = cyp + a.yedge[min(j + 1, oh)];  // grid[j+1][i+1].y  (:172-173)
      const bool y_in = (y1 >= 0 && y1 < H) || (y0 >= 0 && y0 < H);  // :199-200
      const int py = clampi(y1, 1, H - 1);                            // :202, :204
      const int my = clampi(y0, 0, py - 1);
      FOV_CHECK(py, H, 201);
      FOV_CHECK(my, H, 202);
      SampleRow d;
      d.top_off = (uint32_t)my * row_words;
      d.bot_off = (uint32_t)py * row_words;
      d.py = py;
      d.dyf = ((py - my) << 2) | ((y_in && j < oh) ? 1 : 0) | ((my == clampi(y0, 1, H - 1)) ? 2 : 0);
      srow[t] = d;
    }
  }
  __syncthreads();
  const int j0 = jb + warp * kSampleRows;
  if (j0 >= oh) return;  // warp-uniform
  const SampleRow *rows = srow + warp * kSampleRows;

  // ---- x edges of this lane's pixel (lanes past the last pixel compute harmless in-range values)
  const bool has_px = lane < kSampleCols && i < ow;
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
  const bool live = has_px && x_in;
  FOV_CHECK(ic + 1, ow + 1, 203);
  FOV_CHECK(px, W, 204);
  FOV_CHECK(mx, W, 205);
  const uint32_t colL = (uint32_t)mx * 3u, colR = (uint32_t)px * 3u;
  const uint32_t dx = (uint32_t)(px - mx);

  SampleRow d[kSampleRows];
#pragma unroll
  for (int r = 0; r < kSampleRows; ++r) d[r] = rows[r];
  uint32_t *orow = reinterpret_cast<uint32_t *>(a.out + (size_t)f * a.out_stride) +
                   (size_t)j0 * a.o_linesize_px + i;

  pdl_wait();  // everything above came from the edge tables; the SAT and the frames follow
  if (a.src != nullptr) {
    // 1:1 warps: every live box is one pixel wide and every sampled row one pixel high.
    bool unit = !live || dx == 1u;
#pragma unroll
    for (int r = 0; r < kSampleRows; ++r) unit = unit && (!(d[r].dyf & 1) || (d[r].dyf >> 2) == 1);
    if (__all_sync(0xffffffffu, unit)) {
      // box (mx, px] x (my, py] = the pixel (px, py): S(py,px) - S(my,px) - S(py,mx) + S(my,mx)
      const uint32_t *sp = reinterpret_cast<const uint32_t *>(
                               reinterpret_cast<const uint8_t *>(a.src) + (size_t)f * a.src_stride) + px;
      uint32_t v[kSampleRows];
#pragma unroll
      for (int r = 0; r < kSampleRows; ++r) {
        const bool on = live && (d[r].dyf & 1);
        v[r] = on ? __ldg(sp + (size_t)d[r].py * a.src_linesize_px) : 0u;
      }
#pragma unroll
      for (int r = 0; r < kSampleRows; ++r)
        if (live && (d[r].dyf & 1)) {
          if (kWholePx) orow[(size_t)r * a.o_linesize_px] = v[r] & 0x00ffffffu;
          else store_xyz(orow + (size_t)r * a.o_linesize_px, v[r]);
        }
      return;
    }
  }

  const uint32_t *sat =
      reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(a.sat) +
                                         (size_t)f * a.sat_stride);

  // Gather the left corner of all kSampleRows+1 horizontal edges up front.  Edge r+1 is the bottom
  // edge of row r and, away from the frame border, the top edge of row r+1.
  Rgb32 L[kSampleRows + 1];
  L[0] = ld_sat(sat, d[0].top_off + colL);
#pragma unroll
  for (int r = 0; r < kSampleRows; ++r) L[r + 1] = ld_sat(sat, d[r].bot_off + colL);

#pragma unroll
  for (int r = 0; r < kSampleRows; ++r, orow += a.o_linesize_px) {
    if (!(d[r].dyf & 1)) continue;  // warp-uniform: the whole row keeps its contents
    Rgb32 tl = L[r];
    if (r > 0 && !(d[r].dyf & 2)) tl = ld_sat(sat, d[r].top_off + colL);  // warp-uniform, borders
    const Rgb32 bl = L[r + 1];
    Rgb32 tr = shfl_down1(tl);
    Rgb32 br = shfl_down1(bl);
    if (own_right) {  // seam / border lanes and the last column gather their own right corners
      tr = ld_sat(sat, d[r].top_off + colR);
      br = ld_sat(sat, d[r].bot_off + colR);
    }
    if (live) {
      uint32_t s0 = br.r - tr.r + tl.r - bl.r;  // :212-217
      uint32_t s1 = br.g - tr.g + tl.g - bl.g;
      uint32_t s2 = br.b - tr.b + tl.b - bl.b;
      const uint32_t area = dx * (uint32_t)(d[r].dyf >> 2);  // :211
      if (area != 1u) {
        float rcp;  // ~1 ulp is plenty: udiv_exact corrects the quotient by one either way
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"((float)area));
        s0 = udiv_exact(s0, area, rcp);
        s1 = udiv_exact(s1, area, rcp);
        s2 = udiv_exact(s2, area, rcp);
      }
      const uint32_t rgb = (s0 & 0xffu) | ((s1 & 0xffu) << 8) | ((s2 & 0xffu) << 16);
      if (kWholePx) *orow = rgb;
      else store_xyz(orow, rgb);
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

constexpr int kInterpPx = 4;         // consecutive pixels per lane: one 16-byte store per row
// Rows per warp are a template parameter of the kernel (32, 16 or 8): the x-axis work is done once
// per warp, so large batches want tall tiles, while a single frame is less than one wave of CTAs
// and finishes sooner with more, shorter ones.
constexpr int kInterpMaxCols = 136;  // widest reduced-column window the generic path stages
#ifndef FOV360_INTERP_CHUNK
#define FOV360_INTERP_CHUNK 4
#endif
constexpr int kInterpChunk = FOV360_INTERP_CHUNK;  // rows per pass of the periphery path
constexpr int kInterpWarps = 4;      // warps per CTA, stacked vertically
constexpr int kCopyRows = 4;         // copy rows requested per batch: 6 or 8 spill and are slower
#ifndef FOV360_INTERP_MIN_CTAS
#define FOV360_INTERP_MIN_CTAS 5
#endif

// Row descriptor, resolved by one lane per row, read back as one shared-memory broadcast.
struct __align__(16) RowSel {
  int off_lo, off_hi;  // word offsets of the two reduced rows (equal when ty is 0 or 1)
  float ty;            // vertical ratio
  int info;            // exact-hit reduced row (or -1) << 1 | "row pair differs from the row above"
};

template <int C>
__device__ __forceinline__ float vmix_channel(uint32_t a, uint32_t b, float ty) {
  return mix_rn(byte_to_float<C>(a), byte_to_float<C>(b), ty);
}

// One reduced-buffer tap pair of a column as floats: p and (q - p), the two operands of the
// vertical mix that do not depend on the output row.
struct TapPair {
  float p[3], d[3];
};

__device__ __forceinline__ void convert_tap_pair(TapPair &t, uint32_t p, uint32_t q) {
  t.p[0] = byte_to_float<0>(p), t.p[1] = byte_to_float<1>(p), t.p[2] = byte_to_float<2>(p);
  t.d[0] = __fsub_rn(byte_to_float<0>(q), t.p[0]);
  t.d[1] = __fsub_rn(byte_to_float<1>(q), t.p[1]);
  t.d[2] = __fsub_rn(byte_to_float<2>(q), t.p[2]);
}

// mix(p, q, t) = p + (q - p) * t with every operation rounded separately (no FMA), matching the
// oracle's scalar float arithmetic (:143-150); d = q - p.  With q == p (or t == 0) it returns p.
__device__ __forceinline__ float lerp_rn(float p, float d, float t) {
#ifdef FOV360_FUSED_LERP
  return __fmaf_rn(d, t, p);  // <= 1 LSB after truncation, not bit-exact
#else
  return __fadd_rn(p, __fmul_rn(d, t));
#endif
}

__device__ __forceinline__ uint32_t lerp_px(const float4 v, const float4 d, float t) {
  return pack_rgb0(trunc_bits(lerp_rn(v.x, d.x, t)), trunc_bits(lerp_rn(v.y, d.y, t)),
                   trunc_bits(lerp_rn(v.z, d.z, t)));
}

// lerp_px on a staged column: v = {V.x, V.y | V.z, V.w}, d likewise, t2 = {t, t}.  The .w halves
// (the exact-hit samples' 4th bytes) ride through the arithmetic unused.
__device__ __forceinline__ uint32_t lerp_px2(const ulonglong2 v, const ulonglong2 d, f32x2 t2) {
  uint32_t c0, c1, c2, c3;
  unpack2(trunc_bits2(add2_rn(v.x, mul2_rn(d.x, t2))), c0, c1);
  unpack2(trunc_bits2(add2_rn(v.y, mul2_rn(d.y, t2))), c2, c3);
  return pack_rgb0(c0, c1, c2);
}

// One pixel of the periphery path from its staged column: colour from V + D * t, 4th byte from
// V.w (exact hit on the column itself) or D.w (on the column to the right).
__device__ __forceinline__ uint32_t staged_px(const ulonglong2 v, const ulonglong2 d, float t,
                                              uint32_t keep_lo, uint32_t keep_hi) {
  uint32_t vz, vw, dz, dw;
  unpack2(v.y, vz, vw);
  unpack2(d.y, dz, dw);
  return lerp_px2(v, d, pack2(t, t)) | (vw & keep_lo) | (dw & keep_hi);
}

// The same from a staged column v and its right neighbour n (D = n - v is formed here, bit for bit
// what pass 1 would have stored): 4th byte from v.w or n.w.
__device__ __forceinline__ uint32_t staged_px_vn(const ulonglong2 v, const ulonglong2 d,
                                                 const ulonglong2 n, float t, uint32_t keep_lo,
                                                 uint32_t keep_hi) {
  uint32_t vz, vw, nz, nw;
  unpack2(v.y, vz, vw);
  unpack2(n.y, nz, nw);
  return lerp_px2(v, d, pack2(t, t)) | (vw & keep_lo) | (nw & keep_hi);
}
__device__ __forceinline__ ulonglong2 sub_cols(const ulonglong2 n, const ulonglong2 v) {
  return make_ulonglong2(sub2_rn(n.x, v.x), sub2_rn(n.y, v.y));  // the .w halves are never used
}

// interpolate_rect, one warp = 128 columns x kInterpRows rows.
//
// Everything that depends on x only (table entry, wrap, border fix-ups, clamped reduced columns,
// ratio) is resolved once per lane; the y entries are resolved by one lane per row.  The bilinear
// tap is separable exactly as the reference evaluates it - mix vertically at the two columns, then
// mix horizontally.  A ratio of exactly 0 or 1 makes mix() return one operand unchanged, so those
// taps select that operand (both taps become the selected row / column): bit-identical.
//
// The log-rectilinear map is 1:1 around the gaze and 6-10x compressed outside, on each axis, so a
// warp falls into one of these cases:
//  * all 128 columns 1:1: no horizontal mix.  Rows that are 1:1 too are copies, loaded four rows
//    at a time; the others mix two reduced rows whose converted taps (p, q - p) stay in registers
//    while the row pair repeats (6-10 output rows).
//  * at most 32 reduced columns under the 128 pixels (the periphery): lane c owns window column c.
//    Per chunk of rows, pass 1 requests every tap of the chunk, forms the vertical mix V[c] and
//    D[c] = V[c+1] - V[c] in shared memory, pass 2 forms each pixel as V[lo] + D[lo] * tx.
//    Neither pass waits on global memory or synchronises inside.
//    A pixel that hits a sample on both axes is a copy of all 4 bytes of that sample (:67-72): its
//    colour is what the mixes select anyway, so V carries the sample's 4th byte along.
//  * anything else (the warp straddles the edge of the 1:1 band, an exact hit that is not the
//    selected tap, the +-W seam): generic row-by-row paths.
template <int kInterpRows, bool kDevGaze>
__global__ void __launch_bounds__(32 * kInterpWarps, FOV360_INTERP_MIN_CTAS)
    sat_interpolate_rect_kernel(const InterpArgs a, const GazeBatch g) {
  // staging area of a warp: kInterpChunk rows of V (34 slots each) in the periphery path, one row
  // of V and D (kInterpMaxCols each) in the wide-window path
  constexpr int kStage =
      kInterpChunk * 34 > 2 * kInterpMaxCols ? kInterpChunk * 34 : 2 * kInterpMaxCols;
  __shared__ float4 vstage[kInterpWarps][kStage];
  __shared__ RowSel rowsel[kInterpWarps][kInterpRows];
  __shared__ int4 xsel[kInterpPx][32];
  pdl_trigger();
  const int lane = threadIdx.x, warp = threadIdx.y;
  const int x4 = (blockIdx.x * 32 + lane) * kInterpPx;
  const int y0 = (blockIdx.y * kInterpWarps + warp) * kInterpRows;
  const int f = blockIdx.z;
  const int W = a.W, H = a.H, ow = a.ow, oh = a.oh;
  // kDevGaze is a template parameter: the run-time form of this choice cost the by-value path 1 %.
  // A device-side gaze array may itself have been produced on this stream (a copy, or the caller's
  // own kernel), so those instantiations wait before they read it and overlap the launch only.
  if (kDevGaze) pdl_wait();
  const int cxp = gaze_px(kDevGaze ? __ldg(g.dev + 2 * f) : g.xy[2 * f], W);
  const int cyp = gaze_px(kDevGaze ? __ldg(g.dev + 2 * f + 1) : g.xy[2 * f + 1], H);
  const uint32_t *red = reinterpret_cast<const uint32_t *>(a.red + (size_t)f * a.red_stride);

  // Both table entries a lane needs - the x entry of its pixel slot and the y entry of its row - are
  // requested before either is used: the prologue is two L2 latencies deep otherwise.
  static_assert(kInterpRows <= 32, "one lane per row of the warp's tile");
  const bool my_row = lane < kInterpRows;  // tiles shorter than a warp: the other lanes idle
  const int yrow = min(y0 + min(lane, kInterpRows - 1), H - 1);
  const InterpEntry ey = load_entry(a.ly + (clampi(yrow - cyp, -H, H) + H));

  // ---- x axis: resolved once per CTA (its warps cover the same 128 columns): warp k resolves
  // pixel k of every lane --------------------------------------------------------------------
  static_assert(kInterpWarps == kInterpPx, "one warp per pixel slot");
  {
    int x = min(x4 + warp, W - 1);  // lanes past the right edge repeat the last pixel (never stored)
    bool wrapped = false;           // :26-33
    if (x - cxp > W / 2) {
      x -= W;
      wrapped = true;
    } else if (x - cxp < -(W / 2)) {
      x += W;
      wrapped = true;
    }
    const int dx = clampi(x - cxp, -W, W);
    FOV_CHECK(dx + W, 2 * W + 1, 211);
    const AxisSel sx = resolve_axis(load_entry(a.lx + (dx + W)), cxp, W, ow, wrapped);
    FOV_CHECK(sx.lo, ow, 212);
    FOV_CHECK(sx.hi, ow, 213);
    FOV_CHECK(sx.exact_idx, ow, 214);
    xsel[warp][lane] = make_int4(sx.lo, sx.hi, __float_as_int(sx.ratio), sx.exact ? sx.exact_idx : -1);
  }
  __syncthreads();
  if (y0 >= H) return;  // warp-uniform

  // mix(l, r, 0) returns l exactly, so a zero ratio needs one tap.  mix(l, r, 1) = l + (r - l) is
  // r only when l and r are integer-valued (copy rows); after a vertical mix it can round one ulp
  // away from r, so ratio-one pixels keep both taps.
  int xlo[kInterpPx], xhi[kInterpPx], xex[kInterpPx];
  float xr[kInterpPx];
  bool unit_x = true, simple = true, need_left = false;
  int cmin = 0x7fffffff, cmax = -1;
#pragma unroll
  for (int k = 0; k < kInterpPx; ++k) {
    const int4 sx = xsel[k][lane];
    xr[k] = __int_as_float(sx.z);
    const bool zero = xr[k] == 0.0f, one = xr[k] == 1.0f;
    xlo[k] = sx.x;
    xhi[k] = zero ? sx.x : sx.y;
    xex[k] = sx.w;
    // 1:1 path: every pixel is its primary column (hi for ratio one, lo for ratio zero), and the
    // left tap of a ratio-one pixel is the primary column of the pixel before it
    unit_x = unit_x && (zero || one);
    if (one && xhi[k] != xlo[k]) {
      need_left = true;
      if (k > 0) unit_x = unit_x && xlo[k] == (xr[k - 1] == 1.0f ? xhi[k - 1] : xlo[k - 1]);
    }
    // fast paths: an exact hit must be the primary tap
    simple = simple && (xex[k] < 0 || (zero && xex[k] == xlo[k]) || (one && xex[k] == xhi[k]));
    cmin = min(cmin, min(xlo[k], xhi[k]));
    cmax = max(cmax, max(xlo[k], xhi[k]));
  }
  unit_x = __all_sync(0xffffffffu, unit_x);
  need_left = __any_sync(0xffffffffu, need_left);
  cmin = __reduce_min_sync(0xffffffffu, cmin);
  cmax = __reduce_max_sync(0xffffffffu, cmax);
  const int ncols = cmax - cmin + 1;

  // ---- y axis: one lane per row --------------------------------------------------------------
  {
    const AxisSel sy = resolve_axis(ey, cyp, H, oh, false);
    FOV_CHECK(sy.lo, oh, 215);
    FOV_CHECK(sy.hi, oh, 216);
    FOV_CHECK(sy.exact_idx, oh, 217);
    const bool deg = sy.ratio == 0.0f || sy.ratio == 1.0f;
    const int sel = sy.ratio == 1.0f ? sy.hi : sy.lo;
    const int rlo = deg ? sel : sy.lo, rhi = deg ? sel : sy.hi;
    const int yex = sy.exact ? sy.exact_idx : -1;
    simple = simple && (!my_row || yex < 0 || (yex == rlo && rhi == rlo));
    const int pair = rlo | (rhi << 16);
    const int above = __shfl_up_sync(0xffffffffu, pair, 1);
    RowSel mine;
    mine.off_lo = rlo * ow;
    mine.off_hi = rhi * ow;
    mine.ty = sy.ratio;
    mine.info = (yex << 1) | ((lane == 0 || above != pair) ? 1 : 0);
    if (my_row) rowsel[warp][lane] = mine;
  }

  uint32_t *orow = reinterpret_cast<uint32_t *>(a.out + (size_t)f * a.out_stride) +
                   (size_t)y0 * W + x4;
  const bool in_x = x4 < W;
  const bool vec_ok = (x4 + kInterpPx <= W) && ((reinterpret_cast<uintptr_t>(orow) & 15) == 0) &&
                      ((W & 3) == 0);
  const int nrows = min(kInterpRows, H - y0);
  // the fast paths below also assume one aligned 16-byte store per lane and row
  simple = __all_sync(0xffffffffu, simple && vec_ok);  // (orders the rowsel writes, too)
#ifdef FOV360_INTERP_FORCE_GENERIC
  simple = false;
#endif
  pdl_wait();  // both axes are resolved from the tables; the reduced buffer is read from here on

  auto store_row = [&](const uint32_t (&px)[kInterpPx]) {
    if (vec_ok) {
      __stcs(reinterpret_cast<uint4 *>(orow), make_uint4(px[0], px[1], px[2], px[3]));
    } else if (in_x) {
#pragma unroll
      for (int k = 0; k < kInterpPx; ++k)
        if (x4 + k < W) orow[k] = px[k];
    }
  };

  if (simple && unit_x) {
    // ---- every pixel sits on its primary reduced column: vertical mix (or copy) only, plus for
    // ratio-one pixels the rounding of l + (r - l) with l = the column to the left ---------------
    const uint32_t *col[kInterpPx];
    uint32_t keep[kInterpPx];  // bytes an exact-hit row copies: all 4 where x is an exact hit too
    bool one[kInterpPx];
#pragma unroll
    for (int k = 0; k < kInterpPx; ++k) {
      one[k] = xr[k] == 1.0f && xhi[k] != xlo[k];
      col[k] = red + (xr[k] == 1.0f ? xhi[k] : xlo[k]);
      keep[k] = xex[k] >= 0 ? 0xffffffffu : 0x00ffffffu;
    }
    const uint32_t *col_left = red + xlo[0];  // left tap of pixel 0 (when it is a ratio-one pixel)
    const bool all_one = __all_sync(0xffffffffu, one[0] && one[1] && one[2] && one[3]);
    TapPair tpl = {};
    f32x2 tp2[3][2] = {}, td2[3][2] = {};  // p and q - p of channel c, pixels (h, h + 2)
    int r = 0;
    while (r < nrows) {
      const RowSel rs = rowsel[warp][r];
      if (rs.off_lo == rs.off_hi) {
        // Copy rows (integer-valued taps: l + (r - l) == r) come in long runs: four rows of loads
        // in flight per batch.
        int nb = 1;
        RowSel b[kCopyRows];
        b[0] = rs;
#pragma unroll
        for (int j = 1; j < kCopyRows; ++j) {
          b[j] = rowsel[warp][min(r + j, kInterpRows - 1)];
          if (nb == j && r + j < nrows && b[j].off_lo == b[j].off_hi) nb = j + 1;
        }
        uint32_t px[kCopyRows][kInterpPx];
#pragma unroll
        for (int j = 0; j < kCopyRows; ++j) {
          if (j < nb) {
            const bool yhit = b[j].info >= 0;  // exact-hit row index is not -1
#pragma unroll
            for (int k = 0; k < kInterpPx; ++k)
              px[j][k] = __ldg(col[k] + b[j].off_lo) & (yhit ? keep[k] : 0x00ffffffu);
          }
        }
#pragma unroll
        for (int j = 0; j < kCopyRows; ++j) {
          if (j < nb) {
            store_row(px[j]);
            orow += W;
          }
        }
        r += nb;
      } else {
        // Packed form: half h of channel c holds pixels (h, h + 2) of the lane, so the left
        // neighbours of pixels (1, 3) are the pair of pixels (0, 2) itself.
        if (rs.info & 1) {  // warp-uniform: new reduced row pair
          uint32_t rp[kInterpPx], rq[kInterpPx];
#pragma unroll
          for (int k = 0; k < kInterpPx; ++k) {
            rp[k] = __ldg(col[k] + rs.off_lo);
            rq[k] = __ldg(col[k] + rs.off_hi);
          }
          if (need_left) convert_tap_pair(tpl, __ldg(col_left + rs.off_lo), __ldg(col_left + rs.off_hi));
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            tp2[0][h] = bytes_to_float2<0>(rp[h], rp[h + 2]);
            tp2[1][h] = bytes_to_float2<1>(rp[h], rp[h + 2]);
            tp2[2][h] = bytes_to_float2<2>(rp[h], rp[h + 2]);
            td2[0][h] = sub2_rn(bytes_to_float2<0>(rq[h], rq[h + 2]), tp2[0][h]);
            td2[1][h] = sub2_rn(bytes_to_float2<1>(rq[h], rq[h + 2]), tp2[1][h]);
            td2[2][h] = sub2_rn(bytes_to_float2<2>(rq[h], rq[h + 2]), tp2[2][h]);
          }
        }
        {
          const f32x2 ty2 = pack2(rs.ty, rs.ty);
          f32x2 v2[3][2];
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int h = 0; h < 2; ++h) v2[c][h] = add2_rn(tp2[c][h], mul2_rn(td2[c][h], ty2));
          uint32_t bits[kInterpPx][3];
          if (all_one) {  // warp-uniform: every pixel is mix(l, r, 1) = l + (r - l)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float lc = lerp_rn(tpl.p[c], tpl.d[c], rs.ty);
              uint32_t v1, v3;
              unpack2(v2[c][1], v1, v3);
              const f32x2 l02 = pack2(lc, __uint_as_float(v1));  // left of pixels (0, 2)
              const f32x2 o02 = trunc_bits2(add2_rn(l02, sub2_rn(v2[c][0], l02)));
              const f32x2 o13 = trunc_bits2(add2_rn(v2[c][0], sub2_rn(v2[c][1], v2[c][0])));
              unpack2(o02, bits[0][c], bits[2][c]);
              unpack2(o13, bits[1][c], bits[3][c]);
            }
          } else if (need_left) {  // warp-uniform: the warp that holds the gaze column
            float l[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) l[c] = lerp_rn(tpl.p[c], tpl.d[c], rs.ty);
            uint32_t vb[kInterpPx][3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              unpack2(v2[c][0], vb[0][c], vb[2][c]);
              unpack2(v2[c][1], vb[1][c], vb[3][c]);
            }
#pragma unroll
            for (int k = 0; k < kInterpPx; ++k)
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                const float vk = __uint_as_float(vb[k][c]);
                bits[k][c] = trunc_bits(one[k] ? __fadd_rn(l[c], __fsub_rn(vk, l[c])) : vk);
                l[c] = vk;
              }
          } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              unpack2(trunc_bits2(v2[c][0]), bits[0][c], bits[2][c]);
              unpack2(trunc_bits2(v2[c][1]), bits[1][c], bits[3][c]);
            }
          }
          uint32_t px[kInterpPx];
#pragma unroll
          for (int k = 0; k < kInterpPx; ++k) px[k] = pack_rgb0(bits[k][0], bits[k][1], bits[k][2]);
          store_row(px);
          orow += W;
          ++r;
        }
      }
    }
    return;
  }

  float4 *vs = vstage[warp];
  if (simple && ncols <= 32) {
    // ---- periphery: lane c owns window column c ---------------------------------------------
    const float4 *pv[kInterpPx];
    uint32_t keep_lo[kInterpPx], keep_hi[kInterpPx];  // 4th byte of an exact hit: from V.w or D.w
#pragma unroll
    for (int k = 0; k < kInterpPx; ++k) {
      pv[k] = vs + (xlo[k] - cmin);
      keep_lo[k] = (xex[k] >= 0 && xex[k] == xlo[k]) ? 0xff000000u : 0u;
      keep_hi[k] = (xex[k] >= 0 && xex[k] != xlo[k]) ? 0xff000000u : 0u;
      if (xhi[k] == xlo[k]) xr[k] = 0.0f;  // V + D * 0 = V: the selected column, exactly
    }
    // In the periphery 4 consecutive pixels sit on at most two reduced columns - those of the
    // lane's first and last pixel: pass 2 then reads 2 staged columns per row instead of 4 (the
    // pass is bound by shared-memory bandwidth, not by issue slots).
    bool use_b[kInterpPx];
    bool two = true;
#pragma unroll
    for (int k = 1; k < kInterpPx - 1; ++k) {
      use_b[k] = xlo[k] != xlo[0];
      two = two && (xlo[k] == xlo[0] || xlo[k] == xlo[kInterpPx - 1]);
    }
    const bool two_cols = __all_sync(0xffffffffu, two);
    // lanes past the window repeat its last column: every V and D stays finite, and the D of the
    // last column (0, or never multiplied by a non-zero tx) needs no special case
    const uint32_t *rcol = red + cmin + min(lane, ncols - 1);
    uint4 *orow4 = reinterpret_cast<uint4 *>(orow);  // `simple` implies aligned 16-byte stores
    TapPair tp = {};
    uint32_t rawp[kInterpChunk], rawq[kInterpChunk];
    auto request = [&](int r0) {
#pragma unroll
      for (int j = 0; j < kInterpChunk; ++j) {
        const RowSel rs = rowsel[warp][min(r0 + j, kInterpRows - 1)];
        rawp[j] = __ldg(rcol + rs.off_lo);
        rawq[j] = __ldg(rcol + rs.off_hi);
      }
    };
    // packed pass 1: {V.x, V.y} and {V.z, alpha} are the two 8-byte halves of the staged float4.
    // Only V is staged (kStride slots per row: the window's 32 columns and one copy of the last for
    // the right neighbour of column 31); pass 2 forms D = V[c+1] - V[c] for the columns it uses.
    constexpr int kStride = 34;
    static_assert(kInterpChunk * kStride <= kStage, "staged rows must fit the staging area");
    auto stage_row = [&](int j, const RowSel rs) {
      const f32x2 ty2 = pack2(rs.ty, rs.ty);
      const f32x2 v01 = add2_rn(pack2(tp.p[0], tp.p[1]), mul2_rn(pack2(tp.d[0], tp.d[1]), ty2));
      const float v2 = lerp_rn(tp.p[2], tp.d[2], rs.ty);
      const uint32_t alpha = rs.info >= 0 ? (rawp[j] & 0xff000000u) : 0u;
      const ulonglong2 val = make_ulonglong2(v01, pack2(v2, __uint_as_float(alpha)));
      ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(vs) + j * kStride;
      dst[lane] = val;
      if (lane == 31) dst[32] = val;
    };
    request(0);
    for (int r0 = 0; r0 < nrows; r0 += kInterpChunk) {
      RowSel rs[kInterpChunk];
      int fresh_later = 0;
#pragma unroll
      for (int j = 0; j < kInterpChunk; ++j) {
        rs[j] = rowsel[warp][min(r0 + j, kInterpRows - 1)];
        if (j > 0) fresh_later |= rs[j].info;
      }
      __syncwarp();  // pass 2 of the previous chunk has read its V and D
      // pass 1.  In the periphery a reduced row pair lasts 6-10 rows, so most chunks convert taps
      // once (or not at all); the general form converts under a predicate in every row.
      if (!(fresh_later & 1)) {  // warp-uniform
        if (rs[0].info & 1) convert_tap_pair(tp, rawp[0], rawq[0]);
#pragma unroll
        for (int j = 0; j < kInterpChunk; ++j) stage_row(j, rs[j]);
      } else {
#pragma unroll
        for (int j = 0; j < kInterpChunk; ++j) {
          if (rs[j].info & 1) convert_tap_pair(tp, rawp[j], rawq[j]);
          stage_row(j, rs[j]);
        }
      }
      if (r0 + kInterpChunk < nrows) request(r0 + kInterpChunk);  // in flight during pass 2
      __syncwarp();
#pragma unroll
      for (int j = 0; j < kInterpChunk; ++j) {
        uint32_t px[kInterpPx];
        if (two_cols) {  // warp-uniform: columns of pixels 1 and 2 are those of pixel 0 or pixel 3
          const ulonglong2 va = *reinterpret_cast<const ulonglong2 *>(pv[0] + j * kStride);
          const ulonglong2 na = *reinterpret_cast<const ulonglong2 *>(pv[0] + j * kStride + 1);
          const ulonglong2 vb = *reinterpret_cast<const ulonglong2 *>(pv[3] + j * kStride);
          const ulonglong2 nb = *reinterpret_cast<const ulonglong2 *>(pv[3] + j * kStride + 1);
          const ulonglong2 da = sub_cols(na, va), db = sub_cols(nb, vb);
#pragma unroll
          for (int k = 0; k < kInterpPx; ++k) {
            const bool b = k == 3 || (k > 0 && use_b[k]);
            px[k] = staged_px_vn(make_ulonglong2(b ? vb.x : va.x, b ? vb.y : va.y),
                                 make_ulonglong2(b ? db.x : da.x, b ? db.y : da.y),
                                 make_ulonglong2(0ull, b ? nb.y : na.y), xr[k], keep_lo[k], keep_hi[k]);
          }
        } else {
#pragma unroll
          for (int k = 0; k < kInterpPx; ++k) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(pv[k] + j * kStride);
            const ulonglong2 n = *reinterpret_cast<const ulonglong2 *>(pv[k] + j * kStride + 1);
            px[k] = staged_px_vn(v, sub_cols(n, v), n, xr[k], keep_lo[k], keep_hi[k]);
          }
        }
        if (r0 + j < nrows) __stcs(orow4, make_uint4(px[0], px[1], px[2], px[3]));
        orow4 += W / 4;
      }
    }
    return;
  }

  if (ncols <= kInterpMaxCols) {
    // ---- wider windows (the warp straddles the edge of the 1:1 band), exact hits off the
    // selected taps: stage V and D row by row, patch exact hits from the reduced buffer --------
    const float4 *pv[kInterpPx];
#pragma unroll
    for (int k = 0; k < kInterpPx; ++k) {
      pv[k] = vs + (xlo[k] - cmin);
      if (xhi[k] == xlo[k]) xr[k] = 0.0f;
    }
    for (int r = 0; r < nrows; ++r, orow += W) {
      const RowSel rs = rowsel[warp][r];
      const uint32_t *ra = red + rs.off_lo + cmin, *rb = red + rs.off_hi + cmin;
      __syncwarp();  // the previous row's pixels have been formed
      // an iteration stages 31 columns; lane 31 only supplies V[c+1] to lane 30
      for (int c0 = 0; c0 < ncols; c0 += 31) {
        const int c = min(c0 + lane, ncols - 1);
        TapPair t;
        convert_tap_pair(t, __ldg(ra + c), __ldg(rb + c));
        const float v0 = lerp_rn(t.p[0], t.d[0], rs.ty);
        const float v1 = lerp_rn(t.p[1], t.d[1], rs.ty);
        const float v2 = lerp_rn(t.p[2], t.d[2], rs.ty);
        const float n0 = __shfl_down_sync(0xffffffffu, v0, 1);
        const float n1 = __shfl_down_sync(0xffffffffu, v1, 1);
        const float n2 = __shfl_down_sync(0xffffffffu, v2, 1);
        if (lane < 31 && c0 + lane < ncols) {
          vs[c] = make_float4(v0, v1, v2, 0.f);
          vs[kInterpMaxCols + c] =
              make_float4(__fsub_rn(n0, v0), __fsub_rn(n1, v1), __fsub_rn(n2, v2), 0.f);
        }
      }
      __syncwarp();
      uint32_t px[kInterpPx];
#pragma unroll
      for (int k = 0; k < kInterpPx; ++k) px[k] = lerp_px(pv[k][0], pv[k][kInterpMaxCols], xr[k]);
      const int yex = rs.info >> 1;
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

  // ---- seam warps (the window spans the whole reduced width): gather every tap directly ------
  for (int r = 0; r < nrows; ++r, orow += W) {
    const RowSel rs = rowsel[warp][r];
    const float ty = rs.ty;
    const int yex = rs.info >> 1;
    const uint32_t *ra = red + rs.off_lo, *rb = red + rs.off_hi;
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
// interpolate_rect + gnomonic fused: the viewport rendered straight from the reduced buffer.
//
// out(i, j) = interpolate_rect(reduced)[gnomonic source pixel of (i, j)] without ever forming the
// full-resolution frame (what a head-mounted client displays is a viewport,
// so the 4*W*H-byte un-warped frame is pure traffic).  One thread per viewport pixel evaluates the
// inverse gnomonic map (projection_common.cuh) and then exactly the per-pixel form of
// interpolate_rect_kernel - table entry per axis, border fix-ups, exact-hit copy, vertical mixes
// then the horizontal one - so the result equals the two reference kernels run back to back.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sat_interpolate_gnomonic_kernel(
    uint32_t *__restrict__ out, int tw, int th, const InterpArgs a, float gx, float gy,
    const GnomonicView view) {
  const int i = blockIdx.x * 32 + threadIdx.x;
  const int j = blockIdx.y * 8 + threadIdx.y;
  if (i >= tw || j >= th) return;
  const int W = a.W, H = a.H, ow = a.ow, oh = a.oh;
  int x, y;
  gnomonic_source(i, j, tw, th, W, H, view, x, y);
  const int cxp = gaze_px(gx, W), cyp = gaze_px(gy, H);
  bool wrapped = false;  // sat_decoder_interpolate_kernel.cl:26-33
  if (x - cxp > W / 2) {
    x -= W;
    wrapped = true;
  } else if (x - cxp < -(W / 2)) {
    x += W;
    wrapped = true;
  }
  const AxisSel sx =
      resolve_axis(load_entry(a.lx + (clampi(x - cxp, -W, W) + W)), cxp, W, ow, wrapped);
  const AxisSel sy = resolve_axis(load_entry(a.ly + (clampi(y - cyp, -H, H) + H)), cyp, H, oh, false);
  const uint32_t *red = reinterpret_cast<const uint32_t *>(a.red);
  uint32_t px;
  if (sx.exact && sy.exact) {  // :67-72: all 4 bytes of the sample
    px = __ldg(red + (size_t)sy.exact_idx * ow + sx.exact_idx);
  } else {
    const int xlo = sx.lo, xhi = sx.hi, ylo = sy.lo, yhi = sy.hi;
    const uint32_t *ra = red + (size_t)ylo * ow, *rb = red + (size_t)yhi * ow;
    const uint32_t tl = __ldg(ra + xlo), tr = __ldg(ra + xhi);
    const uint32_t bl = __ldg(rb + xlo), br = __ldg(rb + xhi);
    const float tx = sx.ratio, ty = sy.ratio;
    px = pack_rgb0(
        trunc_bits(mix_rn(vmix_channel<0>(tl, bl, ty), vmix_channel<0>(tr, br, ty), tx)),
        trunc_bits(mix_rn(vmix_channel<1>(tl, bl, ty), vmix_channel<1>(tr, br, ty), tx)),
        trunc_bits(mix_rn(vmix_channel<2>(tl, bl, ty), vmix_channel<2>(tr, br, ty), tx)));
  }
  out[(size_t)j * tw + i] = px;
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

FOV_DEFINE_BOUNDS_READER(bounds_read_sat_decode)

cudaError_t launch_sat_sample_rect(const LaunchCtx &lc, int n, uint8_t *out, size_t out_stride, int ow,
                                   int oh, int out_linesize, const uint32_t *sat,
                                   size_t sat_stride, int W, int H, const int16_t *xedge,
                                   const int16_t *yedge, const GazeBatch &gaze,
                                   const uint8_t *src, size_t src_stride, int src_linesize) {
  SampleArgs a;
  a.src = reinterpret_cast<const uint32_t *>(src);
  a.src_stride = src_stride;
  a.src_linesize_px = src_linesize / 4;
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
#ifndef FOV360_SAMPLE_ROWS
#define FOV360_SAMPLE_ROWS 4
#endif
  constexpr int kRows = FOV360_SAMPLE_ROWS;  // reduced rows per warp: kRows + 1 edges x 3 words in flight per lane
  const dim3 grid((ow + kSampleCols - 1) / kSampleCols,
                  (oh + kSampleWarps * kRows - 1) / (kSampleWarps * kRows), n),
      block(32, kSampleWarps);
  KernelScope ks(lc, "sat_sample_rect");
  auto kernel = gaze.dev ? (lc.reduced_pad_zero ? sat_sample_rect_kernel<kRows, true, true>
                                                : sat_sample_rect_kernel<kRows, true, false>)
                         : (lc.reduced_pad_zero ? sat_sample_rect_kernel<kRows, false, true>
                                                : sat_sample_rect_kernel<kRows, false, false>);
  return launch_chained(kernel, grid, block, 0, lc.stream, a, gaze);
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
  // Tile height by the amount of work: 32 rows per warp once the batch fills several waves of
  // CTAs, 16 / 8 for a few frames / one small frame (measured on a B200: one 4K frame 0.046 ms with
  // 32 rows, 0.026 with 8; eight 4K frames 0.125 / 0.116 ms with 32 / 16; four 8K frames 0.191 / 0.196).
  const size_t px = (size_t)n * W * H;
  static const int force_rows = [] {
    const char *e = getenv("FOV360_INTERP_ROWS");
    const int v = e ? atoi(e) : 0;
    return (v == 8 || v == 16 || v == 32) ? v : 0;  // the instantiated tile heights
  }();
  const int rows = force_rows ? force_rows : (px < ((size_t)16 << 20) ? 8 : (px < ((size_t)96 << 20) ? 16 : 32));
  const dim3 grid((W + 32 * kInterpPx - 1) / (32 * kInterpPx),
                  (H + kInterpWarps * rows - 1) / (kInterpWarps * rows), n),
      block(32, kInterpWarps);
  KernelScope ks(lc, "sat_interpolate_rect");
  if (gaze.dev) {
    if (rows == 8) return launch_chained(sat_interpolate_rect_kernel<8, true>, grid, block, 0, lc.stream, a, gaze);
    if (rows == 16) return launch_chained(sat_interpolate_rect_kernel<16, true>, grid, block, 0, lc.stream, a, gaze);
    return launch_chained(sat_interpolate_rect_kernel<32, true>, grid, block, 0, lc.stream, a, gaze);
  }
  if (rows == 8) return launch_chained(sat_interpolate_rect_kernel<8, false>, grid, block, 0, lc.stream, a, gaze);
  if (rows == 16) return launch_chained(sat_interpolate_rect_kernel<16, false>, grid, block, 0, lc.stream, a, gaze);
  return launch_chained(sat_interpolate_rect_kernel<32, false>, grid, block, 0, lc.stream, a, gaze);
}

cudaError_t launch_sat_interpolate_gnomonic(const LaunchCtx &lc, uint8_t *out, int tw, int th,
                                            const uint8_t *red, int ow, int oh, int W, int H,
                                            const InterpEntry *lx, const InterpEntry *ly,
                                            float gaze_x, float gaze_y, const GnomonicView &view) {
  InterpArgs a;
  a.out = out;
  a.red = red;
  a.lx = lx;
  a.ly = ly;
  a.out_stride = 0;
  a.red_stride = 0;
  a.W = W;
  a.H = H;
  a.ow = ow;
  a.oh = oh;
  const dim3 grid((tw + 31) / 32, (th + 7) / 8), block(32, 8);
  KernelScope ks(lc, "sat_interpolate_gnomonic");
  sat_interpolate_gnomonic_kernel<<<grid, block, 0, lc.stream>>>(
      reinterpret_cast<uint32_t *>(out), tw, th, a, gaze_x, gaze_y, view);
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
