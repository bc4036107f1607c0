// ImageSampler kernels for sm_100a: the no-SAT baseline path of the reference (log-rect point
// sampling, log-polar sampling, log-polar inverse warp, 3x3 blur of the outer half).
//
// Replaces sample_rect_kernel (image_sampler_sample_rect_kernel.cl:1-46),
// sample_logpolar_kernel / logpolar_gaussian_blur_kernel
// (image_sampler_sample_logpolar_kernel.cl:41-142) and interpolate_logpolar_kernel
// (image_sampler_interpolate_kernel.cl:1-81).  The 2-D int16 grids of the reference are
// replaced by their separable 1-D factors (luts.cc); the grid value is re-formed in
// registers, which removes a 4 B/pixel table read from both samplers.
#include "bounds_check.cuh"
#include <cstdlib>

#include "fov360_internal.h"
#include "pixel_math.cuh"

namespace fov {
namespace {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// (int)(c * dim + delta): float multiply, float add, truncation
// (image_sampler_sample_rect_kernel.cl:26-27, image_sampler_sample_logpolar_kernel.cl:67-70).
__device__ __forceinline__ int gaze_plus(float c, int dim, int delta) {
  return __float2int_rz(__fadd_rn(__fmul_rn(c, (float)dim), (float)delta));
}

// Copies bytes 0..2 of one pixel, leaving byte 3 (and any wider stride) untouched.  4-byte pixels
// are moved as one load and two partial stores: the target pixel is never read back.
__device__ __forceinline__ void copy_rgb(uint8_t *o, const uint8_t *s, bool word_ok) {
  if (word_ok) {
    store_xyz(reinterpret_cast<uint32_t *>(o), __ldg(reinterpret_cast<const uint32_t *>(s)));
  } else {
    o[0] = s[0];
    o[1] = s[1];
    o[2] = s[2];
  }
}

struct GatherArgs {
  uint8_t *out;
  const uint8_t *src;
  int ow, oh, out_linesize, obpp, W, H, src_linesize, sbpp;
  bool word_ok;
  float cx, cy;
};

#ifndef FOV360_GATHER_ROWS
#define FOV360_GATHER_ROWS 4
#endif
constexpr int kGatherRows = FOV360_GATHER_ROWS;  // reduced rows per thread: that many independent gathers in flight

__global__ void __launch_bounds__(256) img_sample_rect_kernel(const GatherArgs a,
                                                              const int16_t *__restrict__ xd,
                                                              const int16_t *__restrict__ yd) {
  pdl_trigger();
  const int i = blockIdx.x * 32 + threadIdx.x;
  const int j0 = (blockIdx.y * 8 + threadIdx.y) * kGatherRows;
  if (i >= a.ow || j0 >= a.oh) return;
  FOV_CHECK(i, a.ow, 101);
  int x = gaze_plus(a.cx, a.W, xd[i]);
  if (x >= a.W)  // :29-33
    x -= a.W;
  else if (x < 0)
    x += a.W;
  if (x < 0 || x >= a.W) return;  // :35: the column keeps its contents
  pdl_wait();  // tables above; the frame and the target buffer below
  const uint8_t *scol = a.src + (size_t)x * a.sbpp;
  uint8_t *ocol = a.out + (size_t)i * a.obpp;
  if (a.word_ok) {
    uint32_t v[kGatherRows];
    bool on[kGatherRows];
#pragma unroll
    for (int r = 0; r < kGatherRows; ++r) {
      const int y = gaze_plus(a.cy, a.H, yd[min(j0 + r, a.oh - 1)]);
      on[r] = j0 + r < a.oh && y >= 0 && y < a.H;  // :35-43
      if (on[r]) FOV_CHECK((size_t)y * a.src_linesize + (size_t)x * a.sbpp + 3, (size_t)a.H * a.src_linesize, 102);
      v[r] = on[r] ? __ldg(reinterpret_cast<const uint32_t *>(scol + (size_t)y * a.src_linesize)) : 0u;
    }
#pragma unroll
    for (int r = 0; r < kGatherRows; ++r)
      if (on[r])
        store_xyz(reinterpret_cast<uint32_t *>(ocol + (size_t)(j0 + r) * a.out_linesize), v[r]);
    return;
  }
  for (int r = 0; r < kGatherRows && j0 + r < a.oh; ++r) {
    const int y = gaze_plus(a.cy, a.H, yd[j0 + r]);
    if (y >= 0 && y < a.H)
      copy_rgb(ocol + (size_t)(j0 + r) * a.out_linesize, scol + (size_t)y * a.src_linesize, false);
  }
}

// Grid value as the reference stores it: int truncation of the float product, narrowed to
// int16 (image_sampler_sample_logpolar_kernel.cl:31-38).
__device__ __forceinline__ int logpolar_delta(float radius, float trig) {
  return (int)(int16_t)__float2int_rz(__fmul_rn(radius, trig));
}

// v % w as C evaluates it (the sign of the dividend for negative v), for 0 < w < 2^15 and
// |v| < 2^22: a float estimate of the quotient, corrected by at most one either way.
__device__ __forceinline__ int mod_width(int v, int w, float rcp_w) {
  if (v < 0) return v % w;  // only reachable for frames narrower than 3277 pixels
  int r = v - __float2int_rz(__fmul_rn((float)v, rcp_w)) * w;
  if (r < 0) r += w;
  if (r >= w) r -= w;
  return r;
}

__global__ void __launch_bounds__(256) img_sample_logpolar_kernel(
    const GatherArgs a, const float *__restrict__ radius, const float *__restrict__ cs,
    const float *__restrict__ sn) {
  pdl_trigger();
  const int i = blockIdx.x * 32 + threadIdx.x;
  const int j0 = (blockIdx.y * 8 + threadIdx.y) * kGatherRows;
  if (i >= a.ow || j0 >= a.oh) return;
  const float r = radius[i];
  const int W10 = 10 * a.W;
  const float rcpW = __frcp_rn((float)a.W);
  uint8_t *ocol = a.out + (size_t)i * a.obpp;
  const uint8_t *sp[kGatherRows];
  bool on[kGatherRows];
#pragma unroll
  for (int k = 0; k < kGatherRows; ++k) {
    const int j = min(j0 + k, a.oh - 1);
    int x = gaze_plus(a.cx, a.W, logpolar_delta(r, cs[j]));
    int y = gaze_plus(a.cy, a.H, logpolar_delta(r, sn[j]));
    x = mod_width(x + W10, a.W, rcpW);  // :73
    y = clampi(y, 0, a.H - 1);
    on[k] = j0 + k < a.oh && x >= 0;  // :76-77 (x stays negative only for x < -10 W)
    FOV_CHECK(j, a.oh, 111);
    if (on[k]) FOV_CHECK(x, a.W, 112);
    if (on[k]) FOV_CHECK((size_t)y * a.src_linesize + (size_t)x * a.sbpp + (a.word_ok ? 3 : 2), (size_t)a.H * a.src_linesize, 113);
    sp[k] = a.src + (size_t)y * a.src_linesize + (size_t)max(x, 0) * a.sbpp;
  }
  pdl_wait();  // tables above; the frame and the target buffer below
  if (a.word_ok) {
    uint32_t v[kGatherRows];
#pragma unroll
    for (int k = 0; k < kGatherRows; ++k)
      v[k] = on[k] ? __ldg(reinterpret_cast<const uint32_t *>(sp[k])) : 0u;
#pragma unroll
    for (int k = 0; k < kGatherRows; ++k)
      if (on[k])
        store_xyz(reinterpret_cast<uint32_t *>(ocol + (size_t)(j0 + k) * a.out_linesize), v[k]);
    return;
  }
#pragma unroll
  for (int k = 0; k < kGatherRows; ++k)
    if (on[k]) copy_rgb(ocol + (size_t)(j0 + k) * a.out_linesize, sp[k], false);
}

__global__ void __launch_bounds__(256) img_logpolar_grid_expand_kernel(
    int16_t *grid, int ow, int oh, const float *__restrict__ radius, const float *__restrict__ cs,
    const float *__restrict__ sn) {
  const int i = blockIdx.x * 32 + threadIdx.x;
  const int j = blockIdx.y * 8 + threadIdx.y;
  if (i >= ow || j >= oh) return;
  const float r = radius[i];
  grid[((size_t)j * ow + i) * 2] = (int16_t)logpolar_delta(r, cs[j]);
  grid[((size_t)j * ow + i) * 2 + 1] = (int16_t)logpolar_delta(r, sn[j]);
}

// ---------------------------------------------------------------------------------------
// interpolate_logpolar (image_sampler_interpolate_kernel.cl:1-81): inverse log-polar warp.
//
// The map is not separable, so each full-resolution pixel needs its own radius logarithm and
// angle.  What the reference evaluates with library calls per pixel (pow, sqrt, log, atan, fmod,
// exp, cos, sin - several of them promoted to double by its typing) is organised here as:
//  * i_f (:28-33) depends on the integer d2 = dx^2 + dy^2 only.  ln(d2) comes from a host-built
//    table of {1/c, K ln c} per (exponent, top 7 mantissa bits) and a cubic in r = d2/c - 1
//    (|r| <= 2^-8), six double-precision operations; the error (~1e-11 of an index) is far below
//    the double -> float rounding the reference applies last, so i_f is the reference's float
//    except on about one pixel in 10^7.
//  * j_f (:36-40) is evaluated in single precision: octant reduction with the reciprocals of |dx|
//    (once per thread) and |dy| (once per row), a degree-7 polynomial in a^2 for atan(a) on [0, 1],
//    and one FMA that applies sign, quadrant, the oh / 2 pi scale and the reference's "+ 2 oh".
//    The reference itself rounds atanf, the scaled angle and the "+ 2 oh" sum to float, which
//    quantises j_f to one ulp of ~2.5 oh; the single-precision value lands on the same quantum for
//    ~93 % of the pixels and on a neighbouring one otherwise (|difference| <= `zone`).
//  * round(j_f) only matters when the forward map of the rounded indices hits the pixel (:46-55).
//    Where j_f lies within `zone` of n + 0.5 both candidates are tested, and only if one of them
//    hits is j_f re-evaluated exactly as the reference's typing prescribes (float division,
//    correctly rounded float arctangent, double scaling) - about 0.05 % of the pixels.
//  * floor / ceil of j_f next to an integer need no such care: a ratio of ~1 between rows
//    (n - 1, n) and a ratio of ~0 between rows (n, n + 1) give the same pixel to within the
//    contract (<= 1 LSB).
//  * the exact-hit test (:46-55) keeps the reference's double arithmetic; its exp / cos / sin only
//    depend on the rounded indices and come from host-built tables.
//  * an exact hit is a bilinear tap whose four corners are the hit sample (mix(a, a, t) = a
//    exactly) - no divergence; the sample's 4th byte is OR-ed in.
// One lane owns a column of the CTA's tile and walks down its rows two at a time, so everything
// that depends on x only is computed once per thread, the per-row values once per CTA (shared
// memory), and the single-precision arithmetic of the two rows is issued as packed pairs.
// ---------------------------------------------------------------------------------------
struct LpInterpArgs {
  uint32_t *out;
  const uint32_t *red;
  const double2 *lntab;  // [kLnExponents * 128] {1 / c, K ln c}, K = ow / 20
  const double *radius;  // [ow] (double)expf(10.0f * i / ow)                      (:47)
  const double2 *dir;    // [oh] cos, sin of (float)j / oh * 2.0f * M_PI            (:48, :51)
  int W, H, ow, oh, rows;
  float cx, cy;
  double k0, k1, k2, k3;  // K, -K/2, K/3, -K/4
  float turns;            // (float)(oh / (2 pi))
  float zone;
  uint32_t magic;         // 0x4B000000 (see bytes_to_float2_m)
};

constexpr int kLpThreads = 256;
constexpr int kLpMaxRows = 64;

struct __align__(16) LpRow {
  int dy, dy2;
  float ady, rdy;
};

// j_f exactly as the reference's typing evaluates it (:36-44): the slow path of ambiguous pixels.
__device__ __noinline__ float logpolar_jf_exact(int dx, int dy, int oh) {
  const double kPi = 3.14159265358979323846, kPi2 = 1.57079632679489661923;
  if (dx == 0)  // :41-43
    return (float)((kPi2 + kPi * (double)(dy < 0)) * ((double)oh / (2.0 * kPi)));
  const float q = __fdiv_rn((float)dy, (float)dx);
  const float at = (float)atan((double)q);
  const float j_f = (float)(((double)at + kPi * (double)(dx < 0)) * ((double)(float)oh / (2.0 * kPi)));
  // fmod(j_f + 2 oh, oh), :39: the argument is a float in (1.75 oh, 2.75 oh), so the remainder is
  // one or two subtractions of oh, each exact
  const float val = __fadd_rn(j_f, (float)(2 * oh));
  return val >= (float)(2 * oh) ? __fsub_rn(val, (float)(2 * oh)) : __fsub_rn(val, (float)oh);
}

__device__ __forceinline__ bool logpolar_hits(const LpInterpArgs &a, int i, int j, double cxw,
                                              double cyh, int x, int y) {
  FOV_CHECK(i, a.ow, 122);
  FOV_CHECK(j, a.oh, 123);
  const double rad = __ldg(a.radius + i);
  const double2 cs = __ldg(a.dir + j);
  const int calc_x = __double2int_rz(__dadd_rn(cxw, __dmul_rn(rad, cs.x)));  // :46-48
  const int calc_y = __double2int_rz(__dadd_rn(cyh, __dmul_rn(rad, cs.y)));  // :49-51
  return calc_x == x && calc_y == y;                                          // :53
}

#ifndef FOV360_LP_MIN_CTAS
#define FOV360_LP_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(kLpThreads, FOV360_LP_MIN_CTAS) img_interpolate_logpolar_kernel(const LpInterpArgs a) {
  __shared__ LpRow srows[kLpMaxRows];
  pdl_trigger();
  const int W = a.W, H = a.H, ow = a.ow, oh = a.oh;
  const int xx = blockIdx.x * kLpThreads + threadIdx.x;
  const int y0 = blockIdx.y * a.rows;
  const int nrows = min(a.rows, H - y0);
  const float cxw_f = __fmul_rn(a.cx, (float)W), cyh_f = __fmul_rn(a.cy, (float)H);
  const int cxp = __float2int_rz(cxw_f), cyp = __float2int_rz(cyh_f);  // :19-20
  if (threadIdx.x < a.rows) {
    const int dy = min(y0 + (int)threadIdx.x, H - 1) - cyp;
    LpRow r;
    r.dy = dy;
    r.dy2 = dy * dy;
    r.ady = (float)abs(dy);
    r.rdy = dy ? __frcp_rn(r.ady) : 0.0f;
    srows[threadIdx.x] = r;
  }
  __syncthreads();
  if (xx >= W) return;

  // ---- per column -----------------------------------------------------------------------------
  int x = xx;
  if (x - cxp > W / 2)  // :21-25
    x -= W;
  else if (x - cxp < -(W / 2))
    x += W;
  const int dx = x - cxp;
  const int dx2 = dx * dx;
  const float adx = (float)abs(dx), fdx2 = (float)dx2;
  const float rdx = dx ? __frcp_rn(adx) : 0.0f;
  const bool negx = dx < 0;
  const double cxw = (double)cxw_f, cyh = (double)cyh_f;
  const float oh1 = (float)oh, oh2 = (float)(2 * oh), ohq = 0.25f * oh1;
  const float base_x = negx ? __fadd_rn(oh2, 0.5f * oh1) : oh2;  // 2 oh + pi (dx < 0), in turns
  const float base_up = __fadd_rn(base_x, ohq), base_dn = __fsub_rn(base_x, ohq);
  uint32_t *op = a.out + (size_t)y0 * W + xx;
  const uint32_t magic = a.magic;
  const float ow_top = (float)(ow - 1);

  pdl_wait();  // row and column set-up above; the reduced buffer and the target below
  for (int r0 = 0; r0 < nrows; r0 += 2, op += 2 * (size_t)W) {
    float i_f[2], j_f[2], tt[2], aa[2];
    bool swp[2], neg[2], centre[2];
    int yy[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const LpRow rw = srows[min(r0 + k, nrows - 1)];
      yy[k] = rw.dy + cyp;
      // ---- i_f: K ln(d2), K = ow / 20 (:28-33) ----
      const int d2 = dx2 + rw.dy2;
      centre[k] = d2 == 0;
      const float d2f = fmaxf(__fmaf_rn(rw.ady, rw.ady, fdx2), 1.0f);
      FOV_CHECK((int)(__float_as_uint(d2f) >> 16) - (127 << 7), kLnExponents * 128, 121);
      const double2 te = __ldg(a.lntab + ((int)(__float_as_uint(d2f) >> 16) - (127 << 7)));
      const double dd = __hiloint2double(0x43300000, d2) - 4503599627370496.0;  // (double)d2
      const double rr = fma(dd, te.x, -1.0);
      double q = fma(rr, a.k3, a.k2);
      q = fma(rr, q, a.k1);
      q = fma(rr, q, a.k0);
      // clamp(round / floor / ceil(i_f), 0, ow - 1) (:34, :59, :61) = the same of min(i_f, ow - 1):
      // i_f >= 0, and beyond ow - 1 all three indices are ow - 1 whatever the ratio
      i_f[k] = fminf((float)fma(rr, q, te.y), ow_top);
      // ---- octant reduction of the angle ----
      swp[k] = rw.ady > adx;
      neg[k] = (rw.dy < 0) != negx;
      aa[k] = __fmul_rn(fminf(adx, rw.ady), swp[k] ? rw.rdy : rdx);
    }
    {  // atan(a) = a P(a^2) on [0, 1], |error| < 4e-8 before rounding; both rows at once
      const f32x2 A = pack2(aa[0], aa[1]);
      const f32x2 S = mul2_rn(A, A);
      auto c2 = [](float c) { return pack2(c, c); };
      f32x2 P = fma2_rn(c2(-0.004054544493556023f), S, c2(0.02186289243400097f));
      P = fma2_rn(P, S, c2(-0.05591226741671562f));
      P = fma2_rn(P, S, c2(0.09642196446657181f));
      P = fma2_rn(P, S, c2(-0.1390863060951233f));
      P = fma2_rn(P, S, c2(0.19946566224098206f));
      P = fma2_rn(P, S, c2(-0.33329862356185913f));
      P = fma2_rn(P, S, c2(0.9999993443489075f));
      uint32_t t0, t1;
      unpack2(mul2_rn(P, A), t0, t1);
      tt[0] = __uint_as_float(t0), tt[1] = __uint_as_float(t1);
    }
    int ti0[2], ti1[2], tj0[2], tj1[2];
    float ir[2], jr[2];
    bool hit[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      // |atan(dy/dx)| is t, or pi/2 - t when |dy| > |dx|; its sign is that of dy * dx; pi is added
      // for dx < 0; all of it in turns of oh, together with the "+ 2 oh" of :39
      const float sp = (neg[k] != swp[k]) ? -a.turns : a.turns;
      const float base = swp[k] ? (neg[k] ? base_dn : base_up) : base_x;
      const float val = __fmaf_rn(sp, tt[k], base);
      j_f[k] = val >= oh2 ? __fsub_rn(val, oh2) : __fsub_rn(val, oh1);
      float fj = floorf(j_f[k]);
      jr[k] = __fsub_rn(j_f[k], fj);
      int jj = __float2int_rz(fj);
      const float fi = floorf(i_f[k]);
      ir[k] = __fsub_rn(i_f[k], fi);
      int ii = __float2int_rz(fi);
      int i = ii + (ir[k] >= 0.5f ? 1 : 0);  // :34
      if (fabsf(jr[k] - 0.5f) < a.zone || centre[k]) {
        // round(j_f) is not certain: it matters only if a candidate is an exact hit.  (At the gaze
        // pixel itself d2 = 0 and the logarithm above is meaningless: its index is not used.)
        bool exact = centre[k];
        if (!exact)
          exact = logpolar_hits(a, i, jj, cxw, cyh, x, yy[k]) ||  // j_f < oh
                  logpolar_hits(a, i, min(jj + 1, oh - 1), cxw, cyh, x, yy[k]);
        if (exact) {
          if (centre[k]) ir[k] = 0.0f, ii = 0, i = 0;  // i_f = 0 at the gaze pixel itself (:28-29)
          fj = floorf(j_f[k] = logpolar_jf_exact(dx, yy[k] - cyp, oh));
          jr[k] = __fsub_rn(j_f[k], fj);
          jj = __float2int_rz(fj);
        }
      }
      const int j = min(jj + (jr[k] >= 0.5f ? 1 : 0), oh - 1);  // :44
      hit[k] = logpolar_hits(a, i, j, cxw, cyh, x, yy[k]);
      const int min_i = ii;                             // :59
      const int max_i = ii + (ir[k] > 0.0f ? 1 : 0);    // :61
      const int jn = jj + (jr[k] > 0.0f ? 1 : 0);                   // :60, :62: j_f + oh is exact
      ti0[k] = hit[k] ? i : min_i;
      ti1[k] = hit[k] ? i : max_i;
      tj0[k] = hit[k] ? j : jj;
      tj1[k] = hit[k] ? j : (jn >= oh ? jn - oh : jn);
    }
    uint32_t tl[2], tr[2], bl[2], br[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      FOV_CHECK(tj0[k], oh, 124);
      FOV_CHECK(tj1[k], oh, 125);
      FOV_CHECK(ti0[k], ow, 126);
      FOV_CHECK(ti1[k], ow, 127);
      const uint32_t ra = (uint32_t)(tj0[k] * ow), rb = (uint32_t)(tj1[k] * ow);  // ow * oh < 2^31
      tl[k] = __ldg(a.red + (ra + (uint32_t)ti0[k]));
      tr[k] = __ldg(a.red + (ra + (uint32_t)ti1[k]));
      bl[k] = __ldg(a.red + (rb + (uint32_t)ti0[k]));
      br[k] = __ldg(a.red + (rb + (uint32_t)ti1[k]));
    }
    // :69-79: vertical mixes, then the horizontal one, every operation rounded on its own
    const f32x2 JR = pack2(jr[0], jr[1]), IR = pack2(ir[0], ir[1]);
    uint32_t c[3][2];
#define FOV_LP_CHANNEL(C)                                                                  \
  {                                                                                        \
    const f32x2 TL = bytes_to_float2_m<C>(tl[0], tl[1], magic);                            \
    const f32x2 TR = bytes_to_float2_m<C>(tr[0], tr[1], magic);                            \
    const f32x2 BL = bytes_to_float2_m<C>(bl[0], bl[1], magic);                            \
    const f32x2 BR = bytes_to_float2_m<C>(br[0], br[1], magic);                            \
    const f32x2 L = add2_rn(TL, mul2_rn(sub2_rn(BL, TL), JR));                             \
    const f32x2 R = add2_rn(TR, mul2_rn(sub2_rn(BR, TR), JR));                             \
    unpack2(trunc_bits2(add2_rn(L, mul2_rn(sub2_rn(R, L), IR))), c[C][0], c[C][1]);        \
  }
    FOV_LP_CHANNEL(0)
    FOV_LP_CHANNEL(1)
    FOV_LP_CHANNEL(2)
#undef FOV_LP_CHANNEL
#pragma unroll
    for (int k = 0; k < 2; ++k)
      if (r0 + k < nrows)
        __stcs(op + (size_t)k * W,
               pack_rgb0(c[0][k], c[1][k], c[2][k]) | (hit[k] ? (tl[k] & 0xff000000u) : 0u));
  }
}

// logpolar_gaussian_blur_kernel (image_sampler_sample_logpolar_kernel.cl:88-142), any geometry:
// one thread per pixel.
__global__ void __launch_bounds__(256) img_logpolar_blur_kernel(uint32_t *__restrict__ out, int ow,
                                                                int oh,
                                                                const uint32_t *__restrict__ src) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * 32 + threadIdx.x;
  const int j = blockIdx.y * 8 + threadIdx.y;
  if (i >= ow || j >= oh) return;
  const size_t t = (size_t)j * ow + i;
  if (i < ow / 2) {  // :138-139
    out[t] = __ldg(src + t);
    return;
  }
  const float P1 = 0.3377, P2 = 0.1217, P3 = 0.0439;  // :111
  const int jm = max(j - 1, 0), jp = min(j + 1, oh - 1);
  const int im = max(i - 1, 0), ip = min(i + 1, ow - 1);
  const uint32_t c11 = __ldg(src + (size_t)jm * ow + im), c12 = __ldg(src + (size_t)jm * ow + i),
                 c13 = __ldg(src + (size_t)jm * ow + ip), c21 = __ldg(src + (size_t)j * ow + im),
                 c22 = __ldg(src + t), c23 = __ldg(src + (size_t)j * ow + ip),
                 c31 = __ldg(src + (size_t)jp * ow + im), c32 = __ldg(src + (size_t)jp * ow + i),
                 c33 = __ldg(src + (size_t)jp * ow + ip);
  uint32_t v = 0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    auto ch = [c](uint32_t p) { return (float)((p >> (8 * c)) & 0xffu); };
    const float corners = __fadd_rn(__fadd_rn(__fadd_rn(ch(c11), ch(c13)), ch(c31)), ch(c33));
    const float edges = __fadd_rn(__fadd_rn(__fadd_rn(ch(c12), ch(c21)), ch(c23)), ch(c32));
    const float s = __fadd_rn(__fadd_rn(__fmul_rn(P3, corners), __fmul_rn(P2, edges)),
                              __fmul_rn(P1, ch(c22)));  // :123-136
    v |= ((uint32_t)__float2int_rz(s) & 0xffu) << (8 * c);
  }
  out[t] = v;
}

// The same blur for rows of 16-byte aligned pixel quads (every geometry the reference uses: the
// reduced sizes are multiples of 16).  A lane owns 4 consecutive pixels and walks down a band of
// rows with a rolling three-row window in registers; the two halo pixels of a row come from the
// neighbouring lanes by shuffle.  The tap sums (:123-133) are sums of at most four bytes: exact in
// any order, so they are formed as integers, two channels per register (R | B << 16, and G), and
// vertically first (top + bottom of a column is shared by three output pixels); only the three
// weighted products and their two sums (:123-136) are float operations, rounded one by one like the
// reference's.
#ifndef FOV360_BLUR_ROWS
#define FOV360_BLUR_ROWS 4
#endif
constexpr int kBlurRows = FOV360_BLUR_ROWS;   // rows per warp
constexpr int kBlurChunk = 4;  // rows requested together
constexpr int kBlurWarps = 4;

struct BlurCol {  // one row of the window: 6 columns (left halo, 4 own, right halo)
  uint32_t rb[6], g[6];
};

__device__ __forceinline__ void blur_split(BlurCol &c, int k, uint32_t p) {
  c.rb[k] = p & 0x00ff00ffu;
  c.g[k] = (p >> 8) & 0xffu;
}

// 16-bit sum (<= 1020) in the low / high half of v as a float: splice into the mantissa of 2^23.
__device__ __forceinline__ float half_lo_to_float(uint32_t v) {
  return __fsub_rn(__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7610u)), 8388608.0f);
}
__device__ __forceinline__ float half_hi_to_float(uint32_t v) {
  return __fsub_rn(__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7632u)), 8388608.0f);
}

__global__ void __launch_bounds__(32 * kBlurWarps, 6) img_logpolar_blur4_kernel(
    uint4 *__restrict__ out, int ow, int oh, const uint32_t *__restrict__ src) {
  pdl_trigger();
  pdl_wait();  // nothing to prepare: every value comes from the predecessor's buffer
  const int lane = threadIdx.x, warp = threadIdx.y;
  const int ow4 = ow >> 2;
  const int g = blockIdx.x * 32 + lane;
  const int gc = min(g, ow4 - 1);  // lanes past the last quad repeat it (shuffles stay full-warp)
  const int i0 = gc * 4;
  const int j0 = (blockIdx.y * kBlurWarps + warp) * kBlurRows;
  if (j0 >= oh) return;  // warp-uniform
  const int nrows = min(kBlurRows, oh - j0);
  const int half = ow / 2;
  const bool active = g < ow4;
  uint4 *orow = out + (size_t)j0 * ow4 + gc;
  // a warp whose 128 columns all lie in the inner half only copies (:138-139)
  if ((blockIdx.x * 32 + 31) * 4 + 3 < half) {
    for (int r = 0; r < nrows; ++r, orow += ow4)
      if (active) __stcs(orow, __ldg(reinterpret_cast<const uint4 *>(src) + (size_t)(j0 + r) * ow4 + gc));
    return;
  }
  const float P1 = 0.3377, P2 = 0.1217, P3 = 0.0439;  // :111
  const bool first = lane == 0, last = lane == 31 || gc == ow4 - 1;
  const int il = max(i0 - 1, 0), irr = min(i0 + 4, ow - 1);  // clamp-to-edge halo columns (:112-121)
  struct Raw6 {
    uint32_t p[6];  // left halo, the lane's 4 pixels, right halo
  };
  auto load_row = [&](int j) {
    FOV_CHECK(j, oh, 131);
    FOV_CHECK(gc * 4 + 3, ow, 132);
    FOV_CHECK(il, ow, 133);
    FOV_CHECK(irr, ow, 134);
    const uint32_t *row = src + (size_t)j * ow;
    const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(row) + gc);
    Raw6 r;
    r.p[0] = __shfl_up_sync(0xffffffffu, raw.w, 1);
    r.p[5] = __shfl_down_sync(0xffffffffu, raw.x, 1);
    if (first) r.p[0] = __ldg(row + il);
    if (last) r.p[5] = __ldg(row + irr);
    r.p[1] = raw.x, r.p[2] = raw.y, r.p[3] = raw.z, r.p[4] = raw.w;
    return r;
  };
  auto split = [](const Raw6 &r) {
    BlurCol c;
#pragma unroll
    for (int k = 0; k < 6; ++k) blur_split(c, k, r.p[k]);
    return c;
  };
  BlurCol top = split(load_row(max(j0 - 1, 0)));
  Raw6 raw_mid = load_row(j0);
  BlurCol mid = split(raw_mid);
  for (int r0 = 0; r0 < nrows; r0 += kBlurChunk) {
    // the next kBlurChunk rows are requested before any of them is used
    Raw6 nxt[kBlurChunk];
#pragma unroll
    for (int c = 0; c < kBlurChunk; ++c) nxt[c] = load_row(min(j0 + r0 + c + 1, oh - 1));
#pragma unroll
    for (int c = 0; c < kBlurChunk; ++c, orow += ow4) {
      const BlurCol bot = split(nxt[c]);
      uint32_t vrb[6], vg[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        vrb[k] = top.rb[k] + bot.rb[k];
        vg[k] = top.g[k] + bot.g[k];
      }
      uint32_t px[4];
      const uint32_t centre[4] = {raw_mid.p[1], raw_mid.p[2], raw_mid.p[3], raw_mid.p[4]};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t crb = vrb[k] + vrb[k + 2], cg = vg[k] + vg[k + 2];  // corners
        const uint32_t erb = vrb[k + 1] + mid.rb[k] + mid.rb[k + 2];       // edges
        const uint32_t eg = vg[k + 1] + mid.g[k] + mid.g[k + 2];
        const float s0 = __fadd_rn(__fadd_rn(__fmul_rn(P3, half_lo_to_float(crb)),
                                             __fmul_rn(P2, half_lo_to_float(erb))),
                                   __fmul_rn(P1, byte_to_float<0>(centre[k])));
        const float s1 = __fadd_rn(__fadd_rn(__fmul_rn(P3, half_lo_to_float(cg)),
                                             __fmul_rn(P2, half_lo_to_float(eg))),
                                   __fmul_rn(P1, byte_to_float<1>(centre[k])));
        const float s2 = __fadd_rn(__fadd_rn(__fmul_rn(P3, half_hi_to_float(crb)),
                                             __fmul_rn(P2, half_hi_to_float(erb))),
                                   __fmul_rn(P1, byte_to_float<2>(centre[k])));
        const uint32_t blurred = pack_rgb0(trunc_bits(s0), trunc_bits(s1), trunc_bits(s2));
        px[k] = i0 + k < half ? centre[k] : blurred;  // :138-139
      }
      if (active && r0 + c < nrows) __stcs(orow, make_uint4(px[0], px[1], px[2], px[3]));
      top = mid;
      mid = bot;
      raw_mid = nxt[c];
    }
  }
}

GatherArgs make_gather(uint8_t *out, int ow, int oh, int out_linesize, const uint8_t *src, int W,
                       int H, int src_linesize, float cx, float cy) {
  GatherArgs a;
  a.out = out;
  a.src = src;
  a.ow = ow;
  a.oh = oh;
  a.out_linesize = out_linesize;
  a.obpp = out_linesize / ow;
  a.W = W;
  a.H = H;
  a.src_linesize = src_linesize;
  a.sbpp = src_linesize / W;
  a.word_ok = a.obpp == 4 && a.sbpp == 4 && (out_linesize % 4) == 0 && (src_linesize % 4) == 0 &&
              ((uintptr_t)out % 4) == 0 && ((uintptr_t)src % 4) == 0;
  a.cx = cx;
  a.cy = cy;
  return a;
}

}  // namespace

FOV_DEFINE_BOUNDS_READER(bounds_read_image_sampler)

cudaError_t launch_img_sample_rect(const LaunchCtx &lc, uint8_t *out, int ow, int oh, int out_linesize,
                                   const uint8_t *src, int W, int H, int src_linesize,
                                   const int16_t *xd, const int16_t *yd, float cx, float cy) {
  const dim3 grid((ow + 31) / 32, (oh + 8 * kGatherRows - 1) / (8 * kGatherRows)), block(32, 8);
  KernelScope ks(lc, "img_sample_rect");
  return launch_chained(img_sample_rect_kernel, grid, block, 0, lc.stream,
                        make_gather(out, ow, oh, out_linesize, src, W, H, src_linesize, cx, cy), xd, yd);
}

cudaError_t launch_img_sample_logpolar(const LaunchCtx &lc, uint8_t *out, int ow, int oh,
                                       int out_linesize, const uint8_t *src, int W, int H,
                                       int src_linesize, const float *radius, const float *cs,
                                       const float *sn, float cx, float cy) {
  const dim3 grid((ow + 31) / 32, (oh + 8 * kGatherRows - 1) / (8 * kGatherRows)), block(32, 8);
  KernelScope ks(lc, "img_sample_logpolar");
  return launch_chained(img_sample_logpolar_kernel, grid, block, 0, lc.stream,
                        make_gather(out, ow, oh, out_linesize, src, W, H, src_linesize, cx, cy), radius,
                        cs, sn);
}

cudaError_t launch_img_interpolate_logpolar(const LaunchCtx &lc, uint8_t *out, int W, int H,
                                            const uint8_t *red, int ow, int oh, float cx, float cy,
                                            const LogpolarGrid &grid) {
  LpInterpArgs a;
  a.out = reinterpret_cast<uint32_t *>(out);
  a.red = reinterpret_cast<const uint32_t *>(red);
  a.lntab = grid.d_lntab;
  a.radius = grid.d_radius_f64;
  a.dir = grid.d_dir;
  a.W = W;
  a.H = H;
  a.ow = ow;
  a.oh = oh;
  a.cx = cx;
  a.cy = cy;
  const double K = ow / 20.0;
  a.k0 = K;
  a.k1 = -K / 2.0;
  a.k2 = K / 3.0;
  a.k3 = -K / 4.0;
  const double turns = oh / (2.0 * 3.14159265358979323846);
  a.turns = (float)turns;
  a.zone = logpolar_round_zone(oh);
  a.magic = 0x4B000000u;
  // Rows per CTA, measured on a B200 (tools/sweep_tile_rows.sh, profiles/r02_tile_rows.txt): rows
  // near the gaze cost more than the others (both rounding candidates tested, exact re-evaluation),
  // so short tiles balance better than the per-column set-up they repeat costs - 1080p 0.0219 ms
  // with 8 rows against 0.0244 with 16, 4K 0.0535 (10) against 0.0603 (32), 8K 0.1807 (20) against
  // 0.1833 (32).  A model that sized the tiles to fill whole waves of CTAs (one 4K frame = exactly
  // one wave of 50-row tiles) was 7 % slower than 32 rows.  FOV360_LP_ROWS overrides for sweeps.
  {
    static const int force = [] {
      const char *e = getenv("FOV360_LP_ROWS");
      const int v = e ? atoi(e) : 0;
      return (v >= 2 && v <= kLpMaxRows) ? v : 0;
    }();
    const size_t px = (size_t)W * H;
    a.rows = force ? force : (px < ((size_t)3 << 20) ? 8 : (px < ((size_t)12 << 20) ? 10 : 20));
  }
  const dim3 grid_dim((W + kLpThreads - 1) / kLpThreads, (H + a.rows - 1) / a.rows);
  KernelScope ks(lc, "img_interpolate_logpolar");
  return launch_chained(img_interpolate_logpolar_kernel, grid_dim, dim3(kLpThreads), 0, lc.stream, a);
}

cudaError_t launch_img_logpolar_blur(const LaunchCtx &lc, uint8_t *out, int ow, int oh,
                                     const uint8_t *src) {
  KernelScope ks(lc, "img_logpolar_blur");
  if ((ow & 3) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)src & 15) == 0) {
    const dim3 grid((ow / 4 + 31) / 32, (oh + kBlurWarps * kBlurRows - 1) / (kBlurWarps * kBlurRows)),
        block(32, kBlurWarps);
    return launch_chained(img_logpolar_blur4_kernel, grid, block, 0, lc.stream,
                          reinterpret_cast<uint4 *>(out), ow, oh,
                          reinterpret_cast<const uint32_t *>(src));
  } else {
    const dim3 grid((ow + 31) / 32, (oh + 7) / 8), block(32, 8);
    return launch_chained(img_logpolar_blur_kernel, grid, block, 0, lc.stream,
                          reinterpret_cast<uint32_t *>(out), ow, oh,
                          reinterpret_cast<const uint32_t *>(src));
  }
}

cudaError_t launch_img_logpolar_grid_expand(const LaunchCtx &lc, int16_t *grid, int ow, int oh,
                                            const float *radius, const float *cs,
                                            const float *sn) {
  const dim3 grid_dim((ow + 31) / 32, (oh + 7) / 8), block(32, 8);
  KernelScope ks(lc, "img_logpolar_grid_expand");
  img_logpolar_grid_expand_kernel<<<grid_dim, block, 0, lc.stream>>>(grid, ow, oh, radius, cs, sn);
  return cudaGetLastError();
}

}  // namespace fov
