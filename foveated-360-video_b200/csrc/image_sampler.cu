// ImageSampler kernels for sm_100a: the no-SAT baseline path of the reference (log-rect point
// sampling, log-polar sampling, log-polar inverse warp, 3x3 blur of the outer half).
//
// Replaces sample_rect_kernel (image_sampler_sample_rect_kernel.cl:1-46),
// sample_logpolar_kernel / logpolar_gaussian_blur_kernel
// (image_sampler_sample_logpolar_kernel.cl:41-142) and interpolate_logpolar_kernel
// (image_sampler_interpolate_kernel.cl:1-81).  The 2-D int16 grids of the reference are
// replaced by their separable 1-D factors (luts.cc); the grid value is re-formed in
// registers, which removes a 4 B/pixel table read from both samplers.
#include "fov360_internal.h"

namespace fov {
namespace {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// (int)(c * dim + delta): float multiply, float add, truncation
// (image_sampler_sample_rect_kernel.cl:26-27, image_sampler_sample_logpolar_kernel.cl:67-70).
__device__ __forceinline__ int gaze_plus(float c, int dim, int delta) {
  return __float2int_rz(__fadd_rn(__fmul_rn(c, (float)dim), (float)delta));
}

// Copies bytes 0..2 of one pixel, leaving byte 3 (and any wider stride) untouched.
__device__ __forceinline__ void copy_rgb(uint8_t *o, const uint8_t *s, bool word_ok) {
  if (word_ok) {
    const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(s));
    uint32_t *ow = reinterpret_cast<uint32_t *>(o);
    *ow = (*ow & 0xff000000u) | (v & 0x00ffffffu);
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

__global__ void __launch_bounds__(256) img_sample_rect_kernel(const GatherArgs a,
                                                              const int16_t *__restrict__ xd,
                                                              const int16_t *__restrict__ yd) {
  const int i = blockIdx.x * 32 + threadIdx.x;
  const int j = blockIdx.y * 8 + threadIdx.y;
  if (i >= a.ow || j >= a.oh) return;
  int x = gaze_plus(a.cx, a.W, xd[i]);
  const int y = gaze_plus(a.cy, a.H, yd[j]);
  if (x >= a.W)  // :29-33
    x -= a.W;
  else if (x < 0)
    x += a.W;
  if (x >= 0 && x < a.W && y >= 0 && y < a.H)  // :35-43
    copy_rgb(a.out + (size_t)j * a.out_linesize + (size_t)i * a.obpp,
             a.src + (size_t)y * a.src_linesize + (size_t)x * a.sbpp, a.word_ok);
}

// Grid value as the reference stores it: int truncation of the float product, narrowed to
// int16 (image_sampler_sample_logpolar_kernel.cl:31-38).
__device__ __forceinline__ int logpolar_delta(float radius, float trig) {
  return (int)(int16_t)__float2int_rz(__fmul_rn(radius, trig));
}

__global__ void __launch_bounds__(256) img_sample_logpolar_kernel(
    const GatherArgs a, const float *__restrict__ radius, const float *__restrict__ cs,
    const float *__restrict__ sn) {
  const int i = blockIdx.x * 32 + threadIdx.x;
  const int j = blockIdx.y * 8 + threadIdx.y;
  if (i >= a.ow || j >= a.oh) return;
  const float r = radius[i];
  int x = gaze_plus(a.cx, a.W, logpolar_delta(r, cs[j]));
  int y = gaze_plus(a.cy, a.H, logpolar_delta(r, sn[j]));
  x = (x + 10 * a.W) % a.W;  // :73
  y = clampi(y, 0, a.H - 1);
  if (x >= 0 && x < a.W)  // :76-77 (x can stay negative only for |x| > 10 W)
    copy_rgb(a.out + (size_t)j * a.out_linesize + (size_t)i * a.obpp,
             a.src + (size_t)y * a.src_linesize + (size_t)x * a.sbpp, a.word_ok);
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

__device__ __forceinline__ float mix_rn(float a, float b, float t) {
  return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), t));
}

__device__ __forceinline__ uint32_t lerp_pixel(uint32_t tl, uint32_t tr, uint32_t bl, uint32_t br,
                                               float tx, float ty) {
  uint32_t outp = 0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float l = mix_rn((float)((tl >> (8 * c)) & 0xffu), (float)((bl >> (8 * c)) & 0xffu), ty);
    const float r = mix_rn((float)((tr >> (8 * c)) & 0xffu), (float)((br >> (8 * c)) & 0xffu), ty);
    outp |= ((uint32_t)__float2int_rz(mix_rn(l, r, tx)) & 0xffu) << (8 * c);
  }
  return outp;
}

// interpolate_logpolar_kernel (image_sampler_interpolate_kernel.cl:1-81).  Not separable: the
// radius logarithm and the angle arctangent are evaluated per pixel (where the reference's typing
// promotes to double the same is done here; single-precision libm calls are evaluated in double
// and rounded once, which reproduces a - nearly always - correctly rounded host libm result).
// The forward map of the exact-hit check (:46-51) only depends on the rounded indices, so its
// exp / cos / sin come from the host-built radius[ow] and direction[oh] tables.
__global__ void __launch_bounds__(256) img_interpolate_logpolar_kernel(
    uint32_t *__restrict__ out, int W, int H, const uint32_t *__restrict__ red, int ow, int oh,
    float cx, float cy, const float *__restrict__ radius, const double2 *__restrict__ dir) {
  const int xx = blockIdx.x * 32 + threadIdx.x;
  const int yy = blockIdx.y * 8 + threadIdx.y;
  if (xx >= W || yy >= H) return;
  const double kPi = 3.14159265358979323846, kPi2 = 1.57079632679489661923;
  const int cxp = __float2int_rz(__fmul_rn(cx, (float)W));  // :19-20
  const int cyp = __float2int_rz(__fmul_rn(cy, (float)H));
  int x = xx;
  const int y = yy;
  if (x - cxp > W / 2)  // :21-25
    x -= W;
  else if (x - cxp < -(W / 2))
    x += W;
  const int dx = x - cxp, dy = y - cyp;
  float i_f = 0.0f;
  if (!(dx == 0 && dy == 0)) {  // :28-33
    const double d2 = (double)dx * (double)dx + (double)dy * (double)dy;
    i_f = (float)((double)ow * (log(sqrt(d2)) / (double)10.0f));
  }
  const int i = clampi((int)roundf(i_f), 0, ow - 1);  // :34
  float j_f;
  if (dx != 0) {  // :36-40
    const float q = __fdiv_rn((float)dy, (float)dx);
    const float at = (float)atan((double)q);
    j_f = (float)(((double)at + kPi * (double)(dx < 0)) * ((double)(float)oh / (2.0 * kPi)));
    // fmod(j_f + 2 oh, oh), :39: the argument is a float in (1.75 oh, 2.75 oh), so the remainder is
    // one or two subtractions of oh, each exact in double (24-bit operand, oh < 2^15)
    double wrapped = (double)__fadd_rn(j_f, (float)(2 * oh));
    const double period = (double)oh;
    while (wrapped >= period) wrapped -= period;
    j_f = (float)wrapped;
  } else {  // :41-43
    j_f = (float)((kPi2 + kPi * (double)(dy < 0)) * ((double)oh / (2.0 * kPi)));
  }
  const int j = clampi((int)roundf(j_f), 0, oh - 1);  // :44
  const float rad = __ldg(radius + i);  // expf(10.0f * powf((float)i / ow, alpha)), :47
  const double2 cs = __ldg(dir + j);    // cos, sin of (float)j / oh * 2.0f * M_PI
  const int calc_x = __double2int_rz(__dadd_rn((double)__fmul_rn(cx, (float)W), __dmul_rn((double)rad, cs.x)));
  const int calc_y = __double2int_rz(__dadd_rn((double)__fmul_rn(cy, (float)H), __dmul_rn((double)rad, cs.y)));
  uint32_t v;
  if (calc_x == x && calc_y == y) {  // :53-55
    v = __ldg(red + (size_t)j * ow + i);
  } else {  // :59-79
    const int min_i = clampi((int)floorf(i_f), 0, ow - 1);
    const int min_j = (int)floorf(__fadd_rn(j_f, (float)oh)) % oh;
    const int max_i = clampi((int)ceilf(i_f), 0, ow - 1);
    const int max_j = (int)ceilf(__fadd_rn(j_f, (float)oh)) % oh;
    const float ir = __fsub_rn(i_f, floorf(i_f)), jr = __fsub_rn(j_f, floorf(j_f));
    v = lerp_pixel(__ldg(red + (size_t)min_j * ow + min_i), __ldg(red + (size_t)min_j * ow + max_i),
                   __ldg(red + (size_t)max_j * ow + min_i), __ldg(red + (size_t)max_j * ow + max_i),
                   ir, jr);
  }
  out[(size_t)yy * W + xx] = v;
}

// logpolar_gaussian_blur_kernel (image_sampler_sample_logpolar_kernel.cl:88-142).
__global__ void __launch_bounds__(256) img_logpolar_blur_kernel(uint32_t *__restrict__ out, int ow,
                                                                int oh,
                                                                const uint32_t *__restrict__ src) {
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

cudaError_t launch_img_sample_rect(const LaunchCtx &lc, uint8_t *out, int ow, int oh, int out_linesize,
                                   const uint8_t *src, int W, int H, int src_linesize,
                                   const int16_t *xd, const int16_t *yd, float cx, float cy) {
  const dim3 grid((ow + 31) / 32, (oh + 7) / 8), block(32, 8);
  KernelScope ks(lc, "img_sample_rect");
  img_sample_rect_kernel<<<grid, block, 0, lc.stream>>>(
      make_gather(out, ow, oh, out_linesize, src, W, H, src_linesize, cx, cy), xd, yd);
  return cudaGetLastError();
}

cudaError_t launch_img_sample_logpolar(const LaunchCtx &lc, uint8_t *out, int ow, int oh,
                                       int out_linesize, const uint8_t *src, int W, int H,
                                       int src_linesize, const float *radius, const float *cs,
                                       const float *sn, float cx, float cy) {
  const dim3 grid((ow + 31) / 32, (oh + 7) / 8), block(32, 8);
  KernelScope ks(lc, "img_sample_logpolar");
  img_sample_logpolar_kernel<<<grid, block, 0, lc.stream>>>(
      make_gather(out, ow, oh, out_linesize, src, W, H, src_linesize, cx, cy), radius, cs, sn);
  return cudaGetLastError();
}

cudaError_t launch_img_interpolate_logpolar(const LaunchCtx &lc, uint8_t *out, int W, int H,
                                            const uint8_t *red, int ow, int oh, float cx,
                                            float cy, const float *radius, const double2 *dir) {
  const dim3 grid((W + 31) / 32, (H + 7) / 8), block(32, 8);
  KernelScope ks(lc, "img_interpolate_logpolar");
  img_interpolate_logpolar_kernel<<<grid, block, 0, lc.stream>>>(
      reinterpret_cast<uint32_t *>(out), W, H, reinterpret_cast<const uint32_t *>(red), ow, oh, cx,
      cy, radius, dir);
  return cudaGetLastError();
}

cudaError_t launch_img_logpolar_blur(const LaunchCtx &lc, uint8_t *out, int ow, int oh,
                                     const uint8_t *src) {
  const dim3 grid((ow + 31) / 32, (oh + 7) / 8), block(32, 8);
  KernelScope ks(lc, "img_logpolar_blur");
  img_logpolar_blur_kernel<<<grid, block, 0, lc.stream>>>(reinterpret_cast<uint32_t *>(out), ow, oh,
                                                  reinterpret_cast<const uint32_t *>(src));
  return cudaGetLastError();
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
