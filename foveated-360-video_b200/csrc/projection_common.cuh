// Inverse gnomonic projection shared by the plain viewport gather (projections.cu) and the fused
// reduced-buffer -> viewport warp (sat_decode.cu).  Replaces the body of gnomonic_kernel
// (projections_program.cl:7-47).
#pragma once
#include <stdint.h>

#include "fov360_internal.h"

namespace fov {

// The reference calls the float overloads (atanf, sinf, ...).  Evaluating the double function on
// the float argument and rounding once gives the correctly rounded float result; it differs from
// a float libm only where that libm itself is off by an ulp.
__device__ __forceinline__ float f_atan(float v) { return (float)atan((double)v); }
__device__ __forceinline__ float f_sin(float v) { return (float)sin((double)v); }
__device__ __forceinline__ float f_cos(float v) { return (float)cos((double)v); }
__device__ __forceinline__ float f_asin(float v) { return (float)asin((double)v); }
__device__ __forceinline__ float f_atan2(float a, float b) { return (float)atan2((double)a, (double)b); }

// Source pixel (sx, sy) of viewport pixel (i, j); every float operation is rounded separately,
// the mixed float/double typing follows the OpenCL C source (PI and PI_2 are double literals).
__device__ __forceinline__ void gnomonic_source(int i, int j, int tw, int th, int W, int H,
                                                const GnomonicView v, int &sx, int &sy) {
  const double PI = 3.141592653589793, PI_2 = 1.5707963267948966;
  const float uu = __fdiv_rn((float)i, (float)tw), vv = __fdiv_rn((float)j, (float)th);  // :21
  const float x = __fmul_rn(6.0f, __fsub_rn(uu, 0.5f));  // :19, :22-23
  const float y = __fmul_rn(3.0f, __fsub_rn(vv, 0.5f));
  const float rho = __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));  // :29
  const float c = f_atan(rho);                                                 // :30
  const float sc = f_sin(c), cc = f_cos(c);
  float phi = f_asin(__fadd_rn(__fmul_rn(cc, v.sin_phi1),
                               __fdiv_rn(__fmul_rn(__fmul_rn(y, sc), v.cos_phi1), rho)));  // :31
  float lambda = __fadd_rn(
      v.lambda0, f_atan2(__fmul_rn(x, sc),
                         __fsub_rn(__fmul_rn(__fmul_rn(rho, v.cos_phi1), cc),
                                   __fmul_rn(__fmul_rn(y, v.sin_phi1), sc))));  // :32-34
  phi = (float)fmod(((double)phi + PI_2) + 10 * PI, 2 * PI);      // :35
  lambda = (float)fmod(((double)lambda + PI) + 10 * PI, 2 * PI);  // :36
  float su = (float)((double)lambda / (2.0 * PI)), sv = (float)((double)phi / PI);  // :37
  su = fminf(fmaxf(su, 0.0f), 0.999f);  // :38 (fmin(fmax()): the NaN of the centre pixel becomes 0)
  sv = fminf(fmaxf(sv, 0.0f), 0.999f);
  sx = __float2int_rz(__fmul_rn(su, (float)W));  // :40-41
  sy = __float2int_rz(__fmul_rn(sv, (float)H));
}

}  // namespace fov
