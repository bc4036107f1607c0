// Pixel arithmetic shared by the inverse-warp kernels (sat_decode.cu, image_sampler.cu): byte <->
// float conversions that stay off the conversion pipe, the reference's un-fused mix(), and packed
// fp32 pairs (sm_100 FADD2 / FMUL2).
#pragma once
#include <stdint.h>

namespace fov {

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

// Packed fp32 pairs (sm_100 FMUL2 / FADD2): two independent IEEE single operations per instruction,
// each rounded exactly like its scalar form.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into
// FFMA2 even though both carry an explicit rounding mode; a multiply with .ftz is never contracted
// with an add without it, and no operand or product here is subnormal (bytes, ratios k/n, and
// their differences are 0 or >= 2^-17 in magnitude), so .ftz changes nothing.
using f32x2 = unsigned long long;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, uint32_t &lo, uint32_t &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 mul2_rn(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2_rn(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 sub2_rn(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// both halves: low mantissa bits of v + 2^23 rounded toward zero (trunc_bits)
__device__ __forceinline__ f32x2 trunc_bits2(f32x2 v) {
  f32x2 r;
  asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(v), "l"(0x4B0000004B000000ull));
  return r;
}
// byte C of two pixels as a float pair (byte_to_float twice, one packed subtract)
template <int C>
__device__ __forceinline__ f32x2 bytes_to_float2(uint32_t a, uint32_t b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};"
      : "=l"(r)
      : "r"(__byte_perm(a, 0x4B000000u, 0x7440u | C)), "r"(__byte_perm(b, 0x4B000000u, 0x7440u | C)));
  return sub2_rn(r, 0x4B0000004B000000ull);
}

// `.xyz =` store (sat_decoder_sample_rect_kernel.cl:212, image_sampler_sample_rect_kernel.cl:38-40): bytes 0..2 of the target pixel, byte 3 keeps its contents.  Two partial
// stores (u16 + u8) instead of a read-modify-write of the pixel: the old pixel is never loaded, so
// no warp waits on the target buffer.
__device__ __forceinline__ void store_xyz(uint32_t *px, uint32_t rgb) {
  asm volatile("st.global.u16 [%0], %1;" ::"l"(px), "h"((unsigned short)(rgb & 0xffffu)) : "memory");
  asm volatile("st.global.u8 [%0+2], %1;" ::"l"(px), "r"((rgb >> 16) & 0xffu) : "memory");
}

// bytes_to_float2 with the 2^23 bit pattern passed in a register (a kernel argument, opaque to
// ptxas): the selector then is the immediate operand of the PRMT.  With a literal pattern ptxas
// parks the selectors in uniform registers and copies one into a vector register per PRMT.
template <int C>
__device__ __forceinline__ f32x2 bytes_to_float2_m(uint32_t a, uint32_t b, uint32_t magic) {
  uint32_t lo, hi;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(lo) : "r"(a), "r"(magic), "n"(0x7440 | C));
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(b), "r"(magic), "n"(0x7440 | C));
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return sub2_rn(r, 0x4B0000004B000000ull);
}

// fused multiply-add on both halves (one rounding each): only for arithmetic whose result is not
// bound to the reference bit for bit (polynomial approximations).
__device__ __forceinline__ f32x2 fma2_rn(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

}  // namespace fov
