// TEST INFRASTRUCTURE ONLY - not part of the product path.
//
// Minimal OpenCL-C 1.2 -> C++17 host shim.  It lets g++ compile the reference's
// own kernel sources (read in place from /root/reference/src/*.cl at build time;
// never copied into this repo) so that oracle/_ref/libfovref.so *is* the
// reference's arithmetic executed on the host.  The only textual rewrite applied
// to the sources is the OpenCL vector-literal cast `(int2)(a, b)` -> `int2(a, b)`
// (done by oracle/build_oracle.py into a temp dir that is deleted afterwards),
// because in C++ the former parses as a cast of a comma expression.
//
// NDRange semantics: the driver (ref_entry.cc) sets the work-item ids in
// thread-local storage and calls the kernel function once per work-item.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#define __kernel
#define __global
#define __constant const
#define __local

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;

// ---- work-item functions -------------------------------------------------
struct ClshimItem {
  int gid[3];
  int gsz[3];
  int lid[3];
  int lsz[3];
};
extern thread_local ClshimItem clshim_item;
static inline int get_global_id(int d) { return clshim_item.gid[d]; }
static inline int get_global_size(int d) { return clshim_item.gsz[d]; }
static inline int get_local_id(int d) { return clshim_item.lid[d]; }
static inline int get_local_size(int d) { return clshim_item.lsz[d]; }

// ---- vector types ----------------------------------------------------------
struct int2 {
  int x, y;
  int2() : x(0), y(0) {}
  // OpenCL converts each scalar initialiser to the component type (float -> int
  // truncates toward zero), which is what the template does here.
  template <class A, class B>
  int2(A a, B b) : x((int)a), y((int)b) {}
};
static inline int2 operator+(int2 a, int2 b) { return int2(a.x + b.x, a.y + b.y); }

struct short2 {
  short x, y;
};
static inline int2 convert_int2(short2 s) { return int2((int)s.x, (int)s.y); }

struct float2 {
  float x, y;
  float2() : x(0), y(0) {}
  // OpenCL converts each scalar initialiser to float; one scalar is replicated.
  template <class A>
  explicit float2(A s) : x((float)s), y((float)s) {}
  template <class A, class B>
  float2(A a, B b) : x((float)a), y((float)b) {}
};
static inline float2 operator*(float2 a, float2 b) { return float2(a.x * b.x, a.y * b.y); }
static inline float2 operator-(float2 a, float2 b) { return float2(a.x - b.x, a.y - b.y); }

struct uint3 {
  uint x, y, z;
  uint3() : x(0), y(0), z(0) {}
  explicit uint3(uint s) : x(s), y(s), z(s) {}
  uint3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
};
static inline uint3 operator+(uint3 a, uint3 b) { return uint3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline uint3 operator-(uint3 a, uint3 b) { return uint3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline uint3 operator/(uint3 a, uint3 b) { return uint3(a.x / b.x, a.y / b.y, a.z / b.z); }

struct float3 {
  float x, y, z;
  float3() : x(0), y(0), z(0) {}
  float3(float a, float b, float c) : x(a), y(b), z(c) {}
};
static inline float3 operator+(float3 a, float3 b) { return float3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline float3 operator-(float3 a, float3 b) { return float3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline float3 operator*(float s, float3 a) { return float3(s * a.x, s * a.y, s * a.z); }
static inline float3 operator*(float3 a, float s) { return float3(a.x * s, a.y * s, a.z * s); }

// OpenCL 3-component vectors have the size and alignment of 4-component ones.
struct alignas(4) uchar3 {
  uchar x, y, z, pad;
  uchar3() : x(0), y(0), z(0), pad(0) {}
  uchar3(uchar a, uchar b, uchar c) : x(a), y(b), z(c), pad(0) {}
};
static_assert(sizeof(uchar3) == 4, "uchar3 must be 4 bytes");

// uchar4 whose `.xyz` swizzle is assignable and leaves `.w` untouched.
struct alignas(4) uchar4 {
  struct XYZ {
    uchar v[3];
    XYZ &operator=(const uchar3 &c) {
      v[0] = c.x;
      v[1] = c.y;
      v[2] = c.z;
      return *this;
    }
  };
  union {
    struct {
      uchar x, y, z, w;
    };
    XYZ xyz;
  };
};
static_assert(sizeof(uchar4) == 4, "uchar4 must be 4 bytes");

// ---- conversions / loads ---------------------------------------------------
static inline uint3 vload3(size_t offset, const uint *p) {
  return uint3(p[3 * offset], p[3 * offset + 1], p[3 * offset + 2]);
}
// convert_uchar3 without _sat: integer sources wrap modulo 256, float sources
// truncate toward zero (values on this path are always within [0, 255]).
static inline uchar3 convert_uchar3(uint3 v) { return uchar3((uchar)v.x, (uchar)v.y, (uchar)v.z); }
static inline uchar3 convert_uchar3(float3 v) {
  return uchar3((uchar)(int)v.x, (uchar)(int)v.y, (uchar)(int)v.z);
}
static inline float3 convert_float3(uchar3 v) { return float3((float)v.x, (float)v.y, (float)v.z); }

// ---- common / math builtins -------------------------------------------------
// mix(x, y, a) = x + (y - x) * a  (OpenCL 1.2 spec 6.12.4); evaluated as two
// separately rounded float operations (the build uses -ffp-contract=off).
static inline float3 mix(float3 a, float3 b, float t) {
  return float3(a.x + (b.x - a.x) * t, a.y + (b.y - a.y) * t, a.z + (b.z - a.z) * t);
}
static inline int clamp(int v, int lo, int hi) { return std::min(std::max(v, lo), hi); }
static inline uint clamp(uint v, uint lo, uint hi) { return std::min(std::max(v, lo), hi); }
static inline float clamp(float v, float lo, float hi) { return std::fmin(std::fmax(v, lo), hi); }
static inline float2 clamp(float2 v, float2 lo, float2 hi) {
  return float2(clamp(v.x, lo.x, hi.x), clamp(v.y, lo.y, hi.y));
}
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline float max(float a, float b) { return std::fmax(a, b); }
static inline float min(float a, float b) { return std::fmin(a, b); }

using std::abs;
using std::asin;
using std::atan;
using std::atan2;
using std::ceil;
using std::cos;
using std::exp;
using std::floor;
using std::fmod;
using std::log;
using std::pow;
using std::round;
using std::sin;
using std::sqrt;
