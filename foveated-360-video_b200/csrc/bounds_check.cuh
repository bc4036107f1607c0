// Index checks of our own for the gather kernels (compute-sanitizer is closed on the GPU pool this
// library is developed on).  With -DFOV360_BOUNDS_CHECK every table index and every gathered pixel
// coordinate is compared with its limit before it is used; a violation is counted (and the first
// site recorded) in a per-translation-unit device counter that fov_debug_bounds_violations() reads.
// Without the define the checks compile to nothing.  tools/build_variant.sh check
// "-DFOV360_BOUNDS_CHECK" builds the checking library; tests/test_gpu_bounds.py runs the whole
// geometry matrix through it.
#pragma once
#include <cuda_runtime.h>

namespace fov {
#ifdef FOV360_BOUNDS_CHECK
namespace {
__device__ unsigned int g_fov_bounds[2];  // [0] violations, [1] site of the first one
}
__device__ __forceinline__ void bounds_check(long long idx, long long limit, unsigned site) {
  if (idx < 0 || idx >= limit)
    if (atomicAdd(&g_fov_bounds[0], 1u) == 0) g_fov_bounds[1] = site;
}
#define FOV_CHECK(idx, limit, site) ::fov::bounds_check((long long)(idx), (long long)(limit), (site))
#define FOV_DEFINE_BOUNDS_READER(name)                                   \
  void name(unsigned out[2]) {                                           \
    if (cudaMemcpyFromSymbol(out, g_fov_bounds, 8) != cudaSuccess) out[0] = ~0u, out[1] = 0; \
  }
#else
#define FOV_CHECK(idx, limit, site) ((void)0)
#define FOV_DEFINE_BOUNDS_READER(name) \
  void name(unsigned out[2]) { out[0] = out[1] = 0; }
#endif
}  // namespace fov
