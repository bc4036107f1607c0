// Projections::GnomonicProjection for sm_100a: renders a viewport out of an equirectangular frame.
//
// Replaces gnomonic_kernel (projections_program.cl:7-47; launcher projections.cc:51-86).  One
// thread per viewport pixel: seven transcendentals (evaluated in double, see
// projection_common.cuh), then one 4-byte gather.  Bound by the FP64/SFU work, not by memory:
// the viewport is small (a 1920x1080 viewport moves 16 MB) next to the per-frame foveation path.
#include "fov360_internal.h"
#include "projection_common.cuh"

namespace fov {
namespace {

__global__ void __launch_bounds__(256) gnomonic_kernel(uint32_t *__restrict__ out, int tw, int th,
                                                       const uint32_t *__restrict__ src, int W,
                                                       int H, const GnomonicView v) {
  const int i = blockIdx.x * 32 + threadIdx.x;
  const int j = blockIdx.y * 8 + threadIdx.y;
  if (i >= tw || j >= th) return;
  int sx, sy;
  gnomonic_source(i, j, tw, th, W, H, v, sx, sy);
  // dense uchar3 arrays on both sides (4-byte pixels, width-based indexing, :40-44); the whole
  // 4-byte element is copied
  out[(size_t)j * tw + i] = __ldg(src + (size_t)sy * W + sx);
}

}  // namespace

cudaError_t launch_gnomonic(const LaunchCtx &lc, uint8_t *out, int tw, int th, const uint8_t *src,
                            int W, int H, const GnomonicView &view) {
  const dim3 grid((tw + 31) / 32, (th + 7) / 8), block(32, 8);
  KernelScope ks(lc, "gnomonic");
  gnomonic_kernel<<<grid, block, 0, lc.stream>>>(reinterpret_cast<uint32_t *>(out), tw, th,
                                                 reinterpret_cast<const uint32_t *>(src), W, H,
                                                 view);
  return cudaGetLastError();
}

}  // namespace fov
