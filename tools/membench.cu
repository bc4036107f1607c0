// Streaming-bandwidth probes for the access patterns the SAT kernels use (not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/membench.bin tools/membench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

// read 16 B, write 48 B per thread, fully linear (thread i -> px 4i..4i+3)
__global__ void rw_linear(const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(in + i));
    __stcs(out + 3 * i + 0, make_uint4(v.x, v.y, v.z, v.w));
    __stcs(out + 3 * i + 1, make_uint4(v.y, v.z, v.w, v.x));
    __stcs(out + 3 * i + 2, make_uint4(v.z, v.w, v.x, v.y));
  }
}

// same bytes, SAT tiling: warp = 128 px strip (512 B in / 1536 B out per row), R rows per warp,
// rows W pixels apart
__global__ void rw_tiled(const uint4 *__restrict__ in, uint4 *__restrict__ out, int W, int H, int R,
                         int warps_per_cta) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ns = W / 128;
  const int nsc = (ns + warps_per_cta - 1) / warps_per_cta;
  const int s = (blockIdx.x % nsc) * warps_per_cta + warp;
  const int b = blockIdx.x / nsc;
  if (s >= ns) return;
  const size_t x4 = (size_t)s * 32 + lane;
  for (int y = b * R; y < min(b * R + R, H); ++y) {
    const size_t i = (size_t)y * (W / 4) + x4;
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(in + i));
    __stcs(out + 3 * i + 0, make_uint4(v.x, v.y, v.z, v.w));
    __stcs(out + 3 * i + 1, make_uint4(v.y, v.z, v.w, v.x));
    __stcs(out + 3 * i + 2, make_uint4(v.z, v.w, v.x, v.y));
  }
}

// tiled, but every store instruction of a warp covers 512 contiguous bytes (lane-contiguous)
__global__ void rw_tiled_coalesced(const uint4 *__restrict__ in, uint4 *__restrict__ out, int W,
                                   int H, int R, int warps_per_cta) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ns = W / 128;
  const int nsc = (ns + warps_per_cta - 1) / warps_per_cta;
  const int s = (blockIdx.x % nsc) * warps_per_cta + warp;
  const int b = blockIdx.x / nsc;
  if (s >= ns) return;
  for (int y = b * R; y < min(b * R + R, H); ++y) {
    const size_t i = (size_t)y * (W / 4) + (size_t)s * 32;
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(in + i + lane));
    uint4 *o = out + 3 * i;
    __stcs(o + lane, make_uint4(v.x, v.y, v.z, v.w));
    __stcs(o + 32 + lane, make_uint4(v.y, v.z, v.w, v.x));
    __stcs(o + 64 + lane, make_uint4(v.z, v.w, v.x, v.y));
  }
}

// tiled, rows staged in shared memory and written with one 1536-byte cp.async.bulk per warp-row
__global__ void rw_tiled_tma(const uint4 *__restrict__ in, uint4 *__restrict__ out, int W, int H,
                             int R, int warps_per_cta) {
  __shared__ __align__(128) uint8_t stage[8 * 4 * 1536];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ns = W / 128;
  const int nsc = (ns + warps_per_cta - 1) / warps_per_cta;
  const int s = (blockIdx.x % nsc) * warps_per_cta + warp;
  const int b = blockIdx.x / nsc;
  if (s >= ns) return;
  int buf = 0;
  for (int y = b * R; y < min(b * R + R, H); ++y) {
    const size_t i = (size_t)y * (W / 4) + (size_t)s * 32;
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(in + i + lane));
    uint8_t *sb = stage + (warp * 4 + buf) * 1536;
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
    __syncwarp();
    uint4 *sd = reinterpret_cast<uint4 *>(sb + lane * 48);
    sd[0] = make_uint4(v.x, v.y, v.z, v.w);
    sd[1] = make_uint4(v.y, v.z, v.w, v.x);
    sd[2] = make_uint4(v.z, v.w, v.x, v.y);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 1536;\n\t"
                   "cp.async.bulk.commit_group;" ::"l"(out + 3 * i),
                   "r"((uint32_t)__cvta_generic_to_shared(sb))
                   : "memory");
    buf = (buf + 1) & 3;
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
}

__global__ void copy_linear(const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) out[i] = in[i];
}

__global__ void write_only(uint4 *__restrict__ out, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) __stcs(out + i, make_uint4(i, 1, 2, 3));
}

int main() {
  const int W = 7680, H = 3840, F = 8;
  const size_t npx = (size_t)W * H * F, n4 = npx / 4;
  uint4 *in, *out;
  cudaMalloc(&in, npx * 4);
  cudaMalloc(&out, npx * 12);
  cudaMemset(in, 1, npx * 4);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  auto timeit = [&](const char *name, double bytes, auto launch) {
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    printf("%-34s %8.3f ms  %7.1f GB/s  (%s)\n", name, ms / 10, bytes / (ms / 10) / 1e6,
           cudaGetErrorString(cudaGetLastError()));
  };
  // copy inside the big buffer: first third -> second third (npx*4 bytes each way)
  timeit("copy linear (0.94GB rd + 0.94GB wr)", npx * 4.0 * 2,
         [&] { copy_linear<<<148 * 16, 256>>>(out, out + n4, n4); });
  timeit("write-only linear (2.83GB)", npx * 12.0, [&] { write_only<<<148 * 16, 256>>>(out, n4 * 3); });
  timeit("read4+write12 linear", npx * 16.0, [&] { rw_linear<<<148 * 16, 256>>>(in, out, n4); });
  for (int nw : {6, 8})
    for (int R : {32, 64}) {
      char name[64];
      const int nsc = (W / 128 + nw - 1) / nw, nb = (H * F + R - 1) / R;
      snprintf(name, sizeof name, "r4+w12 tiled coalesced NW=%d R=%d", nw, R);
      timeit(name, npx * 16.0, [&] { rw_tiled_coalesced<<<nsc * nb, nw * 32>>>(in, out, W, H * F, R, nw); });
      snprintf(name, sizeof name, "r4+w12 tiled TMA-store NW=%d R=%d", nw, R);
      timeit(name, npx * 16.0, [&] { rw_tiled_tma<<<nsc * nb, nw * 32>>>(in, out, W, H * F, R, nw); });
    }
  for (int nw : {6})
    for (int R : {32}) {
      char name[64];
      snprintf(name, sizeof name, "read4+write12 tiled NW=%d R=%d", nw, R);
      const int nsc = (W / 128 + nw - 1) / nw, nb = (H * F + R - 1) / R;
      timeit(name, npx * 16.0, [&] { rw_tiled<<<nsc * nb, nw * 32>>>(in, out, W, H * F, R, nw); });
    }
  return 0;
}
