// SAT build for sm_100a: RGB0 u8 frame -> packed u32[H][W][3] summed-area table.
//
// Replaces the reference's copy_image_kernel + scan_rows_kernel + scan_columns_kernel
// (sat_encoder_encode_kernels.cl:1-20,44-74; 40 B/pixel of traffic on top of the u8 read,
// one work-item per row / per column).  Design (see DESIGN.md section "SAT build"):
//
//   reduce-then-scan over warp tiles of R rows x 128 pixels, no inter-CTA spinning:
//     K1 sat_reduce   reads the frame once (4 B/px) and emits, per tile, the column sums
//                     over the band, the per-row sums over the strip and the tile total;
//     K2 sat_carry    turns those into carries: exclusive scan of the column sums over
//                     bands, exclusive scan of the row sums over strips, and the 2-D
//                     exclusive prefix of the tile totals (the corner term);
//     K3 sat_scan     re-reads the frame (L2 hit up to ~4K, 4 B/px of DRAM at 8K), does a
//                     warp-shuffle inclusive scan per row, accumulates down the band in
//                     registers and writes the SAT exactly once (12 B/px).
//   DRAM traffic: 16 B/px + (4 B/px when the frame does not survive in L2) + carry tables
//   (a few % of the frame).  All sums are plain u32 adds, i.e. they wrap mod 2^32 exactly
//   like the reference's `uint currentSum`.
//
//   Each lane owns 4 consecutive pixels (one 16-byte load, three 16-byte stores).  The SAT
//   row segment of a warp (1536 B) is either stored straight from registers or staged in
//   shared memory and written with one cp.async.bulk (TMA bulk store) per row.
#include <cstdio>
#include <cstdlib>

#include "fov360_internal.h"
#include "sat_common.cuh"

namespace fov {
namespace {

constexpr int kWarps = 8;  // warps per CTA (side by side along x)

struct SatPlan {
  int n, W, H, linesize, bpp;
  int R;   // rows per band
  int nb;  // bands
  int ns;  // strips
  size_t colsum_off, rowsum_off, tile_off, per_frame;  // scratch layout (bytes)
};

SatPlan make_plan(int n, int W, int H, int linesize, int sm_count) {
  SatPlan p;
  p.n = n;
  p.W = W;
  p.H = H;
  p.linesize = linesize;
  p.bpp = linesize / W;
  p.ns = (W + kStripPx - 1) / kStripPx;
  // Aim at >= ~2 waves of 16 resident warps per SM; taller bands mean smaller carry tables.
  const long long warp_rows = (long long)n * H * p.ns;
  const long long target_tiles = (long long)sm_count * 32;
  long long R = warp_rows / target_tiles;
  R = (R / 8) * 8;
  if (R < 8) R = 8;
  if (R > 128) R = 128;
  if (const char *e = getenv("FOV360_SAT_BAND_ROWS")) {
    int v = atoi(e);
    if (v >= 1 && v <= 4096) R = v;
  }
  p.R = (int)R;
  p.nb = (H + p.R - 1) / p.R;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  p.colsum_off = 0;
  p.rowsum_off = al((size_t)p.nb * W * 3 * 4);
  p.tile_off = p.rowsum_off + al((size_t)p.ns * H * 16);
  p.per_frame = p.tile_off + al((size_t)p.nb * p.ns * 16);
  return p;
}

struct SatArgs {
  const uint8_t *src;
  uint32_t *sat;
  uint8_t *scratch;
  size_t src_stride, sat_stride, scratch_stride;
  size_t colsum_off, rowsum_off, tile_off;
  int W, H, linesize, bpp, R, nb, ns;
};

// ---------------------------------------------------------------------------------------
// K1: per-tile reductions.
// ---------------------------------------------------------------------------------------
template <bool FAST>
__global__ void __launch_bounds__(kWarps * 32) sat_reduce_kernel(const SatArgs a) {
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (s >= a.ns) return;
  const int b = blockIdx.y, f = blockIdx.z;
  const int x0 = s * kStripPx + lane * 4;
  const int y0 = b * a.R;
  const int y1 = min(y0 + a.R, a.H);
  const uint8_t *src = a.src + (size_t)f * a.src_stride;
  uint8_t *scr = a.scratch + (size_t)f * a.scratch_stride;
  uint4 *rowsum = reinterpret_cast<uint4 *>(scr + a.rowsum_off) + (size_t)s * a.H;

  uint32_t acc[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i] = 0;

  for (int yc = y0; yc < y1; yc += 32) {
    uint32_t k0 = 0, k1 = 0, k2 = 0;
    const int rows = min(32, y1 - yc);
#pragma unroll 4
    for (int r = 0; r < rows; ++r) {
      uint32_t p[12];
      unpack_px4(load_px4<FAST>(src + (size_t)(yc + r) * a.linesize, x0, a.W, a.bpp), p);
#pragma unroll
      for (int i = 0; i < 12; ++i) acc[i] += p[i];
      const uint32_t t0 = __reduce_add_sync(0xffffffffu, p[0] + p[3] + p[6] + p[9]);
      const uint32_t t1 = __reduce_add_sync(0xffffffffu, p[1] + p[4] + p[7] + p[10]);
      const uint32_t t2 = __reduce_add_sync(0xffffffffu, p[2] + p[5] + p[8] + p[11]);
      if (lane == r) {
        k0 = t0;
        k1 = t1;
        k2 = t2;
      }
    }
    if (lane < rows) rowsum[yc + lane] = make_uint4(k0, k1, k2, 0);
  }

  // Column sums of this band (raw, not yet scanned over bands).
  uint32_t *colsum = reinterpret_cast<uint32_t *>(scr + a.colsum_off) + ((size_t)b * a.W) * 3;
  if (FAST) {
    if (x0 < a.W) {
      uint4 *d = reinterpret_cast<uint4 *>(colsum + (size_t)x0 * 3);
      d[0] = make_uint4(acc[0], acc[1], acc[2], acc[3]);
      d[1] = make_uint4(acc[4], acc[5], acc[6], acc[7]);
      d[2] = make_uint4(acc[8], acc[9], acc[10], acc[11]);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (x0 + k < a.W) {
        colsum[(size_t)(x0 + k) * 3 + 0] = acc[3 * k + 0];
        colsum[(size_t)(x0 + k) * 3 + 1] = acc[3 * k + 1];
        colsum[(size_t)(x0 + k) * 3 + 2] = acc[3 * k + 2];
      }
  }
  // Tile total.
  const uint32_t t0 = __reduce_add_sync(0xffffffffu, acc[0] + acc[3] + acc[6] + acc[9]);
  const uint32_t t1 = __reduce_add_sync(0xffffffffu, acc[1] + acc[4] + acc[7] + acc[10]);
  const uint32_t t2 = __reduce_add_sync(0xffffffffu, acc[2] + acc[5] + acc[8] + acc[11]);
  if (lane == 0)
    reinterpret_cast<uint4 *>(scr + a.tile_off)[(size_t)b * a.ns + s] = make_uint4(t0, t1, t2, 0);
}

// ---------------------------------------------------------------------------------------
// K2: carries.  Three independent roles share one launch:
//   blocks [0, nA)        rowsum[s][y]  -> exclusive scan over strips s         (left carry)
//   blocks [nA, nA+nB)    colsum[b][e]  -> exclusive scan over bands b          (top carry)
//   blocks [nA+nB, +n)    tile[b][s]    -> 2-D exclusive prefix                 (corner)
// ---------------------------------------------------------------------------------------
constexpr int kCarryThreads = 256;

__global__ void __launch_bounds__(kCarryThreads) sat_carry_kernel(const SatArgs a, int n, int nA,
                                                                  int nB) {
  const int blk = blockIdx.x;
  if (blk < nA) {
    const long long t = (long long)blk * kCarryThreads + threadIdx.x;
    if (t >= (long long)n * a.H) return;
    const int f = (int)(t / a.H), y = (int)(t % a.H);
    uint4 *rs = reinterpret_cast<uint4 *>(a.scratch + (size_t)f * a.scratch_stride + a.rowsum_off);
    uint32_t r0 = 0, r1 = 0, r2 = 0;
#pragma unroll 4
    for (int s = 0; s < a.ns; ++s) {
      uint4 *q = rs + (size_t)s * a.H + y;
      const uint4 v = *q;
      *q = make_uint4(r0, r1, r2, 0);
      r0 += v.x;
      r1 += v.y;
      r2 += v.z;
    }
  } else if (blk < nA + nB) {
    const int row_elems = a.W * 3;
    const long long t = (long long)(blk - nA) * kCarryThreads + threadIdx.x;
    if (t >= (long long)n * row_elems) return;
    const int f = (int)(t / row_elems), e = (int)(t % row_elems);
    uint32_t *cs =
        reinterpret_cast<uint32_t *>(a.scratch + (size_t)f * a.scratch_stride + a.colsum_off);
    uint32_t run = 0;
#pragma unroll 4
    for (int b = 0; b < a.nb; ++b) {
      uint32_t *q = cs + (size_t)b * row_elems + e;
      const uint32_t v = *q;
      *q = run;
      run += v;
    }
  } else {
    const int f = blk - nA - nB;
    uint4 *tt = reinterpret_cast<uint4 *>(a.scratch + (size_t)f * a.scratch_stride + a.tile_off);
    // exclusive over bands, per strip
    for (int s = threadIdx.x; s < a.ns; s += kCarryThreads) {
      uint32_t r0 = 0, r1 = 0, r2 = 0;
      for (int b = 0; b < a.nb; ++b) {
        uint4 *q = tt + (size_t)b * a.ns + s;
        const uint4 v = *q;
        *q = make_uint4(r0, r1, r2, 0);
        r0 += v.x;
        r1 += v.y;
        r2 += v.z;
      }
    }
    __syncthreads();
    // exclusive over strips, per band
    for (int b = threadIdx.x; b < a.nb; b += kCarryThreads) {
      uint32_t r0 = 0, r1 = 0, r2 = 0;
      for (int s = 0; s < a.ns; ++s) {
        uint4 *q = tt + (size_t)b * a.ns + s;
        const uint4 v = *q;
        *q = make_uint4(r0, r1, r2, 0);
        r0 += v.x;
        r1 += v.y;
        r2 += v.z;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// K3: scan + write.
// ---------------------------------------------------------------------------------------
template <bool FAST, bool TMA_STORE>
__global__ void __launch_bounds__(kWarps * 32) sat_scan_kernel(const SatArgs a) {
  __shared__ __align__(128) uint8_t stage[TMA_STORE ? kWarps * kStageBufs * kRowBytes : 16];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int s = blockIdx.x * kWarps + warp;
  if (s >= a.ns) return;
  const int b = blockIdx.y, f = blockIdx.z;
  const int x0 = s * kStripPx + lane * 4;
  const int y0 = b * a.R;
  const int y1 = min(y0 + a.R, a.H);
  const uint8_t *src = a.src + (size_t)f * a.src_stride;
  const uint8_t *scr = a.scratch + (size_t)f * a.scratch_stride;
  uint32_t *sat = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(a.sat) +
                                               (size_t)f * a.sat_stride);
  const uint4 *rowcarry = reinterpret_cast<const uint4 *>(scr + a.rowsum_off) + (size_t)s * a.H;
  const bool in_x = x0 < a.W;

  // Top carry: strip-local inclusive scan along x of the band-exclusive column sums.
  uint32_t acc[12];
  {
    const uint32_t *e = reinterpret_cast<const uint32_t *>(scr + a.colsum_off) +
                        ((size_t)b * a.W) * 3;
    if (FAST) {
      uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0, v2 = v0;
      if (in_x) {
        const uint4 *q = reinterpret_cast<const uint4 *>(e + (size_t)x0 * 3);
        v0 = q[0];
        v1 = q[1];
        v2 = q[2];
      }
      acc[0] = v0.x, acc[1] = v0.y, acc[2] = v0.z, acc[3] = v0.w;
      acc[4] = v1.x, acc[5] = v1.y, acc[6] = v1.z, acc[7] = v1.w;
      acc[8] = v2.x, acc[9] = v2.y, acc[10] = v2.z, acc[11] = v2.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          acc[3 * k + c] = (x0 + k < a.W) ? e[(size_t)(x0 + k) * 3 + c] : 0u;
    }
#pragma unroll
    for (int k = 1; k < 4; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[3 * k + c] += acc[3 * (k - 1) + c];
    uint32_t i0 = acc[9], i1 = acc[10], i2 = acc[11];
    const uint32_t t0 = i0, t1 = i1, t2 = i2;
    warp_scan3(i0, i1, i2, lane);
    i0 -= t0, i1 -= t1, i2 -= t2;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      acc[3 * k + 0] += i0;
      acc[3 * k + 1] += i1;
      acc[3 * k + 2] += i2;
    }
  }
  // Corner term, injected through lane 0's first row carry.
  uint4 corner = make_uint4(0, 0, 0, 0);
  if (lane == 0)
    corner = reinterpret_cast<const uint4 *>(scr + a.tile_off)[(size_t)b * a.ns + s];

  const int strip_px = min(kStripPx, a.W - s * kStripPx);
  uint8_t *my_stage = stage + (TMA_STORE ? (size_t)warp * kStageBufs * kRowBytes : 0);
  int buf = 0;

  constexpr int U = 4;
  uint4 p[U];
  uint4 cr[U];
  // prefetch first group
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int y = min(y0 + u, y1 - 1);
    p[u] = load_px4<FAST>(src + (size_t)y * a.linesize, x0, a.W, a.bpp);
    cr[u] = (lane == 0) ? rowcarry[y] : make_uint4(0, 0, 0, 0);
  }
  if (lane == 0) {
    cr[0].x += corner.x;
    cr[0].y += corner.y;
    cr[0].z += corner.z;
  }

  for (int y = y0; y < y1; y += U) {
    uint4 q[U];
    uint4 qc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      q[u] = p[u];
      qc[u] = cr[u];
    }
    // prefetch the next group while this one is being scanned
    if (y + U < y1) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int yn = min(y + U + u, y1 - 1);
        p[u] = load_px4<FAST>(src + (size_t)yn * a.linesize, x0, a.W, a.bpp);
        cr[u] = (lane == 0) ? rowcarry[yn] : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (y + u < y1) {
        uint32_t v[12];
        unpack_px4(q[u], v);
        // thread-local inclusive prefix over the lane's 4 pixels (+ left carry in lane 0)
        v[0] += qc[u].x, v[1] += qc[u].y, v[2] += qc[u].z;
#pragma unroll
        for (int k = 1; k < 4; ++k)
#pragma unroll
          for (int c = 0; c < 3; ++c) v[3 * k + c] += v[3 * (k - 1) + c];
        uint32_t i0 = v[9], i1 = v[10], i2 = v[11];
        const uint32_t t0 = i0, t1 = i1, t2 = i2;
        warp_scan3(i0, i1, i2, lane);
        i0 -= t0, i1 -= t1, i2 -= t2;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          acc[3 * k + 0] += v[3 * k + 0] + i0;
          acc[3 * k + 1] += v[3 * k + 1] + i1;
          acc[3 * k + 2] += v[3 * k + 2] + i2;
        }
        uint32_t *drow = sat + ((size_t)(y + u) * a.W) * 3;
        if (TMA_STORE) {
          // stage the 1536-byte row segment, then one bulk async store per row
          uint8_t *sb = my_stage + (size_t)buf * kRowBytes;
          if (lane == 0)
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kStageBufs - 1) : "memory");
          __syncwarp();
          uint4 *sd = reinterpret_cast<uint4 *>(sb + lane * 48);
          sd[0] = make_uint4(acc[0], acc[1], acc[2], acc[3]);
          sd[1] = make_uint4(acc[4], acc[5], acc[6], acc[7]);
          sd[2] = make_uint4(acc[8], acc[9], acc[10], acc[11]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            asm volatile(
                "cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\t"
                "cp.async.bulk.commit_group;" ::"l"(drow + (size_t)s * kStripPx * 3),
                "r"(smem_u32(sb)), "r"(strip_px * 12)
                : "memory");
          }
          buf = (buf + 1 == kStageBufs) ? 0 : buf + 1;
        } else if (FAST) {
          if (in_x) {
            uint4 *d = reinterpret_cast<uint4 *>(drow + (size_t)x0 * 3);
            __stcs(d + 0, make_uint4(acc[0], acc[1], acc[2], acc[3]));
            __stcs(d + 1, make_uint4(acc[4], acc[5], acc[6], acc[7]));
            __stcs(d + 2, make_uint4(acc[8], acc[9], acc[10], acc[11]));
          }
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (x0 + k < a.W) {
              drow[(size_t)(x0 + k) * 3 + 0] = acc[3 * k + 0];
              drow[(size_t)(x0 + k) * 3 + 1] = acc[3 * k + 1];
              drow[(size_t)(x0 + k) * 3 + 2] = acc[3 * k + 2];
            }
        }
      }
    }
  }
  if (TMA_STORE) {
    // shared memory must outlive the in-flight bulk reads
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
  }
}

}  // namespace

size_t sat_scratch_bytes(int n, int W, int H) {
  // worst case over the band heights make_plan() may pick (R >= 8 unless overridden to less)
  size_t worst = 0;
  for (int sm : {1, 148, 1 << 20}) {
    SatPlan p = make_plan(n, W, H, 4 * W, sm);
    if (p.per_frame > worst) worst = p.per_frame;
  }
  return worst * (size_t)n;
}

cudaError_t launch_sat_encode(const LaunchCtx &lc, int n, uint32_t *sat, size_t sat_stride,
                              const uint8_t *src, size_t src_stride, int W, int H, int linesize,
                              void *scratch) {
  cudaStream_t st = lc.stream;
  const SatPlan p = make_plan(n, W, H, linesize, lc.sm_count);
  SatArgs a;
  a.src = src;
  a.sat = sat;
  a.scratch = static_cast<uint8_t *>(scratch);
  a.src_stride = src_stride;
  a.sat_stride = sat_stride;
  a.scratch_stride = p.per_frame;
  a.colsum_off = p.colsum_off;
  a.rowsum_off = p.rowsum_off;
  a.tile_off = p.tile_off;
  a.W = W;
  a.H = H;
  a.linesize = linesize;
  a.bpp = p.bpp;
  a.R = p.R;
  a.nb = p.nb;
  a.ns = p.ns;

  const bool fast = p.bpp == 4 && (W % 4) == 0 && (linesize % 16) == 0 &&
                    ((uintptr_t)src % 16) == 0 && (src_stride % 16) == 0 &&
                    ((uintptr_t)sat % 16) == 0 && (sat_stride % 16) == 0;
  static const bool tma_store = [] {
    const char *e = getenv("FOV360_SAT_TMA_STORE");
    return e ? atoi(e) != 0 : true;
  }();

  const dim3 grid((p.ns + kWarps - 1) / kWarps, p.nb, n);
  const dim3 block(kWarps * 32);
  {
    KernelScope ks(lc, "sat_reduce");
    if (fast)
      sat_reduce_kernel<true><<<grid, block, 0, st>>>(a);
    else
      sat_reduce_kernel<false><<<grid, block, 0, st>>>(a);
  }

  const int nA = (int)(((long long)n * H + kCarryThreads - 1) / kCarryThreads);
  const int nB = (int)(((long long)n * W * 3 + kCarryThreads - 1) / kCarryThreads);
  {
    KernelScope ks(lc, "sat_carry");
    sat_carry_kernel<<<nA + nB + n, kCarryThreads, 0, st>>>(a, n, nA, nB);
  }

  KernelScope ks(lc, "sat_scan");
  if (fast && tma_store)
    sat_scan_kernel<true, true><<<grid, block, 0, st>>>(a);
  else if (fast)
    sat_scan_kernel<true, false><<<grid, block, 0, st>>>(a);
  else
    sat_scan_kernel<false, false><<<grid, block, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace fov
