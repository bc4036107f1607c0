// Single-pass SAT build for sm_100a: RGB0 u8 frame -> packed u32[H][W][3] summed-area table with
// 16 B/pixel of DRAM traffic (4 read + 12 written) - the algorithmic minimum.
//
// Replaces the reference's copy_image_kernel + scan_rows_kernel + scan_columns_kernel
// (sat_encoder_encode_kernels.cl:1-20,44-74: 4 + 12 + 24 + 24 = 64 B/pixel over three passes,
// one work-item per row / per column).  See DESIGN.md "SAT build".
//
// Decomposition.  The frame is cut into tiles of R rows x (NW warps x 128 pixels); one CTA owns
// one tile and each warp one 128-pixel strip of it (a lane owns 4 consecutive pixels: one 16-byte
// load, three 16-byte stores per row).  For a pixel (y, x) of tile (band b, strip s):
//
//   S(y, x) = T_b(x)                                  SAT row just above the tile
//           + sum_{y' = y0..y} [ left(y') + local row prefix(y', x) ]
//
// where left(y') = sum of row y' over all columns left of the warp strip.  A tile therefore needs
// two carries from other tiles:
//   * left(y')  - a vector over the tile's R rows: every CTA publishes its per-row sums right
//                 after reading its tile; a CTA adds up the vectors of the (few) CTAs to its left
//                 in the same band.  No chain: the addends do not depend on anybody's carries.
//   * T_b(x)    - a vector over the strip's columns, carried down each warp-strip column by
//                 decoupled look-back: a warp publishes its band aggregate
//                 G_b(x) = T_{b+1}(x) - T_b(x) as soon as its left carry is known (state AGG),
//                 walks up the column adding aggregates until it meets a strip whose T is final
//                 (state INC), and then publishes its own inclusive value.  The inclusive value
//                 T_{b+1}(x) IS the last SAT row of the tile, so it is written straight into the
//                 output and costs no extra traffic; only the aggregates use scratch memory.
// Tiles are handed out through an atomic ticket in (band, frame, strip) order, so every tile a
// CTA waits for has an earlier ticket and is already resident or finished: no deadlock, no
// co-residency assumption beyond what a running CTA guarantees.
//
// The tile is read twice - once to reduce (DRAM), once to scan (L2 hit a few microseconds
// later) - and written once through a per-warp shared-memory ring with cp.async.bulk (TMA bulk
// store, L2 evict-first so the output stream does not push the input tiles out of L2).
// All sums are plain u32 adds: they wrap mod 2^32 exactly like the reference's `uint currentSum`.
#include <cstdio>
#include <cstdlib>

#include "fov360_internal.h"
#include "sat_common.cuh"

namespace fov {
namespace {

#ifndef FOV360_ONEPASS_MIN_CTAS
#define FOV360_ONEPASS_MIN_CTAS 3
#endif

constexpr int kMaxWarps = 8;
constexpr int kMaxBandRows = 64;
constexpr uint32_t kAgg = 1, kInc = 2;  // column-carry states (low 2 bits of a flag word)

struct OnePassArgs {
  const uint8_t *src;
  uint32_t *sat;
  size_t src_stride, sat_stride;
  int W, H, linesize;
  int n, R, nb, ns, nsc;  // frames, band rows, bands, warp strips, CTA strips
  uint32_t epoch, total_tiles;
  int debug_nowait;  // timing experiments only: skip the carry waits (WRONG results)
  uint32_t *counters;  // [0] ticket, [1] finished CTAs
  uint32_t *flag_left;  // [tile]               == epoch when rowagg[tile] is published
  uint4 *rowagg;        // [tile][kMaxBandRows] per-row sums of a CTA tile
  uint32_t *flag_col;   // [tile][NW]           epoch << 2 | state
  uint32_t *colagg;     // [tile][NW][384]      band aggregate G of a warp strip
};

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void st_release(uint32_t *p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// 128-bit streaming load with an explicit L2 eviction policy (createpolicy handle).
__device__ __forceinline__ uint4 ldg_hint(const uint4 *p, uint64_t policy) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(policy));
  return r;
}

__device__ __forceinline__ uint4 load_row4(const uint8_t *row, int x0, int W, uint64_t policy) {
  uint4 v = make_uint4(0, 0, 0, 0);
  if (x0 < W) v = ldg_hint(reinterpret_cast<const uint4 *>(row + (size_t)x0 * 4), policy);
  return v;
}

__device__ __forceinline__ void add12(uint32_t (&d)[12], const uint4 a, const uint4 b,
                                      const uint4 c) {
  d[0] += a.x, d[1] += a.y, d[2] += a.z, d[3] += a.w;
  d[4] += b.x, d[5] += b.y, d[6] += b.z, d[7] += b.w;
  d[8] += c.x, d[9] += c.y, d[10] += c.z, d[11] += c.w;
}

// STORE: 1 = cp.async.bulk (TMA) store of the staged row, 2 = staged row re-read lane-contiguously
// and written with three fully coalesced 16-byte stores per lane.
template <int STORE, int MIN_CTAS, int kLoadDepth, int U>
__global__ void __launch_bounds__(kMaxWarps * 32, MIN_CTAS) sat_onepass_kernel(const OnePassArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t s_ticket;
  __shared__ uint4 s_lsum;
  const int NW = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint8_t *stage = smem;  // [NW][kStageBufs][kRowBytes]
  constexpr bool TMA_STORE = STORE == 1;
  uint4 *s_rs = reinterpret_cast<uint4 *>(smem + (size_t)NW * kStageBufs * kRowBytes);
  uint4 *s_left = s_rs + NW * kMaxBandRows;  // [kMaxBandRows] carry from the CTAs to the left
  uint4 *s_wt = s_left + kMaxBandRows;       // [kMaxWarps]    per-warp tile totals

  if (threadIdx.x == 0) s_ticket = atomicAdd(&a.counters[0], 1u);
  __syncthreads();
  const uint32_t tile = s_ticket;  // (band, frame, strip) order
  const int s = (int)(tile % (uint32_t)a.nsc);
  const int f = (int)((tile / (uint32_t)a.nsc) % (uint32_t)a.n);
  const int b = (int)(tile / ((uint32_t)a.nsc * (uint32_t)a.n));
  const int strip = s * NW + warp;
  const bool active = strip < a.ns;
  const int x0 = strip * kStripPx + lane * 4;
  const bool in_x = x0 < a.W;
  const int y0 = b * a.R;
  const int y1 = min(y0 + a.R, a.H);
  const int rows = y1 - y0;
  const uint8_t *src = a.src + (size_t)f * a.src_stride;
  uint32_t *sat = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(a.sat) +
                                               (size_t)f * a.sat_stride);

  // The strip is read twice: keep it in L2 after the first read (evict-last), release it on the
  // second (evict-first); the SAT rows stream out evict-first.
  uint64_t pol_keep, pol_stream;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));

  // ---- phase A: read the strip once; column sums per lane, row sums per row ------------------
  uint32_t acc[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i] = 0;
  for (int yc = y0; yc < y1; yc += 32) {
    uint32_t k0 = 0, k1 = 0, k2 = 0;
    const int nr = min(32, y1 - yc);
    for (int r8 = 0; r8 < nr; r8 += kLoadDepth) {
      uint4 q[kLoadDepth];
#pragma unroll
      for (int u = 0; u < kLoadDepth; ++u) {
        const int y = min(yc + r8 + u, y1 - 1);
        q[u] = load_row4(src + (size_t)y * a.linesize, x0, a.W, pol_keep);
      }
#pragma unroll
      for (int u = 0; u < kLoadDepth; ++u) {
        if (r8 + u < nr) {
          uint32_t p[12];
          unpack_px4(q[u], p);
#pragma unroll
          for (int i = 0; i < 12; ++i) acc[i] += p[i];
          const uint32_t t0 = __reduce_add_sync(0xffffffffu, p[0] + p[3] + p[6] + p[9]);
          const uint32_t t1 = __reduce_add_sync(0xffffffffu, p[1] + p[4] + p[7] + p[10]);
          const uint32_t t2 = __reduce_add_sync(0xffffffffu, p[2] + p[5] + p[8] + p[11]);
          if (lane == r8 + u) {
            k0 = t0;
            k1 = t1;
            k2 = t2;
          }
        }
      }
    }
    if (lane < nr) s_rs[warp * kMaxBandRows + (yc - y0) + lane] = make_uint4(k0, k1, k2, 0);
  }
  {
    const uint32_t t0 = __reduce_add_sync(0xffffffffu, acc[0] + acc[3] + acc[6] + acc[9]);
    const uint32_t t1 = __reduce_add_sync(0xffffffffu, acc[1] + acc[4] + acc[7] + acc[10]);
    const uint32_t t2 = __reduce_add_sync(0xffffffffu, acc[2] + acc[5] + acc[8] + acc[11]);
    if (lane == 0) s_wt[warp] = make_uint4(t0, t1, t2, 0);
  }
  __syncthreads();

  // ---- row sums: exclusive prefix over the CTA's warps, publish the CTA totals ---------------
  if ((int)threadIdx.x < rows) {
    const int r = threadIdx.x;
    uint32_t r0 = 0, r1 = 0, r2 = 0;
    for (int w = 0; w < NW; ++w) {
      const uint4 v = s_rs[w * kMaxBandRows + r];
      s_rs[w * kMaxBandRows + r] = make_uint4(r0, r1, r2, 0);
      r0 += v.x, r1 += v.y, r2 += v.z;
    }
    s_left[r] = make_uint4(0, 0, 0, 0);
    if (s + 1 < a.nsc) {  // somebody to the right will want it
      __stcg(&a.rowagg[(size_t)tile * kMaxBandRows + r], make_uint4(r0, r1, r2, 0));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && s + 1 < a.nsc) st_release(&a.flag_left[tile], a.epoch);

  // ---- left carry: add up the row sums of every CTA to the left in this band -----------------
  for (int q = warp; q < s; q += NW) {
    const uint32_t pt = tile - (uint32_t)s + (uint32_t)q;
    if (lane == 0)
      while (ld_acquire(&a.flag_left[pt]) != a.epoch && !a.debug_nowait) __nanosleep(40);
    __syncwarp();
    for (int r = lane; r < rows; r += 32) {
      const uint4 v = __ldcg(&a.rowagg[(size_t)pt * kMaxBandRows + r]);
      uint32_t *d = reinterpret_cast<uint32_t *>(&s_left[r]);
      atomicAdd(d + 0, v.x);
      atomicAdd(d + 1, v.y);
      atomicAdd(d + 2, v.z);
    }
  }
  __syncthreads();
  if (warp == 0) {
    uint32_t l0 = 0, l1 = 0, l2 = 0;
    for (int r = lane; r < rows; r += 32) {
      const uint4 v = s_left[r];
      l0 += v.x, l1 += v.y, l2 += v.z;
    }
    l0 = __reduce_add_sync(0xffffffffu, l0);
    l1 = __reduce_add_sync(0xffffffffu, l1);
    l2 = __reduce_add_sync(0xffffffffu, l2);
    if (lane == 0) s_lsum = make_uint4(l0, l1, l2, 0);
  }
  __syncthreads();

  if (active) {
    // ---- band aggregate G(x) = everything this band adds to the SAT row below it -------------
    uint32_t gsum[12];
    {
      uint32_t b0 = s_lsum.x, b1 = s_lsum.y, b2 = s_lsum.z;
      for (int w = 0; w < warp; ++w) {
        const uint4 v = s_wt[w];
        b0 += v.x, b1 += v.y, b2 += v.z;
      }
#pragma unroll
      for (int i = 0; i < 12; ++i) gsum[i] = acc[i];
#pragma unroll
      for (int k = 1; k < 4; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) gsum[3 * k + c] += gsum[3 * (k - 1) + c];
      uint32_t i0 = gsum[9], i1 = gsum[10], i2 = gsum[11];
      const uint32_t t0 = i0, t1 = i1, t2 = i2;
      warp_scan3(i0, i1, i2, lane);
      i0 += b0 - t0, i1 += b1 - t1, i2 += b2 - t2;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        gsum[3 * k + 0] += i0;
        gsum[3 * k + 1] += i1;
        gsum[3 * k + 2] += i2;
      }
    }

    // ---- column carry by decoupled look-back up the warp-strip column ------------------------
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[i] = 0;  // becomes T_b(x)
    const uint32_t col = tile * (uint32_t)NW + (uint32_t)warp;
    const uint32_t col_step = (uint32_t)a.n * (uint32_t)a.nsc * (uint32_t)NW;  // one band up
    const bool more_bands = b + 1 < a.nb;
    if (b > 0) {
      if (more_bands) {
        if (in_x) {
          uint4 *d = reinterpret_cast<uint4 *>(a.colagg + (size_t)col * (kStripPx * 3)) + lane * 3;
          __stcg(d + 0, make_uint4(gsum[0], gsum[1], gsum[2], gsum[3]));
          __stcg(d + 1, make_uint4(gsum[4], gsum[5], gsum[6], gsum[7]));
          __stcg(d + 2, make_uint4(gsum[8], gsum[9], gsum[10], gsum[11]));
        }
        __syncwarp();
        if (lane == 0) st_release(&a.flag_col[col], (a.epoch << 2) | kAgg);
      }
      int k = b - 1;
      uint32_t pc = col - col_step;
      while (true) {
        uint32_t st = 0;
        if (lane == 0)
          while (((st = ld_acquire(&a.flag_col[pc])) >> 2) != a.epoch && !a.debug_nowait)
            __nanosleep(20);
        if (a.debug_nowait) st = kInc;
        st = __shfl_sync(0xffffffffu, st, 0);
        if ((st & 3u) == kInc) {  // T_{k+1} = last SAT row of band k: final
          if (in_x) {
            const int yr = min((k + 1) * a.R, a.H) - 1;
            const uint4 *p = reinterpret_cast<const uint4 *>(sat + ((size_t)yr * a.W + x0) * 3);
            add12(acc, __ldcg(p), __ldcg(p + 1), __ldcg(p + 2));
          }
          break;
        }
        if (in_x) {
          const uint4 *p =
              reinterpret_cast<const uint4 *>(a.colagg + (size_t)pc * (kStripPx * 3)) + lane * 3;
          add12(acc, __ldcg(p), __ldcg(p + 1), __ldcg(p + 2));
        }
        if (k == 0) break;  // T_0 = 0
        --k;
        pc -= col_step;
      }
    }
    if (more_bands) {
      // publish the inclusive value = the tile's last SAT row (phase C rewrites the same words)
      if (in_x) {
        uint4 *d = reinterpret_cast<uint4 *>(sat + ((size_t)(y1 - 1) * a.W + x0) * 3);
        __stcg(d + 0, make_uint4(acc[0] + gsum[0], acc[1] + gsum[1], acc[2] + gsum[2], acc[3] + gsum[3]));
        __stcg(d + 1, make_uint4(acc[4] + gsum[4], acc[5] + gsum[5], acc[6] + gsum[6], acc[7] + gsum[7]));
        __stcg(d + 2, make_uint4(acc[8] + gsum[8], acc[9] + gsum[9], acc[10] + gsum[10], acc[11] + gsum[11]));
      }
      __syncwarp();
      if (lane == 0) st_release(&a.flag_col[col], (a.epoch << 2) | kInc);
    }

    // ---- phase C: re-read the strip (L2), scan each row, accumulate down, write once ----------
    const uint4 *rc = s_rs + warp * kMaxBandRows;  // exclusive prefix over the CTA's warps
    const int strip_px = min(kStripPx, a.W - strip * kStripPx);
    uint8_t *my_stage = stage + (size_t)warp * kStageBufs * kRowBytes;
    const uint64_t policy = pol_stream;
    int buf = 0;
    uint4 p[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      p[u] = load_row4(src + (size_t)min(y0 + u, y1 - 1) * a.linesize, x0, a.W, pol_stream);
    for (int y = y0; y < y1; y += U) {
      uint4 q[U];
#pragma unroll
      for (int u = 0; u < U; ++u) q[u] = p[u];
      if (y + U < y1) {
#pragma unroll
        for (int u = 0; u < U; ++u)
          p[u] = load_row4(src + (size_t)min(y + U + u, y1 - 1) * a.linesize, x0, a.W, pol_stream);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (y + u < y1) {
          uint32_t v[12];
          unpack_px4(q[u], v);
          if (lane == 0) {  // carry-in of this row: CTAs to the left + warps to the left
            const uint4 c1 = rc[y + u - y0], c2 = s_left[y + u - y0];
            v[0] += c1.x + c2.x, v[1] += c1.y + c2.y, v[2] += c1.z + c2.z;
          }
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
                  "cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;\n\t"
                  "cp.async.bulk.commit_group;" ::"l"(drow + (size_t)strip * kStripPx * 3),
                  "r"(smem_u32(sb)), "r"(strip_px * 12), "l"(policy)
                  : "memory");
            }
            buf = (buf + 1 == kStageBufs) ? 0 : buf + 1;
          } else {
            // transpose through shared memory so that each store instruction of the warp covers
            // 512 contiguous bytes (lane-strided 48-byte stores run at ~60 % of this)
            uint8_t *sb = my_stage + (size_t)buf * kRowBytes;
            uint4 *sd = reinterpret_cast<uint4 *>(sb + lane * 48);
            sd[0] = make_uint4(acc[0], acc[1], acc[2], acc[3]);
            sd[1] = make_uint4(acc[4], acc[5], acc[6], acc[7]);
            sd[2] = make_uint4(acc[8], acc[9], acc[10], acc[11]);
            __syncwarp();
            const uint4 *sl = reinterpret_cast<const uint4 *>(sb) + lane;
            uint4 *d = reinterpret_cast<uint4 *>(drow + (size_t)strip * kStripPx * 3) + lane;
            const int n16 = strip_px * 3 / 4;  // 16-byte words in this row segment
#pragma unroll
            for (int k = 0; k < 3; ++k)
              if (k * 32 + lane < n16) __stcs(d + k * 32, sl[k * 32]);
            buf ^= 1;
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

  // ---- the last CTA to finish re-arms the ticket counter for the next launch ------------------
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&a.counters[1], 1u) == a.total_tiles - 1) {
      a.counters[0] = 0;
      a.counters[1] = 0;
      __threadfence();
    }
  }
}

int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return e ? atoi(e) : dflt;
}

}  // namespace

SatOnePassPlan sat_onepass_plan(int n, int W, int H) {
  SatOnePassPlan p;
  p.ns = (W + kStripPx - 1) / kStripPx;
  // warps per CTA: fewest idle warp slots in the last CTA of a band, widest CTA on ties
  int best = 8, best_waste = 1 << 30;
  for (int nw = kMaxWarps; nw >= 4; --nw) {
    const int waste = ((p.ns + nw - 1) / nw) * nw - p.ns;
    if (waste < best_waste) best = nw, best_waste = waste;
  }
  static const int force_nw = env_int("FOV360_SAT_WARPS", 0);
  if (force_nw >= 1 && force_nw <= kMaxWarps) best = force_nw;
  p.NW = best;
  p.nsc = (p.ns + p.NW - 1) / p.NW;
  static const int band = env_int("FOV360_SAT_BAND_ROWS", 32);
  p.R = band < 1 ? 1 : (band > kMaxBandRows ? kMaxBandRows : band);
  p.nb = (H + p.R - 1) / p.R;
  const size_t tiles = (size_t)n * p.nb * p.nsc;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  p.off_counters = 0;
  p.off_flag_left = 256;
  p.off_flag_col = p.off_flag_left + al(tiles * 4);
  p.off_rowagg = p.off_flag_col + al(tiles * p.NW * 4);
  p.off_colagg = p.off_rowagg + al(tiles * kMaxBandRows * 16);
  p.bytes = p.off_colagg + al(tiles * p.NW * kStripPx * 12);
  // Flags are matched against a per-launch epoch, so they never need clearing between launches
  // of one layout; the counters and both flag arrays must be zeroed when the layout changes.
  p.clear_bytes = p.off_rowagg;
  return p;
}

bool sat_onepass_eligible(const uint32_t *sat, size_t sat_stride, const uint8_t *src,
                          size_t src_stride, int W, int H, int linesize) {
  static const int disabled = env_int("FOV360_SAT_THREE_PASS", 0);
  (void)H;
  return !disabled && linesize / W == 4 && (W % 4) == 0 && (linesize % 16) == 0 &&
         ((uintptr_t)src % 16) == 0 && (src_stride % 16) == 0 && ((uintptr_t)sat % 16) == 0 &&
         (sat_stride % 16) == 0;
}

cudaError_t launch_sat_onepass(const LaunchCtx &lc, int n, uint32_t *sat, size_t sat_stride,
                               const uint8_t *src, size_t src_stride, int W, int H, int linesize,
                               void *scratch, uint32_t epoch) {
  const SatOnePassPlan p = sat_onepass_plan(n, W, H);
  uint8_t *base = static_cast<uint8_t *>(scratch);
  OnePassArgs a;
  a.src = src;
  a.sat = sat;
  a.src_stride = src_stride;
  a.sat_stride = sat_stride;
  a.W = W;
  a.H = H;
  a.linesize = linesize;
  a.n = n;
  a.R = p.R;
  a.nb = p.nb;
  a.ns = p.ns;
  a.nsc = p.nsc;
  a.epoch = epoch;
  static const int nowait = env_int("FOV360_SAT_DEBUG_NOWAIT", 0);
  a.debug_nowait = nowait;
  a.total_tiles = (uint32_t)((size_t)n * p.nb * p.nsc);
  a.counters = reinterpret_cast<uint32_t *>(base + p.off_counters);
  a.flag_left = reinterpret_cast<uint32_t *>(base + p.off_flag_left);
  a.rowagg = reinterpret_cast<uint4 *>(base + p.off_rowagg);
  a.flag_col = reinterpret_cast<uint32_t *>(base + p.off_flag_col);
  a.colagg = reinterpret_cast<uint32_t *>(base + p.off_colagg);

  static const bool tma_store = env_int("FOV360_SAT_TMA_STORE", 0) != 0;
  const size_t carry_smem = (size_t)(p.NW * kMaxBandRows + kMaxBandRows + kMaxWarps) * 16;
  const size_t smem = carry_smem + (size_t)p.NW * kStageBufs * kRowBytes;
  const int max_smem = (kMaxWarps * kMaxBandRows + kMaxBandRows + kMaxWarps) * 16 +
                       kMaxWarps * kStageBufs * kRowBytes;
  KernelScope ks(lc, "sat_onepass");
  static const int variant = env_int("FOV360_SAT_VARIANT", 0);
#define FOV_LAUNCH(TMA, MINC, DEPTH, UU)                                                          \
  do {                                                                                             \
    cudaFuncSetAttribute(sat_onepass_kernel<TMA, MINC, DEPTH, UU>,                                 \
                         cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);                   \
    sat_onepass_kernel<TMA, MINC, DEPTH, UU><<<a.total_tiles, p.NW * 32, smem, lc.stream>>>(a);    \
  } while (0)
  if (tma_store)
    FOV_LAUNCH(1, 3, 8, 4);
  else if (variant == 1)
    FOV_LAUNCH(2, 4, 8, 2);
  else if (variant == 2)
    FOV_LAUNCH(2, 3, 16, 4);
  else
    FOV_LAUNCH(2, 3, 8, 4);
#undef FOV_LAUNCH
  return cudaGetLastError();
}

}  // namespace fov
