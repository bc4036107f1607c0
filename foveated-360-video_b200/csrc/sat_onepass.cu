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
//                 (state INC), and then publishes its own inclusive value T_{b+1}(x) (the last
//                 SAT row of the tile) in the same scratch slot.
// Both carries travel as self-validating 16-byte units {three words, tag} written and polled with
// single-copy-atomic 128-bit accesses: no release fence, no separate flag (see ld_unit below).
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
#include <mutex>

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
constexpr uint32_t kRow = 3;             // row-carry units: tags of the two kinds never coincide

struct OnePassArgs {
  const uint8_t *src;
  uint32_t *sat;
  size_t src_stride, sat_stride;
  int W, H, linesize;
  int n, R, nb, ns, nsc;  // frames, band rows, bands, warp strips, CTA strips
  uint32_t total_tiles;
  uint32_t *counters;  // [0] ticket, [1] finished CTAs, [2] epoch of the last completed launch
  uint4 *rowagg;       // [tile][kMaxBandRows]  {r, g, b, epoch << 2 | kRow}: per-row sums of a CTA tile
  uint4 *colagg;       // [tile][NW][4][32]     {v0, v1, v2, epoch << 2 | state}: column carry
#ifdef FOV360_SAT_TRACE
  long long *trace;  // [tile][16] clock64, then %globaltimer, at the 8 phase boundaries (tools/sat_trace.cu only)
#endif
};

#ifdef FOV360_SAT_TRACE
__device__ __forceinline__ long long trace_globaltimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define FOV_TRACE(i)                                                  \
  if (threadIdx.x == 0) {                                             \
    a.trace[(size_t)tile * 16 + (i)] = clock64();                     \
    a.trace[(size_t)tile * 16 + 8 + (i)] = trace_globaltimer();       \
  }
#else
#define FOV_TRACE(i)
#endif

// Carries travel between CTAs as self-validating 16-byte units {three words, tag}: one
// single-copy-atomic 128-bit store publishes data and tag together, the consumer polls the unit
// itself.  No release fence (MEMBAR.GPU has to drain every store the SM has in flight - several
// microseconds while the neighbours stream SAT rows), no separate flag round trip.
__device__ __forceinline__ uint4 ld_unit(const uint4 *p) {
  uint64_t lo, hi;
  asm volatile(
      "{\n\t.reg .b128 q;\n\tld.relaxed.gpu.global.b128 q, [%2];\n\tmov.b128 {%0, %1}, q;\n\t}"
      : "=l"(lo), "=l"(hi)
      : "l"(p)
      : "memory");
  return make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
}

__device__ __forceinline__ void st_unit(uint4 *p, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  const uint64_t lo = (uint64_t)x | ((uint64_t)y << 32), hi = (uint64_t)z | ((uint64_t)w << 32);
  asm volatile(
      "{\n\t.reg .b128 q;\n\tmov.b128 q, {%1, %2};\n\tst.relaxed.gpu.global.b128 [%0], q;\n\t}" ::"l"(p),
      "l"(lo), "l"(hi)
      : "memory");
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

// kLoadDepth rows are in flight per lane while the strip is reduced, U while it is scanned.
template <int MIN_CTAS, int kLoadDepth, int U>
__global__ void __launch_bounds__(kMaxWarps * 32, MIN_CTAS) sat_onepass_kernel(const OnePassArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t s_ticket, s_epoch;
  constexpr int kBufs = kStageBufs;
  const int NW = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint8_t *stage = smem;  // [NW][kBufs][kRowBytes]
  uint4 *s_rs = reinterpret_cast<uint4 *>(smem + (size_t)NW * kBufs * kRowBytes);
  const int RS = a.R;                  // row stride of the per-warp row-sum table
  uint4 *s_left = s_rs + NW * RS;      // [R] carry from the CTAs to the left

  pdl_trigger();
  if ((int)threadIdx.x < RS) s_left[threadIdx.x] = make_uint4(0, 0, 0, 0);
  pdl_wait();  // the frames (and the scratch of an earlier SAT build) are final from here on
  if (threadIdx.x == 0) {
    s_ticket = atomicAdd(&a.counters[0], 1u);
    // The launch epoch lives on the device: the last CTA of a launch stores the epoch it used, the
    // next launch uses that + 1.  Nothing the host passes changes from launch to launch, so the
    // launch can be replayed from a captured CUDA graph.
    s_epoch = __ldcg(&a.counters[2]) + 1u;
  }
  __syncthreads();
  const uint32_t tile = s_ticket;  // (band, frame, strip) order
  const uint32_t epoch = s_epoch;
  FOV_TRACE(0);
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
  // Sums of at most 64 rows (columns) or 128 pixels (rows) of bytes fit 16 bits, so two channels
  // share a register: (R | B << 16) and G.
  uint32_t crb[4] = {0, 0, 0, 0}, cg[4] = {0, 0, 0, 0};
  for (int yc = y0; yc < y1; yc += 32) {
    uint32_t krb = 0, kg = 0;
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
          const uint32_t w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
          uint32_t trb = 0, tg = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t rb = w[k] & 0x00ff00ffu, g = __byte_perm(w[k], 0, 0x4441);
            crb[k] += rb, cg[k] += g;
            trb += rb, tg += g;
          }
          trb = __reduce_add_sync(0xffffffffu, trb);
          tg = __reduce_add_sync(0xffffffffu, tg);
          if (lane == r8 + u) krb = trb, kg = tg;
        }
      }
    }
    if (lane < nr)
      s_rs[warp * RS + (yc - y0) + lane] = make_uint4(krb & 0xffffu, kg, krb >> 16, 0);
  }
  FOV_TRACE(1);
  __syncthreads();
  FOV_TRACE(2);

  // ---- row sums: exclusive prefix over the CTA's warps, publish the CTA totals ---------------
  if ((int)threadIdx.x < rows) {
    const int r = threadIdx.x;
    uint32_t r0 = 0, r1 = 0, r2 = 0;
    for (int w = 0; w < NW; ++w) {
      const uint4 v = s_rs[w * RS + r];
      s_rs[w * RS + r] = make_uint4(r0, r1, r2, 0);
      r0 += v.x, r1 += v.y, r2 += v.z;
    }
    if (s + 1 < a.nsc)  // somebody to the right will want it
      st_unit(&a.rowagg[(size_t)tile * kMaxBandRows + r], r0, r1, r2, (epoch << 2) | kRow);
  }
  FOV_TRACE(3);

  // ---- left carry: add up the row sums of every CTA to the left in this band -----------------
  // warp w takes CTAs w, w + NW, ...; lane r polls the units of rows r and r + 32
  if (warp < s) {
    uint32_t l[2][3] = {{0, 0, 0}, {0, 0, 0}};
    for (int q = warp; q < s; q += NW) {
      const uint4 *pu = a.rowagg + (size_t)(tile - (uint32_t)s + (uint32_t)q) * kMaxBandRows;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = lane + 32 * h;
        if (r < rows) {
          uint4 v = ld_unit(pu + r);
          while (v.w != ((epoch << 2) | kRow)) {
            __nanosleep(20);
            v = ld_unit(pu + r);
          }
          l[h][0] += v.x, l[h][1] += v.y, l[h][2] += v.z;
        }
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lane + 32 * h;
      if (r < rows) {
        uint32_t *d = reinterpret_cast<uint32_t *>(&s_left[r]);
        atomicAdd(d + 0, l[h][0]);
        atomicAdd(d + 1, l[h][1]);
        atomicAdd(d + 2, l[h][2]);
      }
    }
  }
  __syncthreads();
  FOV_TRACE(4);

  if (active) {
    // ---- carry-in of each row (CTAs to the left + warps to the left) and its sum over the rows
    uint4 *rc = s_rs + warp * RS;
    uint32_t b0 = 0, b1 = 0, b2 = 0;
    for (int r = lane; r < rows; r += 32) {
      const uint4 c1 = rc[r], c2 = s_left[r];
      const uint4 c = make_uint4(c1.x + c2.x, c1.y + c2.y, c1.z + c2.z, 0);
      rc[r] = c;
      b0 += c.x, b1 += c.y, b2 += c.z;
    }
    b0 = __reduce_add_sync(0xffffffffu, b0);
    b1 = __reduce_add_sync(0xffffffffu, b1);
    b2 = __reduce_add_sync(0xffffffffu, b2);  // also orders the rc[] writes before phase C

    // ---- band aggregate G(x) = everything this band adds to the SAT row below it -------------
    uint32_t gsum[12];
    {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        gsum[3 * k + 0] = crb[k] & 0xffffu, gsum[3 * k + 1] = cg[k], gsum[3 * k + 2] = crb[k] >> 16;
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
    // A lane owns 4 units (12 words); each unit is followed on its own until it meets an
    // inclusive value, so a predecessor caught between its AGG and INC stores is harmless.
    uint32_t acc[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[i] = 0;  // becomes T_b(x)
    const uint32_t col = tile * (uint32_t)NW + (uint32_t)warp;
    const uint32_t col_step = (uint32_t)a.n * (uint32_t)a.nsc * (uint32_t)NW;  // one band up
    const bool more_bands = b + 1 < a.nb;
    uint4 *my_units = a.colagg + (size_t)col * 128 + lane;
    if (b > 0) {
      if (more_bands && in_x) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          st_unit(my_units + 32 * k, gsum[3 * k], gsum[3 * k + 1], gsum[3 * k + 2],
                  (epoch << 2) | kAgg);
      }
      uint32_t open = in_x ? 0xfu : 0u;  // units still looking for an inclusive value
      uint32_t pc = col - col_step;
      while (__any_sync(0xffffffffu, open != 0)) {
        const uint4 *pu = a.colagg + (size_t)pc * 128 + lane;
        uint32_t pending = open;
        while (pending) {
          uint4 u[4];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (pending & (1u << k)) u[k] = ld_unit(pu + 32 * k);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if ((pending & (1u << k)) && (u[k].w >> 2) == epoch) {
              acc[3 * k + 0] += u[k].x, acc[3 * k + 1] += u[k].y, acc[3 * k + 2] += u[k].z;
              pending &= ~(1u << k);
              if ((u[k].w & 3u) == kInc) open &= ~(1u << k);
            }
          }
          if (pending) __nanosleep(20);
        }
        pc -= col_step;  // band 0 only ever publishes inclusive values: the walk ends there
      }
    }
    if (more_bands && in_x) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        st_unit(my_units + 32 * k, acc[3 * k] + gsum[3 * k], acc[3 * k + 1] + gsum[3 * k + 1],
                acc[3 * k + 2] + gsum[3 * k + 2], (epoch << 2) | kInc);
    }
    FOV_TRACE(5);

    // ---- phase C: re-read the strip (L2), scan each row, accumulate down, write once ----------
    const int strip_px = min(kStripPx, a.W - strip * kStripPx);
    uint8_t *my_stage = stage + (size_t)warp * kBufs * kRowBytes;
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
          // row prefix inside the 128-pixel strip in packed 16-bit fields (<= 128 * 255)
          const uint32_t w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
          uint32_t rb[4], g[4];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            rb[k] = w[k] & 0x00ff00ffu, g[k] = __byte_perm(w[k], 0, 0x4441);
#pragma unroll
          for (int k = 1; k < 4; ++k) rb[k] += rb[k - 1], g[k] += g[k - 1];
          uint32_t irb = rb[3], ig = g[3];
          warp_scan2(irb, ig, lane);
          irb -= rb[3], ig -= g[3];
          const uint4 c = rc[y + u - y0];
          const uint32_t cr = (irb & 0xffffu) + c.x, cgn = ig + c.y, cb = (irb >> 16) + c.z;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            acc[3 * k + 0] += (rb[k] & 0xffffu) + cr;
            acc[3 * k + 1] += g[k] + cgn;
            acc[3 * k + 2] += (rb[k] >> 16) + cb;
          }
          uint32_t *drow = sat + ((size_t)(y + u) * a.W) * 3;
          uint8_t *sb = my_stage + (size_t)buf * kRowBytes;
          uint4 *sd = reinterpret_cast<uint4 *>(sb + lane * 48);
          // stage the 1536-byte row segment, then one bulk async store per row (lane-strided
          // 48-byte register stores reach ~60 % of the bandwidth of 512-byte-contiguous ones)
          if (lane == 0)
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kBufs - 1) : "memory");
          __syncwarp();
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
          buf = (buf + 1 == kBufs) ? 0 : buf + 1;
        }
      }
    }
    // shared memory must outlive the in-flight bulk reads
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
  }

  // ---- the last CTA to finish re-arms the ticket counter for the next launch ------------------
  FOV_TRACE(6);
  __syncthreads();
  FOV_TRACE(7);
  if (threadIdx.x == 0) {
    if (atomicAdd(&a.counters[1], 1u) == a.total_tiles - 1) {
      a.counters[0] = 0;
      a.counters[1] = 0;
      a.counters[2] = epoch;  // every other CTA has finished: nobody reads it before the next launch
    }
  }
}

#ifdef FOV360_SAT_TRACE
long long *g_sat_trace = nullptr;
#endif

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
  static const int band = env_int("FOV360_SAT_BAND_ROWS", 24);
  p.R = band < 1 ? 1 : (band > kMaxBandRows ? kMaxBandRows : band);
  // one thread per band row publishes the row carries: a forced narrow CTA bounds the band height
  if (p.R > p.NW * 32) p.R = p.NW * 32;
  p.nb = (H + p.R - 1) / p.R;
  const size_t tiles = (size_t)n * p.nb * p.nsc;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  p.off_counters = 0;
  p.off_rowagg = 256;
  p.off_colagg = p.off_rowagg + al(tiles * kMaxBandRows * 16);
  p.bytes = p.off_colagg + al(tiles * p.NW * 128 * 16);
  // Every carry unit is tagged with the launch epoch (kept in counters[2]), which only grows:
  // a unit left behind by any earlier launch, of this or another tile layout, can never carry the
  // current tag.  The scratch is zeroed when it is (re)allocated, after the three-kernel fallback
  // has used the same bytes, and when the epoch wraps.
  p.clear_bytes = p.bytes;
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
                               void *scratch) {
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
  a.total_tiles = (uint32_t)((size_t)n * p.nb * p.nsc);
  a.counters = reinterpret_cast<uint32_t *>(base + p.off_counters);
  a.rowagg = reinterpret_cast<uint4 *>(base + p.off_rowagg);
  a.colagg = reinterpret_cast<uint4 *>(base + p.off_colagg);
#ifdef FOV360_SAT_TRACE
  a.trace = g_sat_trace;
#endif

  const size_t smem = (size_t)(p.NW * p.R + p.R) * 16 + (size_t)p.NW * kStageBufs * kRowBytes;
  constexpr int kMaxSmem = (kMaxWarps * kMaxBandRows + kMaxBandRows) * 16 +
                           kMaxWarps * kStageBufs * kRowBytes;
  // 3 CTAs of 256 threads per SM bound the registers at 80; the 192-thread CTAs 8K frames use then
  // run 4 per SM.  8 rows in flight per lane while reducing, 4 while scanning.
  auto kernel = sat_onepass_kernel<3, 8, 4>;
  // once per device, whichever connection thread gets here first
  static std::once_flag attr_once[64];
  std::call_once(attr_once[lc.device & 63], [&] {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  });
  KernelScope ks(lc, "sat_onepass");
  return launch_chained(kernel, dim3(a.total_tiles), dim3(p.NW * 32), smem, lc.stream, a);
}

}  // namespace fov
