// Phase timeline of the single-pass SAT kernel (not part of the library): compiles sat_onepass.cu
// with FOV360_SAT_TRACE, runs it on synthetic frames and prints the average clock cycles a CTA
// spends between the trace points.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include \
//        -I foveated-360-video_b200/csrc -o tools/sat_trace.bin tools/sat_trace.cu
#define FOV360_SAT_TRACE 1
#include "../foveated-360-video_b200/csrc/sat_onepass.cu"

#include <algorithm>
#include <vector>

cudaEvent_t fov::Profiler::get() { return nullptr; }  // profiling stays off here
void fov::Profiler::recycle() {}
void fov::Profiler::collect() {}
bool fov::pdl_enabled() { return false; }

int main(int argc, char **argv) {
  // usage: sat_trace.bin [frames] [W] [H]
  const int F = argc > 1 ? atoi(argv[1]) : 8;
  const int W = argc > 2 ? atoi(argv[2]) : 7680, H = argc > 3 ? atoi(argv[3]) : 3840;
  using namespace fov;
  const SatOnePassPlan p = sat_onepass_plan(F, W, H);
  const size_t tiles = (size_t)F * p.nb * p.nsc;
  uint8_t *src;
  uint32_t *sat;
  void *scratch;
  cudaMalloc(&src, (size_t)W * H * 4 * F);
  cudaMalloc(&sat, (size_t)W * H * 12 * F);
  cudaMalloc(&scratch, p.bytes);
  cudaMalloc(&g_sat_trace, tiles * 16 * sizeof(long long));
  cudaMemset(src, 1, (size_t)W * H * 4 * F);
  cudaMemset(scratch, 0, p.bytes);
  LaunchCtx lc;
  cudaStreamCreate(&lc.stream);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i)
    launch_sat_onepass(lc, F, sat, (size_t)W * H * 12, src, (size_t)W * H * 4, W, H, W * 4, scratch);
  cudaEventRecord(e0, lc.stream);
  for (int i = 0; i < 10; ++i)
    launch_sat_onepass(lc, F, sat, (size_t)W * H * 12, src, (size_t)W * H * 4, W, H, W * 4, scratch);
  cudaEventRecord(e1, lc.stream);
  cudaStreamSynchronize(lc.stream);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("NW=%d R=%d tiles=%zu  %.4f ms/launch  (%s)\n", p.NW, p.R, tiles, ms / 10,
         cudaGetErrorString(cudaGetLastError()));
  std::vector<long long> t(tiles * 16);
  cudaMemcpy(t.data(), g_sat_trace, t.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  static const char *names[7] = {"phase A loads+reduce", "barrier after A", "row sums + publish",
                                 "left carry + barrier", "gsum + look-back + INC", "phase C scan+store",
                                 "final barrier"};
  double tot = 0;
  for (int i = 0; i < 7; ++i) {
    std::vector<long long> d(tiles);
    for (size_t k = 0; k < tiles; ++k) d[k] = t[k * 16 + i + 1] - t[k * 16 + i];
    std::sort(d.begin(), d.end());
    double s = 0;
    for (auto v : d) s += v;
    tot += s / tiles;
    printf("  %-24s mean %8.0f  p50 %8lld  p90 %8lld  p99 %8lld cycles\n", names[i], s / tiles,
           d[tiles / 2], d[tiles * 9 / 10], d[tiles * 99 / 100]);
  }
  printf("  CTA lifetime mean %.0f cycles = %.1f us at 1.965 GHz\n", tot, tot / 1965.0);
  // wall-clock timeline of the last launch (%globaltimer, ns): when tiles start and finish
  long long t0 = t[8], t1 = 0;
  for (size_t k = 0; k < tiles; ++k) t0 = std::min(t0, t[k * 16 + 8]), t1 = std::max(t1, t[k * 16 + 15]);
  printf("  kernel span first-start -> last-end %.2f us\n", (t1 - t0) / 1e3);
  const size_t per_band = tiles / p.nb;
  for (int b = 0; b < p.nb; b += std::max(1, p.nb / 16)) {
    long long s0 = 1ll << 62, e1 = 0, c0 = 0;
    for (size_t k = b * per_band; k < (b + 1) * per_band; ++k) {
      s0 = std::min(s0, t[k * 16 + 8]);
      e1 = std::max(e1, t[k * 16 + 15]);
      c0 = std::max(c0, t[k * 16 + 8 + 5]);
    }
    printf("    band %3d: starts %7.2f us, carries known %7.2f us, ends %7.2f us\n", b, (s0 - t0) / 1e3,
           (c0 - t0) / 1e3, (e1 - t0) / 1e3);
  }
  return 0;
}
