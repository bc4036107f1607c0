// Internal declarations shared by the translation units of libfov360.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "fov360.h"

namespace fov {

constexpr int kMaxBatchPerLaunch = 64;  // gaze points carried by value in kernel params

// ---- per-axis lookup tables (built on the host, see luts.cc) -----------------------------

// One entry of the inverse-warp table, indexed by (pos - centre) + n_full.
// Everything interpolate_rect_kernel derives from one axis, minus the two gaze-dependent
// border fix-ups (sat_decoder_interpolate_kernel.cl:105-116) that the kernel applies inline.
struct __align__(16) InterpEntry {
  int16_t idx_exact;  // clamp(u + n_red/2, 0, n_red-1)               (:69-70)
  int16_t min_u;      // min(u, u+du)                                    (:100-103)
  int16_t max_u;      // max(u, u+du)
  int16_t exact;      // delta_calculated == delta                      (:67)
  int16_t rel_lo;     // min(d_min, d_calc): lo = centre + rel_lo       (:91-98)
  int16_t rel_hi;     // max(d_min, d_calc)
  float ratio;        // clamp((d - rel_lo) / (rel_hi - rel_lo), 0, 1)  (:135-142)
};
static_assert(sizeof(InterpEntry) == 16, "InterpEntry must be 16 bytes");

struct SatGrid {  // SATDecoder grid, separable form
  int ow, oh, W, H;
  int16_t *d_xedge = nullptr;  // [ow+1]
  int16_t *d_yedge = nullptr;  // [oh+1]
  std::vector<int16_t> h_xedge, h_yedge;
};

struct InterpLut {  // SATDecoder inverse warp, separable form
  int W, H, ow, oh;
  InterpEntry *d_x = nullptr;  // [2W+1], index = dx + W
  InterpEntry *d_y = nullptr;  // [2H+1], index = dy + H
};

struct ImgGrid {  // ImageSampler log-rect grid, separable form
  int ow, oh, W, H;
  int16_t *d_xd = nullptr;  // [ow]
  int16_t *d_yd = nullptr;  // [oh]
  std::vector<int16_t> h_xd, h_yd;
};

struct LogpolarGrid {  // ImageSampler log-polar grid, separable factors
  int ow, oh;
  float *d_radius = nullptr;  // [ow]
  float *d_cos = nullptr;     // [oh]
  float *d_sin = nullptr;     // [oh]
  double2 *d_dir = nullptr;   // [oh] (cos, sin) of the inverse warp's double-typed angle
  double *d_radius_f64 = nullptr;  // [ow] radius widened to double (operand of the exact-hit test)
  double2 *d_lntab = nullptr;      // [kLnExponents * 128] {1 / c, (ow / 20) ln c} (build_logpolar_lntab)
  std::vector<float> h_radius, h_cos, h_sin;
};

// Host-side table builders (luts.cc).  Plain C++ + libm, no CUDA.
void build_sat_grid_edges(int ow, int oh, int W, int H, std::vector<int16_t> &xe,
                          std::vector<int16_t> &ye);
void build_interp_axis(int n_full, int n_red, std::vector<InterpEntry> &lut);
void build_img_grid_axes(int ow, int oh, int W, int H, std::vector<int16_t> &xd,
                         std::vector<int16_t> &yd);
void build_logpolar_axes(int ow, int oh, std::vector<float> &radius, std::vector<float> &cs,
                         std::vector<float> &sn);
void build_logpolar_directions(int oh, std::vector<double2> &dir);
constexpr int kLnExponents = 31;  // d2 = dx^2 + dy^2 < 2^31 for every accepted geometry
void build_logpolar_lntab(int ow, std::vector<double2> &tab);
float logpolar_round_zone(int oh);

// ---- launch context: stream + optional per-kernel event timing ------------------------------

// Per-kernel CUDA-event timing (the analogue of CL_QUEUE_PROFILING_ENABLE, which the reference
// never switches on, opencl_manager.cc:55).  Off by default: no events are recorded.
constexpr size_t kMaxPendingEvents = 4096;
struct Profiler {
  struct Pending {
    const char *name;
    cudaEvent_t a, b;
  };
  struct Total {
    double ms = 0;
    uint64_t launches = 0;
  };
  bool enabled = false;
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> pool;
  std::map<std::string, Total> totals;
  cudaEvent_t get();
  void collect();  // requires the stream to be idle
  void recycle();  // folds the pairs that have completed; never waits
  void release();
};

struct LaunchCtx {
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  int device = 0;
  Profiler *prof = nullptr;
  uint64_t *launches = nullptr;
  bool reduced_pad_zero = false;  // FOV_OPT_REDUCED_PAD_ZERO: sample_rect may write whole pixels
};

// RAII bracket around ONE kernel launch: counts it and, when profiling, times it.
class KernelScope {
 public:
  KernelScope(const LaunchCtx &lc, const char *name) : lc_(lc) {
    if (lc.launches) ++*lc.launches;
    if (lc.prof && lc.prof->enabled) {
      // a long profiled run must not grow without bound: fold completed pairs now and then
      if (lc.prof->pending.size() >= kMaxPendingEvents) lc.prof->recycle();
      p_.name = name;
      p_.a = lc.prof->get();
      p_.b = lc.prof->get();
      on_ = p_.a && p_.b && cudaEventRecord(p_.a, lc.stream) == cudaSuccess;
      if (!on_) give_back();
    }
  }
  ~KernelScope() {
    if (!on_) return;
    if (cudaEventRecord(p_.b, lc_.stream) == cudaSuccess)
      lc_.prof->pending.push_back(p_);
    else
      give_back();
  }

 private:
  void give_back() {
    if (p_.a) lc_.prof->pool.push_back(p_.a);
    if (p_.b) lc_.prof->pool.push_back(p_.b);
  }
  const LaunchCtx &lc_;
  Profiler::Pending p_{};
  bool on_ = false;
};

// ---- programmatic dependent launch (the encode -> sample -> interpolate chain) ----------------
// A kernel launched with launch_chained() may become resident while its predecessor in the stream
// is still draining: CTA scheduling and whatever it does before pdl_wait() (table look-ups, shared
// memory set-up - nothing a predecessor could have produced) overlap the predecessor's tail.
// Every chained kernel executes pdl_wait() before it touches frame data, so the ordering the
// in-order queue promises (sat_encoder.cc:67-135 ... sat_decoder.cc:887-927 rely on it) still holds
// transitively; pdl_trigger() at its start lets its own successor do the same.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();  // capi.cu: on unless FOV360_NO_PDL is set

template <class... KArgs, class... Args>
cudaError_t launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                           cudaStream_t stream, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
#endif

// Index-check counters of the gather kernels (bounds_check.cuh; zero unless the library was built
// with -DFOV360_BOUNDS_CHECK): {violations, site of the first one} per translation unit.
void bounds_read_image_sampler(unsigned out[2]);
void bounds_read_sat_decode(unsigned out[2]);

// ---- kernel launchers (one per .cu) ----------------------------------------------------------

struct GazeBatch {
  float xy[2 * kMaxBatchPerLaunch];
  // When set, the kernels read the gaze of frame f from dev[2f], dev[2f+1] (device memory) instead
  // of xy: nothing in the launch then depends on the gaze, so a captured CUDA graph replays with
  // whatever the buffer holds at that time (fov_*_dev entry points).
  const float *dev = nullptr;
};

// SAT build scratch: carry tables sized for (n, W, H); owned by the context.
struct SatScratch {
  void *base = nullptr;
  size_t bytes = 0;
};
size_t sat_scratch_bytes(int n, int W, int H);
cudaError_t launch_sat_encode(const LaunchCtx &lc, int n, uint32_t *sat, size_t sat_stride,
                              const uint8_t *src, size_t src_stride, int W, int H, int linesize,
                              void *scratch);

// Single-pass SAT build (sat_onepass.cu): tile geometry and scratch layout for (n, W, H).
struct SatOnePassPlan {
  int NW, R, nb, ns, nsc;
  size_t off_counters, off_rowagg, off_colagg, bytes, clear_bytes;
};
SatOnePassPlan sat_onepass_plan(int n, int W, int H);
bool sat_onepass_eligible(const uint32_t *sat, size_t sat_stride, const uint8_t *src,
                          size_t src_stride, int W, int H, int linesize);
cudaError_t launch_sat_onepass(const LaunchCtx &lc, int n, uint32_t *sat, size_t sat_stride,
                               const uint8_t *src, size_t src_stride, int W, int H, int linesize,
                               void *scratch);

cudaError_t launch_sat_sample_rect(const LaunchCtx &lc, int n, uint8_t *out, size_t out_stride, int ow,
                                   int oh, int out_linesize, const uint32_t *sat,
                                   size_t sat_stride, int W, int H, const int16_t *xedge,
                                   const int16_t *yedge, const GazeBatch &gaze,
                                   const uint8_t *src = nullptr, size_t src_stride = 0,
                                   int src_linesize = 0);
cudaError_t launch_sat_interpolate_rect(const LaunchCtx &lc, int n, uint8_t *out, size_t out_stride,
                                        int W, int H, const uint8_t *red, size_t red_stride,
                                        int ow, int oh, const InterpEntry *lx,
                                        const InterpEntry *ly, const GazeBatch &gaze);
cudaError_t launch_sat_decode(const LaunchCtx &lc, uint8_t *out, int out_linesize, const uint32_t *sat,
                              int W, int H);

cudaError_t launch_img_sample_rect(const LaunchCtx &lc, uint8_t *out, int ow, int oh, int out_linesize,
                                   const uint8_t *src, int W, int H, int src_linesize,
                                   const int16_t *xd, const int16_t *yd, float cx, float cy);
cudaError_t launch_img_sample_logpolar(const LaunchCtx &lc, uint8_t *out, int ow, int oh,
                                       int out_linesize, const uint8_t *src, int W, int H,
                                       int src_linesize, const float *radius, const float *cs,
                                       const float *sn, float cx, float cy);
cudaError_t launch_img_interpolate_logpolar(const LaunchCtx &lc, uint8_t *out, int W, int H,
                                            const uint8_t *red, int ow, int oh, float cx,
                                            float cy, const LogpolarGrid &grid);
cudaError_t launch_img_logpolar_blur(const LaunchCtx &lc, uint8_t *out, int ow, int oh,
                                     const uint8_t *src);
cudaError_t launch_img_logpolar_grid_expand(const LaunchCtx &lc, int16_t *grid, int ow, int oh,
                                            const float *radius, const float *cs,
                                            const float *sn);

// Projections (projections.cu, sat_decode.cu).  Per-call constants of the inverse gnomonic map:
// phi1 / lambda0 are exact IEEE double arithmetic narrowed to float (projections_program.cl:26,28);
// the sine / cosine of phi1 are evaluated once on the host with the libm the parity oracle uses.
struct GnomonicView {
  float lambda0;             // (cx - 0.5) * 2 * PI
  float sin_phi1, cos_phi1;  // of phi1 = (cy - 0.5) * PI
};
GnomonicView make_gnomonic_view(float cx, float cy);  // luts.cc: host libm, like the tables
cudaError_t launch_gnomonic(const LaunchCtx &lc, uint8_t *out, int tw, int th, const uint8_t *src,
                            int W, int H, const GnomonicView &view);
cudaError_t launch_sat_interpolate_gnomonic(const LaunchCtx &lc, uint8_t *out, int tw, int th,
                                            const uint8_t *red, int ow, int oh, int W, int H,
                                            const InterpEntry *lx, const InterpEntry *ly,
                                            float gaze_x, float gaze_y, const GnomonicView &view);

// RGB0 -> YUV420P / NV12 (color_convert.cu).  nv12: `u` is the interleaved plane, `v` unused.
cudaError_t launch_rgb0_to_yuv(const LaunchCtx &lc, bool nv12, int n, uint8_t *y, size_t y_stride,
                               int y_ls, uint8_t *u, int u_ls, uint8_t *v, int v_ls, size_t c_stride,
                               const uint8_t *src, size_t src_stride, int src_ls, int W, int H);
// YUV420P / NV12 -> RGB0 (color_convert.cu).  nv12: `u` is the interleaved plane, `v` unused.
cudaError_t launch_yuv_to_rgb0(const LaunchCtx &lc, bool nv12, int n, uint8_t *dst,
                               size_t dst_stride, int dst_ls, const uint8_t *y, size_t y_stride,
                               int y_ls, const uint8_t *u, int u_ls, const uint8_t *v, int v_ls,
                               size_t c_stride, int W, int H);

}  // namespace fov
