// C ABI of libfov360.so (see include/fov360.h): context / memory management, table caches and
// argument validation in front of the kernel launchers.  There is no CPU fallback anywhere in
// this file: without a CUDA device fov_ctx_create() fails and nothing else can be called.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <unordered_set>

#include "fov360_internal.h"

using namespace fov;

struct fov_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  std::string last_error;
  uint64_t launches = 0;
  Profiler prof;
  // per-kernel event timing is off while a graph is being captured (the events would be replayed)
  bool reduced_pad_zero = false;  // fov_ctx_set_option(FOV_OPT_REDUCED_PAD_ZERO)
  LaunchCtx lc() {
    return LaunchCtx{stream, sm_count, device, capturing ? nullptr : &prof, &launches, reduced_pad_zero};
  }
  SatScratch sat_scratch;
  bool sat_scratch_dirty = true;  // holds bytes that are not carry units of an earlier epoch
  uint32_t sat_epoch = 0;         // one-pass launches since the last clear (the epoch itself is
                                  // device-resident, sat_onepass.cu); bounds the 30-bit tag
  std::map<std::tuple<int, int, int, int>, SatGrid> sat_grids;
  std::map<std::tuple<int, int, int, int>, InterpLut> interp_luts;
  std::map<std::tuple<int, int, int, int>, ImgGrid> img_grids;
  std::map<std::tuple<int, int>, LogpolarGrid> lp_grids;
  // The reference objects hold ONE current grid; lazily-initialised samplers use it.
  const SatGrid *cur_sat_grid = nullptr;
  // device allocations handed out by fov_malloc and not yet freed: released with the context
  std::unordered_set<void *> allocations;
  // CUDA-graph capture of a call sequence (fov_graph_*): while capturing nothing may allocate,
  // clear or wait, so tables and scratch must exist already (run the sequence once beforehand)
  bool capturing = false;
  uint64_t capture_launches0 = 0;
  uint32_t capture_sat0 = 0;
  uint64_t scratch_generation = 0;  // bumped when the SAT scratch moves: captured graphs hold its address
};

struct fov_graph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  uint64_t kernels = 0;        // kernel launches per replay
  uint32_t sat_launches = 0;   // one-pass SAT launches per replay (each consumes one epoch)
  uint64_t scratch_generation = 0;
};

namespace fov {

bool pdl_enabled() {
  static const bool on = getenv("FOV360_NO_PDL") == nullptr;
  return on;
}

cudaEvent_t Profiler::get() {
  if (!pool.empty()) {
    cudaEvent_t e = pool.back();
    pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  if (cudaEventCreate(&e) != cudaSuccess) return nullptr;  // the launch is then not timed
  return e;
}

// Pairs whose second event has completed are folded into the totals without waiting for anything.
void Profiler::recycle() {
  size_t done = 0;
  while (done < pending.size() && cudaEventQuery(pending[done].b) == cudaSuccess) ++done;
  if (!done) return;
  std::vector<Pending> rest(pending.begin() + done, pending.end());
  pending.resize(done);
  collect();
  pending.swap(rest);
}

void Profiler::collect() {
  for (const Pending &p : pending) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      Total &t = totals[p.name];
      t.ms += ms;
      t.launches += 1;
    }
    pool.push_back(p.a);
    pool.push_back(p.b);
  }
  pending.clear();
}

void Profiler::release() {
  collect();
  for (cudaEvent_t e : pool) cudaEventDestroy(e);
  pool.clear();
}

}  // namespace fov

namespace {

// errors raised without a context (creation failures, calls on a null context): per thread, like
// errno - connection threads fail independently
thread_local std::string g_global_error;

int fail(fov_ctx *ctx, int code, const std::string &msg) {
  if (ctx)
    ctx->last_error = msg;
  else
    g_global_error = msg;
  return code;
}

int cuda_fail(fov_ctx *ctx, cudaError_t e, const char *what) {
  return fail(ctx, (int)e, std::string(what) + ": " + cudaGetErrorName(e) + " (" +
                               cudaGetErrorString(e) + ")");
}

#define FOV_REQUIRE_CTX(ctx) \
  if (!(ctx)) return fail(nullptr, FOV_ERR_NO_CONTEXT, "fov360: not initialized with a CUDA context")
#define FOV_NOT_WHILE_CAPTURING(ctx, what)                                                        \
  if ((ctx)->capturing)                                                                           \
    return fail(ctx, FOV_ERR_INVALID, std::string(what) +                                         \
                ": not possible while a graph is being captured - run the same call sequence "   \
                "once before fov_graph_begin_capture")
#define FOV_CUDA(ctx, expr, what)                        \
  do {                                                   \
    cudaError_t e_ = (expr);                             \
    if (e_ != cudaSuccess) return cuda_fail(ctx, e_, what); \
  } while (0)

// Makes the context's device current for the duration of a call and restores the caller's: a
// server thread that also drives NVENC / NVDEC on another GPU keeps its own current device.
struct DeviceGuard {
  explicit DeviceGuard(const fov_ctx *c) {
    if (cudaGetDevice(&prev_) != cudaSuccess) prev_ = -1;
    if (prev_ != c->device) {
      cudaSetDevice(c->device);
      switched_ = true;
    }
  }
  ~DeviceGuard() {
    if (switched_ && prev_ >= 0) cudaSetDevice(prev_);
  }
  DeviceGuard(const DeviceGuard &) = delete;
  DeviceGuard &operator=(const DeviceGuard &) = delete;

 private:
  int prev_ = -1;
  bool switched_ = false;
};

template <class T>
cudaError_t upload(fov_ctx *ctx, T **dptr, const std::vector<T> &h) {
  cudaError_t e = cudaMalloc(reinterpret_cast<void **>(dptr), h.size() * sizeof(T));
  if (e != cudaSuccess) return e;
  // Table uploads are init-time; a blocking copy keeps the host vectors' lifetime trivial.
  e = cudaMemcpyAsync(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(ctx->stream);
}

bool dims_ok(int ow, int oh, int W, int H) {
  // int16 tables (like the reference's grid) hold deltas up to ~1.01*dim.
  return ow > 0 && oh > 0 && W > 0 && H > 0 && W <= 30000 && H <= 30000 && ow <= 30000 &&
         oh <= 30000;
}

int get_sat_grid(fov_ctx *ctx, int ow, int oh, int W, int H, const SatGrid **out) {
  if (!dims_ok(ow, oh, W, H)) return fail(ctx, FOV_ERR_INVALID, "sat grid: invalid dimensions");
  auto key = std::make_tuple(ow, oh, W, H);
  auto it = ctx->sat_grids.find(key);
  if (it == ctx->sat_grids.end()) {
    FOV_NOT_WHILE_CAPTURING(ctx, "sat grid creation");
    SatGrid g;
    g.ow = ow, g.oh = oh, g.W = W, g.H = H;
    build_sat_grid_edges(ow, oh, W, H, g.h_xedge, g.h_yedge);
    FOV_CUDA(ctx, upload(ctx, &g.d_xedge, g.h_xedge), "sat grid upload");
    FOV_CUDA(ctx, upload(ctx, &g.d_yedge, g.h_yedge), "sat grid upload");
    it = ctx->sat_grids.emplace(key, std::move(g)).first;
  }
  *out = &it->second;
  return FOV_OK;
}

int get_interp_lut(fov_ctx *ctx, int W, int H, int ow, int oh, const InterpLut **out) {
  if (!dims_ok(ow, oh, W, H)) return fail(ctx, FOV_ERR_INVALID, "interp lut: invalid dimensions");
  auto key = std::make_tuple(W, H, ow, oh);
  auto it = ctx->interp_luts.find(key);
  if (it == ctx->interp_luts.end()) {
    FOV_NOT_WHILE_CAPTURING(ctx, "interp lut creation");
    InterpLut l;
    l.W = W, l.H = H, l.ow = ow, l.oh = oh;
    std::vector<InterpEntry> hx, hy;
    build_interp_axis(W, ow, hx);
    build_interp_axis(H, oh, hy);
    FOV_CUDA(ctx, upload(ctx, &l.d_x, hx), "interp lut upload");
    FOV_CUDA(ctx, upload(ctx, &l.d_y, hy), "interp lut upload");
    it = ctx->interp_luts.emplace(key, l).first;
  }
  *out = &it->second;
  return FOV_OK;
}

int get_img_grid(fov_ctx *ctx, int ow, int oh, int W, int H, const ImgGrid **out) {
  if (!dims_ok(ow, oh, W, H)) return fail(ctx, FOV_ERR_INVALID, "img grid: invalid dimensions");
  auto key = std::make_tuple(ow, oh, W, H);
  auto it = ctx->img_grids.find(key);
  if (it == ctx->img_grids.end()) {
    FOV_NOT_WHILE_CAPTURING(ctx, "img grid creation");
    ImgGrid g;
    g.ow = ow, g.oh = oh, g.W = W, g.H = H;
    build_img_grid_axes(ow, oh, W, H, g.h_xd, g.h_yd);
    FOV_CUDA(ctx, upload(ctx, &g.d_xd, g.h_xd), "img grid upload");
    FOV_CUDA(ctx, upload(ctx, &g.d_yd, g.h_yd), "img grid upload");
    it = ctx->img_grids.emplace(key, std::move(g)).first;
  }
  *out = &it->second;
  return FOV_OK;
}

int get_lp_grid(fov_ctx *ctx, int ow, int oh, const LogpolarGrid **out) {
  if (!dims_ok(ow, oh, 1, 1)) return fail(ctx, FOV_ERR_INVALID, "logpolar grid: invalid dimensions");
  auto key = std::make_tuple(ow, oh);
  auto it = ctx->lp_grids.find(key);
  if (it == ctx->lp_grids.end()) {
    FOV_NOT_WHILE_CAPTURING(ctx, "logpolar grid creation");
    LogpolarGrid g;
    g.ow = ow, g.oh = oh;
    build_logpolar_axes(ow, oh, g.h_radius, g.h_cos, g.h_sin);
    FOV_CUDA(ctx, upload(ctx, &g.d_radius, g.h_radius), "logpolar grid upload");
    FOV_CUDA(ctx, upload(ctx, &g.d_cos, g.h_cos), "logpolar grid upload");
    FOV_CUDA(ctx, upload(ctx, &g.d_sin, g.h_sin), "logpolar grid upload");
    std::vector<double2> dir;
    build_logpolar_directions(oh, dir);
    FOV_CUDA(ctx, upload(ctx, &g.d_dir, dir), "logpolar grid upload");
    std::vector<double> radius64(g.h_radius.begin(), g.h_radius.end());
    FOV_CUDA(ctx, upload(ctx, &g.d_radius_f64, radius64), "logpolar grid upload");
    std::vector<double2> lntab;
    build_logpolar_lntab(ow, lntab);
    FOV_CUDA(ctx, upload(ctx, &g.d_lntab, lntab), "logpolar grid upload");
    it = ctx->lp_grids.emplace(key, std::move(g)).first;
  }
  *out = &it->second;
  return FOV_OK;
}

// Scratch of the path the call takes (one-pass plan or the three-kernel fallback): they never run
// in the same call, and the larger of the two would double the footprint of a batch of 8K frames.
int ensure_sat_scratch(fov_ctx *ctx, int n, int W, int H, bool onepass) {
  const size_t need = onepass ? sat_onepass_plan(n, W, H).bytes : sat_scratch_bytes(n, W, H);
  if (need <= ctx->sat_scratch.bytes) return FOV_OK;
  FOV_NOT_WHILE_CAPTURING(ctx, "sat scratch allocation");
  ++ctx->scratch_generation;
  if (ctx->sat_scratch.base) {
    FOV_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "sat scratch resize");
    FOV_CUDA(ctx, cudaFree(ctx->sat_scratch.base), "sat scratch free");
    ctx->sat_scratch = SatScratch();
  }
  FOV_CUDA(ctx, cudaMalloc(&ctx->sat_scratch.base, need), "sat scratch alloc");
  ctx->sat_scratch.bytes = need;
  ctx->sat_scratch_dirty = true;  // fresh memory: anything could look like a tag
  return FOV_OK;
}

}  // namespace

extern "C" {

int fov_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

fov_ctx *fov_ctx_create(int device, int *err) {
  auto set = [&](int code, const std::string &m) {
    g_global_error = m;
    if (err) *err = code;
    return (fov_ctx *)nullptr;
  };
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return set(FOV_ERR_NO_DEVICE,
               std::string("fov360: no usable CUDA device (there is no CPU fallback): ") +
                   (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  if (device < 0 || device >= n) return set(FOV_ERR_INVALID, "fov360: device index out of range");
  std::unique_ptr<fov_ctx> ctx(new fov_ctx);
  ctx->device = device;
  DeviceGuard g(ctx.get());  // the caller's current device is restored on return
  if ((e = cudaGetLastError()) != cudaSuccess)
    return set((int)e, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess)
    return set((int)e, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
  if (err) *err = FOV_OK;
  return ctx.release();
}

void fov_ctx_destroy(fov_ctx *ctx) {
  if (!ctx) return;
  DeviceGuard g(ctx);
  cudaStreamSynchronize(ctx->stream);
  for (auto &kv : ctx->sat_grids) {
    cudaFree(kv.second.d_xedge);
    cudaFree(kv.second.d_yedge);
  }
  for (auto &kv : ctx->interp_luts) {
    cudaFree(kv.second.d_x);
    cudaFree(kv.second.d_y);
  }
  for (auto &kv : ctx->img_grids) {
    cudaFree(kv.second.d_xd);
    cudaFree(kv.second.d_yd);
  }
  for (auto &kv : ctx->lp_grids) {
    cudaFree(kv.second.d_radius);
    cudaFree(kv.second.d_cos);
    cudaFree(kv.second.d_sin);
    cudaFree(kv.second.d_dir);
    cudaFree(kv.second.d_radius_f64);
    cudaFree(kv.second.d_lntab);
  }
  if (ctx->sat_scratch.base) cudaFree(ctx->sat_scratch.base);
  for (void *p : ctx->allocations) cudaFree(p);  // buffers the caller never returned
  ctx->prof.release();
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char *fov_last_error_string(const fov_ctx *ctx) {
  return ctx ? ctx->last_error.c_str() : g_global_error.c_str();
}

int fov_ctx_device(const fov_ctx *ctx) { return ctx ? ctx->device : -1; }
void *fov_ctx_stream(const fov_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
uint64_t fov_ctx_launch_count(const fov_ctx *ctx) { return ctx ? ctx->launches : 0; }

int fov_ctx_set_option(fov_ctx *ctx, int option, int value) {
  FOV_REQUIRE_CTX(ctx);
  if (ctx->capturing)
    return fail(ctx, FOV_ERR_INVALID, "fov_ctx_set_option: set options before fov_graph_begin_capture");
  if (option != FOV_OPT_REDUCED_PAD_ZERO)
    return fail(ctx, FOV_ERR_INVALID, "fov_ctx_set_option: unknown option " + std::to_string(option));
  ctx->reduced_pad_zero = value != 0;
  return FOV_OK;
}

int fov_ctx_get_option(const fov_ctx *ctx, int option, int *value) {
  if (!ctx || !value || option != FOV_OPT_REDUCED_PAD_ZERO) return FOV_ERR_INVALID;
  *value = ctx->reduced_pad_zero ? 1 : 0;
  return FOV_OK;
}

int fov_sync(fov_ctx *ctx) {
  FOV_REQUIRE_CTX(ctx);
  DeviceGuard g(ctx);
  FOV_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "fov_sync");
  return FOV_OK;
}

int fov_profile_enable(fov_ctx *ctx, int on) {
  FOV_REQUIRE_CTX(ctx);
  DeviceGuard g(ctx);
  FOV_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "fov_profile_enable");
  ctx->prof.collect();
  ctx->prof.enabled = on != 0;
  return FOV_OK;
}

int fov_profile_reset(fov_ctx *ctx) {
  FOV_REQUIRE_CTX(ctx);
  DeviceGuard g(ctx);
  FOV_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "fov_profile_reset");
  ctx->prof.collect();
  ctx->prof.totals.clear();
  return FOV_OK;
}

int fov_profile_count(fov_ctx *ctx) {
  if (!ctx) return 0;
  DeviceGuard g(ctx);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return 0;
  ctx->prof.collect();
  return (int)ctx->prof.totals.size();
}

int fov_profile_get(fov_ctx *ctx, int index, char *name, size_t name_cap, double *total_ms,
                    uint64_t *launches) {
  FOV_REQUIRE_CTX(ctx);
  if (index < 0 || index >= (int)ctx->prof.totals.size())
    return fail(ctx, FOV_ERR_INVALID, "fov_profile_get: index out of range");
  auto it = ctx->prof.totals.begin();
  std::advance(it, index);
  if (name && name_cap) {
    strncpy(name, it->first.c_str(), name_cap - 1);
    name[name_cap - 1] = 0;
  }
  if (total_ms) *total_ms = it->second.ms;
  if (launches) *launches = it->second.launches;
  return FOV_OK;
}

int fov_malloc(fov_ctx *ctx, void **dptr, size_t nbytes) {
  FOV_REQUIRE_CTX(ctx);
  if (!dptr) return fail(ctx, FOV_ERR_INVALID, "fov_malloc: null output pointer");
  DeviceGuard g(ctx);
  FOV_CUDA(ctx, cudaMalloc(dptr, nbytes ? nbytes : 1), "fov_malloc");
  ctx->allocations.insert(*dptr);
  return FOV_OK;
}

// Stream-ordered: the buffer is released once the work already queued on the context's stream
// has run; the host does not wait (clReleaseMemObject semantics).
int fov_free(fov_ctx *ctx, void *dptr) {
  FOV_REQUIRE_CTX(ctx);
  if (!dptr) return FOV_OK;
  DeviceGuard g(ctx);
  ctx->allocations.erase(dptr);
  if (cudaFreeAsync(dptr, ctx->stream) == cudaSuccess) return FOV_OK;
  (void)cudaGetLastError();  // not a stream-orderable allocation on this driver: wait, then free
  FOV_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "fov_free");
  FOV_CUDA(ctx, cudaFree(dptr), "fov_free");
  return FOV_OK;
}

int fov_memset(fov_ctx *ctx, void *dptr, int byte, size_t nbytes) {
  FOV_REQUIRE_CTX(ctx);
  DeviceGuard g(ctx);
  FOV_CUDA(ctx, cudaMemsetAsync(dptr, byte, nbytes, ctx->stream), "fov_memset");
  return FOV_OK;
}

int fov_memcpy_h2d(fov_ctx *ctx, void *dst, const void *src, size_t nbytes) {
  FOV_REQUIRE_CTX(ctx);
  DeviceGuard g(ctx);
  FOV_CUDA(ctx, cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, ctx->stream),
           "fov_memcpy_h2d");
  FOV_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "fov_memcpy_h2d");
  return FOV_OK;
}

int fov_memcpy_d2h(fov_ctx *ctx, void *dst, const void *src, size_t nbytes) {
  FOV_REQUIRE_CTX(ctx);
  DeviceGuard g(ctx);
  FOV_CUDA(ctx, cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToHost, ctx->stream),
           "fov_memcpy_d2h");
  FOV_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "fov_memcpy_d2h");
  return FOV_OK;
}

int fov_memcpy_h2d_async(fov_ctx *ctx, void *dst, const void *src, size_t nbytes) {
  FOV_REQUIRE_CTX(ctx);
  DeviceGuard g(ctx);
  FOV_CUDA(ctx, cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, ctx->stream),
           "fov_memcpy_h2d_async");
  return FOV_OK;
}

int fov_memcpy_d2h_async(fov_ctx *ctx, void *dst, const void *src, size_t nbytes) {
  FOV_REQUIRE_CTX(ctx);
  DeviceGuard g(ctx);
  FOV_CUDA(ctx, cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToHost, ctx->stream),
           "fov_memcpy_d2h_async");
  return FOV_OK;
}

int fov_host_alloc(fov_ctx *ctx, void **hptr, size_t nbytes) {
  FOV_REQUIRE_CTX(ctx);
  if (!hptr) return fail(ctx, FOV_ERR_INVALID, "fov_host_alloc: null output pointer");
  DeviceGuard g(ctx);
  FOV_CUDA(ctx, cudaHostAlloc(hptr, nbytes ? nbytes : 1, cudaHostAllocDefault), "fov_host_alloc");
  return FOV_OK;
}

int fov_host_free(fov_ctx *ctx, void *hptr) {
  FOV_REQUIRE_CTX(ctx);
  DeviceGuard g(ctx);
  FOV_CUDA(ctx, cudaFreeHost(hptr), "fov_host_free");
  return FOV_OK;
}

// ---- SATEncoder ---------------------------------------------------------------------------

int fov_sat_encode_batched(fov_ctx *ctx, int n, uint32_t *sat, size_t sat_stride,
                           const uint8_t *src, size_t src_stride, int W, int H, int linesize) {
  FOV_REQUIRE_CTX(ctx);
  if (n <= 0 || !sat || !src || W <= 0 || H <= 0 || linesize < 3 * W || linesize / W < 3)
    return fail(ctx, FOV_ERR_INVALID, "fov_sat_encode: invalid arguments");
  if (((uintptr_t)sat % 4) != 0 || (sat_stride % 4) != 0)
    return fail(ctx, FOV_ERR_INVALID, "fov_sat_encode: SAT buffer must be 4-byte aligned");
  if (n > 65535) return fail(ctx, FOV_ERR_INVALID, "fov_sat_encode: batch too large");
  DeviceGuard g(ctx);
  const bool onepass = sat_onepass_eligible(sat, sat_stride, src, src_stride, W, H, linesize);
  int rc = ensure_sat_scratch(ctx, n, W, H, onepass);
  if (rc) return rc;
  if (onepass) {
    if (ctx->sat_scratch_dirty || ctx->sat_epoch >= 0x3ffffff0u) {
      FOV_NOT_WHILE_CAPTURING(ctx, "sat scratch clear");
      // Ticket counters and carry tags start from zero.  Alternating geometries or batch sizes need
      // no clear: tags carry a context-wide epoch, so units of other layouts are simply stale.
      FOV_CUDA(ctx, cudaMemsetAsync(ctx->sat_scratch.base, 0, ctx->sat_scratch.bytes, ctx->stream),
               "sat scratch clear");
      ctx->sat_scratch_dirty = false;
      ctx->sat_epoch = 0;
    }
    FOV_CUDA(ctx,
             launch_sat_onepass(ctx->lc(), n, sat, sat_stride, src, src_stride, W, H, linesize,
                                ctx->sat_scratch.base),
             "sat encode launch");
    ++ctx->sat_epoch;
    return FOV_OK;
  }
  // Unaligned or 3-byte-pixel sources: the three-kernel reduce / carry / scan path.
  ctx->sat_scratch_dirty = true;  // it reuses the same scratch bytes
  FOV_CUDA(ctx,
           launch_sat_encode(ctx->lc(), n, sat, sat_stride, src, src_stride, W, H, linesize,
                             ctx->sat_scratch.base),
           "sat encode launch");
  return FOV_OK;
}

int fov_sat_encode(fov_ctx *ctx, uint32_t *sat, const uint8_t *src, int W, int H, int linesize) {
  return fov_sat_encode_batched(ctx, 1, sat, 0, src, 0, W, H, linesize);
}

// ---- SATDecoder ---------------------------------------------------------------------------

int fov_sat_grid_init(fov_ctx *ctx, int ow, int oh, int W, int H) {
  FOV_REQUIRE_CTX(ctx);
  DeviceGuard g(ctx);
  const SatGrid *grid = nullptr;
  int rc = get_sat_grid(ctx, ow, oh, W, H, &grid);
  if (rc) return rc;
  ctx->cur_sat_grid = grid;
  const InterpLut *lut = nullptr;
  return get_interp_lut(ctx, W, H, ow, oh, &lut);  // warm the inverse-warp table as well
}

int fov_sat_grid_export(fov_ctx *ctx, int16_t *host_grid, int ow, int oh, int W, int H) {
  FOV_REQUIRE_CTX(ctx);
  if (!host_grid) return fail(ctx, FOV_ERR_INVALID, "fov_sat_grid_export: null pointer");
  DeviceGuard g(ctx);
  const SatGrid *grid = nullptr;
  int rc = get_sat_grid(ctx, ow, oh, W, H, &grid);
  if (rc) return rc;
  // Round-trip through the device copies so the export checks what the kernels actually read.
  std::vector<int16_t> xe(ow + 1), ye(oh + 1);
  FOV_CUDA(ctx, cudaMemcpy(xe.data(), grid->d_xedge, xe.size() * 2, cudaMemcpyDeviceToHost),
           "grid export");
  FOV_CUDA(ctx, cudaMemcpy(ye.data(), grid->d_yedge, ye.size() * 2, cudaMemcpyDeviceToHost),
           "grid export");
  for (int ty = 0; ty <= oh; ++ty)
    for (int tx = 0; tx <= ow; ++tx) {
      host_grid[((size_t)ty * (ow + 1) + tx) * 2] = xe[tx];
      host_grid[((size_t)ty * (ow + 1) + tx) * 2 + 1] = ye[ty];
    }
  return FOV_OK;
}

namespace {
// Where the per-frame gaze comes from: a host array (copied into the launch parameters) or a device
// array the kernels read themselves (graph-replayable launches).
struct GazeSrc {
  const float *host = nullptr;
  const float *dev = nullptr;
  bool ok() const { return host != nullptr || dev != nullptr; }
  GazeBatch batch(int f0, int m) const {
    GazeBatch gz;
    memset(gz.xy, 0, sizeof(gz.xy));
    if (dev)
      gz.dev = dev + 2 * (size_t)f0;
    else
      memcpy(gz.xy, host + 2 * (size_t)f0, sizeof(float) * 2 * m);
    return gz;
  }
};

// sample_rect for n frames.  `src` (optional) are the RGB0 frames the SATs were built from by this
// library in the same call sequence: the kernel may read 1x1 boxes from them (identical bits).
int sample_rect_batched(fov_ctx *ctx, int n, uint8_t *out, size_t out_stride, int ow, int oh,
                        int out_linesize, const uint32_t *sat, size_t sat_stride, int W, int H,
                        const GazeSrc &gaze, const uint8_t *src, size_t src_stride,
                        int src_linesize) {
  FOV_REQUIRE_CTX(ctx);
  if (n <= 0 || !out || !sat || !gaze.ok() || ow <= 0 || oh <= 0 || out_linesize < 4 * ow ||
      (out_linesize % 4) != 0 || ((uintptr_t)out % 4) != 0 || (out_stride % 4) != 0 ||
      ((uintptr_t)sat % 4) != 0 || (sat_stride % 4) != 0 || W < 2 || H < 2)
    return fail(ctx, FOV_ERR_INVALID, "fov_sat_sample_rect: invalid arguments");
  DeviceGuard g(ctx);
  const SatGrid *grid = nullptr;
  int rc = get_sat_grid(ctx, ow, oh, W, H, &grid);  // lazy init, sat_decoder.cc:312-317
  if (rc) return rc;
  ctx->cur_sat_grid = grid;
  for (int f0 = 0; f0 < n; f0 += kMaxBatchPerLaunch) {
    const int m = n - f0 < kMaxBatchPerLaunch ? n - f0 : kMaxBatchPerLaunch;
    const GazeBatch gz = gaze.batch(f0, m);
    FOV_CUDA(ctx,
             launch_sat_sample_rect(
                 ctx->lc(), m, out + (size_t)f0 * out_stride, out_stride, ow, oh, out_linesize,
                 reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(sat) +
                                                    (size_t)f0 * sat_stride),
                 sat_stride, W, H, grid->d_xedge, grid->d_yedge, gz,
                 src ? src + (size_t)f0 * src_stride : nullptr, src_stride, src_linesize),
             "sample_rect launch");
  }
  return FOV_OK;
}

int interpolate_rect_batched(fov_ctx *ctx, int n, uint8_t *out, size_t out_stride, int W, int H,
                             const uint8_t *red, size_t red_stride, int ow, int oh,
                             const GazeSrc &gaze) {
  FOV_REQUIRE_CTX(ctx);
  if (n <= 0 || !out || !red || !gaze.ok() || ((uintptr_t)out % 4) != 0 || (out_stride % 4) != 0 ||
      ((uintptr_t)red % 4) != 0 || (red_stride % 4) != 0)
    return fail(ctx, FOV_ERR_INVALID, "fov_sat_interpolate_rect: invalid arguments");
  DeviceGuard g(ctx);
  const InterpLut *lut = nullptr;
  int rc = get_interp_lut(ctx, W, H, ow, oh, &lut);
  if (rc) return rc;
  for (int f0 = 0; f0 < n; f0 += kMaxBatchPerLaunch) {
    const int m = n - f0 < kMaxBatchPerLaunch ? n - f0 : kMaxBatchPerLaunch;
    const GazeBatch gz = gaze.batch(f0, m);
    FOV_CUDA(ctx,
             launch_sat_interpolate_rect(ctx->lc(), m, out + (size_t)f0 * out_stride, out_stride,
                                         W, H, red + (size_t)f0 * red_stride, red_stride, ow, oh,
                                         lut->d_x, lut->d_y, gz),
             "interpolate_rect launch");
  }
  return FOV_OK;
}

int encode_sample_batched(fov_ctx *ctx, int n, uint8_t *reduced, size_t red_stride, uint32_t *sat,
                          size_t sat_stride, const uint8_t *src, size_t src_stride, int W, int H,
                          int linesize, int ow, int oh, const GazeSrc &gaze) {
  int rc = fov_sat_encode_batched(ctx, n, sat, sat_stride, src, src_stride, W, H, linesize);
  if (rc) return rc;
  // RGB0 frames with 4-byte aligned rows let sample_rect read its 1x1 boxes from the frame
  static const bool no_hint = getenv("FOV360_SAMPLE_NO_SRC") != nullptr;
  const bool hint = !no_hint && linesize / W == 4 && (linesize % 4) == 0 && ((uintptr_t)src % 4) == 0 &&
                    (src_stride % 4) == 0;
  return sample_rect_batched(ctx, n, reduced, red_stride, ow, oh, 4 * ow, sat, sat_stride, W, H,
                             gaze, hint ? src : nullptr, src_stride, linesize);
}
}  // namespace

int fov_sat_sample_rect_batched(fov_ctx *ctx, int n, uint8_t *out, size_t out_stride, int ow,
                                int oh, int out_linesize, const uint32_t *sat, size_t sat_stride,
                                int W, int H, const float *gaze_xy) {
  return sample_rect_batched(ctx, n, out, out_stride, ow, oh, out_linesize, sat, sat_stride, W, H,
                             GazeSrc{gaze_xy, nullptr}, nullptr, 0, 0);
}

int fov_sat_sample_rect(fov_ctx *ctx, uint8_t *out, int ow, int oh, int out_linesize,
                        const uint32_t *sat, int W, int H, float cx, float cy) {
  const float gz[2] = {cx, cy};
  return fov_sat_sample_rect_batched(ctx, 1, out, 0, ow, oh, out_linesize, sat, 0, W, H, gz);
}

int fov_sat_interpolate_rect_batched(fov_ctx *ctx, int n, uint8_t *out, size_t out_stride, int W,
                                     int H, const uint8_t *red, size_t red_stride, int ow, int oh,
                                     const float *gaze_xy) {
  return interpolate_rect_batched(ctx, n, out, out_stride, W, H, red, red_stride, ow, oh,
                                  GazeSrc{gaze_xy, nullptr});
}

int fov_sat_interpolate_rect(fov_ctx *ctx, uint8_t *out, int W, int H, int out_linesize,
                             const uint8_t *red, int ow, int oh, int red_linesize, float cx,
                             float cy) {
  (void)out_linesize;  // unused by the reference kernel as well (sat_decoder.cc:902-912)
  (void)red_linesize;
  const float gz[2] = {cx, cy};
  return fov_sat_interpolate_rect_batched(ctx, 1, out, 0, W, H, red, 0, ow, oh, gz);
}

int fov_sat_decode(fov_ctx *ctx, uint8_t *out, int out_linesize, const uint32_t *sat, int W,
                   int H) {
  FOV_REQUIRE_CTX(ctx);
  if (!out || !sat || W <= 0 || H <= 0 || out_linesize / W < 3)
    return fail(ctx, FOV_ERR_INVALID, "fov_sat_decode: invalid arguments");
  DeviceGuard g(ctx);
  FOV_CUDA(ctx, launch_sat_decode(ctx->lc(), out, out_linesize, sat, W, H), "decode launch");
  return FOV_OK;
}

int fov_sat_encode_sample_batched(fov_ctx *ctx, int n, uint8_t *reduced, size_t red_stride,
                                  uint32_t *sat, size_t sat_stride, const uint8_t *src,
                                  size_t src_stride, int W, int H, int linesize, int ow, int oh,
                                  const float *gaze_xy) {
  return encode_sample_batched(ctx, n, reduced, red_stride, sat, sat_stride, src, src_stride, W, H,
                               linesize, ow, oh, GazeSrc{gaze_xy, nullptr});
}

int fov_sat_foveate_batched(fov_ctx *ctx, int n, uint8_t *full_out, size_t full_stride,
                            uint8_t *reduced, size_t red_stride, uint32_t *sat, size_t sat_stride,
                            const uint8_t *src, size_t src_stride, int W, int H, int linesize,
                            int ow, int oh, const float *gaze_xy) {
  const GazeSrc gaze{gaze_xy, nullptr};
  int rc = encode_sample_batched(ctx, n, reduced, red_stride, sat, sat_stride, src, src_stride, W,
                                 H, linesize, ow, oh, gaze);
  if (rc) return rc;
  return interpolate_rect_batched(ctx, n, full_out, full_stride, W, H, reduced, red_stride, ow, oh,
                                  gaze);
}

// The same two sequences with the gaze in DEVICE memory (2n floats): the launches carry nothing that
// changes from frame to frame, so they can be captured once and replayed (fov_graph_*).
int fov_sat_encode_sample_batched_dev(fov_ctx *ctx, int n, uint8_t *reduced, size_t red_stride,
                                      uint32_t *sat, size_t sat_stride, const uint8_t *src,
                                      size_t src_stride, int W, int H, int linesize, int ow, int oh,
                                      const float *gaze_dev) {
  return encode_sample_batched(ctx, n, reduced, red_stride, sat, sat_stride, src, src_stride, W, H,
                               linesize, ow, oh, GazeSrc{nullptr, gaze_dev});
}

int fov_sat_foveate_batched_dev(fov_ctx *ctx, int n, uint8_t *full_out, size_t full_stride,
                                uint8_t *reduced, size_t red_stride, uint32_t *sat,
                                size_t sat_stride, const uint8_t *src, size_t src_stride, int W,
                                int H, int linesize, int ow, int oh, const float *gaze_dev) {
  const GazeSrc gaze{nullptr, gaze_dev};
  int rc = encode_sample_batched(ctx, n, reduced, red_stride, sat, sat_stride, src, src_stride, W,
                                 H, linesize, ow, oh, gaze);
  if (rc) return rc;
  return interpolate_rect_batched(ctx, n, full_out, full_stride, W, H, reduced, red_stride, ow, oh,
                                  gaze);
}

// ---- CUDA-graph capture of a call sequence (run_satlogrectilinear.cc:926-943 as one submission) --

int fov_graph_begin_capture(fov_ctx *ctx) {
  FOV_REQUIRE_CTX(ctx);
  if (ctx->capturing) return fail(ctx, FOV_ERR_INVALID, "fov_graph_begin_capture: already capturing");
  DeviceGuard g(ctx);
  FOV_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal),
           "fov_graph_begin_capture");
  ctx->capturing = true;
  ctx->capture_launches0 = ctx->launches;
  ctx->capture_sat0 = ctx->sat_epoch;
  return FOV_OK;
}

int fov_graph_end_capture(fov_ctx *ctx, fov_graph **out) {
  FOV_REQUIRE_CTX(ctx);
  if (!ctx->capturing) return fail(ctx, FOV_ERR_INVALID, "fov_graph_end_capture: not capturing");
  DeviceGuard g(ctx);
  ctx->capturing = false;
  std::unique_ptr<fov_graph> gr(new fov_graph);
  cudaError_t e = cudaStreamEndCapture(ctx->stream, &gr->graph);
  if (e == cudaSuccess) e = cudaGraphInstantiate(&gr->exec, gr->graph, 0);
  if (e != cudaSuccess) {
    if (gr->graph) cudaGraphDestroy(gr->graph);
    (void)cudaGetLastError();
    if (out) *out = nullptr;
    return cuda_fail(ctx, e, "fov_graph_end_capture");
  }
  gr->kernels = ctx->launches - ctx->capture_launches0;
  gr->sat_launches = ctx->sat_epoch - ctx->capture_sat0;
  gr->scratch_generation = ctx->scratch_generation;
  // the captured launches did not run: give their counts back
  ctx->launches = ctx->capture_launches0;
  ctx->sat_epoch = ctx->capture_sat0;
  if (!out) return fail(ctx, FOV_ERR_INVALID, "fov_graph_end_capture: null output pointer");
  *out = gr.release();
  return FOV_OK;
}

int fov_graph_launch(fov_ctx *ctx, fov_graph *graph) {
  FOV_REQUIRE_CTX(ctx);
  if (!graph || !graph->exec) return fail(ctx, FOV_ERR_INVALID, "fov_graph_launch: no graph");
  if (ctx->capturing) return fail(ctx, FOV_ERR_INVALID, "fov_graph_launch: a capture is in progress");
  if (graph->scratch_generation != ctx->scratch_generation)
    return fail(ctx, FOV_ERR_INVALID,
                "fov_graph_launch: the context's SAT scratch moved since the capture (a larger call "
                "ran in between): capture the sequence again");
  DeviceGuard g(ctx);
  if (graph->sat_launches &&
      (ctx->sat_scratch_dirty || (uint64_t)ctx->sat_epoch + graph->sat_launches >= 0x3ffffff0u)) {
    // the three-kernel fallback used the scratch since the capture, or the 30-bit epoch tag is about
    // to wrap: clear the carry units and the device-side counter before the replay
    FOV_CUDA(ctx, cudaMemsetAsync(ctx->sat_scratch.base, 0, ctx->sat_scratch.bytes, ctx->stream),
             "sat scratch clear");
    ctx->sat_scratch_dirty = false;
    ctx->sat_epoch = 0;
  }
  FOV_CUDA(ctx, cudaGraphLaunch(graph->exec, ctx->stream), "fov_graph_launch");
  ctx->launches += graph->kernels;
  ctx->sat_epoch += graph->sat_launches;
  return FOV_OK;
}

void fov_graph_destroy(fov_ctx *ctx, fov_graph *graph) {
  if (!graph) return;
  if (ctx) {
    DeviceGuard g(ctx);
    cudaStreamSynchronize(ctx->stream);
  }
  if (graph->exec) cudaGraphExecDestroy(graph->exec);
  if (graph->graph) cudaGraphDestroy(graph->graph);
  delete graph;
}

// ---- Projections ----------------------------------------------------------------------------

int fov_gnomonic(fov_ctx *ctx, uint8_t *out, int tw, int th, int out_linesize, const uint8_t *src,
                 int W, int H, int src_linesize, float cx, float cy) {
  FOV_REQUIRE_CTX(ctx);
  (void)out_linesize;  // never reach the reference kernel either (projections.cc:65-71)
  (void)src_linesize;
  if (!out || !src || tw <= 0 || th <= 0 || W <= 0 || H <= 0 || ((uintptr_t)out % 4) != 0 ||
      ((uintptr_t)src % 4) != 0)
    return fail(ctx, FOV_ERR_INVALID, "fov_gnomonic: invalid arguments");
  DeviceGuard g(ctx);
  FOV_CUDA(ctx, launch_gnomonic(ctx->lc(), out, tw, th, src, W, H, make_gnomonic_view(cx, cy)),
           "gnomonic launch");
  return FOV_OK;
}

int fov_sat_interpolate_gnomonic(fov_ctx *ctx, uint8_t *out, int tw, int th, const uint8_t *red,
                                 int ow, int oh, int W, int H, float gaze_x, float gaze_y,
                                 float view_x, float view_y) {
  FOV_REQUIRE_CTX(ctx);
  if (!out || !red || tw <= 0 || th <= 0 || ow <= 0 || oh <= 0 || W <= 0 || H <= 0 ||
      ((uintptr_t)out % 4) != 0 || ((uintptr_t)red % 4) != 0)
    return fail(ctx, FOV_ERR_INVALID, "fov_sat_interpolate_gnomonic: invalid arguments");
  DeviceGuard g(ctx);
  const InterpLut *lut = nullptr;
  int rc = get_interp_lut(ctx, W, H, ow, oh, &lut);
  if (rc) return rc;
  FOV_CUDA(ctx,
           launch_sat_interpolate_gnomonic(ctx->lc(), out, tw, th, red, ow, oh, W, H, lut->d_x,
                                           lut->d_y, gaze_x, gaze_y,
                                           make_gnomonic_view(view_x, view_y)),
           "interpolate_gnomonic launch");
  return FOV_OK;
}

// ---- ImageSampler ---------------------------------------------------------------------------

int fov_img_grid_init(fov_ctx *ctx, int ow, int oh, int W, int H) {
  FOV_REQUIRE_CTX(ctx);
  DeviceGuard g(ctx);
  const ImgGrid *grid = nullptr;
  return get_img_grid(ctx, ow, oh, W, H, &grid);
}

int fov_img_grid_export(fov_ctx *ctx, int16_t *host_grid, int ow, int oh, int W, int H) {
  FOV_REQUIRE_CTX(ctx);
  if (!host_grid) return fail(ctx, FOV_ERR_INVALID, "fov_img_grid_export: null pointer");
  DeviceGuard g(ctx);
  const ImgGrid *grid = nullptr;
  int rc = get_img_grid(ctx, ow, oh, W, H, &grid);
  if (rc) return rc;
  std::vector<int16_t> xd(ow), yd(oh);
  FOV_CUDA(ctx, cudaMemcpy(xd.data(), grid->d_xd, xd.size() * 2, cudaMemcpyDeviceToHost),
           "grid export");
  FOV_CUDA(ctx, cudaMemcpy(yd.data(), grid->d_yd, yd.size() * 2, cudaMemcpyDeviceToHost),
           "grid export");
  for (int j = 0; j < oh; ++j)
    for (int i = 0; i < ow; ++i) {
      host_grid[((size_t)j * ow + i) * 2] = xd[i];
      host_grid[((size_t)j * ow + i) * 2 + 1] = yd[j];
    }
  return FOV_OK;
}

static bool gather_args_ok(const uint8_t *out, int ow, int oh, int out_linesize, const uint8_t *src,
                           int W, int H, int src_linesize) {
  return out && src && ow > 0 && oh > 0 && W > 0 && H > 0 && out_linesize / ow >= 3 &&
         src_linesize / W >= 3;
}

int fov_img_sample_rect(fov_ctx *ctx, uint8_t *out, int ow, int oh, int out_linesize,
                        const uint8_t *src, int W, int H, int src_linesize, float cx, float cy) {
  FOV_REQUIRE_CTX(ctx);
  if (!gather_args_ok(out, ow, oh, out_linesize, src, W, H, src_linesize))
    return fail(ctx, FOV_ERR_INVALID, "fov_img_sample_rect: invalid arguments");
  DeviceGuard g(ctx);
  const ImgGrid *grid = nullptr;
  int rc = get_img_grid(ctx, ow, oh, W, H, &grid);
  if (rc) return rc;
  FOV_CUDA(ctx,
           launch_img_sample_rect(ctx->lc(), out, ow, oh, out_linesize, src, W, H, src_linesize,
                                  grid->d_xd, grid->d_yd, cx, cy),
           "img sample_rect launch");
  return FOV_OK;
}

int fov_img_logpolar_grid_init(fov_ctx *ctx, int ow, int oh, int W, int H) {
  FOV_REQUIRE_CTX(ctx);
  (void)W;  // the log-polar radius is in source pixels and does not depend on W, H
  (void)H;
  DeviceGuard g(ctx);
  const LogpolarGrid *grid = nullptr;
  return get_lp_grid(ctx, ow, oh, &grid);
}

int fov_img_logpolar_grid_export(fov_ctx *ctx, int16_t *host_grid, int ow, int oh) {
  FOV_REQUIRE_CTX(ctx);
  if (!host_grid) return fail(ctx, FOV_ERR_INVALID, "fov_img_logpolar_grid_export: null pointer");
  DeviceGuard g(ctx);
  const LogpolarGrid *grid = nullptr;
  int rc = get_lp_grid(ctx, ow, oh, &grid);
  if (rc) return rc;
  int16_t *d = nullptr;
  const size_t bytes = (size_t)ow * oh * 2 * sizeof(int16_t);
  FOV_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&d), bytes), "logpolar export alloc");
  cudaError_t e = launch_img_logpolar_grid_expand(ctx->lc(), d, ow, oh, grid->d_radius,
                                                  grid->d_cos, grid->d_sin);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(host_grid, d, bytes, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  FOV_CUDA(ctx, e, "logpolar export");
  return FOV_OK;
}

int fov_img_sample_logpolar(fov_ctx *ctx, uint8_t *out, int ow, int oh, int out_linesize,
                            const uint8_t *src, int W, int H, int src_linesize, float cx,
                            float cy) {
  FOV_REQUIRE_CTX(ctx);
  if (!gather_args_ok(out, ow, oh, out_linesize, src, W, H, src_linesize))
    return fail(ctx, FOV_ERR_INVALID, "fov_img_sample_logpolar: invalid arguments");
  DeviceGuard g(ctx);
  const LogpolarGrid *grid = nullptr;
  int rc = get_lp_grid(ctx, ow, oh, &grid);
  if (rc) return rc;
  FOV_CUDA(ctx,
           launch_img_sample_logpolar(ctx->lc(), out, ow, oh, out_linesize, src, W, H,
                                      src_linesize, grid->d_radius, grid->d_cos, grid->d_sin, cx,
                                      cy),
           "img sample_logpolar launch");
  return FOV_OK;
}

int fov_img_interpolate_logpolar(fov_ctx *ctx, uint8_t *out, int W, int H, int out_linesize,
                                 const uint8_t *red, int ow, int oh, int red_linesize, float cx,
                                 float cy) {
  FOV_REQUIRE_CTX(ctx);
  (void)out_linesize;  // unused by the reference kernel as well (image_sampler.cc:794-803)
  (void)red_linesize;
  if (!out || !red || W <= 0 || H <= 0 || ow <= 0 || oh <= 0 || ((uintptr_t)out % 4) != 0 ||
      ((uintptr_t)red % 4) != 0)
    return fail(ctx, FOV_ERR_INVALID, "fov_img_interpolate_logpolar: invalid arguments");
  DeviceGuard g(ctx);
  const LogpolarGrid *grid = nullptr;  // radius / direction tables of the exact-hit check
  int rc = get_lp_grid(ctx, ow, oh, &grid);
  if (rc) return rc;
  FOV_CUDA(ctx,
           launch_img_interpolate_logpolar(ctx->lc(), out, W, H, red, ow, oh, cx, cy, *grid),
           "img interpolate_logpolar launch");
  return FOV_OK;
}

int fov_img_logpolar_blur(fov_ctx *ctx, uint8_t *out, int ow, int oh, int linesize,
                          const uint8_t *src) {
  FOV_REQUIRE_CTX(ctx);
  (void)linesize;  // dense uchar3 addressing in the reference kernel
  if (!out || !src || ow <= 0 || oh <= 0 || ((uintptr_t)out % 4) != 0 || ((uintptr_t)src % 4) != 0)
    return fail(ctx, FOV_ERR_INVALID, "fov_img_logpolar_blur: invalid arguments");
  DeviceGuard g(ctx);
  FOV_CUDA(ctx, launch_img_logpolar_blur(ctx->lc(), out, ow, oh, src), "img blur launch");
  return FOV_OK;
}

// ---- VideoEncoder colour conversion ------------------------------------------------------------

namespace {
int yuv_convert(fov_ctx *ctx, const char *who, bool nv12, int n, uint8_t *y, size_t y_stride,
                int y_ls, uint8_t *u, int u_ls, uint8_t *v, int v_ls, size_t c_stride,
                const uint8_t *src, size_t src_stride, int src_ls, int W, int H) {
  FOV_REQUIRE_CTX(ctx);
  const int cw = nv12 ? W : W / 2;  // bytes per chroma row
  if (!y || !u || (!nv12 && !v) || !src || n < 1 || W <= 0 || H <= 0 || y_ls < W || u_ls < cw ||
      (!nv12 && v_ls < cw) || src_ls < 4 * W || (src_ls & 3) || ((uintptr_t)src & 3) ||
      (src_stride & 3))
    return fail(ctx, FOV_ERR_INVALID, std::string(who) + ": invalid arguments");
  if ((W & 1) || (H & 1) || H < 8)
    return fail(ctx, FOV_ERR_UNSUPPORTED,
                std::string(who) + ": width and height must be even and height >= 8");
  DeviceGuard g(ctx);
  for (int f0 = 0; f0 < n; f0 += 65535) {  // gridDim.z limit
    const int nf = std::min(n - f0, 65535);
    FOV_CUDA(ctx,
             launch_rgb0_to_yuv(ctx->lc(), nv12, nf, y + (size_t)f0 * y_stride, y_stride, y_ls,
                                u + (size_t)f0 * c_stride, u_ls,
                                v ? v + (size_t)f0 * c_stride : nullptr, v_ls, c_stride,
                                src + (size_t)f0 * src_stride, src_stride, src_ls, W, H),
             who);
  }
  return FOV_OK;
}
}  // namespace

int fov_rgb0_to_yuv420p(fov_ctx *ctx, uint8_t *y, int y_linesize, uint8_t *u, int u_linesize,
                        uint8_t *v, int v_linesize, const uint8_t *src, int src_linesize,
                        int width, int height) {
  return yuv_convert(ctx, "fov_rgb0_to_yuv420p", false, 1, y, 0, y_linesize, u, u_linesize, v,
                     v_linesize, 0, src, 0, src_linesize, width, height);
}

int fov_rgb0_to_nv12(fov_ctx *ctx, uint8_t *y, int y_linesize, uint8_t *uv, int uv_linesize,
                     const uint8_t *src, int src_linesize, int width, int height) {
  return yuv_convert(ctx, "fov_rgb0_to_nv12", true, 1, y, 0, y_linesize, uv, uv_linesize, nullptr,
                     0, 0, src, 0, src_linesize, width, height);
}

int fov_rgb0_to_yuv420p_batched(fov_ctx *ctx, int n, uint8_t *y, size_t y_stride, int y_linesize,
                                uint8_t *u, uint8_t *v, size_t chroma_stride, int chroma_linesize,
                                const uint8_t *src, size_t src_stride, int src_linesize, int width,
                                int height) {
  return yuv_convert(ctx, "fov_rgb0_to_yuv420p_batched", false, n, y, y_stride, y_linesize, u,
                     chroma_linesize, v, chroma_linesize, chroma_stride, src, src_stride,
                     src_linesize, width, height);
}

int fov_rgb0_to_nv12_batched(fov_ctx *ctx, int n, uint8_t *y, size_t y_stride, int y_linesize,
                             uint8_t *uv, size_t uv_stride, int uv_linesize, const uint8_t *src,
                             size_t src_stride, int src_linesize, int width, int height) {
  return yuv_convert(ctx, "fov_rgb0_to_nv12_batched", true, n, y, y_stride, y_linesize, uv,
                     uv_linesize, nullptr, 0, uv_stride, src, src_stride, src_linesize, width,
                     height);
}

// ---- VideoDecoder colour conversion ------------------------------------------------------------

namespace {
int rgb_convert(fov_ctx *ctx, const char *who, bool nv12, int n, uint8_t *dst, size_t dst_stride,
                int dst_ls, const uint8_t *y, size_t y_stride, int y_ls, const uint8_t *u, int u_ls,
                const uint8_t *v, int v_ls, size_t c_stride, int W, int H) {
  FOV_REQUIRE_CTX(ctx);
  const int cw = nv12 ? W : W / 2;  // bytes per chroma row
  if (!dst || !y || !u || (!nv12 && !v) || n < 1 || W <= 0 || H <= 0 || y_ls < W || u_ls < cw ||
      (!nv12 && v_ls < cw) || dst_ls < 4 * W || (dst_ls & 3) || ((uintptr_t)dst & 3) ||
      (dst_stride & 3))
    return fail(ctx, FOV_ERR_INVALID, std::string(who) + ": invalid arguments");
  if ((W & 1) || (H & 1))
    return fail(ctx, FOV_ERR_UNSUPPORTED, std::string(who) + ": width and height must be even");
  DeviceGuard g(ctx);
  for (int f0 = 0; f0 < n; f0 += 65535) {  // gridDim.z limit
    const int nf = std::min(n - f0, 65535);
    FOV_CUDA(ctx,
             launch_yuv_to_rgb0(ctx->lc(), nv12, nf, dst + (size_t)f0 * dst_stride, dst_stride,
                                dst_ls, y + (size_t)f0 * y_stride, y_stride, y_ls,
                                u + (size_t)f0 * c_stride, u_ls,
                                v ? v + (size_t)f0 * c_stride : nullptr, v_ls, c_stride, W, H),
             who);
  }
  return FOV_OK;
}
}  // namespace

int fov_yuv420p_to_rgb0(fov_ctx *ctx, uint8_t *dst, int dst_linesize, const uint8_t *y,
                        int y_linesize, const uint8_t *u, int u_linesize, const uint8_t *v,
                        int v_linesize, int width, int height) {
  return rgb_convert(ctx, "fov_yuv420p_to_rgb0", false, 1, dst, 0, dst_linesize, y, 0, y_linesize,
                     u, u_linesize, v, v_linesize, 0, width, height);
}

int fov_nv12_to_rgb0(fov_ctx *ctx, uint8_t *dst, int dst_linesize, const uint8_t *y, int y_linesize,
                     const uint8_t *uv, int uv_linesize, int width, int height) {
  return rgb_convert(ctx, "fov_nv12_to_rgb0", true, 1, dst, 0, dst_linesize, y, 0, y_linesize, uv,
                     uv_linesize, nullptr, 0, 0, width, height);
}

int fov_yuv420p_to_rgb0_batched(fov_ctx *ctx, int n, uint8_t *dst, size_t dst_stride,
                                int dst_linesize, const uint8_t *y, size_t y_stride, int y_linesize,
                                const uint8_t *u, const uint8_t *v, size_t chroma_stride,
                                int chroma_linesize, int width, int height) {
  return rgb_convert(ctx, "fov_yuv420p_to_rgb0_batched", false, n, dst, dst_stride, dst_linesize, y,
                     y_stride, y_linesize, u, chroma_linesize, v, chroma_linesize, chroma_stride,
                     width, height);
}

int fov_nv12_to_rgb0_batched(fov_ctx *ctx, int n, uint8_t *dst, size_t dst_stride, int dst_linesize,
                             const uint8_t *y, size_t y_stride, int y_linesize, const uint8_t *uv,
                             size_t uv_stride, int uv_linesize, int width, int height) {
  return rgb_convert(ctx, "fov_nv12_to_rgb0_batched", true, n, dst, dst_stride, dst_linesize, y,
                     y_stride, y_linesize, uv, uv_linesize, nullptr, 0, uv_stride, width, height);
}

int fov_debug_bounds_violations(fov_ctx *ctx, unsigned *count, unsigned *first_site) {
  FOV_REQUIRE_CTX(ctx);
  DeviceGuard g(ctx);
  FOV_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "fov_debug_bounds_violations");
  unsigned a[2] = {0, 0}, b[2] = {0, 0};
  bounds_read_image_sampler(a);
  bounds_read_sat_decode(b);
  if (count) *count = a[0] + b[0];
  if (first_site) *first_site = a[0] ? a[1] : b[1];
#ifdef FOV360_BOUNDS_CHECK
  return 1;  // a checking build
#else
  return FOV_OK;
#endif
}

int fov_reduced_dim(int full_dim) {
  // 16 * ceil(dim / 1.8 / 16), run_satlogrectilinear.cc:113-114
  return 16 * (int)std::ceil(full_dim / 1.8 / 16);
}

}  // extern "C"
