/*
 * fov360.h - C ABI of the B200-native foveation transform (libfov360.so).
 *
 * This is the drop-in boundary for the per-frame foveation hot path of
 * AugmentariumLab/foveated-360-video.  It replaces the reference's OpenCL layer
 * (src/opencl_manager.{h,cc} + cl::Buffer / cl::copy at the call sites) and is what
 * the C++ classes in include/fov360/<class>.h (SATEncoder / SATDecoder / ImageSampler /
 * OpenCLManager, same names and signatures as the reference) are written against.
 *
 * Conventions (identical to the reference unless stated):
 *   - widths/heights are in pixels, linesizes in BYTES, gaze is two floats in [0,1]
 *     (x right, y down), pixel centre = (int)(c * dim)  (sat_decoder_sample_rect_kernel.cl:176);
 *   - frames are RGB0 u8 (4 B/pixel; channel 3 is padding), the SAT is dense packed
 *     u32[H][W][3] holding wrapping (mod 2^32) sums (sat_encoder.cc:77);
 *   - every operation is enqueued asynchronously on the context's stream (the
 *     reference's in-order command queue); results are visible after fov_sync() or a
 *     blocking fov_memcpy_*;
 *   - a context is thread-compatible (one context per thread), not thread-safe;
 *   - all pointers passed to the fov_sat_* / fov_img_* entry points are DEVICE pointers
 *     (fov_malloc) - they play the role of the reference's cl_mem handles;
 *   - return value 0 = success, otherwise an FOV_ERR_* code (negative) or a positive
 *     cudaError_t; fov_last_error_string() describes the last failure of a context.
 *
 * There is deliberately no CPU fallback: if no CUDA device is usable, fov_ctx_create
 * fails and every other entry point returns FOV_ERR_NO_CONTEXT.
 */
#ifndef FOV360_H_
#define FOV360_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FOV360_VERSION 100

enum {
  FOV_OK = 0,
  FOV_ERR_NO_CONTEXT = -1,   /* ctx == NULL (the reference's "Not initialized with OpenCL") */
  FOV_ERR_INVALID = -2,      /* bad size / pointer / linesize                                */
  FOV_ERR_NO_DEVICE = -3,    /* no usable CUDA device: there is no CPU fallback              */
  FOV_ERR_GRID = -4,         /* grid / LUT not initialised and cannot be derived             */
  FOV_ERR_UNSUPPORTED = -5
};

typedef struct fov_ctx fov_ctx;

/* ---- device runtime: replaces OpenCLManager (opencl_manager.h:8-22, .cc:7-67) ---------- */

/* OpenCLManager::InitializeContext: binds `device`, creates the in-order stream. */
fov_ctx *fov_ctx_create(int device, int *err);
void fov_ctx_destroy(fov_ctx *ctx);
/* OpenCLManager::GetCLErrorString analogue; ctx may be NULL (global creation errors). */
const char *fov_last_error_string(const fov_ctx *ctx);
int fov_device_count(void);
int fov_ctx_device(const fov_ctx *ctx);
/* The context's cudaStream_t (as void*), for callers that record their own CUDA events. */
void *fov_ctx_stream(const fov_ctx *ctx);
/* clFlush + clFinish (video_server.cc:302-303). */
int fov_sync(fov_ctx *ctx);
/* Number of kernels launched through this context so far (bench accounting). */
uint64_t fov_ctx_launch_count(const fov_ctx *ctx);

/* Context options: promises a caller may make that the reference's interface has no way to state.
 * All default to 0, i.e. exactly the reference's behaviour for arbitrary buffer contents.
 *
 * FOV_OPT_REDUCED_PAD_ZERO: byte 3 of every pixel of the reduced buffers handed to
 *   SampleFrameRectGPU (all fov_sat_sample_rect* / encode_sample* / foveate* entry points) is 0 and
 *   may be (re)written as 0 - e.g. the buffer was cleared once when it was allocated, as a buffer
 *   that goes on to a video encoder normally is.  sample_rect then writes each sampled pixel as one
 *   32-bit word (r, g, b, 0) instead of the reference's 3-byte .xyz store
 *   (sat_decoder_sample_rect_kernel.cl:240).  Under the promise the buffer contents are identical;
 *   pixels whose box misses the frame are still left untouched.  Kernel time -7 % at 8K x 16. */
enum { FOV_OPT_REDUCED_PAD_ZERO = 1 };
int fov_ctx_set_option(fov_ctx *ctx, int option, int value);
int fov_ctx_get_option(const fov_ctx *ctx, int option, int *value);

/* Per-kernel device timing (the role of CL_QUEUE_PROFILING_ENABLE + clGetEventProfilingInfo;
 * the reference creates its queue without it, opencl_manager.cc:55).  While enabled, every kernel
 * launched through the context is bracketed by CUDA events on the context's stream; totals are
 * kept per kernel name.  fov_profile_count() synchronises the stream and returns the number of
 * distinct kernels seen; fov_profile_get() reads entry `index` (name, summed ms, launches). */
int fov_profile_enable(fov_ctx *ctx, int on);
int fov_profile_reset(fov_ctx *ctx);
int fov_profile_count(fov_ctx *ctx);
int fov_profile_get(fov_ctx *ctx, int index, char *name, size_t name_cap, double *total_ms,
                    uint64_t *launches);

/* cl::Buffer(context, CL_MEM_READ_WRITE, n) (video_server.cc:224-232). */
int fov_malloc(fov_ctx *ctx, void **dptr, size_t nbytes);
int fov_free(fov_ctx *ctx, void *dptr);
int fov_memset(fov_ctx *ctx, void *dptr, int byte, size_t nbytes);
/* Blocking copies = cl::copy (video_server.cc:297-299, 342-345). */
int fov_memcpy_h2d(fov_ctx *ctx, void *dst_dev, const void *src_host, size_t nbytes);
int fov_memcpy_d2h(fov_ctx *ctx, void *dst_host, const void *src_dev, size_t nbytes);
/* Stream-ordered copies for pipelined callers (host memory should be pinned). */
int fov_memcpy_h2d_async(fov_ctx *ctx, void *dst_dev, const void *src_host, size_t nbytes);
int fov_memcpy_d2h_async(fov_ctx *ctx, void *dst_host, const void *src_dev, size_t nbytes);
int fov_host_alloc(fov_ctx *ctx, void **hptr, size_t nbytes); /* pinned host memory */
int fov_host_free(fov_ctx *ctx, void *hptr);

/* ---- SATEncoder (sat_encoder.h:39-40) ------------------------------------------------- */

/* SATEncoder::EncodeFrameGPU (sat_encoder.cc:67-135; copy_image_kernel + scan_rows_kernel +
 * scan_columns_kernel, sat_encoder_encode_kernels.cl:1-20,44-74).
 * src: u8[H][src_linesize], pixel stride src_linesize / W (3 or 4), channels 0..2 used.
 * sat: u32[H][W][3], S[y][x][c] = sum_{y'<=y, x'<=x} I[y'][x'][c] mod 2^32. */
int fov_sat_encode(fov_ctx *ctx, uint32_t *sat, const uint8_t *src, int width, int height,
                   int src_linesize);
/* n independent frames, frame f at base + f*stride (strides in BYTES). One launch set. */
int fov_sat_encode_batched(fov_ctx *ctx, int n, uint32_t *sat, size_t sat_stride,
                           const uint8_t *src, size_t src_stride, int width, int height,
                           int src_linesize);

/* ---- SATDecoder (sat_decoder.h:48-51, 63-66, 78-82) ------------------------------------- */

/* SATDecoder::InitializeGrid (sat_decoder.cc:139-170; create_grid_kernel,
 * sat_decoder_sample_rect_kernel.cl:243-295).  The reference stores an int16
 * [(oh+1)][(ow+1)][2] table; its x component depends on the column only and its y
 * component on the row only, so this library keeps the two 1-D edge tables.  They are
 * computed on the host at init time (truncating float formulas: bit-exactness requires
 * the same libm as the oracle) and cached per (ow, oh, W, H). */
int fov_sat_grid_init(fov_ctx *ctx, int out_width, int out_height, int src_width, int src_height);
/* Expands the cached tables into the reference's int16[(oh+1)][(ow+1)][2] layout (host memory). */
int fov_sat_grid_export(fov_ctx *ctx, int16_t *host_grid, int out_width, int out_height,
                        int src_width, int src_height);

/* SATDecoder::SampleFrameRectGPU (sat_decoder.cc:301-348; sample_rect_kernel,
 * sat_decoder_sample_rect_kernel.cl:138-241).  `src_width/src_height` are the reference's
 * codec_ctx->width/height.  out: uchar4[oh][out_linesize/4]; only bytes 0..2 of a pixel
 * are written, and only when its box touches the frame - other bytes keep their contents.
 * The grid is built lazily when absent (sat_decoder.cc:312-317). */
int fov_sat_sample_rect(fov_ctx *ctx, uint8_t *out, int out_width, int out_height,
                        int out_linesize, const uint32_t *sat, int src_width, int src_height,
                        float center_x, float center_y);
/* gaze_xy: HOST array of 2*n floats (x0,y0,x1,y1,...). */
int fov_sat_sample_rect_batched(fov_ctx *ctx, int n, uint8_t *out, size_t out_stride,
                                int out_width, int out_height, int out_linesize,
                                const uint32_t *sat, size_t sat_stride, int src_width,
                                int src_height, const float *gaze_xy);

/* SATDecoder::InterpolateFrameRectGPU (sat_decoder.cc:887-927; interpolate_rect_kernel,
 * sat_decoder_interpolate_kernel.cl:1-152).  Inverse log-rectilinear warp + bilinear.
 * As in the reference, both buffers are addressed as dense 4-byte-pixel arrays and the two
 * linesize arguments are NOT used (sat_decoder.cc:902-912); all 4 bytes of every target
 * pixel are written. */
int fov_sat_interpolate_rect(fov_ctx *ctx, uint8_t *out, int out_width, int out_height,
                             int out_linesize, const uint8_t *reduced, int red_width,
                             int red_height, int red_linesize, float center_x, float center_y);
int fov_sat_interpolate_rect_batched(fov_ctx *ctx, int n, uint8_t *out, size_t out_stride,
                                     int out_width, int out_height, const uint8_t *reduced,
                                     size_t red_stride, int red_width, int red_height,
                                     const float *gaze_xy);

/* SATDecoder::DecodeFrameGPU (sat_decoder.cc:176-210; decode_kernel,
 * sat_decoder_decode_kernel.cl:1-58): exact SAT -> image inverse (1x1 boxes), clamped to
 * [0,255]; out pixel stride = out_linesize / W, bytes 0..2 written. */
int fov_sat_decode(fov_ctx *ctx, uint8_t *out, int out_linesize, const uint32_t *sat, int width,
                   int height);

/* The offline runner's per-frame sequence (run_satlogrectilinear.cc:926-943:
 * EncodeFrameGPU -> SampleFrameRectGPU -> InterpolateFrameRectGPU) for n independent frames
 * with per-frame gaze, enqueued as one batched launch set.  Frame f of every buffer lives at
 * base + f*stride (strides in BYTES); `sat` is scratch the caller owns (n SATs), `reduced`
 * receives the foveated buffers (what the server would hand to the video encoder) and
 * `full_out` the un-warped frames (what the client displays).  Semantics per frame are exactly
 * those of the three single-frame calls. */
int fov_sat_foveate_batched(fov_ctx *ctx, int n, uint8_t *full_out, size_t full_stride,
                            uint8_t *reduced, size_t red_stride, uint32_t *sat, size_t sat_stride,
                            const uint8_t *src, size_t src_stride, int src_width, int src_height,
                            int src_linesize, int red_width, int red_height, const float *gaze_xy);

/* The server's per-frame sequence (video_server.cc:300-338: EncodeFrameGPU ->
 * SampleFrameRectGPU; the inverse warp runs on the client) for n independent frames / streams
 * with per-frame gaze.  Same layout rules as fov_sat_foveate_batched, of which this is the first
 * two stages. */
int fov_sat_encode_sample_batched(fov_ctx *ctx, int n, uint8_t *reduced, size_t red_stride,
                                  uint32_t *sat, size_t sat_stride, const uint8_t *src,
                                  size_t src_stride, int src_width, int src_height,
                                  int src_linesize, int red_width, int red_height,
                                  const float *gaze_xy);

/* The two sequences above with the gaze array in DEVICE memory (2n floats, frame f at 2f, 2f+1):
 * nothing in the launches changes from frame to frame - the gaze is read by the kernels, the SAT
 * build keeps its launch epoch on the device - so a sequence can be captured once as a CUDA graph
 * and replayed (the 3-stage pipeline as a single submission).  The array
 * is read in stream order: whatever wrote it earlier on the context's stream (fov_memcpy_h2d_async,
 * or a kernel of the caller's on fov_ctx_stream) is complete before the kernels read it. */
int fov_sat_encode_sample_batched_dev(fov_ctx *ctx, int n, uint8_t *reduced, size_t red_stride,
                                      uint32_t *sat, size_t sat_stride, const uint8_t *src,
                                      size_t src_stride, int src_width, int src_height,
                                      int src_linesize, int red_width, int red_height,
                                      const float *gaze_xy_dev);
int fov_sat_foveate_batched_dev(fov_ctx *ctx, int n, uint8_t *full_out, size_t full_stride,
                                uint8_t *reduced, size_t red_stride, uint32_t *sat,
                                size_t sat_stride, const uint8_t *src, size_t src_stride,
                                int src_width, int src_height, int src_linesize, int red_width,
                                int red_height, const float *gaze_xy_dev);

/* CUDA-graph capture of whatever is enqueued on the context between begin and end (kernels,
 * fov_memset, fov_memcpy_*_async).  While capturing nothing may allocate, clear or wait: run the
 * same sequence once before fov_graph_begin_capture so that tables and scratch exist (the call
 * fails with FOV_ERR_INVALID and a message otherwise).  Use the *_dev entry points for anything
 * whose gaze changes between replays and update the device array with fov_memcpy_h2d_async before
 * fov_graph_launch.  A graph is tied to its context and to the buffers named at capture time; it
 * is refused (FOV_ERR_INVALID) after the context's SAT scratch had to grow.  No reference
 * counterpart (the OpenCL queue of the reference takes one clEnqueueNDRangeKernel per stage). */
typedef struct fov_graph fov_graph;
int fov_graph_begin_capture(fov_ctx *ctx);
int fov_graph_end_capture(fov_ctx *ctx, fov_graph **graph);
int fov_graph_launch(fov_ctx *ctx, fov_graph *graph);
void fov_graph_destroy(fov_ctx *ctx, fov_graph *graph);

/* ---- ImageSampler (image_sampler.h:57-65, 74-78, 84-91) --------------------------------- */

/* ImageSampler::InitializeGrid (image_sampler.cc:170-202; create_grid_kernel,
 * image_sampler_sample_rect_kernel.cl:48-88): raw log-rect deltas, kept as two 1-D tables. */
int fov_img_grid_init(fov_ctx *ctx, int out_width, int out_height, int src_width, int src_height);
int fov_img_grid_export(fov_ctx *ctx, int16_t *host_grid /* [oh][ow][2] */, int out_width,
                        int out_height, int src_width, int src_height);
/* ImageSampler::SampleFrameRectGPU (image_sampler.cc:249-299; sample_rect_kernel, :1-46). */
int fov_img_sample_rect(fov_ctx *ctx, uint8_t *out, int out_width, int out_height,
                        int out_linesize, const uint8_t *src, int src_width, int src_height,
                        int src_linesize, float center_x, float center_y);

/* ImageSampler::InitializeLogpolarGrid (image_sampler.cc:204-247; create_logpolar_grid_kernel,
 * image_sampler_sample_logpolar_kernel.cl:5-39): (int)(exp(10 i/ow) * (cos,sin)(2 pi j/oh)).
 * Kept as radius[ow], cos[oh], sin[oh] float tables; the product is formed on the device. */
int fov_img_logpolar_grid_init(fov_ctx *ctx, int out_width, int out_height, int src_width,
                               int src_height);
int fov_img_logpolar_grid_export(fov_ctx *ctx, int16_t *host_grid /* [oh][ow][2] */, int out_width,
                                 int out_height);
/* ImageSampler::SampleFrameLogPolarGPU (image_sampler.cc:577-621; sample_logpolar_kernel, :41-86). */
int fov_img_sample_logpolar(fov_ctx *ctx, uint8_t *out, int out_width, int out_height,
                            int out_linesize, const uint8_t *src, int src_width, int src_height,
                            int src_linesize, float center_x, float center_y);
/* ImageSampler::InterpolateFrameLogPolarGPU (image_sampler.cc:780-818;
 * interpolate_logpolar_kernel, image_sampler_interpolate_kernel.cl:1-81); linesizes unused. */
int fov_img_interpolate_logpolar(fov_ctx *ctx, uint8_t *out, int out_width, int out_height,
                                 int out_linesize, const uint8_t *reduced, int red_width,
                                 int red_height, int red_linesize, float center_x, float center_y);
/* ImageSampler::ApplyLogPolarGaussianBlur (image_sampler.cc:820-857;
 * logpolar_gaussian_blur_kernel, image_sampler_sample_logpolar_kernel.cl:88-142). */
int fov_img_logpolar_blur(fov_ctx *ctx, uint8_t *out, int width, int height, int linesize,
                          const uint8_t *src);

/* ---- Projections (projections.h:20-35) ----------------------------------------------------- */

/* Projections::GnomonicProjection (projections.cc:51-86; gnomonic_kernel,
 * projections_program.cl:7-47): inverse gnomonic projection of an out_width x out_height viewport
 * (fixed tangent-plane extent 6 x 3) centred on (center_x, center_y) in [0,1]^2 out of a
 * src_width x src_height equirectangular frame.  Both buffers are dense arrays of 4-byte pixels
 * (the kernel's uchar3); the linesize arguments are accepted and ignored like the reference's. */
int fov_gnomonic(fov_ctx *ctx, uint8_t *out, int out_width, int out_height, int out_linesize,
                 const uint8_t *src, int src_width, int src_height, int src_linesize,
                 float center_x, float center_y);
/* No reference counterpart: fov_sat_interpolate_rect followed by
 * fov_gnomonic in one kernel - the viewport is rendered straight from the reduced buffer and the
 * full-resolution frame is never formed.  (gaze_x, gaze_y) is the foveation centre the reduced
 * buffer was sampled with, (view_x, view_y) the viewport centre. */
int fov_sat_interpolate_gnomonic(fov_ctx *ctx, uint8_t *out, int out_width, int out_height,
                                 const uint8_t *reduced, int red_width, int red_height,
                                 int full_width, int full_height, float gaze_x, float gaze_y,
                                 float view_x, float view_y);

/* ---- VideoEncoder colour conversion (video_encoder.cc:380-398) ------- */

/* The reference hands the foveated RGB0 buffer to NVENC through
 * sws_getContext(w, h, RGB0, w, h, YUV420P, SWS_BILINEAR) + sws_scale on the host and
 * av_hwframe_transfer_data (video_encoder.cc:389-398).  These entry points produce the same planes
 * on the device - bit-exact with libswscale's C arithmetic (SWS_BITEXACT; the x86 SIMD path the
 * plain call takes differs by at most 1 LSB in chroma) - straight into the surface the hardware
 * encoder reads: y/u/v + linesizes are AVFrame::data[0..2] / linesize[0..2] of the AV_PIX_FMT_CUDA
 * frame (sw_format YUV420P, video_encoder.cc:549).  BT.601 limited range, chroma = horizontal
 * pair sums through the matrix, then the vertical taps 1/8 3/8 3/8 1/8 on rows 2j-1..2j+2.
 * width and height must be even and height >= 8 (libswscale builds other filters below that):
 * FOV_ERR_UNSUPPORTED otherwise.  src: RGB0 u8[height][src_linesize]. */
int fov_rgb0_to_yuv420p(fov_ctx *ctx, uint8_t *y, int y_linesize, uint8_t *u, int u_linesize,
                        uint8_t *v, int v_linesize, const uint8_t *src, int src_linesize,
                        int width, int height);
/* Same samples in NVENC's native NV12 layout: uv = u8[height/2][uv_linesize], U and V interleaved. */
int fov_rgb0_to_nv12(fov_ctx *ctx, uint8_t *y, int y_linesize, uint8_t *uv, int uv_linesize,
                     const uint8_t *src, int src_linesize, int width, int height);
/* n frames in one launch (the serving configuration: one reduced buffer per stream); frame f of a
 * plane lives at base + f*stride (BYTES); u and v share chroma_stride and chroma_linesize. */
int fov_rgb0_to_yuv420p_batched(fov_ctx *ctx, int n, uint8_t *y, size_t y_stride, int y_linesize,
                                uint8_t *u, uint8_t *v, size_t chroma_stride, int chroma_linesize,
                                const uint8_t *src, size_t src_stride, int src_linesize, int width,
                                int height);
int fov_rgb0_to_nv12_batched(fov_ctx *ctx, int n, uint8_t *y, size_t y_stride, int y_linesize,
                             uint8_t *uv, size_t uv_stride, int uv_linesize, const uint8_t *src,
                             size_t src_stride, int src_linesize, int width, int height);

/* ---- VideoDecoder colour conversion (video_decoder.cc:165-170, :222) --- */

/* The reference turns every decoded frame into RGB0 with
 * sws_getContext(w, h, YUV420P, w, h, RGB0, SWS_BILINEAR) + sws_scale on the host and uploads it
 * (video_server.cc:291-299).  These entry points do the conversion on the device, so a frame
 * decoded there (NVDEC produces NV12) feeds fov_sat_encode directly.  Bit-exact with libswscale's
 * yuv2rgb converter: chroma replicated over 2x2 blocks, 16-bit fixed-point BT.601 limited-range
 * matrix, 4th byte written as 255.  width and height must be even (FOV_ERR_UNSUPPORTED otherwise).
 * dst: RGB0 u8[height][dst_linesize] (4-byte aligned). */
int fov_yuv420p_to_rgb0(fov_ctx *ctx, uint8_t *dst, int dst_linesize, const uint8_t *y,
                        int y_linesize, const uint8_t *u, int u_linesize, const uint8_t *v,
                        int v_linesize, int width, int height);
int fov_nv12_to_rgb0(fov_ctx *ctx, uint8_t *dst, int dst_linesize, const uint8_t *y, int y_linesize,
                     const uint8_t *uv, int uv_linesize, int width, int height);
/* n frames in one launch; frame f of a plane lives at base + f*stride (BYTES). */
int fov_yuv420p_to_rgb0_batched(fov_ctx *ctx, int n, uint8_t *dst, size_t dst_stride,
                                int dst_linesize, const uint8_t *y, size_t y_stride, int y_linesize,
                                const uint8_t *u, const uint8_t *v, size_t chroma_stride,
                                int chroma_linesize, int width, int height);
int fov_nv12_to_rgb0_batched(fov_ctx *ctx, int n, uint8_t *dst, size_t dst_stride, int dst_linesize,
                             const uint8_t *y, size_t y_stride, int y_linesize, const uint8_t *uv,
                             size_t uv_stride, int uv_linesize, int width, int height);

/* ---- parameters.h semantics ------------------------------------------------------------ */

/* REDUCED_BUFFER_WIDTH/HEIGHT (parameters.h:8-9) for 1920x1080, and the runner's general rule
 * 16*ceil(dim/1.8/16) (run_satlogrectilinear.cc:113-114). */
int fov_reduced_dim(int full_dim);

/* Development aid (no reference counterpart): libraries built with -DFOV360_BOUNDS_CHECK compare
 * every table index and gathered coordinate of the sampling / warp kernels with its limit; this
 * returns the number of violations since the library was loaded and the site of the first one.
 * Returns 1 from a checking build, 0 (and a count of 0) from a normal one. */
int fov_debug_bounds_violations(fov_ctx *ctx, unsigned *count, unsigned *first_site);

#ifdef __cplusplus
}
#endif
#endif /* FOV360_H_ */
