// TEST INFRASTRUCTURE ONLY - not part of the product path.
//
// Host driver for the reference's own OpenCL kernels compiled through clshim.h.
// Each ref_* entry point reproduces the NDRange shape and the argument order of
// the corresponding reference launcher (cited per function) and runs the kernel
// body once per work-item.  Work-items of these kernels are independent, so the
// outer NDRange dimension is distributed over host threads with OpenMP.
//
// The kernel sources are #included from a temporary directory created by
// oracle/build_oracle.py (sed-rewritten copies of /root/reference/src/*.cl that
// are deleted after the build); only the resulting .so lands in oracle/_ref/.
#include "clshim.h"

#include <omp.h>

#include <vector>

thread_local ClshimItem clshim_item;

namespace ref_sat_enc {
#include "sat_encoder_encode_kernels.cl.inc"
}
namespace ref_sat_dec {
#include "sat_decoder_decode_kernel.cl.inc"
}
namespace ref_sat_smp {
#include "sat_decoder_sample_rect_kernel.cl.inc"
}
namespace ref_sat_itp {
#include "sat_decoder_interpolate_kernel.cl.inc"
}
namespace ref_img_smp {
#include "image_sampler_sample_rect_kernel.cl.inc"
}
namespace ref_img_lp {
#include "image_sampler_sample_logpolar_kernel.cl.inc"
}
namespace ref_img_itp {
#include "image_sampler_interpolate_kernel.cl.inc"
}
namespace ref_proj {  // last: its #define PI / PI_2 / DEG2RAD / RAD2DEG are not namespaced
#include "projections_program.cl.inc"
}

// SATEncoder::EncodeFrameCPU (sat_encoder.cc:137-185), the reference's host-side SAT build: the
// function's own text, extracted by build_oracle.py, compiled against the two struct members it
// reads.  Scalar, single-threaded, column-major loops - exactly as the reference ships it.
namespace ref_cpu {
struct AVCodecContext {
  int width, height;
};
struct AVFrame {
  uint8_t *data[8];
  int linesize[8];
};
class SATEncoder {
 public:
  void EncodeFrameCPU(uint32_t *target_frame, AVCodecContext *codec_ctx, AVFrame *frame);
};
#include "encode_frame_cpu.cc.inc"
}  // namespace ref_cpu

namespace {

int g_threads = 0;  // 0 = OpenMP default

inline int roundup8(int v) { return 8 * ((v + 7) / 8); }

// Run f() for every work-item of a 2-D NDRange (gx, gy) with 8x8 work-groups.
template <class F>
void ndrange2(int gx, int gy, F f) {
  const int nt = g_threads > 0 ? g_threads : omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nt)
  for (int y = 0; y < gy; ++y) {
    ClshimItem &it = clshim_item;
    it.gsz[0] = gx;
    it.gsz[1] = gy;
    it.gsz[2] = 1;
    it.lsz[0] = 8;
    it.lsz[1] = 8;
    it.lsz[2] = 1;
    it.gid[1] = y;
    it.lid[1] = y % 8;
    it.gid[2] = it.lid[2] = 0;
    for (int x = 0; x < gx; ++x) {
      it.gid[0] = x;
      it.lid[0] = x % 8;
      f();
    }
  }
}

template <class F>
void ndrange1(int gx, F f) {
  const int nt = g_threads > 0 ? g_threads : omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nt)
  for (int x = 0; x < gx; ++x) {
    ClshimItem &it = clshim_item;
    it.gsz[0] = gx;
    it.gsz[1] = it.gsz[2] = 1;
    it.lsz[0] = 8;
    it.lsz[1] = it.lsz[2] = 1;
    it.gid[0] = x;
    it.lid[0] = x % 8;
    it.gid[1] = it.gid[2] = it.lid[1] = it.lid[2] = 0;
    f();
  }
}

}  // namespace

extern "C" {

void ref_set_threads(int n) { g_threads = n; }
int ref_get_threads(void) { return g_threads > 0 ? g_threads : omp_get_max_threads(); }

// SATEncoder::EncodeFrameCPU, sat_encoder.cc:137-185.
void ref_sat_encode_cpu(uint32_t *sat, const uint8_t *src, int W, int H, int src_linesize) {
  ref_cpu::AVCodecContext ctx{W, H};
  ref_cpu::AVFrame frame = {};
  frame.data[0] = const_cast<uint8_t *>(src);
  frame.linesize[0] = src_linesize;
  ref_cpu::SATEncoder().EncodeFrameCPU(sat, &ctx, &frame);
}

// SATEncoder::EncodeFrameGPU, sat_encoder.cc:67-135 (three in-order launches).
void ref_sat_encode(uint32_t *sat, const uint8_t *src, int W, int H, int src_linesize) {
  const int target_linesize = 3 * W;  // sat_encoder.cc:77, in u32 elements
  ndrange2(W, H, [&] {
    ref_sat_enc::copy_image_kernel(sat, target_linesize, const_cast<uint8_t *>(src), W, H,
                                   src_linesize);
  });
  ndrange1(H, [&] { ref_sat_enc::scan_rows_kernel(sat, W, H, target_linesize); });
  ndrange1(W, [&] { ref_sat_enc::scan_columns_kernel(sat, W, H, target_linesize); });
}

// SATDecoder::InitializeGrid, sat_decoder.cc:139-170.  grid: int16[(oh+1)][(ow+1)][2].
void ref_sat_create_grid(int16_t *grid, int ow, int oh, int W, int H) {
  ndrange2(roundup8(ow + 1), roundup8(oh + 1),
           [&] { ref_sat_smp::create_grid_kernel(grid, ow, oh, W, H); });
}

// SATDecoder::SampleFrameRectGPU, sat_decoder.cc:301-348.
void ref_sat_sample_rect(uint8_t *out, int ow, int oh, int out_linesize, const uint32_t *sat,
                         int W, int H, const int16_t *grid, float cx, float cy) {
  float2 center = {cx, cy};
  ndrange2(roundup8(ow), roundup8(oh), [&] {
    ref_sat_smp::sample_rect_kernel(reinterpret_cast<uchar4 *>(out), ow, oh, out_linesize,
                                    const_cast<uint32_t *>(sat), W, H,
                                    const_cast<int16_t *>(grid), center);
  });
}

// SATDecoder::InterpolateFrameRectGPU, sat_decoder.cc:887-927 (linesizes unused).
void ref_sat_interpolate_rect(uint8_t *out, int W, int H, const uint8_t *reduced, int ow, int oh,
                              float cx, float cy) {
  float2 center = {cx, cy};
  ndrange2(roundup8(W), roundup8(H), [&] {
    ref_sat_itp::interpolate_rect_kernel(
        reinterpret_cast<uchar3 *>(out), W, H,
        reinterpret_cast<uchar3 *>(const_cast<uint8_t *>(reduced)), ow, oh, center);
  });
}

// SATDecoder::DecodeFrameGPU, sat_decoder.cc:176-210 (evident intent: a 2-D
// (W, H) NDRange; the reference passes work_dim = 0 and therefore never runs).
void ref_sat_decode(uint8_t *out, int out_linesize, const uint32_t *sat, int W, int H) {
  const int source_linesize = 3 * W;
  ndrange2(W, H, [&] {
    ref_sat_dec::decode_kernel(out, out_linesize, const_cast<uint32_t *>(sat), W, H,
                               source_linesize);
  });
}

// ImageSampler::InitializeGrid, image_sampler.cc:170-202.  grid: int16[oh][ow][2].
// The kernel's row guard tests `j >= output_width` (image_sampler_sample_rect_kernel.cl:64),
// so rows up to roundup8(oh)-1 may be written; run into a padded scratch and copy.
void ref_img_create_grid(int16_t *grid, int ow, int oh, int W, int H) {
  std::vector<int16_t> scratch((size_t)roundup8(oh) * ow * 2 + 16, 0);
  ndrange2(roundup8(ow), roundup8(oh),
           [&] { ref_img_smp::create_grid_kernel(scratch.data(), ow, oh, W, H); });
  std::memcpy(grid, scratch.data(), (size_t)oh * ow * 2 * sizeof(int16_t));
}

// ImageSampler::SampleFrameRectGPU, image_sampler.cc:249-299 (global = ow x oh exactly).
void ref_img_sample_rect(uint8_t *out, int ow, int oh, int out_linesize, const uint8_t *src, int W,
                         int H, int src_linesize, const int16_t *grid, float cx, float cy) {
  ndrange2(ow, oh, [&] {
    ref_img_smp::sample_rect_kernel(out, ow, oh, out_linesize, const_cast<uint8_t *>(src), W, H,
                                    src_linesize, const_cast<int16_t *>(grid), cx, cy);
  });
}

// ImageSampler::InitializeLogpolarGrid, image_sampler.cc:204-247.  grid: int16[oh][ow][2].
void ref_img_create_logpolar_grid(int16_t *grid, int ow, int oh, int W, int H) {
  ndrange2(roundup8(ow), roundup8(oh),
           [&] { ref_img_lp::create_logpolar_grid_kernel(grid, ow, oh, W, H); });
}

// ImageSampler::SampleFrameLogPolarGPU, image_sampler.cc:577-621.
void ref_img_sample_logpolar(uint8_t *out, int ow, int oh, int out_linesize, const uint8_t *src,
                             int W, int H, int src_linesize, const int16_t *grid, float cx,
                             float cy) {
  ndrange2(roundup8(ow), roundup8(oh), [&] {
    ref_img_lp::sample_logpolar_kernel(out, ow, oh, out_linesize, const_cast<uint8_t *>(src), W,
                                       H, src_linesize, const_cast<int16_t *>(grid), cx, cy);
  });
}

// ImageSampler::InterpolateFrameLogPolarGPU, image_sampler.cc:780-818 (global = W x H exactly).
void ref_img_interpolate_logpolar(uint8_t *out, int W, int H, const uint8_t *reduced, int ow,
                                  int oh, float cx, float cy) {
  float2 center = {cx, cy};
  ndrange2(W, H, [&] {
    ref_img_itp::interpolate_logpolar_kernel(
        reinterpret_cast<uchar3 *>(out), W, H,
        reinterpret_cast<uchar3 *>(const_cast<uint8_t *>(reduced)), ow, oh, center);
  });
}

// ImageSampler::ApplyLogPolarGaussianBlur, image_sampler.cc:820-857 (global = ow x oh exactly).
void ref_img_logpolar_blur(uint8_t *out, int ow, int oh, int linesize, const uint8_t *src) {
  ndrange2(ow, oh, [&] {
    ref_img_lp::logpolar_gaussian_blur_kernel(reinterpret_cast<uchar3 *>(out), ow, oh, linesize,
                                              reinterpret_cast<uchar3 *>(const_cast<uint8_t *>(src)));
  });
}

// Projections::GnomonicProjection, projections.cc:51-86 (8x8 work-groups over the viewport;
// the linesize arguments never reach the kernel).
void ref_gnomonic(uint8_t *out, int tw, int th, const uint8_t *src, int W, int H, float cx,
                  float cy) {
  float2 center = {cx, cy};
  ndrange2(roundup8(tw), roundup8(th), [&] {
    ref_proj::gnomonic_kernel(reinterpret_cast<uchar3 *>(out), tw, th,
                              reinterpret_cast<uchar3 *>(const_cast<uint8_t *>(src)), W, H,
                              center);
  });
}
}  // extern "C"
