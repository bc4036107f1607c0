// The reference's call sequences, written against include/fov360/*.h exactly as the reference
// writes them against its own headers:
//   server loop   video_server.cc:224-232, 296-345   (EncodeFrameGPU -> SampleFrameRectGPU)
//   client loop   video_client.cc:303-319            (InterpolateFrameRectGPU)
//   offline loop  run_satlogrectilinear.cc:915-949   (encode -> sample -> interpolate)
// Prints FNV-1a-64 hashes of every buffer as one JSON object; tests/test_cpp_dropin.py compares
// them with tests/golden/golden.json (generated from the reference's own kernels).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fov360/image_sampler.h"
#include "fov360/parameters.h"
#include "fov360/projections.h"
#include "fov360/sat_decoder.h"
#include "fov360/sat_encoder.h"
#include "fov360/video_frame_converter.h"

struct CodecCtxStub {  // stands in for AVCodecContext: only width/height are read
  int width, height;
};

static uint64_t fnv1a64(const void *p, size_t n) {
  const uint8_t *b = static_cast<const uint8_t *>(p);
  uint64_t h = 0xcbf29ce484222325ull;
  for (size_t i = 0; i < n; ++i) {
    h ^= b[i];
    h *= 0x100000001b3ull;
  }
  return h;
}

int main(int argc, char **argv) {
  const int W = argc > 1 ? atoi(argv[1]) : 1920, H = argc > 2 ? atoi(argv[2]) : 1080;
  const uint32_t seed = argc > 3 ? (uint32_t)atoi(argv[3]) : 12345u;
  const float cx = argc > 4 ? (float)atof(argv[4]) : 0.65f, cy = argc > 5 ? (float)atof(argv[5]) : 0.75f;
  const int ow = ReducedBufferDim(W), oh = ReducedBufferDim(H);
  const int linesize = 4 * W, out_linesize = 4 * ow;

  // synthetic RGB0 frame (SURVEY 8(c) generator)
  std::vector<uint8_t> frame((size_t)linesize * H);
  uint32_t s = seed;
  for (size_t i = 0; i < frame.size(); ++i) {
    s = s * 1664525u + 1013904223u;
    frame[i] = ((i & 3) == 3) ? 0 : (uint8_t)(s >> 24);
  }

  OpenCLManager cl_manager;
  cl_manager.InitializeContext();
  SATEncoder sat_encoder(&cl_manager);
  SATDecoder sat_decoder(&cl_manager);
  ImageSampler image_sampler(&cl_manager);
  Projections projections(&cl_manager);
  VideoFrameConverter converter(&cl_manager);
  CodecCtxStub codec_ctx{W, H};

  cl::Buffer cl_source_buffer(cl_manager.context, CL_MEM_READ_WRITE, (size_t)linesize * H);
  cl::Buffer cl_sat_buffer(cl_manager.context, CL_MEM_READ_WRITE, (size_t)12 * W * H);
  cl::Buffer cl_out_buffer(cl_manager.context, CL_MEM_READ_WRITE, (size_t)out_linesize * oh);
  cl::Buffer cl_full_buffer(cl_manager.context, CL_MEM_READ_WRITE, (size_t)linesize * H);
  cl::Buffer cl_lp_buffer(cl_manager.context, CL_MEM_READ_WRITE, (size_t)out_linesize * oh);
  const int vw = 960, vh = 540;  // viewport of the gnomonic golden vectors
  cl::Buffer cl_view_buffer(cl_manager.context, CL_MEM_READ_WRITE, (size_t)4 * vw * vh);
  cl::Buffer cl_view2_buffer(cl_manager.context, CL_MEM_READ_WRITE, (size_t)4 * vw * vh);

  std::vector<uint8_t> reduced((size_t)out_linesize * oh, 0), full((size_t)linesize * H, 0),
      logpolar((size_t)out_linesize * oh, 0);
  std::vector<uint32_t> sat((size_t)3 * W * H);

  cl::copy(cl_manager.command_queue, frame.begin(), frame.end(), cl_source_buffer);
  cl::copy(cl_manager.command_queue, reduced.begin(), reduced.end(), cl_out_buffer);  // zero-fill
  cl::copy(cl_manager.command_queue, logpolar.begin(), logpolar.end(), cl_lp_buffer);

  sat_encoder.EncodeFrameGPU(cl_sat_buffer(), cl_source_buffer(), W, H, linesize);
  clFlush(cl_manager.command_queue());
  clFinish(cl_manager.command_queue());
  sat_decoder.SampleFrameRectGPU(cl_out_buffer(), ow, oh, out_linesize, cl_sat_buffer(), &codec_ctx,
                                 cx, cy);
  sat_decoder.InterpolateFrameRectGPU(cl_full_buffer(), W, H, linesize, cl_out_buffer(), ow, oh,
                                      out_linesize, cx, cy);
  image_sampler.SampleFrameLogPolarGPU(cl_lp_buffer(), ow, oh, out_linesize, cl_source_buffer(), W,
                                       H, linesize, cx, cy);

  // server side, video_encoder.cc:380-398: the reduced buffer becomes the encoder's YUV420P
  // surface on the device (planes of one allocation, like a hardware frame)
  cl::Buffer cl_yuv_buffer(cl_manager.context, CL_MEM_READ_WRITE, (size_t)ow * oh * 3 / 2);
  uint8_t *yuv_base = static_cast<uint8_t *>(cl_yuv_buffer());
  uint8_t *const yuv_data[3] = {yuv_base, yuv_base + (size_t)ow * oh,
                                yuv_base + (size_t)ow * oh * 5 / 4};
  const int yuv_linesize[3] = {ow, ow / 2, ow / 2};
  converter.RGB0ToYUV420P(yuv_data, yuv_linesize, cl_out_buffer(), out_linesize, ow, oh);
  std::vector<uint8_t> yuv((size_t)ow * oh * 3 / 2);
  cl::copy(cl_manager.command_queue, cl_yuv_buffer, yuv.begin(), yuv.end());

  // client side, projections.cc:51-86: viewport out of the un-warped frame, and the fused form
  projections.GnomonicProjection(cl_view_buffer(), vw, vh, 4 * vw, cl_full_buffer(), W, H, linesize,
                                 cx, cy);
  projections.InterpolateGnomonicGPU(cl_view2_buffer(), vw, vh, cl_out_buffer(), ow, oh, W, H, cx, cy,
                                     cx, cy);
  std::vector<uint8_t> view((size_t)4 * vw * vh), view2((size_t)4 * vw * vh);
  cl::copy(cl_manager.command_queue, cl_view_buffer, view.begin(), view.end());
  cl::copy(cl_manager.command_queue, cl_view2_buffer, view2.begin(), view2.end());

  cl::copy(cl_manager.command_queue, cl_sat_buffer, sat.begin(), sat.end());
  cl::copy(cl_manager.command_queue, cl_out_buffer, reduced.begin(), reduced.end());
  cl::copy(cl_manager.command_queue, cl_full_buffer, full.begin(), full.end());
  cl::copy(cl_manager.command_queue, cl_lp_buffer, logpolar.begin(), logpolar.end());

  printf("{\"W\": %d, \"H\": %d, \"ow\": %d, \"oh\": %d, \"sat\": \"%016llx\", "
         "\"reduced_zero\": \"%016llx\", \"interp\": \"%016llx\", \"logpolar\": \"%016llx\", "
         "\"view_equal\": %d, \"yuv420p\": \"%016llx\", \"launches\": %llu}\n",
         W, H, ow, oh, (unsigned long long)fnv1a64(sat.data(), sat.size() * 4),
         (unsigned long long)fnv1a64(reduced.data(), reduced.size()),
         (unsigned long long)fnv1a64(full.data(), full.size()),
         (unsigned long long)fnv1a64(logpolar.data(), logpolar.size()), (int)(view == view2),
         (unsigned long long)fnv1a64(yuv.data(), yuv.size()),
         (unsigned long long)fov_ctx_launch_count(cl_manager.handle()));
  return 0;
}
