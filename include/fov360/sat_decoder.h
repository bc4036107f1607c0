// Drop-in for the reference's src/sat_decoder.h (sat_decoder.h:20-83): the GPU methods of
// SATDecoder with the reference's signatures.  `codec_ctx` is only read for ->width / ->height
// (sat_decoder.cc:328-329), so SampleFrameRectGPU accepts any pointer type with those members -
// FFmpeg's AVCodecContext at the real call sites, a plain struct in tests.  The experimental
// reduced-SAT / 360 samplers and the CPU twins are not part of this library.
#pragma once
#include <iostream>

#include "opencl_manager.h"

class SATDecoder {
 public:
  SATDecoder() = default;
  explicit SATDecoder(OpenCLManager *cl_manager) : cl_manager_(cl_manager) {}

  // sat_decoder.cc:139-170
  void InitializeGrid(int target_width, int target_height, int source_width, int source_height) {
    if (!ready()) return;
    report(fov_sat_grid_init(ctx(), target_width, target_height, source_width, source_height),
           "InitializeGrid");
  }

  // sat_decoder.cc:176-210
  void DecodeFrameGPU(cl_mem cl_target_buffer, int target_linesize, cl_mem cl_source_buffer,
                      int width, int height) {
    if (!ready()) return;
    report(fov_sat_decode(ctx(), static_cast<uint8_t *>(cl_target_buffer), target_linesize,
                          static_cast<const uint32_t *>(cl_source_buffer), width, height),
           "DecodeFrameGPU");
  }

  // sat_decoder.cc:301-348
  template <class CodecContext>
  void SampleFrameRectGPU(cl_mem cl_target_buffer, int target_width, int target_height,
                          int target_linesize, cl_mem cl_source_buffer, CodecContext *codec_ctx,
                          float center_x, float center_y) {
    if (!ready()) return;
    report(fov_sat_sample_rect(ctx(), static_cast<uint8_t *>(cl_target_buffer), target_width,
                               target_height, target_linesize,
                               static_cast<const uint32_t *>(cl_source_buffer), codec_ctx->width,
                               codec_ctx->height, center_x, center_y),
           "SampleFrameRectGPU");
  }

  // sat_decoder.cc:887-927 (both linesize arguments are unused there as well)
  void InterpolateFrameRectGPU(cl_mem cl_target_buffer, int target_width, int target_height,
                               int target_linesize, cl_mem cl_source_buffer, int source_width,
                               int source_height, int source_linesize, float center_x,
                               float center_y) {
    if (!ready()) return;
    report(fov_sat_interpolate_rect(ctx(), static_cast<uint8_t *>(cl_target_buffer), target_width,
                                    target_height, target_linesize,
                                    static_cast<const uint8_t *>(cl_source_buffer), source_width,
                                    source_height, source_linesize, center_x, center_y),
           "InterpolateFrameRectGPU");
  }

 private:
  fov_ctx *ctx() const { return cl_manager_->handle(); }
  bool ready() const {
    if (cl_manager_ && cl_manager_->handle()) return true;
    std::cerr << "Not initialized with OpenCL" << std::endl;  // sat_decoder.cc:179-183
    return false;
  }
  void report(int rc, const char *what) const {
    if (rc != FOV_OK) std::cerr << what << " failed: " << fov_last_error_string(ctx()) << std::endl;
  }
  OpenCLManager *cl_manager_ = nullptr;
};
