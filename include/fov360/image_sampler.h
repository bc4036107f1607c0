// Drop-in for the reference's src/image_sampler.h (image_sampler.h:29-102): the GPU methods of
// ImageSampler (log-rect point sampling, log-polar sampling / inverse warp / blur).  The image
// pyramid methods are absent: their kernel source is missing from the reference.
#pragma once
#include <iostream>

#include "opencl_manager.h"

class ImageSampler {
 public:
  ImageSampler() = default;
  explicit ImageSampler(OpenCLManager *cl_manager) : cl_manager_(cl_manager) {}

  void InitializeGrid(int target_width, int target_height, int source_width, int source_height) {
    if (!ready()) return;
    report(fov_img_grid_init(ctx(), target_width, target_height, source_width, source_height),
           "InitializeGrid");
  }
  void InitializeLogpolarGrid(int target_width, int target_height, int source_width,
                              int source_height) {
    if (!ready()) return;
    report(fov_img_logpolar_grid_init(ctx(), target_width, target_height, source_width,
                                      source_height),
           "InitializeLogpolarGrid");
  }
  // image_sampler.cc:249-299
  void SampleFrameRectGPU(cl_mem cl_target_buffer, int target_width, int target_height,
                          int target_linesize, cl_mem cl_source_buffer, int source_width,
                          int source_height, int source_linesize, float center_x, float center_y) {
    if (!ready()) return;
    report(fov_img_sample_rect(ctx(), static_cast<uint8_t *>(cl_target_buffer), target_width,
                               target_height, target_linesize,
                               static_cast<const uint8_t *>(cl_source_buffer), source_width,
                               source_height, source_linesize, center_x, center_y),
           "SampleFrameRectGPU");
  }
  // image_sampler.cc:577-621
  void SampleFrameLogPolarGPU(cl_mem cl_target_buffer, int target_width, int target_height,
                              int target_linesize, cl_mem cl_source_buffer, int source_width,
                              int source_height, int source_linesize, float center_x,
                              float center_y) {
    if (!ready()) return;
    report(fov_img_sample_logpolar(ctx(), static_cast<uint8_t *>(cl_target_buffer), target_width,
                                   target_height, target_linesize,
                                   static_cast<const uint8_t *>(cl_source_buffer), source_width,
                                   source_height, source_linesize, center_x, center_y),
           "SampleFrameLogPolarGPU");
  }
  // image_sampler.cc:780-818
  void InterpolateFrameLogPolarGPU(cl_mem cl_target_buffer, int target_width, int target_height,
                                   int target_linesize, cl_mem cl_source_buffer, int source_width,
                                   int source_height, int source_linesize, float center_x,
                                   float center_y) {
    if (!ready()) return;
    report(fov_img_interpolate_logpolar(ctx(), static_cast<uint8_t *>(cl_target_buffer),
                                        target_width, target_height, target_linesize,
                                        static_cast<const uint8_t *>(cl_source_buffer),
                                        source_width, source_height, source_linesize, center_x,
                                        center_y),
           "InterpolateFrameLogPolarGPU");
  }
  // image_sampler.cc:820-857
  void ApplyLogPolarGaussianBlur(cl_mem cl_target_buffer, int target_width, int target_height,
                                 int target_linesize, cl_mem cl_source_buffer) {
    if (!ready()) return;
    report(fov_img_logpolar_blur(ctx(), static_cast<uint8_t *>(cl_target_buffer), target_width,
                                 target_height, target_linesize,
                                 static_cast<const uint8_t *>(cl_source_buffer)),
           "ApplyLogPolarGaussianBlur");
  }

 private:
  fov_ctx *ctx() const { return cl_manager_->handle(); }
  bool ready() const {
    if (cl_manager_ && cl_manager_->handle()) return true;
    std::cerr << "Not initialized with OpenCL" << std::endl;
    return false;
  }
  void report(int rc, const char *what) const {
    if (rc != FOV_OK) std::cerr << what << " failed: " << fov_last_error_string(ctx()) << std::endl;
  }
  OpenCLManager *cl_manager_ = nullptr;
};
