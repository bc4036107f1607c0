// Drop-in for the reference's src/projections.h (projections.h:20-35): Projections renders a
// viewport out of an equirectangular frame (inverse gnomonic projection).  InterpolateGnomonicGPU
// has no reference counterpart: it is SATDecoder::InterpolateFrameRectGPU followed by
// GnomonicProjection in one kernel, for clients that only display a viewport.
#pragma once
#include <iostream>

#include "opencl_manager.h"

class Projections {
 public:
  Projections() = default;
  explicit Projections(OpenCLManager *cl_manager) : cl_manager_(cl_manager) {}

  // projections.cc:51-86.  The definition's parameter order is (width, height); the linesizes never
  // reach the kernel.
  void GnomonicProjection(cl_mem cl_target_buffer, int target_width, int target_height,
                          int target_linesize, cl_mem cl_source_buffer, int source_width,
                          int source_height, int source_linesize, float center_x, float center_y) {
    if (!ready()) return;
    report(fov_gnomonic(ctx(), static_cast<uint8_t *>(cl_target_buffer), target_width,
                        target_height, target_linesize,
                        static_cast<const uint8_t *>(cl_source_buffer), source_width,
                        source_height, source_linesize, center_x, center_y),
           "GnomonicProjection");
  }

  void InterpolateGnomonicGPU(cl_mem cl_target_buffer, int target_width, int target_height,
                              cl_mem cl_reduced_buffer, int reduced_width, int reduced_height,
                              int full_width, int full_height, float gaze_x, float gaze_y,
                              float view_x, float view_y) {
    if (!ready()) return;
    report(fov_sat_interpolate_gnomonic(ctx(), static_cast<uint8_t *>(cl_target_buffer),
                                        target_width, target_height,
                                        static_cast<const uint8_t *>(cl_reduced_buffer),
                                        reduced_width, reduced_height, full_width, full_height,
                                        gaze_x, gaze_y, view_x, view_y),
           "InterpolateGnomonicGPU");
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
