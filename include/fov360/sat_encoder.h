// Drop-in for the reference's src/sat_encoder.h (sat_encoder.h:21-43): same class and method
// signature for the GPU path.  EncodeFrameCPU is intentionally absent: this library has no CPU
// path.  Errors follow the reference's convention: print to std::cerr and return.
#pragma once
#include <iostream>

#include "opencl_manager.h"

class SATEncoder {
 public:
  SATEncoder() = default;  // disabled object, like the reference's default constructor
  explicit SATEncoder(OpenCLManager *cl_manager) : cl_manager_(cl_manager) {}

  // sat_encoder.cc:67-135.  cl_target_buffer: u32[H][W][3]; cl_source_buffer: RGB0/RGB24 frame.
  void EncodeFrameGPU(cl_mem cl_target_buffer, cl_mem cl_source_buffer, int source_width,
                      int source_height, int source_linesize) {
    if (!cl_manager_ || !cl_manager_->handle()) {
      std::cerr << "Not initialized with OpenCL" << std::endl;  // sat_encoder.cc:70-74
      return;
    }
    const int rc = fov_sat_encode(cl_manager_->handle(), static_cast<uint32_t *>(cl_target_buffer),
                                  static_cast<const uint8_t *>(cl_source_buffer), source_width,
                                  source_height, source_linesize);
    if (rc != FOV_OK)
      std::cerr << "EncodeFrameGPU failed: " << fov_last_error_string(cl_manager_->handle())
                << std::endl;
  }

 private:
  OpenCLManager *cl_manager_ = nullptr;
};
