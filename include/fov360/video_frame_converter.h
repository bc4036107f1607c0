// Device-side replacement for the colour conversion inside the reference's
// VideoEncoder::EncodeFrame (src/video_encoder.cc:380-398): sws_getContext(RGB0 -> YUV420P,
// SWS_BILINEAR) + sws_scale on the host, then av_hwframe_transfer_data.  The reduced buffer is
// already on the device, so the planes are written straight into the AV_PIX_FMT_CUDA frame's
// data[0..2] (sw_format YUV420P, video_encoder.cc:549) - or into an NV12 surface - with
// libswscale's exact C arithmetic.  A maintainer replaces video_encoder.cc:380-398 by one call:
//
//   converter.RGB0ToYUV420P(hw_frame->data, hw_frame->linesize, cl_out_buffer(), 4 * width,
//                           width, height);
//
// The other direction replaces the sws_scale of VideoDecoder::GetFrame (src/video_decoder.cc:165-170,
// :222): a frame decoded on the device (YUV420P planes or an NVDEC NV12 surface) becomes the RGB0
// frame EncodeFrameGPU reads, without the host round trip of video_server.cc:291-299:
//
//   converter.NV12ToRGB0(cl_source_buffer(), 4 * width, nv12_data, nv12_linesize, width, height);
//
// Errors follow the reference's convention: print to std::cerr and return.
#pragma once
#include <iostream>

#include "opencl_manager.h"

class VideoFrameConverter {
 public:
  VideoFrameConverter() = default;
  explicit VideoFrameConverter(OpenCLManager *cl_manager) : cl_manager_(cl_manager) {}

  // data / linesize: AVFrame::data[0..2] and AVFrame::linesize[0..2] of a device frame.
  void RGB0ToYUV420P(uint8_t *const data[3], const int linesize[3], cl_mem cl_source_buffer,
                     int source_linesize, int width, int height) {
    if (!Ready()) return;
    Report("RGB0ToYUV420P",
           fov_rgb0_to_yuv420p(cl_manager_->handle(), data[0], linesize[0], data[1], linesize[1],
                               data[2], linesize[2],
                               static_cast<const uint8_t *>(cl_source_buffer), source_linesize,
                               width, height));
  }

  // data / linesize: the Y plane and the interleaved UV plane of an NV12 surface.
  void RGB0ToNV12(uint8_t *const data[2], const int linesize[2], cl_mem cl_source_buffer,
                  int source_linesize, int width, int height) {
    if (!Ready()) return;
    Report("RGB0ToNV12",
           fov_rgb0_to_nv12(cl_manager_->handle(), data[0], linesize[0], data[1], linesize[1],
                            static_cast<const uint8_t *>(cl_source_buffer), source_linesize, width,
                            height));
  }

  // data / linesize: the three planes of a decoded YUV420P frame on the device.
  void YUV420PToRGB0(cl_mem cl_target_buffer, int target_linesize, const uint8_t *const data[3],
                     const int linesize[3], int width, int height) {
    if (!Ready()) return;
    Report("YUV420PToRGB0",
           fov_yuv420p_to_rgb0(cl_manager_->handle(), static_cast<uint8_t *>(cl_target_buffer),
                               target_linesize, data[0], linesize[0], data[1], linesize[1], data[2],
                               linesize[2], width, height));
  }

  // data / linesize: the Y plane and the interleaved UV plane of an NV12 surface (NVDEC output).
  void NV12ToRGB0(cl_mem cl_target_buffer, int target_linesize, const uint8_t *const data[2],
                  const int linesize[2], int width, int height) {
    if (!Ready()) return;
    Report("NV12ToRGB0",
           fov_nv12_to_rgb0(cl_manager_->handle(), static_cast<uint8_t *>(cl_target_buffer),
                            target_linesize, data[0], linesize[0], data[1], linesize[1], width,
                            height));
  }

 private:
  bool Ready() const {
    if (cl_manager_ && cl_manager_->handle()) return true;
    std::cerr << "Not initialized with OpenCL" << std::endl;
    return false;
  }
  void Report(const char *what, int rc) const {
    if (rc != FOV_OK)
      std::cerr << what << " failed: " << fov_last_error_string(cl_manager_->handle()) << std::endl;
  }
  OpenCLManager *cl_manager_ = nullptr;
};
