// Stand-ins for the reference's NON-foveation collaborators (FFmpeg structs, VideoDecoder,
// VideoEncoder, the per-connection state of VideoServer) so that the reference's own call-site
// text - extracted at build time from $FOV_REF_DIR by oracle/build_callsite_check.py, never copied
// into this repository - compiles against include/fov360/*.h and runs on synthetic frames.
// TEST INFRASTRUCTURE ONLY.
//
//  * VideoDecoder::GetFrame fills the SURVEY 8(c) LCG frame (seed + frame index) into an RGB0
//    AVFrame, `frames` times, then reports end of stream.
//  * VideoEncoder::EncodeFrameToFile records the FNV-1a-64 hash of the frame it is handed.
// "source video" strings have the form  WIDTHxHEIGHT:SEED:FRAMES.
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#define AV_PIX_FMT_RGB0 295
#define AV_NUM_DATA_POINTERS 8

struct AVCodecContext {
  int width = 0, height = 0;
  int pix_fmt = AV_PIX_FMT_RGB0;
};

struct AVFrame {
  uint8_t *data[AV_NUM_DATA_POINTERS] = {};
  int linesize[AV_NUM_DATA_POINTERS] = {};
  int width = 0, height = 0, format = -1;
  int64_t pts = 0, pkt_dts = 0;
};

inline AVFrame *av_frame_alloc() { return new AVFrame(); }
inline int av_frame_get_buffer(AVFrame *f, int /*align*/) {  // packed RGB0, linesize = 4 * width
  f->linesize[0] = 4 * f->width;
  f->data[0] = static_cast<uint8_t *>(calloc((size_t)f->linesize[0] * f->height, 1));
  return f->data[0] ? 0 : -1;
}
inline void av_frame_free(AVFrame **f) {
  if (f && *f) {
    free((*f)->data[0]);
    delete *f;
    *f = nullptr;
  }
}
struct AVFrameDeleter {
  void operator()(AVFrame *f) const { av_frame_free(&f); }
};

inline uint64_t callsite_fnv1a64(const void *p, size_t n) {
  const uint8_t *b = static_cast<const uint8_t *>(p);
  uint64_t h = 0xcbf29ce484222325ull;
  for (size_t i = 0; i < n; ++i) {
    h ^= b[i];
    h *= 0x100000001b3ull;
  }
  return h;
}

class VideoDecoder {
 public:
  AVCodecContext *source_codec_ctx = &ctx_;
  int OpenVideo(const std::string &spec) {
    unsigned w = 0, h = 0, seed = 0, n = 0;
    if (sscanf(spec.c_str(), "%ux%u:%u:%u", &w, &h, &seed, &n) != 4) return -1;
    ctx_.width = (int)w, ctx_.height = (int)h, seed_ = seed, frames_ = (int)n;
    return 0;
  }
  // 0 on success, negative at end of stream (video_decoder.cc:165-222 hands out RGB0 frames)
  int GetFrame(AVFrame *frame, int /*pix_fmt*/) {
    if (next_ >= frames_) return -1;
    if (!frame->data[0]) {
      frame->width = ctx_.width, frame->height = ctx_.height, frame->format = AV_PIX_FMT_RGB0;
      if (av_frame_get_buffer(frame, 1) != 0) return -1;
    }
    uint32_t s = seed_ + (uint32_t)next_;
    const size_t n = (size_t)frame->linesize[0] * frame->height;
    for (size_t i = 0; i < n; ++i) {
      s = s * 1664525u + 1013904223u;
      frame->data[0][i] = ((i & 3) == 3) ? 0 : (uint8_t)(s >> 24);
    }
    frame->pts = frame->pkt_dts = next_++;
    return 0;
  }

 private:
  AVCodecContext ctx_;
  uint32_t seed_ = 12345;
  int frames_ = 0, next_ = 0;
};

inline std::vector<uint64_t> &callsite_encoded_hashes() {
  static std::vector<uint64_t> v;
  return v;
}

class VideoEncoder {
 public:
  VideoEncoder(AVCodecContext *, void *, const std::string &) {}
  int EncodeFrameToFile(AVFrame *frame) {
    if (frame)
      callsite_encoded_hashes().push_back(
          callsite_fnv1a64(frame->data[0], (size_t)frame->linesize[0] * frame->height));
    return 0;
  }
  void WriteTrailerAndCloseFile() {}
};

// What the server loop reads of its per-connection state (video_server.h: ConnectionData).
struct CallsiteConnData {
  std::mutex wait_mutex, center_xy_mutex;
  double center_x = 0.5, center_y = 0.5;
  bool exit_thread = false;
};
