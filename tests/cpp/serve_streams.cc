// A server's worth of connections in one process, written against include/fov360/*.h the way the
// reference writes its per-connection thread (video_server.cc:62-66, 85, 224-232, 287-345), with
// the steps either side of the foveation path on the device as well:
//
//   per connection thread:  placement.Acquire() -> OpenCLManager -> buffers
//   per frame:              NV12 surface -> RGB0            (video_decoder.cc:165-222)
//                           EncodeFrameGPU                  (video_server.cc:300)
//                           SampleFrameRectGPU(gaze[frame]) (video_server.cc:336-339)
//                           reduced RGB0 -> NV12 surface    (video_encoder.cc:380-398)
//
// Every connection replays the gaze trace from its own offset.  Prints one JSON object with the
// FNV-1a-64 hash of every connection's last NV12 surface and the aggregate frame rate;
// tests/test_cpp_serve_streams.py recomputes the hashes with the oracle.  Thread-compatibility is
// the contract under test: one context per thread, nothing shared but the library.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "fov360/gaze_view_points.h"
#include "fov360/parameters.h"
#include "fov360/sat_decoder.h"
#include "fov360/sat_encoder.h"
#include "fov360/session_placement.h"
#include "fov360/video_frame_converter.h"

struct CodecCtxStub {
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

struct Result {
  int device = -1;
  unsigned last_record = 0;
  uint64_t hash = 0;
};

static void Connection(int s, int frames, int W, int H, const GazeViewPoints &trace,
                       SessionPlacement *placement, Result *out) {
  const int ow = ReducedBufferDim(W), oh = ReducedBufferDim(H);
  OpenCLManager cl_manager;
  cl_manager.device_index = placement->Acquire();
  cl_manager.InitializeContext();
  SATEncoder sat_encoder(&cl_manager);
  SATDecoder sat_decoder(&cl_manager);
  VideoFrameConverter converter(&cl_manager);
  CodecCtxStub codec_ctx{W, H};

  // the connection's "decoded" NV12 surface: LCG bytes seeded by the connection id
  std::vector<uint8_t> nv12((size_t)W * H * 3 / 2);
  uint32_t lcg = 1000u + (uint32_t)s;
  for (auto &b : nv12) {
    lcg = lcg * 1664525u + 1013904223u;
    b = (uint8_t)(lcg >> 24);
  }
  cl::Buffer cl_nv12(cl_manager.context, CL_MEM_READ_WRITE, nv12.size());
  cl::Buffer cl_source_buffer(cl_manager.context, CL_MEM_READ_WRITE, (size_t)4 * W * H);
  cl::Buffer cl_sat_buffer(cl_manager.context, CL_MEM_READ_WRITE, (size_t)12 * W * H);
  cl::Buffer cl_out_buffer(cl_manager.context, CL_MEM_READ_WRITE, (size_t)4 * ow * oh);
  cl::Buffer cl_enc_surface(cl_manager.context, CL_MEM_READ_WRITE, (size_t)ow * oh * 3 / 2);
  cl::copy(cl_manager.command_queue, nv12.begin(), nv12.end(), cl_nv12);
  std::vector<uint8_t> zero((size_t)4 * ow * oh, 0);
  cl::copy(cl_manager.command_queue, zero.begin(), zero.end(), cl_out_buffer);

  uint8_t *in_base = static_cast<uint8_t *>(cl_nv12());
  const uint8_t *const in_data[2] = {in_base, in_base + (size_t)W * H};
  const int in_linesize[2] = {W, W};
  uint8_t *out_base = static_cast<uint8_t *>(cl_enc_surface());
  uint8_t *const out_data[2] = {out_base, out_base + (size_t)ow * oh};
  const int out_linesize[2] = {ow, ow};

  unsigned rec = 0;
  for (int f = 0; f < frames; ++f) {
    rec = (unsigned)((f + 7 * s) % trace.points.size());
    const float cx = trace.points[rec].gaze_point[0], cy = trace.points[rec].gaze_point[1];
    converter.NV12ToRGB0(cl_source_buffer(), 4 * W, in_data, in_linesize, W, H);
    sat_encoder.EncodeFrameGPU(cl_sat_buffer(), cl_source_buffer(), W, H, 4 * W);
    sat_decoder.SampleFrameRectGPU(cl_out_buffer(), ow, oh, 4 * ow, cl_sat_buffer(), &codec_ctx, cx, cy);
    converter.RGB0ToNV12(out_data, out_linesize, cl_out_buffer(), 4 * ow, ow, oh);
  }
  std::vector<uint8_t> surface((size_t)ow * oh * 3 / 2);
  cl::copy(cl_manager.command_queue, cl_enc_surface, surface.begin(), surface.end());
  out->device = cl_manager.device_index;
  out->last_record = rec;
  out->hash = fnv1a64(surface.data(), surface.size());
  placement->Release(cl_manager.device_index);
}

int main(int argc, char **argv) {
  if (argc < 6) {
    fprintf(stderr, "usage: %s trace.txt streams frames W H\n", argv[0]);
    return 2;
  }
  const GazeViewPoints trace{std::string(argv[1])};
  const int streams = atoi(argv[2]), frames = atoi(argv[3]), W = atoi(argv[4]), H = atoi(argv[5]);
  if (trace.points.empty() || streams < 1 || frames < 1) return 2;
  SessionPlacement placement(fov_device_count());
  std::vector<Result> results(streams);
  std::vector<std::thread> threads;
  const auto t0 = std::chrono::steady_clock::now();
  for (int s = 0; s < streams; ++s)
    threads.emplace_back(Connection, s, frames, W, H, std::cref(trace), &placement, &results[s]);
  for (auto &t : threads) t.join();
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printf("{\"streams\": %d, \"frames\": %d, \"W\": %d, \"H\": %d, \"devices\": %d, \"seconds\": %.4f, "
         "\"frames_per_s\": %.1f, \"results\": [",
         streams, frames, W, H, placement.DeviceCount(), sec, streams * frames / sec);
  for (int s = 0; s < streams; ++s)
    printf("%s{\"device\": %d, \"record\": %u, \"hash\": \"%016llx\"}", s ? ", " : "", results[s].device,
           results[s].last_record, (unsigned long long)results[s].hash);
  printf("]}\n");
  return 0;
}
