/* The C ABI from plain C (no C++, no Python): capture the offline runner's per-frame sequence
 * (run_satlogrectilinear.cc:926-943) once as a CUDA graph and replay it per frame with a new gaze.
 * Prints the FNV-1a-64 hash of the un-warped frame of every replay and of the same frames computed
 * with three eager calls; tests/test_cpp_graph.py holds the two lists equal to each other and the
 * first one to the golden hashes generated from the reference's kernels.
 *   graph_replay W H SEED  cx0 cy0  cx1 cy1 ... */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "fov360.h"

#define CHECK(call)                                                              \
  do {                                                                           \
    int rc_ = (call);                                                            \
    if (rc_ != FOV_OK) {                                                         \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, fov_last_error_string(ctx)); \
      return 1;                                                                  \
    }                                                                            \
  } while (0)

static uint64_t fnv1a64(const uint8_t *b, size_t n) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 0x100000001b3ull;
  return h;
}

int main(int argc, char **argv) {
  if (argc < 6) return 2;
  const int W = atoi(argv[1]), H = atoi(argv[2]);
  const uint32_t seed = (uint32_t)atoi(argv[3]);
  const int ngaze = (argc - 4) / 2;
  const int ow = fov_reduced_dim(W), oh = fov_reduced_dim(H);
  const size_t fb = (size_t)4 * W * H, sb = (size_t)12 * W * H, rb = (size_t)4 * ow * oh;
  int err = 0;
  fov_ctx *ctx = fov_ctx_create(0, &err);
  if (!ctx) {
    fprintf(stderr, "fov_ctx_create: %s\n", fov_last_error_string(NULL));
    return 1;
  }
  uint8_t *frame = malloc(fb), *out = malloc(fb);
  uint32_t s = seed; /* SURVEY 8(c) generator */
  for (size_t i = 0; i < fb; ++i) {
    s = s * 1664525u + 1013904223u;
    frame[i] = ((i & 3) == 3) ? 0 : (uint8_t)(s >> 24);
  }
  void *src, *sat, *red, *full, *gaze_dev;
  CHECK(fov_malloc(ctx, &src, fb));
  CHECK(fov_malloc(ctx, &sat, sb));
  CHECK(fov_malloc(ctx, &red, rb));
  CHECK(fov_malloc(ctx, &full, fb));
  CHECK(fov_malloc(ctx, &gaze_dev, 8));
  CHECK(fov_memcpy_h2d(ctx, src, frame, fb));
  float gz[2] = {0.5f, 0.5f};
  CHECK(fov_memcpy_h2d(ctx, gaze_dev, gz, 8));
  /* one eager run: tables and scratch come into being outside the capture */
  CHECK(fov_sat_foveate_batched_dev(ctx, 1, full, fb, red, rb, sat, sb, src, fb, W, H, 4 * W, ow, oh,
                                    gaze_dev));
  CHECK(fov_sync(ctx));
  fov_graph *graph = NULL;
  CHECK(fov_graph_begin_capture(ctx));
  CHECK(fov_sat_foveate_batched_dev(ctx, 1, full, fb, red, rb, sat, sb, src, fb, W, H, 4 * W, ow, oh,
                                    gaze_dev));
  CHECK(fov_graph_end_capture(ctx, &graph));
  printf("{\"graph\": [");
  for (int k = 0; k < ngaze; ++k) {
    gz[0] = (float)atof(argv[4 + 2 * k]);
    gz[1] = (float)atof(argv[5 + 2 * k]);
    CHECK(fov_memset(ctx, red, 0, rb)); /* the golden hashes start from a cleared reduced buffer */
    CHECK(fov_memcpy_h2d_async(ctx, gaze_dev, gz, 8));
    CHECK(fov_graph_launch(ctx, graph));
    CHECK(fov_memcpy_d2h(ctx, out, full, fb)); /* blocking: gz may change after this */
    printf("%s\"%016llx\"", k ? ", " : "", (unsigned long long)fnv1a64(out, fb));
  }
  printf("], \"eager\": [");
  for (int k = 0; k < ngaze; ++k) {
    const float cx = (float)atof(argv[4 + 2 * k]), cy = (float)atof(argv[5 + 2 * k]);
    CHECK(fov_memset(ctx, red, 0, rb));
    CHECK(fov_sat_encode(ctx, sat, src, W, H, 4 * W));
    CHECK(fov_sat_sample_rect(ctx, red, ow, oh, 4 * ow, sat, W, H, cx, cy));
    CHECK(fov_sat_interpolate_rect(ctx, full, W, H, 4 * W, red, ow, oh, 4 * ow, cx, cy));
    CHECK(fov_memcpy_d2h(ctx, out, full, fb));
    printf("%s\"%016llx\"", k ? ", " : "", (unsigned long long)fnv1a64(out, fb));
  }
  printf("], \"launches\": %llu}\n", (unsigned long long)fov_ctx_launch_count(ctx));
  fov_graph_destroy(ctx, graph);
  fov_ctx_destroy(ctx);
  free(frame);
  free(out);
  return 0;
}
