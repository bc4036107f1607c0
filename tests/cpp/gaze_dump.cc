// Prints every record a GazeViewPoints parser extracts from a trace file, one JSON array.  Built
// twice by tests/test_gaze_trace.py: against include/fov360/gaze_view_points.h (the drop-in) and,
// where /root/reference exists, against the reference's own src/gaze_view_points.{h,cc}
// (-DUSE_REFERENCE) to generate / cross-check tests/golden/gaze_trace.json.
#include <cstdio>
#ifdef USE_REFERENCE
#include "gaze_view_points.h"
#else
#include "fov360/gaze_view_points.h"
#include "fov360/session_placement.h"
#endif

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  GazeViewPoints gv{std::string(argv[1])};
  printf("[");
  for (size_t i = 0; i < gv.points.size(); ++i) {
    const auto &p = gv.points[i];
    printf("%s[%u, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g]", i ? ", " : "", p.frame,
           p.view_point[0], p.view_point[1], p.gaze_point[0], p.gaze_point[1], p.pred_view_point[0],
           p.pred_view_point[1], p.pred_gaze_point[0], p.pred_gaze_point[1]);
  }
  printf("]\n");
#ifndef USE_REFERENCE
  if (argc > 2) {  // placement self-check: 8 devices, 19 sessions, 3 leave, 2 join
    SessionPlacement pl(8);
    int dev[19];
    for (int s = 0; s < 19; ++s) dev[s] = pl.Acquire();
    for (int s = 0; s < 19; ++s)
      if (dev[s] != s % 8) return 3;
    pl.Release(dev[0]), pl.Release(dev[8]), pl.Release(dev[5]);
    if (pl.Acquire() != 0 || pl.Acquire() != 5 || pl.Acquire() != 0 || pl.Sessions(0) != 3) return 4;
  }
#endif
  return 0;
}
