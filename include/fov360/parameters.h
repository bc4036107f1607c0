// Foveation parameters of the reference (src/parameters.h:8-9, run_satlogrectilinear.cc:113-114).
#pragma once
#include "../fov360.h"

#define REDUCED_BUFFER_WIDTH 1072
#define REDUCED_BUFFER_HEIGHT 608

// 16 * ceil(dim / 1.8 / 16): the reduced-buffer size the offline runner derives per resolution.
inline int ReducedBufferDim(int full_dim) { return fov_reduced_dim(full_dim); }
