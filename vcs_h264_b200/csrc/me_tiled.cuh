// me_tiled.cuh -- tiled step-1 full-search kernel (placeholder: not built yet, AUTO falls back
// to me_generic).
#pragma once
#include "common.cuh"

namespace vcs {
struct MeTiledState {};
inline int me_tiled_launch(MeTiledState &, cudaStream_t, int, const MeGeom &, const FrameAddr &, int,
                           int16_t *, uint32_t *, uint8_t *, int, size_t, int *used, char *, size_t) {
    *used = 0;
    return 0;
}
inline void me_tiled_destroy(MeTiledState &) {}
}  // namespace vcs
