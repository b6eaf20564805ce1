// warp_fast.cu -- staged fast path (placeholder until the tiled kernel lands).
#include "bevk_common.cuh"

int bevk_launch_warp_fast(const BevkWarpParams &, int, int, int, cudaStream_t) { return 0; }
