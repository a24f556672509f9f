// bsx_map_se_wide.cu -- the single-end WGBS kernel for -v >= 8: candidates that pass the 32-base inline context are
// tested against the next 16 bases on either side (bsx_index::d_ctx2) before they touch the reference.
#define BSX_BUILD_SE 1
#define BSX_CALLS 0
#define BSX_RRBS(A) 0
#define BSX_WIDE(A) 1
#define BSX_SE_KERNEL bsx_map_se_wide_kernel
#define BSX_SE_OCC bsx_map_occupancy_se_wide
#define BSX_SE_LAUNCH bsx_launch_map_se_wide
#include "bsx_map_impl.cuh"
