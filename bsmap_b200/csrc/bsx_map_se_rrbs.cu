// bsx_map_se_rrbs.cu -- the single-end RRBS (-D) mapping kernel: same source, RRBS fixed at compile time.
#define BSX_BUILD_SE 1
#define BSX_CALLS 0
#define BSX_RRBS(A) 1
#define BSX_WIDE(A) 0
#define BSX_SE_KERNEL bsx_map_se_rrbs_kernel
#define BSX_SE_OCC bsx_map_occupancy_se_rrbs
#define BSX_SE_LAUNCH bsx_launch_map_se_rrbs
#include "bsx_map_impl.cuh"
