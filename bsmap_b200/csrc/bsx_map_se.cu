// bsx_map_se.cu -- the single-end WGBS mapping kernel (SingleAlign::Do_Batch): everything inlined, RRBS code
// compiled out (the kernel is instruction-cache sensitive: 90 KB of SASS with it).
#define BSX_BUILD_SE 1
#define BSX_CALLS 0
#define BSX_RRBS(A) 0
#define BSX_SE_KERNEL bsx_map_se_wgbs_kernel
#define BSX_SE_OCC bsx_map_occupancy_se_wgbs
#define BSX_SE_LAUNCH bsx_launch_map_se_wgbs
#include "bsx_map_impl.cuh"
