// bsx_map_se.cu -- the single-end mapping kernel (SingleAlign::Do_Batch): everything inlined.
#define BSX_BUILD_SE 1
#define BSX_CALLS 0
#include "bsx_map_impl.cuh"
