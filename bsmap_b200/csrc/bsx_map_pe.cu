// bsx_map_pe.cu -- the paired-end mapping kernel (PairAlign::Do_Batch): big device functions are calls.
#define BSX_BUILD_PE 1
#define BSX_CALLS 1
#include "bsx_map_impl.cuh"
