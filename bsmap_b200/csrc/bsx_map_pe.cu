// bsx_map_pe.cu -- the paired-end mapping kernel (PairAlign::Do_Batch): big device functions are calls.
#define BSX_BUILD_PE 1
#define BSX_CALLS 1
#define BSX_WIDE(A) 0          // the wide-context phase costs the pairing kernel 10 % in registers; pairs at -v >= 8 use 32 bases
#include "bsx_map_impl.cuh"
