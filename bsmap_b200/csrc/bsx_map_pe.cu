// bsx_map_pe.cu -- the paired-end mapping kernel (PairAlign::Do_Batch): inlined whole; every big device function has one
// call site (bsx_map_impl.cuh loops over the mates), so the list walk exists once in the binary.
#define BSX_BUILD_PE 1
#define BSX_CALLS 0
#define BSX_WIDE(A) 0          // the wide-context phase costs the pairing kernel 10 % in registers; pairs at -v >= 8 use 32 bases
#include "bsx_map_impl.cuh"
