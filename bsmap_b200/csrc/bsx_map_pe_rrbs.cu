// bsx_map_pe_rrbs.cu -- the paired-end RRBS (-D) mapping kernel: same source, RRBS fixed at compile time.
#define BSX_BUILD_PE 1
#define BSX_CALLS 0
#define BSX_RRBS(A) 1
#define BSX_WIDE(A) 0
#define BSX_PE_KERNEL bsx_map_pe_rrbs_kernel
#define BSX_PE_OCC bsx_map_occupancy_pe_rrbs
#define BSX_PE_LAUNCH bsx_launch_map_pe_rrbs
#include "bsx_map_impl.cuh"
