// bsx_map_pe_wide.cu -- the paired-end WGBS kernel for indexes built with -v >= 8 (16-byte context entries).
#define BSX_BUILD_PE 1
#define BSX_CALLS 0
#define BSX_RRBS(A) 0
#define BSX_WIDE(A) 1
#define BSX_PE_KERNEL bsx_map_pe_wide_kernel
#define BSX_PE_OCC bsx_map_occupancy_pe_wide
#define BSX_PE_LAUNCH bsx_launch_map_pe_wide
#include "bsx_map_impl.cuh"
