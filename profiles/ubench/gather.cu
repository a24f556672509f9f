// Micro-benchmark: random 16-byte gathers from a 1.5 GB array with different load flavours.
// Question: how many bytes does HBM move per 16-byte gather (sector vs line granularity)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather gather.cu && ./gather
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ld_nc(const uint4 *p) { return __ldg(p); }
__device__ __forceinline__ uint4 ld_ca(const uint4 *p) { return *p; }
__device__ __forceinline__ uint4 ld_cg(const uint4 *p) { return __ldcg(p); }
__device__ __forceinline__ uint4 ld_cs(const uint4 *p) { return __ldcs(p); }
__device__ __forceinline__ uint4 ld_lu(const uint4 *p) { return __ldlu(p); }
__device__ __forceinline__ uint4 ld_cv(const uint4 *p) { return __ldcv(p); }
__device__ __forceinline__ uint4 ld_nc_noalloc(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ld_nc_l2_64(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ld_evict_first(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.L1::evict_first.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ld_b32x4(const uint4 *p) {   // four scalar nc loads
    const uint32_t *q = (const uint32_t *)p;
    return make_uint4(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
}

template <int V>
__global__ void gather(const uint4 *__restrict__ a, uint64_t n_vec, int iters, uint32_t *out) {
    uint64_t x = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    uint32_t acc = 0;
    for (int i = 0; i < iters; i++) {
        x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
        uint64_t idx = ((x * 0x2545F4914F6CDD1Dull) >> 20) % n_vec;
        uint4 v;
        if (V == 0) v = ld_nc(a + idx);
        else if (V == 1) v = ld_ca(a + idx);
        else if (V == 2) v = ld_cg(a + idx);
        else if (V == 3) v = ld_cs(a + idx);
        else if (V == 4) v = ld_lu(a + idx);
        else if (V == 5) v = ld_cv(a + idx);
        else if (V == 6) v = ld_nc_noalloc(a + idx);
        else if (V == 7) v = ld_nc_l2_64(a + idx);
        else if (V == 8) v = ld_evict_first(a + idx);
        else v = ld_b32x4(a + idx);
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678) out[0] = acc;
}

template <int V>
void run(const char *name, const uint4 *a, uint64_t n_vec, uint32_t *out) {
    const int iters = 64, blocks = 148 * 8, threads = 512;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather<V><<<blocks, threads>>>(a, n_vec, 4, out);
    cudaEventRecord(e0);
    gather<V><<<blocks, threads>>>(a, n_vec, iters, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double g = (double)blocks * threads * iters;
    printf("%-22s %8.3f ms  %7.2f G gathers/s  (16 B each: %7.1f GB/s useful)  err=%s\n", name, ms, g / ms * 1e-6, g * 16 / ms * 1e-6,
           cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char **argv) {
    const uint64_t bytes = 1536ull << 20, n_vec = bytes / 16;
    uint4 *a; uint32_t *out;
    cudaMalloc(&a, bytes); cudaMemset(a, 1, bytes); cudaMalloc(&out, 4);
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    printf("cudaLimitMaxL2FetchGranularity default = %zu\n", g);
    if (argc > 1) { cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, atoi(argv[1])); cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("now %zu\n", g); }
    run<0>("ld.global.nc (ldg)", a, n_vec, out);
    run<1>("ld.global (ca)", a, n_vec, out);
    run<2>("ld.global.cg", a, n_vec, out);
    run<3>("ld.global.cs", a, n_vec, out);
    run<4>("ld.global.lu", a, n_vec, out);
    run<5>("ld.global.cv", a, n_vec, out);
    run<6>("nc.L1::no_allocate", a, n_vec, out);
    run<7>("nc.L2::64B", a, n_vec, out);
    run<8>("L1::evict_first", a, n_vec, out);
    run<9>("4 x ldg.b32", a, n_vec, out);
    return 0;
}
