// Micro-benchmark 2: random 16-byte gather rate vs footprint (L2-resident .. HBM) and vs the L2
// fetch-granularity limit; variant = ld.global.nc.L2::64B unless noted.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
template <int V>
__device__ __forceinline__ uint4 ld(const uint4 *p) {
    uint4 v;
    if (V == 0) v = __ldg(p);
    else asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
template <int V, int U>
__global__ void gather(const uint4 *__restrict__ a, uint64_t n_vec, int iters, uint32_t *out) {
    uint64_t x = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    uint32_t acc = 0;
    for (int i = 0; i < iters; i += U) {
        uint64_t idx[U];
#pragma unroll
        for (int u = 0; u < U; u++) { x ^= x >> 12; x ^= x << 25; x ^= x >> 27; idx[u] = ((x * 0x2545F4914F6CDD1Dull) >> 20) % n_vec; }
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = ld<V>(a + idx[u]);
#pragma unroll
        for (int u = 0; u < U; u++) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678) out[0] = acc;
}
template <int V, int U>
void run(const char *name, const uint4 *a, uint64_t bytes, uint32_t *out, int blocks_per_sm) {
    const int iters = 64, blocks = 148 * blocks_per_sm, threads = 256;
    uint64_t n_vec = bytes / 16;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather<V, U><<<blocks, threads>>>(a, n_vec, 8, out);
    cudaEventRecord(e0);
    gather<V, U><<<blocks, threads>>>(a, n_vec, iters, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double g = (double)blocks * threads * iters;
    printf("%-10s U=%d blocks/SM=%2d footprint %6llu MB: %7.2f G gathers/s\n", name, U, blocks_per_sm, (unsigned long long)(bytes >> 20), g / ms * 1e-6);
}
int main(int argc, char **argv) {
    const uint64_t maxb = 8192ull << 20;
    uint4 *a; uint32_t *out;
    cudaMalloc(&a, maxb); cudaMemset(a, 1, maxb); cudaMalloc(&out, 4);
    if (argc > 1) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, atoi(argv[1]));
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity limit = %zu\n", g);
    for (uint64_t mb : {16ull, 48ull, 96ull, 192ull, 775ull, 1550ull, 8192ull}) run<1, 4>("nc.L2::64B", a, mb << 20, out, 8);
    for (int bps : {2, 4, 8}) { run<1, 1>("nc.L2::64B", a, 1550ull << 20, out, bps); run<1, 2>("nc.L2::64B", a, 1550ull << 20, out, bps); run<1, 8>("nc.L2::64B", a, 1550ull << 20, out, bps); }
    run<0, 4>("ldg", a, 1550ull << 20, out, 8);
    run<0, 4>("ldg", a, 48ull << 20, out, 8);
    return 0;
}
