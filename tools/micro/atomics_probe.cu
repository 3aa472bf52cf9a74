// atomics_probe.cu -- how fast can a B200 apply 4-byte increments to a table, as a function of the
// table (slice) size and of where the indices come from (generated in registers vs read from memory)?
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void gen_ids(uint32_t *ids, uint64_t n, uint32_t mask, uint64_t per_bin, int bin_shift)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t v = (uint32_t)mix(i) & mask;
        if (per_bin) v = (v & ((1u << bin_shift) - 1)) | ((uint32_t)(i / per_bin) << bin_shift);
        ids[i] = v;
    }
}

__global__ void apply_mem(const uint32_t *__restrict__ ids, uint64_t n, uint32_t *table)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 4;
    for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i + 4 <= n; i += stride) {
        const uint4 v = __ldg((const uint4 *)(ids + i));
        atomicAdd(table + v.x, 1u);
        atomicAdd(table + v.y, 1u);
        atomicAdd(table + v.z, 1u);
        atomicAdd(table + v.w, 1u);
    }
}

__global__ void apply_reg(uint64_t n, uint32_t mask, uint32_t *table)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        atomicAdd(table + ((uint32_t)mix(i) & mask), 1u);
}

int main()
{
    const uint64_t n = 1ull << 30;
    uint32_t *ids, *table;
    cudaMalloc(&ids, n * 4);
    cudaMalloc(&table, 4ull << 28);
    cudaMemset(table, 0, 4ull << 28);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms;
    for (int bits : {16, 20, 22, 24, 26, 28}) {
        const uint32_t mask = (1u << bits) - 1;
        apply_reg<<<148 * 16, 256>>>(n, mask, table);
        cudaEventRecord(e0);
        apply_reg<<<148 * 16, 256>>>(n, mask, table);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("reg  random over 2^%d counters (%5.0f MB): %8.3f ms  %.2e inc/s\n", bits, 4.0 * (1u << bits) / 1e6, ms, n / (ms * 1e-3));
        gen_ids<<<148 * 16, 256>>>(ids, n, mask, 0, 0);
        apply_mem<<<148 * 16, 256>>>(ids, n, table);
        cudaEventRecord(e0);
        apply_mem<<<148 * 16, 256>>>(ids, n, table);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("mem  random over 2^%d counters            : %8.3f ms  %.2e inc/s\n", bits, ms, n / (ms * 1e-3));
    }
    // binned: 2^28 table, ids grouped into 64 / 256 / 1024 bins (consecutive ids share the high bits)
    for (int p : {6, 8, 10, 12}) {
        const int shift = 28 - p;
        gen_ids<<<148 * 16, 256>>>(ids, n, (1u << 28) - 1, n >> p, shift);
        apply_mem<<<148 * 16, 256>>>(ids, n, table);
        cudaEventRecord(e0);
        apply_mem<<<148 * 16, 256>>>(ids, n, table);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("mem  binned, 2^28 table in %4d bins of %6.2f MB: %8.3f ms  %.2e inc/s\n", 1 << p, 4.0 * (1u << shift) / 1e6, ms, n / (ms * 1e-3));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
