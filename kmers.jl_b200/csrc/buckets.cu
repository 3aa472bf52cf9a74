// buckets.cu -- the canonical k-mer bucket-count table (north_star extension; not in the reference):
// table[fx_hash(canonical k-mer) >> (64 - B)] += 1 over a read set.
//
// Two regimes, both bound by how fast L2 can apply 4-byte increments (2.2e11 / s measured):
//   * the table fits L2 (<= 64 MB): extract_kernel<SINK_BUCKETS> increments it directly;
//   * larger tables (B = 28 is 1 GiB): random increments miss L2 and every one of them becomes a
//     32-byte DRAM sector read + write (measured 24 G k-mers/s).  So the bucket ids are first
//     written out (SINK_IDS), partitioned by their high bits into bins whose table slice is
//     <= 16 MB (per-block counting sort in shared memory, exact placement from a count matrix --
//     no atomics on global cursors), and then applied bin after bin, so that all SMs work on one
//     L2-resident slice of the table at a time.
#include "plan.h"

namespace kmc {

namespace {

constexpr int kBinBlock = 256;
constexpr int kBinIters = 8;
constexpr int kBinPerIter = kBinBlock * 8;          // ids staged per iteration
constexpr int kBinChunk = kBinPerIter * kBinIters;  // ids per block
constexpr int kMaxBins = 1024;

__global__ void __launch_bounds__(kBinBlock) bin_hist_kernel(const uint32_t *__restrict__ ids, uint64_t n, int shift, int n_bins,
                                                            uint64_t n_blocks, uint64_t *__restrict__ matrix)
{
    __shared__ uint32_t s_cnt[kMaxBins];
    for (int b = threadIdx.x; b < n_bins; b += kBinBlock) s_cnt[b] = 0;
    __syncthreads();
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kBinChunk;
    for (int i = 0; i < kBinChunk / kBinBlock; ++i) {
        const uint64_t e = base + static_cast<uint64_t>(i) * kBinBlock + threadIdx.x;
        if (e < n) atomicAdd(&s_cnt[__ldg(ids + e) >> shift], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < n_bins; b += kBinBlock) matrix[static_cast<uint64_t>(b) * n_blocks + blockIdx.x] = s_cnt[b];
}

__global__ void __launch_bounds__(kBinBlock) bin_scatter_kernel(const uint32_t *__restrict__ ids, uint64_t n, int shift, int n_bins,
                                                               uint64_t n_blocks, const uint64_t *__restrict__ offs,
                                                               uint32_t *__restrict__ binned)
{
    __shared__ uint32_t s_cnt[kMaxBins], s_start[kMaxBins], s_off[kMaxBins];
    __shared__ uint64_t s_glob[kMaxBins];
    __shared__ uint32_t s_ids[kBinPerIter];
    __shared__ uint16_t s_bin[kBinPerIter];
    __shared__ uint32_t s_wsum[kBinBlock / 32];
    for (int b = threadIdx.x; b < n_bins; b += kBinBlock) s_glob[b] = offs[static_cast<uint64_t>(b) * n_blocks + blockIdx.x];
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kBinChunk;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int it = 0; it < kBinIters; ++it) {
        const uint64_t it_base = base + static_cast<uint64_t>(it) * kBinPerIter;
        if (it_base >= n) break; // block-uniform
        for (int b = threadIdx.x; b < n_bins; b += kBinBlock) s_cnt[b] = 0;
        __syncthreads();
        uint32_t id[8];
        bool ok[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint64_t e = it_base + static_cast<uint64_t>(j) * kBinBlock + threadIdx.x;
            ok[j] = e < n;
            id[j] = ok[j] ? __ldg(ids + e) : 0u;
            if (ok[j]) atomicAdd(&s_cnt[id[j] >> shift], 1u);
        }
        __syncthreads();
        // exclusive scan of s_cnt[0..n_bins) -> s_start (4 bins per thread, 256 threads cover 1024 bins)
        {
            uint32_t c[4], sum = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int b = 4 * threadIdx.x + q;
                c[q] = b < n_bins ? s_cnt[b] : 0u;
                sum += c[q];
            }
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            uint32_t before = 0;
#pragma unroll
            for (int w = 0; w < kBinBlock / 32; ++w) before += (w < warp) ? s_wsum[w] : 0u;
            uint32_t run = before + incl - sum;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int b = 4 * threadIdx.x + q;
                if (b < n_bins) {
                    s_start[b] = run;
                    s_off[b] = run;
                }
                run += c[q];
            }
        }
        __syncthreads();
        uint32_t total = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (ok[j]) {
                const uint32_t b = id[j] >> shift;
                const uint32_t pos = atomicAdd(&s_off[b], 1u);
                s_ids[pos] = id[j];
                s_bin[pos] = static_cast<uint16_t>(b);
            }
        }
        __syncthreads();
        {
            const uint64_t left = n - it_base;
            total = left < kBinPerIter ? static_cast<uint32_t>(left) : kBinPerIter;
        }
        for (uint32_t t = threadIdx.x; t < total; t += kBinBlock) {
            const uint32_t b = s_bin[t];
            binned[s_glob[b] + (t - s_start[b])] = s_ids[t];
        }
        __syncthreads();
        for (int b = threadIdx.x; b < n_bins; b += kBinBlock) s_glob[b] += s_cnt[b];
        // the next iteration's zeroing of s_cnt is ordered behind this by its own barrier
        __syncthreads();
    }
}

// An increment that MISSES L2 is far more expensive than its 64 bytes of DRAM traffic: the L2 slice's
// atomic unit waits for the fill, so misses serialise at DRAM latency (tools/micro/atomics_probe.cu:
// 2.2e11 increments/s on a warm 16 MB table, 2.3e10/s when every increment misses -- and still only
// 2.7e10/s on binned ids, because each slice is cold when its bin starts).  So every bin first pulls
// its table slice into L2 with plain coalesced loads (bandwidth-bound, 16 MB), then applies its ids.
__global__ void __launch_bounds__(256) warm_slice_kernel(const uint32_t *__restrict__ slice, uint64_t n_counters,
                                                         uint32_t *__restrict__ sink)
{
    const uint4 *p = reinterpret_cast<const uint4 *>(slice);
    const uint64_t n4 = n_counters / 4;
    uint32_t acc = 0;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        uint4 v;
        asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));
        acc |= v.x & v.y & v.z & v.w;
    }
    if (acc == 0xffffffffu) *sink = acc; // keeps the loads alive; a count of 2^32-1 in four neighbours does not happen
}

// the ids of bin `bin` are binned[offs[bin * n_blocks] .. offs[(bin + 1) * n_blocks])
__global__ void __launch_bounds__(256) bin_apply_kernel(const uint32_t *__restrict__ binned, const uint64_t *__restrict__ offs,
                                                        uint64_t n_blocks, int bin, uint32_t *__restrict__ table)
{
    const uint64_t begin = __ldg(offs + static_cast<uint64_t>(bin) * n_blocks);
    const uint64_t end = __ldg(offs + static_cast<uint64_t>(bin + 1) * n_blocks);
    const uint64_t a4 = (begin + 3) & ~3ull; // 16-byte aligned middle part
    const uint64_t tid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t threads = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    if (tid < a4 - begin && begin + tid < end) atomicAdd(table + binned[begin + tid], 1u);
    for (uint64_t i = a4 + tid * 4; i < end; i += threads * 4) {
        if (i + 4 <= end) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(binned + i));
            atomicAdd(table + v.x, 1u);
            atomicAdd(table + v.y, 1u);
            atomicAdd(table + v.z, 1u);
            atomicAdd(table + v.w, 1u);
        } else {
            for (uint64_t j = i; j < end; ++j) atomicAdd(table + binned[j], 1u);
        }
    }
}

} // namespace

// ids[0..n) are bucket ids of `bucket_bits` bits; adds their histogram to table.  tmp: n u32 (binned ids),
// matrix / offs: (n_bins * n_blocks + 1) u64 each, scan_tmp: scan_tmp_elems(n_bins * n_blocks) u64.
cudaError_t binned_count(const uint32_t *ids, uint64_t n, int bucket_bits, uint32_t *table, uint32_t *binned, uint64_t *matrix,
                         uint64_t *offs, uint64_t *scan_tmp, int sm_count, cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    const int p = binned_count_bin_bits(bucket_bits);
    const int n_bins = 1 << p, shift = bucket_bits - p;
    const uint64_t n_blocks = binned_count_blocks(n);
    bin_hist_kernel<<<static_cast<unsigned>(n_blocks), kBinBlock, 0, stream>>>(ids, n, shift, n_bins, n_blocks, matrix);
    cudaError_t e = inclusive_offsets_u64(matrix, offs, static_cast<uint64_t>(n_bins) * n_blocks, scan_tmp, stream);
    if (e != cudaSuccess) return e;
    bin_scatter_kernel<<<static_cast<unsigned>(n_blocks), kBinBlock, 0, stream>>>(ids, n, shift, n_bins, n_blocks, offs, binned);
    // offs has n_bins * n_blocks + 1 entries: offs[(bin + 1) * n_blocks] of the last bin is the total
    const uint64_t slice = 1ull << shift; // counters per bin
    for (int b = 0; b < n_bins; ++b) {
        warm_slice_kernel<<<static_cast<unsigned>(sm_count * 8), 256, 0, stream>>>(table + static_cast<uint64_t>(b) * slice, slice,
                                                                                  reinterpret_cast<uint32_t *>(scan_tmp));
        bin_apply_kernel<<<static_cast<unsigned>(sm_count * 16), 256, 0, stream>>>(binned, offs, n_blocks, b, table);
    }
    return cudaGetLastError();
}

cudaError_t warm_table(const uint32_t *table, uint64_t n_counters, int sm_count, cudaStream_t stream)
{
    static uint32_t *sink = nullptr; // never written in practice (see warm_slice_kernel)
    if (!sink) {
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&sink), 256);
        if (e != cudaSuccess) return e;
    }
    warm_slice_kernel<<<static_cast<unsigned>(sm_count * 8), 256, 0, stream>>>(table, n_counters, sink);
    return cudaGetLastError();
}

int binned_count_bin_bits(int bucket_bits)
{
    int p = bucket_bits - 22; // table slice per bin <= 2^22 counters = 16 MB
    if (p < 6) p = 6;
    if (p > 10) p = 10;
    if (p > bucket_bits) p = bucket_bits;
    return p;
}

uint64_t binned_count_blocks(uint64_t n) { return (n + kBinChunk - 1) / kBinChunk; }

} // namespace kmc
