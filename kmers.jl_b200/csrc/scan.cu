// scan.cu -- small device-side prefix sums used to lay out ragged read sets (window offsets,
// group-slot offsets) and, later, ordered compaction.  Reduce-then-scan in three launches:
// per-tile sums, a single-block scan of the tile sums, per-tile rescan with the carried offset.
// The inputs are per-READ arrays (8 bytes per read against >= 16 bytes per WINDOW of output), so
// this is plumbing, not the hot path.
#include "kmc_internal.h"

namespace kmc {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

uint64_t scan_tmp_elems(uint64_t n) { return (n + kScanTile - 1) / kScanTile + 1; }

__device__ __forceinline__ uint64_t warp_incl_scan(uint64_t v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) >= d) v += t;
    }
    return v;
}

// block-wide exclusive scan of one value per thread; returns the exclusive prefix, *total = block sum
__device__ __forceinline__ uint64_t block_excl_scan(uint64_t v, uint64_t *total)
{
    __shared__ uint64_t warp_sums[kScanThreads / 32];
    __shared__ uint64_t block_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t incl = warp_incl_scan(v);
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint64_t w = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
        uint64_t wi = warp_incl_scan(w);
        if (lane < kScanThreads / 32) warp_sums[lane] = wi - w;
        if (lane == kScanThreads / 32 - 1) block_total = wi;
    }
    __syncthreads();
    uint64_t r = incl - v + warp_sums[warp];
    *total = block_total;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads) tile_sums_kernel(const uint64_t *__restrict__ in, uint64_t n,
                                                                uint64_t *__restrict__ sums)
{
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kScanTile;
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        uint64_t idx = base + static_cast<uint64_t>(i) * kScanThreads + threadIdx.x;
        if (idx < n) s += in[idx];
    }
    uint64_t total;
    block_excl_scan(s, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// single block: exclusive scan of sums[0..m) in place
__global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(uint64_t *sums, uint64_t m)
{
    uint64_t carry = 0;
    for (uint64_t base = 0; base < m; base += kScanThreads) {
        uint64_t idx = base + threadIdx.x;
        uint64_t v = idx < m ? sums[idx] : 0;
        uint64_t total;
        uint64_t ex = block_excl_scan(v, &total);
        if (idx < m) sums[idx] = ex + carry;
        carry += total;
    }
}

// out[0] = 0 ; out[i+1] = inclusive prefix.  Thread t owns kScanItems CONSECUTIVE elements.
__global__ void __launch_bounds__(kScanThreads) tile_rescan_kernel(const uint64_t *__restrict__ in, uint64_t n,
                                                                  const uint64_t *__restrict__ sums,
                                                                  uint64_t *__restrict__ out)
{
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kScanTile + static_cast<uint64_t>(threadIdx.x) * kScanItems;
    uint64_t v[kScanItems];
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        s += v[i];
    }
    uint64_t total;
    uint64_t ex = block_excl_scan(s, &total) + sums[blockIdx.x];
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        ex += v[i];
        if (base + i < n) out[base + i + 1] = ex;
    }
}

cudaError_t inclusive_offsets_u64(const uint64_t *in, uint64_t *out, uint64_t n, uint64_t *tmp, cudaStream_t stream)
{
    if (n == 0) {
        return cudaMemsetAsync(out, 0, sizeof(uint64_t), stream);
    }
    const uint64_t tiles = (n + kScanTile - 1) / kScanTile;
    tile_sums_kernel<<<static_cast<unsigned>(tiles), kScanThreads, 0, stream>>>(in, n, tmp);
    scan_sums_kernel<<<1, kScanThreads, 0, stream>>>(tmp, tiles);
    tile_rescan_kernel<<<static_cast<unsigned>(tiles), kScanThreads, 0, stream>>>(in, n, tmp, out);
    return cudaGetLastError();
}

__global__ void window_counts_kernel(const uint64_t *__restrict__ seq_len, uint64_t n, uint64_t k,
                                     uint64_t *__restrict__ cnt)
{
    uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) {
        uint64_t len = seq_len[i];
        cnt[i] = len >= k ? len - k + 1 : 0; // FwKmers.jl:40-43
    }
}

cudaError_t window_counts(const uint64_t *seq_len, uint64_t n, int k, uint64_t *cnt, cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    window_counts_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(seq_len, n, static_cast<uint64_t>(k), cnt);
    return cudaGetLastError();
}

__global__ void group_slots_kernel(const uint64_t *__restrict__ win_off, uint64_t n, uint64_t g,
                                   uint64_t *__restrict__ slots)
{
    uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) {
        uint64_t a = win_off[i], b = win_off[i + 1];
        slots[i] = (b > a) ? ((b + g - 1) / g - a / g) : 0;
    }
}

cudaError_t group_slots(const uint64_t *win_off, uint64_t n, int g, uint64_t *slots, cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    group_slots_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(win_off, n, static_cast<uint64_t>(g), slots);
    return cudaGetLastError();
}

// tile_first[t] = the sequence that owns work item t * tile_items (largest r with item_off[r] <= item),
// for t in [0, n_tiles]; the extraction kernels narrow their per-item searches with it.
__global__ void tile_first_kernel(const uint64_t *__restrict__ item_off, uint64_t n_seqs, uint64_t tile_items,
                                  uint64_t n_out, uint64_t *__restrict__ tile_first)
{
    const uint64_t t = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n_out) return;
    const uint64_t item = t * tile_items;
    uint64_t lo = 0, hi = n_seqs;
    while (hi - lo > 1) {
        const uint64_t mid = (lo + hi) >> 1;
        if (item_off[mid] <= item) lo = mid; else hi = mid;
    }
    tile_first[t] = lo;
}

cudaError_t tile_first_reads(const uint64_t *item_off, uint64_t n_seqs, uint64_t tile_items, uint64_t n_tiles,
                             uint64_t *tile_first, cudaStream_t stream)
{
    const uint64_t n_out = n_tiles + 1;
    tile_first_kernel<<<static_cast<unsigned>((n_out + 255) / 256), 256, 0, stream>>>(item_off, n_seqs, tile_items, n_out, tile_first);
    return cudaGetLastError();
}

} // namespace kmc
