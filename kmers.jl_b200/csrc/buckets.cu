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
#include <type_traits>

#include "plan.h"

namespace kmc {

namespace {

constexpr int kBinBlock = 256;
constexpr int kBinWarps = kBinBlock / 32;
constexpr int kBinPerThread = 16;                          // ids per thread per iteration (four 128-bit loads)
constexpr int kBinPerIter = kBinBlock * kBinPerThread;     // ids staged per iteration
constexpr int kBinIters = 4;
constexpr int kBinChunk = kBinPerIter * kBinIters;         // ids per block
constexpr int kMinBinBits = 6, kMaxBinBits = 10;

// Shared-memory atomics cost about two cycles per lane on this machine, as much as the L2 increments the
// binning is there to save, so bins are counted and ranked without them, the way a radix sort ranks its
// digits: the lanes of a warp that hold the same bin find each other with one ballot per bin bit, the
// lowest of them adds their number to a counter that only this warp touches, and a lane's rank is the
// counter before that plus the number of peers below it.
template <int P>
__device__ __forceinline__ uint32_t bin_peers(uint32_t bin, bool ok)
{
    uint32_t peers = __ballot_sync(0xffffffffu, ok);
#pragma unroll
    for (int b = 0; b < P; ++b) {
        const bool bit = (bin >> b) & 1u;
        const uint32_t set = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? set : ~set;
    }
    return peers;
}

// Loads the kBinPerThread ids of this thread for the iteration starting at it_base: warp w owns the
// ids [w * 512, w * 512 + 512) of the iteration, as four rows of 128 (one uint4 per lane).  `ids` must
// be 16-byte aligned; the tail of the array is loaded id by id.
__device__ __forceinline__ void load_iter_ids(const uint32_t *__restrict__ ids, uint64_t n, uint64_t it_base,
                                              uint32_t (&id)[kBinPerThread], uint32_t &ok_mask)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ok_mask = 0;
#pragma unroll
    for (int r = 0; r < kBinPerThread / 4; ++r) {
        const uint64_t e = it_base + static_cast<uint64_t>(warp) * (32 * kBinPerThread) + r * 128 + lane * 4;
        if (e + 4 <= n) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(ids + e));
            id[4 * r + 0] = v.x;
            id[4 * r + 1] = v.y;
            id[4 * r + 2] = v.z;
            id[4 * r + 3] = v.w;
            ok_mask |= 0xfu << (4 * r);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool ok = e + q < n;
                id[4 * r + q] = ok ? __ldg(ids + e + q) : 0u;
                ok_mask |= (ok ? 1u : 0u) << (4 * r + q);
            }
        }
    }
}

// matrix[bin * n_blocks + block] = number of ids of bin `bin` in the block's chunk
template <int P>
__global__ void __launch_bounds__(kBinBlock) bin_hist_kernel(const uint32_t *__restrict__ ids, uint64_t n, int shift,
                                                            uint64_t n_blocks, uint64_t *__restrict__ matrix)
{
    constexpr int NB = 1 << P;
    __shared__ uint32_t s_cnt[kBinWarps][NB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b = lane; b < NB; b += 32) s_cnt[warp][b] = 0;
    __syncwarp();
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kBinChunk;
    for (int it = 0; it < kBinIters; ++it) {
        const uint64_t it_base = base + static_cast<uint64_t>(it) * kBinPerIter;
        if (it_base >= n) break; // block-uniform
        uint32_t id[kBinPerThread], ok_mask;
        load_iter_ids(ids, n, it_base, id, ok_mask);
#pragma unroll
        for (int j = 0; j < kBinPerThread; ++j) {
            const bool ok = (ok_mask >> j) & 1u;
            const uint32_t bin = id[j] >> shift;
            const uint32_t peers = bin_peers<P>(bin, ok);
            if (ok && lane == __ffs(peers) - 1) s_cnt[warp][bin] += __popc(peers);
            __syncwarp();
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < NB; b += kBinBlock) {
        uint32_t c = 0;
#pragma unroll
        for (int w = 0; w < kBinWarps; ++w) c += s_cnt[w][b];
        matrix[static_cast<uint64_t>(b) * n_blocks + blockIdx.x] = c;
    }
}

// binned[offs[bin * n_blocks + block] ...) receives the block's ids of bin `bin` (any order within the bin)
template <int P>
__global__ void __launch_bounds__(kBinBlock, 3) bin_scatter_kernel(const uint32_t *__restrict__ ids, uint64_t n, int shift,
                                                               uint64_t n_blocks, const uint64_t *__restrict__ offs,
                                                               uint32_t *__restrict__ binned)
{
    constexpr int NB = 1 << P;
    constexpr int BPT = (NB + kBinBlock - 1) / kBinBlock; // bins per thread in the block scan
    __shared__ uint16_t s_wcnt[kBinWarps][NB]; // per warp: count, then the warp's offset within the bin
    __shared__ uint32_t s_start[NB + 1];       // first staging slot of each bin in this iteration
    __shared__ uint64_t s_glob[NB];            // where the block's next id of each bin goes
    __shared__ uint32_t s_ids[kBinPerIter];
    __shared__ uint32_t s_wsum[kBinWarps];
    for (int b = threadIdx.x; b < NB; b += kBinBlock) s_glob[b] = offs[static_cast<uint64_t>(b) * n_blocks + blockIdx.x];
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kBinChunk;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (int it = 0; it < kBinIters; ++it) {
        const uint64_t it_base = base + static_cast<uint64_t>(it) * kBinPerIter;
        if (it_base >= n) break; // block-uniform
        for (int b = lane; b < NB; b += 32) s_wcnt[warp][b] = 0;
        __syncwarp();
        uint32_t id[kBinPerThread], ok_mask;
        uint16_t rank[kBinPerThread];
        load_iter_ids(ids, n, it_base, id, ok_mask);
#pragma unroll
        for (int j = 0; j < kBinPerThread; ++j) {
            const bool ok = (ok_mask >> j) & 1u;
            const uint32_t bin = id[j] >> shift;
            const uint32_t peers = bin_peers<P>(bin, ok);
            const uint32_t before = ok ? s_wcnt[warp][bin] : 0u;
            rank[j] = static_cast<uint16_t>(before + __popc(peers & lt_mask));
            __syncwarp();
            if (ok && lane == __ffs(peers) - 1) s_wcnt[warp][bin] = static_cast<uint16_t>(before + __popc(peers));
            __syncwarp();
        }
        __syncthreads();
        // per bin: exclusive prefix over the warps (in place) and the total; exclusive scan of the totals -> s_start
        {
            uint32_t tot[BPT], sum = 0;
#pragma unroll
            for (int q = 0; q < BPT; ++q) {
                const int b = BPT * threadIdx.x + q;
                uint32_t run = 0;
                if (b < NB) {
#pragma unroll
                    for (int w = 0; w < kBinWarps; ++w) {
                        const uint32_t c = s_wcnt[w][b];
                        s_wcnt[w][b] = static_cast<uint16_t>(run);
                        run += c;
                    }
                }
                tot[q] = run;
                sum += run;
            }
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            uint32_t run = incl - sum;
#pragma unroll
            for (int w = 0; w < kBinWarps; ++w) run += (w < warp) ? s_wsum[w] : 0u;
#pragma unroll
            for (int q = 0; q < BPT; ++q) {
                const int b = BPT * threadIdx.x + q;
                if (b < NB) s_start[b] = run;
                run += tot[q];
            }
            if (threadIdx.x == kBinBlock - 1) s_start[NB] = run;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kBinPerThread; ++j) {
            if ((ok_mask >> j) & 1u) {
                const uint32_t bin = id[j] >> shift;
                s_ids[s_start[bin] + s_wcnt[warp][bin] + rank[j]] = id[j];
            }
        }
        __syncthreads();
        const uint32_t total = s_start[NB];
        for (uint32_t t = threadIdx.x; t < total; t += kBinBlock) {
            const uint32_t v = s_ids[t];
            const uint32_t b = v >> shift;
            binned[s_glob[b] + (t - s_start[b])] = v;
        }
        __syncthreads();
        for (int b = threadIdx.x; b < NB; b += kBinBlock) s_glob[b] += s_start[b + 1] - s_start[b];
        // the next iteration's writes to s_wcnt / s_start / s_ids are ordered behind this by its barriers
        __syncthreads();
    }
}

// 64 bins (every table up to 2^28 counters): no warp cooperation at all.  A thread counts the bins of
// its own ids in byte counters that only it touches -- the four bins 4g..4g+3 share the word [g][thread],
// so a warp's accesses fall into 32 different banks -- and the value a counter had before an id was
// counted is that id's rank among the thread's ids of the bin.  Threads are ordered by
// (lane, warp) for the prefix sums, which lets a lane sum the words of "its" eight threads without
// bank conflicts.
constexpr int kTpBins = 64, kTpGroups = kTpBins / 4;

__device__ __forceinline__ uint32_t tp_count(uint8_t *cnt, uint32_t bin)
{
    uint8_t *c = cnt + (bin >> 2) * (kBinBlock * 4) + threadIdx.x * 4 + (bin & 3u);
    const uint32_t before = *c;
    *c = static_cast<uint8_t>(before + 1);
    return before;
}

__global__ void __launch_bounds__(kBinBlock) bin_hist64_kernel(const uint32_t *__restrict__ ids, uint64_t n, int shift,
                                                              uint64_t n_blocks, uint64_t *__restrict__ matrix)
{
    __shared__ uint32_t s_cnt[kTpGroups][kBinBlock];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int g = 0; g < kTpGroups; ++g) s_cnt[g][threadIdx.x] = 0;
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kBinChunk;
    for (int it = 0; it < kBinIters; ++it) { // at most kBinIters * kBinPerThread = 64 ids per thread: a byte holds it
        const uint64_t it_base = base + static_cast<uint64_t>(it) * kBinPerIter;
        if (it_base >= n) break; // block-uniform
        uint32_t id[kBinPerThread], ok_mask;
        load_iter_ids(ids, n, it_base, id, ok_mask);
#pragma unroll
        for (int j = 0; j < kBinPerThread; ++j)
            if ((ok_mask >> j) & 1u) tp_count(reinterpret_cast<uint8_t *>(&s_cnt[0][0]), id[j] >> shift);
    }
    __syncthreads();
    // warp w sums groups 2w, 2w+1 over the 256 threads (two 16-bit fields per word: a bin total is <= 16384)
#pragma unroll
    for (int q = 0; q < kTpGroups / kBinWarps; ++q) {
        const int g = warp * (kTpGroups / kBinWarps) + q;
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int i = 0; i < kBinBlock / 32; ++i) {
            const uint32_t x = s_cnt[g][lane + 32 * i];
            lo += x & 0x00ff00ffu;
            hi += (x >> 8) & 0x00ff00ffu;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            lo += __shfl_xor_sync(0xffffffffu, lo, d);
            hi += __shfl_xor_sync(0xffffffffu, hi, d);
        }
        if (lane < 4) {
            const uint32_t v = (lane & 1) ? hi : lo;
            const uint32_t c = (lane & 2) ? (v >> 16) : (v & 0xffffu);
            matrix[static_cast<uint64_t>(4 * g + lane) * n_blocks + blockIdx.x] = c; // bins 4g + {0: lo.lo, 1: hi.lo, 2: lo.hi, 3: hi.hi}
        }
    }
}

struct Scatter64Smem {
    uint32_t cnt[kTpGroups][kBinBlock];      // byte counters, four bins per word
    uint32_t off[kTpGroups * 2][kBinBlock];  // ids of the bin held by the threads before this one: two 16-bit fields per
                                             // word, bin b in word [(b >> 2) * 2 + (b & 1)], field (b >> 1) & 1
    uint32_t ids[kBinPerIter];
    uint64_t dst[kTpBins];                   // binned index of staging slot 0 of each bin (wraps; only sums are used)
    uint64_t glob[kTpBins];                  // where the block's next id of each bin goes
    uint32_t start[kTpBins + 1];
    uint32_t tot[kTpBins];
};

__global__ void __launch_bounds__(kBinBlock, 3) bin_scatter64_kernel(const uint32_t *__restrict__ ids, uint64_t n, int shift,
                                                                    uint64_t n_blocks, const uint64_t *__restrict__ offs,
                                                                    uint32_t *__restrict__ binned)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Scatter64Smem &sm = *reinterpret_cast<Scatter64Smem *>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < kTpBins) sm.glob[threadIdx.x] = offs[static_cast<uint64_t>(threadIdx.x) * n_blocks + blockIdx.x];
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kBinChunk;
    for (int it = 0; it < kBinIters; ++it) {
        const uint64_t it_base = base + static_cast<uint64_t>(it) * kBinPerIter;
        if (it_base >= n) break; // block-uniform
        const bool full = it_base + kBinPerIter <= n; // block-uniform: all but the last iteration of the last block
#pragma unroll
        for (int g = 0; g < kTpGroups; ++g) sm.cnt[g][threadIdx.x] = 0;
        uint32_t id[kBinPerThread], ok_mask;
        uint32_t rank[kBinPerThread / 4]; // four byte ranks per word (a rank is < 16)
        load_iter_ids(ids, n, it_base, id, ok_mask);
        uint8_t *cnt = reinterpret_cast<uint8_t *>(&sm.cnt[0][0]);
        if (full) {
#pragma unroll
            for (int j = 0; j < kBinPerThread; ++j) {
                const uint32_t r = tp_count(cnt, id[j] >> shift) << (8 * (j % 4));
                rank[j / 4] = (j % 4 == 0) ? r : (rank[j / 4] | r);
            }
        } else {
#pragma unroll
            for (int j = 0; j < kBinPerThread; ++j) {
                if (j % 4 == 0) rank[j / 4] = 0;
                if ((ok_mask >> j) & 1u) rank[j / 4] |= tp_count(cnt, id[j] >> shift) << (8 * (j % 4));
            }
        }
        __syncthreads();
        // exclusive prefix of every bin's counters over the threads in (lane, warp) order
#pragma unroll
        for (int q = 0; q < kTpGroups / kBinWarps; ++q) {
            const int g = warp * (kTpGroups / kBinWarps) + q;
            uint32_t lo[kBinBlock / 32], hi[kBinBlock / 32], slo = 0, shi = 0;
#pragma unroll
            for (int i = 0; i < kBinBlock / 32; ++i) {
                const uint32_t x = sm.cnt[g][lane + 32 * i];
                lo[i] = slo;
                hi[i] = shi;
                slo += x & 0x00ff00ffu;
                shi += (x >> 8) & 0x00ff00ffu;
            }
            uint32_t ilo = slo, ihi = shi;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t a = __shfl_up_sync(0xffffffffu, ilo, d), b = __shfl_up_sync(0xffffffffu, ihi, d);
                if (lane >= d) {
                    ilo += a;
                    ihi += b;
                }
            }
            const uint32_t blo = ilo - slo, bhi = ihi - shi;
#pragma unroll
            for (int i = 0; i < kBinBlock / 32; ++i) {
                sm.off[2 * g + 0][lane + 32 * i] = blo + lo[i]; // bins 4g (low field) and 4g + 2 (high field)
                sm.off[2 * g + 1][lane + 32 * i] = bhi + hi[i]; // bins 4g + 1 and 4g + 3
            }
            if (lane == 31) {
                sm.tot[4 * g + 0] = ilo & 0xffffu;
                sm.tot[4 * g + 1] = ihi & 0xffffu;
                sm.tot[4 * g + 2] = ilo >> 16;
                sm.tot[4 * g + 3] = ihi >> 16;
            }
        }
        __syncthreads();
        if (warp == 0) { // exclusive scan of the 64 totals
            const uint32_t t0 = sm.tot[2 * lane], t1 = sm.tot[2 * lane + 1];
            uint32_t incl = t0 + t1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t a = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += a;
            }
            const uint32_t s0 = incl - t0 - t1, s1 = incl - t1;
            sm.start[2 * lane] = s0;
            sm.start[2 * lane + 1] = s1;
            sm.dst[2 * lane] = sm.glob[2 * lane] - s0;
            sm.dst[2 * lane + 1] = sm.glob[2 * lane + 1] - s1;
            sm.glob[2 * lane] += t0;
            sm.glob[2 * lane + 1] += t1;
            if (lane == 31) sm.start[kTpBins] = incl;
        }
        __syncthreads();
        const uint32_t *off_col = &sm.off[0][threadIdx.x];
        auto place = [&](int j) {
            const uint32_t bin = id[j] >> shift;
            const uint32_t w = off_col[(((bin >> 1) & ~1u) | (bin & 1u)) * kBinBlock];
            const uint32_t o = (w >> ((bin & 2u) << 3)) & 0xffffu;
            sm.ids[sm.start[bin] + o + ((rank[j / 4] >> (8 * (j % 4))) & 0xffu)] = id[j];
        };
        if (full) {
#pragma unroll
            for (int j = 0; j < kBinPerThread; ++j) place(j);
        } else {
#pragma unroll
            for (int j = 0; j < kBinPerThread; ++j)
                if ((ok_mask >> j) & 1u) place(j);
        }
        __syncthreads();
        const uint32_t total = sm.start[kTpBins];
        if (full) {
#pragma unroll 4
            for (uint32_t t = threadIdx.x; t < kBinPerIter; t += kBinBlock) {
                const uint32_t v = sm.ids[t];
                binned[sm.dst[v >> shift] + t] = v;
            }
        } else {
            for (uint32_t t = threadIdx.x; t < total; t += kBinBlock) {
                const uint32_t v = sm.ids[t];
                binned[sm.dst[v >> shift] + t] = v;
            }
        }
        // the next iteration's writes to cnt / off / start / dst are ordered behind these reads by its barriers;
        // its writes to ids come after its second barrier
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
    if (bucket_bits < p) return cudaErrorInvalidValue; // the caller bins only tables far larger than 2^p counters
    const int n_bins = 1 << p, shift = bucket_bits - p;
    const uint64_t n_blocks = binned_count_blocks(n);
    auto run = [&](auto tag) -> cudaError_t {
        constexpr int P = decltype(tag)::value;
        bin_hist_kernel<P><<<static_cast<unsigned>(n_blocks), kBinBlock, 0, stream>>>(ids, n, shift, n_blocks, matrix);
        cudaError_t e = inclusive_offsets_u64(matrix, offs, static_cast<uint64_t>(n_bins) * n_blocks, scan_tmp, stream);
        if (e != cudaSuccess) return e;
        bin_scatter_kernel<P><<<static_cast<unsigned>(n_blocks), kBinBlock, 0, stream>>>(ids, n, shift, n_blocks, offs, binned);
        return cudaGetLastError();
    };
    auto run64 = [&]() -> cudaError_t {
        static_assert(kBinIters * kBinPerThread < 256, "byte counters");
        cudaError_t e = cudaFuncSetAttribute(bin_scatter64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(sizeof(Scatter64Smem)));
        if (e != cudaSuccess) return e;
        bin_hist64_kernel<<<static_cast<unsigned>(n_blocks), kBinBlock, 0, stream>>>(ids, n, shift, n_blocks, matrix);
        e = inclusive_offsets_u64(matrix, offs, static_cast<uint64_t>(n_bins) * n_blocks, scan_tmp, stream);
        if (e != cudaSuccess) return e;
        bin_scatter64_kernel<<<static_cast<unsigned>(n_blocks), kBinBlock, sizeof(Scatter64Smem), stream>>>(ids, n, shift, n_blocks,
                                                                                                           offs, binned);
        return cudaGetLastError();
    };
    cudaError_t e;
    switch (p) {
    case 6: e = run64(); break;
    case 7: e = run(std::integral_constant<int, 7>()); break;
    case 8: e = run(std::integral_constant<int, 8>()); break;
    case 9: e = run(std::integral_constant<int, 9>()); break;
    case 10: e = run(std::integral_constant<int, 10>()); break;
    default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
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
    if (p < kMinBinBits) p = kMinBinBits;
    if (p > kMaxBinBits) p = kMaxBinBits;
    return p;
}

uint64_t binned_count_blocks(uint64_t n) { return (n + kBinChunk - 1) / kBinChunk; }

} // namespace kmc
