// buckets.cu -- the canonical k-mer bucket-count table (north_star extension; not in the reference):
// table[fx_hash(canonical k-mer) >> (64 - B)] += 1 over a read set.
//
// Two regimes, both bound by how fast an SM can issue scattered 4-byte increments (1.29 cycles per lane:
// 2.2e11 / s measured; shared-memory atomics are no faster, so there is nothing to privatise):
//   * the table fits L2 (<= 96 MB): extract_kernel<SINK_BUCKETS> increments it directly;
//   * larger tables (B = 28 is 1 GiB): random increments miss L2 and serialise at DRAM latency (measured
//     24 G k-mers/s).  So the bucket ids are first written out (SINK_IDS), partitioned by their high bits into
//     bins whose table slice is <= 16 MB (binning.cuh: no atomics, exact placement from a count matrix), and
//     then applied bin after bin, so that all SMs work on an L2-resident part of the table at a time.  The
//     table is final range by range, which kmc_bucket_count_async reports through events so that the merge of
//     several GPUs' tables can overlap the count.
#include <cstdlib>
#include <type_traits>

#include "binning.cuh"
#include "plan.h"

namespace kmc {

namespace {

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
        // evict_last: the slice is to stay in L2 while its bins are applied, whatever streams through beside it
        asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(p + i), "l"(0x14F0000000000000ull));
        acc |= v.x & v.y & v.z & v.w;
    }
    if (acc == 0xffffffffu) *sink = acc; // keeps the loads alive; a count of 2^32-1 in four neighbours does not happen
}

// the ids of bins [bin, bin_end) are binned[offs[bin * n_blocks] .. offs[bin_end * n_blocks])
__global__ void __launch_bounds__(256) bin_apply_kernel(const uint32_t *__restrict__ binned, const uint64_t *__restrict__ offs,
                                                        uint64_t n_blocks, int bin, int bin_end, uint32_t *__restrict__ table)
{
    const uint64_t begin = __ldg(offs + static_cast<uint64_t>(bin) * n_blocks);
    const uint64_t end = __ldg(offs + static_cast<uint64_t>(bin_end) * n_blocks);
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


// ---------------------------------------------------------------------------------------------
// The fused first half of the binned count for the common case -- one-limb k-mers (K <= 32) over an aligned uniform
// read set (extract_kernels.cuh: AlignedItems; C5 is 150 bp reads, K = 31).  The three passes of the exact path
// (bucket ids written out, per-chunk bin histogram, scatter) read and write 12 bytes per k-mer three times over and
// need the histogram only to know where every chunk's run of every bin goes.  Counting does not care about the order
// inside a bin, so here ONE kernel produces the ids and bins them: a block computes the canonical k-mers and hashes
// of 4096 windows per iteration (16 per thread, in registers), runs the same shared-memory binning step
// (binning.cuh: scatter64_iter) and reserves the place of each of its 64 runs with one atomicAdd on the bin's
// cursor.  The bins have a fixed capacity, mean + 12.5 % + 8192 ids: fx_hash spreads distinct k-mers evenly.  A set in which
// one bucket range attracts much more than its share (reads of one repeated k-mer, say) fills a bin; a run that no longer
// fits goes to the SPILL list instead (one more atomicAdd; the list can hold every id of the call, so nothing is ever
// dropped and there is nothing to fall back to), and the bin ends where that run would have begun.  The spilled ids are
// applied to the table directly, before the bins -- they are few, or they are many and hit the same few counters.
// 4.2 + 1.9 + 7.9 ms of the 30.6 ms per 3 G k-mers become one kernel of 9.7 ms, and the call stays asynchronous.
// ---------------------------------------------------------------------------------------------
constexpr int kStateBinEnd = binning::kTpBins;      // state[64 .. 128): where a full bin ends (~0: it never filled up)
constexpr int kStateSpill = 2 * binning::kTpBins;   // state[128]: ids in the spill list
constexpr int kStateWords = 2 * binning::kTpBins + 2;

struct FusedBins {
    uint32_t *binned;             // 64 bins of `cap` ids each, then the spill list
    uint64_t cap;                 // a multiple of 4 (the apply kernel loads 16 bytes at a time)
    unsigned long long *state;    // [64] ids reserved per bin, [64] bin ends, [1] spill cursor
    int bin_shift;                // bin = id >> bin_shift
};

template <int NX>
__global__ void __launch_bounds__(binning::kBlock, 3) bucket_bin_kernel(const ExtractParams p, const FusedBins f)
{
    using S = binning::Shape<uint32_t>;
    constexpr int G = 8, PT = S::kPerThread;
    static_assert(PT == 2 * G && S::kChunk == kTileItems * G, "a thread bins the ids of two work items per iteration");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    binning::Scatter64Smem<uint32_t> &sm = *reinterpret_cast<binning::Scatter64Smem<uint32_t> *>(smem_raw);
    const uint32_t n_items = static_cast<uint32_t>(p.items);
    const uint32_t tile_base = blockIdx.x * static_cast<uint32_t>(kTileItems);
    if (p.pf_tiles) burst_prefetch<false, G, 2>(p, n_items);
    const AlignedItems<G, 2> items(p);
    const uint32_t tile_end = n_items - tile_base > static_cast<uint32_t>(kTileItems) ? tile_base + kTileItems : n_items;
    const bool inside = items.template loads_inside<NX>(p, tile_end - 1u);
    const binning::IdBin bin_of{f.bin_shift};
    constexpr int kItemsPerIter = 2 * binning::kBlock;
    // where the run of `count` ids of bin b goes
    auto reserve = [&](int b, uint32_t count) -> uint64_t {
        if (count == 0) return 0;
        const uint64_t at = atomicAdd(f.state + b, static_cast<unsigned long long>(count));
        if (at + count > f.cap) {
            // the bin is full: it ends where this run would have begun (every run reserved before it fits, every later
            // one starts beyond the capacity too), and the run goes to the spill list
            atomicMin(f.state + kStateBinEnd + b, static_cast<unsigned long long>(at));
            return binning::kTpBins * f.cap + atomicAdd(f.state + kStateSpill, static_cast<unsigned long long>(count));
        }
        return static_cast<uint64_t>(b) * f.cap + at;
    };
    // the bucket ids of one work item (its G windows)
    auto ids_of = [&](uint32_t item, uint32_t *key) {
        uint32_t x[NX];
        load_block_at<NX>(p, items.bit_of(item), inside, x);
        uint64_t fw[G][1], rv[G][1];
        block_kmers<1, NX, G, true, true, 2>(x, p.s0, p.head_mask, fw, rv);
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const uint64_t c[1] = {fw[j][0] < rv[j][0] ? fw[j][0] : rv[j][0]};
            key[j] = static_cast<uint32_t>(fx_hash<1>(c, 0) >> p.bucket_shift);
        }
    };
#pragma unroll 1
    for (uint32_t it_base = tile_base; it_base < tile_end; it_base += kItemsPerIter) {
        uint32_t key[PT];
        if (tile_end - it_base >= static_cast<uint32_t>(kItemsPerIter)) { // block-uniform: every iteration but the set's last
            ids_of(it_base + threadIdx.x, key);
            ids_of(it_base + binning::kBlock + threadIdx.x, key + G);
            binning::scatter64_iter<uint32_t, true>(sm, key, 0xffffu, bin_of, reserve, f.binned);
        } else {
            uint32_t ok_mask = 0;
#pragma unroll
            for (int j = 0; j < PT; ++j) key[j] = 0;
            if (it_base + threadIdx.x < tile_end) {
                ids_of(it_base + threadIdx.x, key);
                ok_mask = 0xffu;
            }
            if (it_base + binning::kBlock + threadIdx.x < tile_end) {
                ids_of(it_base + binning::kBlock + threadIdx.x, key + G);
                ok_mask |= 0xff00u;
            }
            binning::scatter64_iter<uint32_t, false>(sm, key, ok_mask, bin_of, reserve, f.binned);
        }
    }
}

// The ids of bin b are binned[b * cap .. b * cap + min(reserved, bin end)).  The next 16 bytes of ids are loaded before the
// current four are applied.
__device__ __forceinline__ uint4 ld_ids(const uint32_t *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p), "l"(0x12F0000000000000ull));
    return v;
}
__device__ __forceinline__ void inc_counter(uint32_t *p)
{
    asm volatile("red.global.add.L2::cache_hint.u32 [%0], 1, %1;" ::"l"(p), "l"(0x14F0000000000000ull) : "memory");
}

__global__ void __launch_bounds__(256) bin_apply_fused_kernel(const uint32_t *__restrict__ binned, uint64_t cap,
                                                              const unsigned long long *__restrict__ state, int bin, int bin_end,
                                                              uint32_t *__restrict__ table)
{
    const uint64_t tid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t step = static_cast<uint64_t>(gridDim.x) * blockDim.x * 4;
    for (int b = bin; b < bin_end; ++b) {
        const uint32_t *ids = binned + static_cast<uint64_t>(b) * cap; // 16-byte aligned: cap is a multiple of 4
        const uint64_t reserved = state[b], full_at = state[kStateBinEnd + b];
        const uint64_t n = reserved < full_at ? reserved : full_at, n4 = n & ~3ull;
        uint64_t i = tid * 4;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (i < n4) v = ld_ids(ids + i);
        while (i < n4) {
            const uint64_t nxt = i + step;
            uint4 w = make_uint4(0, 0, 0, 0);
            if (nxt < n4) w = ld_ids(ids + nxt);
            inc_counter(table + v.x);
            inc_counter(table + v.y);
            inc_counter(table + v.z);
            inc_counter(table + v.w);
            v = w;
            i = nxt;
        }
        if (tid < n - n4) inc_counter(table + ids[n4 + tid]);
    }
}

// the spill list: ids whose bin was full, applied to the table as they come (no slice of it is resident for them)
__global__ void __launch_bounds__(256) spill_apply_kernel(const uint32_t *__restrict__ spill, const unsigned long long *__restrict__ state,
                                                          uint32_t *__restrict__ table)
{
    const uint64_t n = state[kStateSpill];
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
        atomicAdd(table + spill[i], 1u);
}

} // namespace

// ids[0..n) are bucket ids of `bucket_bits` bits; adds their histogram to table.  tmp: n u32 (binned ids),
// matrix / offs: (n_bins * n_blocks + 1) u64 each, scan_tmp: scan_tmp_elems(n_bins * n_blocks) u64.
cudaError_t binned_count(const uint32_t *ids, uint64_t n, int bucket_bits, uint32_t *table, uint32_t *binned, uint64_t *matrix,
                         uint64_t *offs, uint64_t *scan_tmp, int sm_count, cudaStream_t stream, uint32_t n_parts, void *const *events)
{
    if (!events) n_parts = 0;
    if (n == 0) {
        for (uint32_t i = 0; i < n_parts; ++i) {
            cudaError_t e = cudaEventRecord(static_cast<cudaEvent_t>(events[i]), stream);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    const int p = binned_count_bin_bits(bucket_bits);
    if (bucket_bits < p) return cudaErrorInvalidValue; // the caller bins only tables far larger than 2^p counters
    const int n_bins = 1 << p, shift = bucket_bits - p;
    const uint64_t n_blocks = binned_count_blocks(n);
    cudaError_t e = binning::partition<uint32_t>(ids, n, p, binning::IdBin{shift}, binned, matrix, offs, scan_tmp, stream);
    if (e != cudaSuccess) return e;
    // offs has n_bins * n_blocks + 1 entries: offs[(bin + 1) * n_blocks] of the last bin is the total
    const uint64_t slice = 1ull << shift; // counters per bin
    const int group = apply_group(2); // bins applied per launch
    uint32_t parts_done = 0;
    for (int b = 0; b < n_bins; b += group) {
        const int b_end = b + group < n_bins ? b + group : n_bins;
        warm_slice_kernel<<<static_cast<unsigned>(sm_count * 8), 256, 0, stream>>>(table + static_cast<uint64_t>(b) * slice,
                                                                                  slice * (b_end - b), reinterpret_cast<uint32_t *>(scan_tmp));
        bin_apply_kernel<<<static_cast<unsigned>(sm_count * 16), 256, 0, stream>>>(binned, offs, n_blocks, b, b_end, table);
        // range i of the table = bins [i * n_bins / n_parts, (i + 1) * n_bins / n_parts): final once they are applied
        while (parts_done < n_parts && static_cast<uint64_t>(parts_done + 1) * n_bins <= static_cast<uint64_t>(b_end) * n_parts) {
            e = cudaEventRecord(static_cast<cudaEvent_t>(events[parts_done++]), stream);
            if (e != cudaSuccess) return e;
        }
    }
    return cudaGetLastError();
}

uint64_t fused_bin_capacity(uint64_t n_ids)
{
    const uint64_t cap = n_ids / binning::kTpBins + n_ids / (8 * binning::kTpBins) + 8192;
    return (cap + 3) & ~3ull;
}
// ids the binned buffer holds: the 64 bins and a spill list long enough for every id of the call
uint64_t fused_bin_buffer_ids(uint64_t n_ids) { return binning::kTpBins * fused_bin_capacity(n_ids) + n_ids + 4096; }
uint64_t fused_bin_state_bytes() { return kStateWords * sizeof(unsigned long long); }

// First half: one kernel bins the bucket ids of every window.  p: the parameter block of an aligned uniform set of
// one-limb k-mers with whole groups (the caller has checked that), bucket_shift set.
cudaError_t fused_bin_ids(ExtractParams p, int nx, int bucket_bits, uint32_t *binned, uint64_t cap, unsigned long long *state,
                          cudaStream_t stream)
{
    const int pbits = binned_count_bin_bits(bucket_bits);
    if (pbits != 6 || bucket_bits < pbits) return cudaErrorInvalidValue;
    const uint64_t tiles = (p.items + kTileItems - 1) / kTileItems;
    if (tiles == 0 || tiles > 0x7fffffffull || p.items >= 0xffffffffull - kTileItems) return cudaErrorInvalidConfiguration;
    set_iteration_strides(p, 8);
    if (!p.aligned || p.al_tail) return cudaErrorInvalidValue; // whole groups only
    p.al_magic = aligned_magic(p.gprm);
    p.pf_tiles = 0;
    if (prefetch_enabled()) {
        const uint64_t per_tile = static_cast<uint64_t>(p.nw32) * 4 / tiles + 1, t = kPfChunkBytes / per_tile;
        p.pf_tiles = static_cast<uint32_t>(t < 2 * kPfLead ? 2 * kPfLead : (t > (1u << 20) ? (1u << 20) : t));
    }
    cudaError_t e = cudaMemsetAsync(state, 0, kStateWords * sizeof(unsigned long long), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(state + kStateBinEnd, 0xff, binning::kTpBins * sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    const FusedBins f{binned, cap, state, bucket_bits - pbits};
    constexpr int smem = static_cast<int>(sizeof(binning::Scatter64Smem<uint32_t>));
    auto launch = [&](auto tag) -> cudaError_t {
        constexpr int NX = decltype(tag)::value;
        cudaError_t e2 = cudaFuncSetAttribute(bucket_bin_kernel<NX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e2 != cudaSuccess) return e2;
        bucket_bin_kernel<NX><<<static_cast<unsigned>(tiles), binning::kBlock, smem, stream>>>(p, f);
        return cudaGetLastError();
    };
    switch (nx) {
    case 1: return launch(std::integral_constant<int, 1>());
    case 2: return launch(std::integral_constant<int, 2>());
    case 3: return launch(std::integral_constant<int, 3>());
    }
    return cudaErrorInvalidValue;
}

// Second half: the spill list, then the bins in table order, as in binned_count.
cudaError_t fused_bin_apply(const uint32_t *binned, uint64_t cap, const unsigned long long *state, int bucket_bits, uint32_t *table,
                            uint32_t *sink, int sm_count, cudaStream_t stream, uint32_t n_parts, void *const *events)
{
    if (!events) n_parts = 0;
    const int pbits = binned_count_bin_bits(bucket_bits);
    const int n_bins = 1 << pbits, shift = bucket_bits - pbits;
    const uint64_t slice = 1ull << shift;
    const int group = apply_group(2);
    spill_apply_kernel<<<static_cast<unsigned>(sm_count * 8), 256, 0, stream>>>(binned + static_cast<uint64_t>(n_bins) * cap, state, table);
    uint32_t parts_done = 0;
    for (int b = 0; b < n_bins; b += group) {
        const int b_end = b + group < n_bins ? b + group : n_bins;
        warm_slice_kernel<<<static_cast<unsigned>(sm_count * 8), 256, 0, stream>>>(table + static_cast<uint64_t>(b) * slice,
                                                                                  slice * (b_end - b), sink);
        bin_apply_fused_kernel<<<static_cast<unsigned>(sm_count * 16), 256, 0, stream>>>(binned, cap, state, b, b_end, table);
        while (parts_done < n_parts && static_cast<uint64_t>(parts_done + 1) * n_bins <= static_cast<uint64_t>(b_end) * n_parts) {
            cudaError_t e = cudaEventRecord(static_cast<cudaEvent_t>(events[parts_done++]), stream);
            if (e != cudaSuccess) return e;
        }
    }
    return cudaGetLastError();
}

// KMC_FUSED_BIN=0 keeps every binned count on the exact three-pass path (A/B measurements, tests of both)
bool fused_bin_enabled()
{
    static const bool on = [] {
        const char *e = getenv("KMC_FUSED_BIN");
        return !(e && e[0] == '0');
    }();
    return on;
}

// Bins applied per launch.  Measured on a B200 (KMC_APPLY_GROUP overrides, for experiments): the bucket table
// (16 MB slices, increments) is fastest with 2 bins = 32 MB per launch (1: +1 %, 4: +6 %, 8 = 128 MB: 2x slower --
// the slices no longer fit L2); the k-mer table (12 MB slices, a load and an increment per k-mer) keeps improving up to
// 8 bins = 96 MB (1: 10.7 ms, 2: 9.8, 4: 9.6, 8: 9.3 ms per 480 M k-mers); 6 leaves room for the k-mer stream itself.
int apply_group(int dflt)
{
    static const int env = [] {
        const char *e = getenv("KMC_APPLY_GROUP");
        const int v = e ? atoi(e) : 0;
        return v < 0 ? 0 : (v > 64 ? 64 : v);
    }();
    return env ? env : dflt;
}

cudaError_t warm_table(const uint32_t *table, uint64_t n_counters, int sm_count, uint32_t *sink, cudaStream_t stream)
{
    warm_slice_kernel<<<static_cast<unsigned>(sm_count * 8), 256, 0, stream>>>(table, n_counters, sink);
    return cudaGetLastError();
}

int binned_count_bin_bits(int bucket_bits)
{
    int p = bucket_bits - 22; // table slice per bin <= 2^22 counters = 16 MB
    if (p < binning::kMinBits) p = binning::kMinBits;
    if (p > binning::kMaxBits) p = binning::kMaxBits;
    return p;
}

uint64_t binned_count_blocks(uint64_t n) { return binning::blocks_for<uint32_t>(n); }

} // namespace kmc
