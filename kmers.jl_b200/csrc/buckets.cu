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
        asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));
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
