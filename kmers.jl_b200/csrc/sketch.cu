// sketch.cu -- consumers of the k-mer stream that never write it:
//
//   kmc_minhash_sketch  bottom-s MinHash sketch under fx_hash -- the reference's example
//                       `sketch(fx_hash, CanonicalDNAMers{16}(seq), 1000)` (docs/src/minhash.md:31-36;
//                       MinHash.jl is not vendored: the published bottom-s definition is restated --
//                       the s smallest DISTINCT hash values over all k-mers, ascending)
//   kmc_composition     k-mer composition vector -- `counts[as_integer(kmer) + 1] += 1` over
//                       FwDNAMers{K} (docs/src/composition.md:28-39), K <= 14
//
// Both run consume_kernel: the same (read, group slot) work items, block load and static funnel
// shifts as extract_kernel (kmer_core.cuh), with the stores replaced by a small functor.
//
// Sketch: (1) histogram of the top 12 hash bits (4096 counters: every block counts in shared memory
// and adds its non-zero counters to the global table once), (2) the host picks the smallest
// threshold bucket t whose cumulative count reaches s (control logic over 4096 integers), (3) a second pass appends every hash whose top bits are <= t to a candidate array
// sized exactly by the histogram, (4) the candidates -- s plus at most one bucket's worth -- are
// sorted and deduplicated on the device (cub radix sort + unique: library plumbing on a tiny
// array), and if duplicates left fewer than s distinct values the threshold moves up and (3)-(4)
// repeat.  The k-mers are read twice (2 x 0.25 B/symbol) and nothing proportional to them is written.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include <vector>

#include "plan.h"

namespace kmc {

namespace {

enum : int { OP_HASH_HIST = 0, OP_HASH_BELOW = 1, OP_COMPOSITION = 2, OP_COMPOSITION_SHARED = 3 };

struct ConsumeParams {
    uint32_t *table;               // OP_HASH_HIST: u32[2^12]; OP_COMPOSITION*: u32[4^K]
    uint32_t shift;                // OP_HASH_HIST / OP_HASH_BELOW: 64 - bits
    uint64_t limit;                // OP_HASH_BELOW: keep h <= limit
    uint64_t *cand;                // OP_HASH_BELOW: candidate array
    unsigned long long *cursor;    // OP_HASH_BELOW: elements appended so far
    uint64_t cand_cap;
    uint32_t table_entries;        // OP_HASH_HIST, OP_COMPOSITION_SHARED: counters kept in shared memory
};

constexpr int kSharedCompositionMax = 4096; // 4^6 counters = 16 KB of shared memory
constexpr int kSketchBits = 12;             // top hash bits of the sketch histogram (also kept in shared memory)
static_assert((1 << kSketchBits) <= kSharedCompositionMax, "the sketch histogram lives in the same shared array");

template <int N, int NX, bool RAGGED, bool CANON, int OP>
__global__ void __launch_bounds__(kBlockThreads) consume_kernel(const ExtractParams p, const ConsumeParams c)
{
    constexpr int G = GroupOf<N>::G;
    __shared__ TileShared<RAGGED> sh;
    constexpr bool SHARED_HIST = (OP == OP_COMPOSITION_SHARED || OP == OP_HASH_HIST);
    __shared__ uint32_t s_hist[SHARED_HIST ? kSharedCompositionMax : 1];
    const uint64_t tile_base = static_cast<uint64_t>(blockIdx.x) * kTileItems;
    if (SHARED_HIST) {
        for (uint32_t i = threadIdx.x; i < c.table_entries; i += kBlockThreads) s_hist[i] = 0;
    }
    TileCursor<RAGGED, G> cur;
    cur.init(p, tile_base, sh, threadIdx.x); // (block-wide barriers inside: also orders the zeroing above)
    if (SHARED_HIST) __syncthreads();

#pragma unroll 1
    for (int it = 0; it < kTileIters; ++it) {
        const uint32_t li = static_cast<uint32_t>(it) * kBlockThreads + threadIdx.x;
        const uint64_t item = tile_base + li;
        if (item >= p.items) break;
        cur.locate(p, item, li, sh);
        const int jlo = cur.jlo, jhi = cur.jhi;
        if (jhi > jlo) {
            uint32_t x[NX];
            load_block<NX>(p.w32, p.nw32, cur.bit(p), x);
            uint64_t fw[G][N], rv[G][N];
            block_kmers<N, NX, G, true, CANON>(x, p.s0, p.head_mask, fw, rv);
#pragma unroll
            for (int j = 0; j < G; ++j) {
                if (j < jlo || j >= jhi) continue;
                bool take_fw = true;
                if (CANON) take_fw = limbs_less<N>(fw[j], rv[j]);
                uint64_t a[N];
#pragma unroll
                for (int i = 0; i < N; ++i) a[i] = take_fw ? fw[j][i] : rv[j][i];
                if (OP == OP_HASH_HIST) {
                    atomicAdd(s_hist + (fx_hash<N>(a, 0) >> c.shift), 1u);
                } else if (OP == OP_HASH_BELOW) {
                    const uint64_t h = fx_hash<N>(a, 0);
                    if (h <= c.limit) {
                        const unsigned long long at = atomicAdd(c.cursor, 1ull);
                        if (at < c.cand_cap) c.cand[at] = h;
                    }
                } else if (OP == OP_COMPOSITION) {
                    atomicAdd(c.table + a[N - 1], 1u); // as_integer(kmer): K <= 14, one limb
                } else {
                    atomicAdd(s_hist + a[N - 1], 1u);
                }
            }
        }
        cur.advance(p);
    }
    if (SHARED_HIST) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < c.table_entries; i += kBlockThreads) {
            const uint32_t v = s_hist[i];
            if (v) atomicAdd(c.table + i, v);
        }
    }
}

using ConsumeLaunchFn = cudaError_t (*)(ExtractParams, ConsumeParams, cudaStream_t);

template <int N, int NX, bool RAGGED, bool CANON, int OP>
cudaError_t launch_consume(ExtractParams p, ConsumeParams c, cudaStream_t stream)
{
    const uint64_t tiles = (p.items + kTileItems - 1) / kTileItems;
    if (tiles == 0) return cudaSuccess;
    if (tiles > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    set_iteration_strides(p);
    consume_kernel<N, NX, RAGGED, CANON, OP><<<static_cast<unsigned>(tiles), kBlockThreads, 0, stream>>>(p, c);
    return cudaGetLastError();
}

template <int N, int NX, int OP>
ConsumeLaunchFn pick_consume(bool ragged, bool canon)
{
    if (ragged) return canon ? &launch_consume<N, NX, true, true, OP> : &launch_consume<N, NX, true, false, OP>;
    return canon ? &launch_consume<N, NX, false, true, OP> : &launch_consume<N, NX, false, false, OP>;
}

template <int N, int OP>
ConsumeLaunchFn pick_consume_nx(int nx, bool ragged, bool canon)
{
    constexpr int NXMAX = (64 * N + 2 * GroupOf<N>::G - 2 + 31) / 32;
    if (nx == NXMAX) return pick_consume<N, NXMAX, OP>(ragged, canon);
    if (nx == NXMAX - 1) return pick_consume<N, (NXMAX - 1 > 0 ? NXMAX - 1 : 1), OP>(ragged, canon);
    if (nx == NXMAX - 2) return pick_consume<N, (NXMAX - 2 > 0 ? NXMAX - 2 : 1), OP>(ragged, canon);
    return nullptr;
}

template <int OP>
ConsumeLaunchFn consume_launcher(const Geometry &ge, bool ragged, bool canon)
{
    switch (ge.n_limbs) {
    case 1: return pick_consume_nx<1, OP>(ge.nx, ragged, canon);
    case 2: return pick_consume_nx<2, OP>(ge.nx, ragged, canon);
    }
    return nullptr;
}

struct AsyncBuf { // stream-ordered temporary
    void *p = nullptr;
    cudaStream_t s = nullptr;
    cudaError_t alloc(uint64_t bytes, cudaStream_t stream)
    {
        s = stream;
        return cudaMallocAsync(&p, bytes ? bytes : 1, stream);
    }
    ~AsyncBuf()
    {
        if (p) cudaFreeAsync(p, s);
    }
};

} // namespace

} // namespace kmc

using namespace kmc;

extern "C" int32_t kmc_minhash_sketch(kmc_ctx *ctx, const kmc_seqs *seqs, int32_t k, int32_t mode, uint64_t s,
                                      uint64_t *out_hashes, kmc_result *result)
{
    int32_t st = check_common(ctx, seqs, k);
    if (st) return st;
    if (!out_hashes || !result) return fail(ctx, KMC_E_BAD_ARG, "out_hashes / result is NULL");
    if (mode != KMC_FW && mode != KMC_CANON) return fail(ctx, KMC_E_BAD_ARG, "mode must be KMC_FW or KMC_CANON");
    if (seqs->src_bits != 2) return fail(ctx, KMC_E_UNSUPPORTED, "the sketch needs a 2-bit source");
    if (k > 64) return fail(ctx, KMC_E_UNSUPPORTED, "the sketch supports K <= 64");
    if (s < 1) return fail(ctx, KMC_E_BAD_ARG, "the sketch size must be at least 1");
    CU(cudaSetDevice(ctx->device));
    memset(result, 0, sizeof *result);
    const Geometry ge = geometry(k);
    const bool canon = mode == KMC_CANON;
    cudaStream_t stream = ctx->stream;
    CU(cudaEventRecord(ctx->ev_k0, stream));
    constexpr int kBits = kSketchBits;
    constexpr uint64_t kBuckets = 1ull << kBits;
    st = ensure_scratch(ctx, layout_scratch_bytes(seqs) + kBuckets * 4 + 1024);
    if (st) return st;
    st = ensure_host_small(ctx);
    if (st) return st;
    Scratch scratch{static_cast<char *>(ctx->scratch), ctx->scratch_bytes, 0};
    uint32_t *table = static_cast<uint32_t *>(scratch.take(kBuckets * 4));
    unsigned long long *cursor = static_cast<unsigned long long *>(scratch.take(16));
    Layout L;
    st = plan_layout(ctx, seqs, k, ge, stream, KnownTotals(), scratch, &L);
    if (st) return st;
    if (L.total == 0) return KMC_OK;
    ExtractParams p = base_params(seqs, k, ge, L, 0);
    const bool ragged = !L.uniform_len;

    // (1) histogram of the top hash bits
    ConsumeParams c{};
    c.table = table;
    c.shift = 64 - kBits;
    c.table_entries = static_cast<uint32_t>(kBuckets);
    CU(cudaMemsetAsync(table, 0, kBuckets * 4, stream));
    ConsumeLaunchFn hist = consume_launcher<OP_HASH_HIST>(ge, ragged, canon);
    ConsumeLaunchFn below = consume_launcher<OP_HASH_BELOW>(ge, ragged, canon);
    if (!hist || !below) return fail(ctx, KMC_E_UNSUPPORTED, "no kernel for this K");
    CU(hist(p, c, stream));
    std::vector<uint32_t> counts(kBuckets);
    CU(cudaMemcpyAsync(counts.data(), table, kBuckets * 4, cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));

    // (2)-(4) threshold bucket, candidates, sort + unique; repeat with a higher threshold while
    // duplicates leave fewer than s distinct values
    uint64_t want = s; // candidates to collect (counting duplicates)
    for (;;) {
        uint64_t cum = 0;
        uint64_t t = 0;
        for (; t < kBuckets; ++t) {
            cum += counts[t];
            if (cum >= want) break;
        }
        const bool all = t >= kBuckets - 1;
        if (t >= kBuckets) t = kBuckets - 1;
        AsyncBuf cand, sorted, uniq, tmp, nsel;
        CU(cand.alloc(cum * 8, stream));
        CU(sorted.alloc(cum * 8, stream));
        CU(uniq.alloc(cum * 8, stream));
        CU(nsel.alloc(8, stream));
        c.limit = all ? ~0ull : (((t + 1) << (64 - kBits)) - 1);
        c.cand = static_cast<uint64_t *>(cand.p);
        c.cursor = cursor;
        c.cand_cap = cum;
        CU(cudaMemsetAsync(cursor, 0, 8, stream));
        CU(below(p, c, stream));
        size_t b1 = 0, b2 = 0;
        CU(cub::DeviceRadixSort::SortKeys(nullptr, b1, static_cast<uint64_t *>(cand.p), static_cast<uint64_t *>(sorted.p), cum, 0,
                                          64, stream));
        CU(cub::DeviceSelect::Unique(nullptr, b2, static_cast<uint64_t *>(sorted.p), static_cast<uint64_t *>(uniq.p),
                                     static_cast<uint64_t *>(nsel.p), cum, stream));
        CU(tmp.alloc(b1 > b2 ? b1 : b2, stream));
        CU(cub::DeviceRadixSort::SortKeys(tmp.p, b1, static_cast<uint64_t *>(cand.p), static_cast<uint64_t *>(sorted.p), cum, 0, 64,
                                          stream));
        CU(cub::DeviceSelect::Unique(tmp.p, b2, static_cast<uint64_t *>(sorted.p), static_cast<uint64_t *>(uniq.p),
                                     static_cast<uint64_t *>(nsel.p), cum, stream));
        CU(cudaMemcpyAsync(&ctx->host_small[100], nsel.p, 8, cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(&ctx->host_small[101], cursor, 8, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        const uint64_t distinct = ctx->host_small[100];
        if (ctx->host_small[101] != cum) return fail(ctx, KMC_E_BAD_ARG, "internal: the two sketch passes disagree");
        if (distinct >= s || all) {
            const uint64_t n = distinct < s ? distinct : s;
            CU(cudaMemcpyAsync(out_hashes, uniq.p, n * 8, cudaMemcpyDeviceToDevice, stream));
            result->n_written = n;
            break;
        }
        want = cum + 2 * (s - distinct) + 1; // at least one more bucket
    }
    CU(cudaEventRecord(ctx->ev_k1, stream));
    CU(cudaStreamSynchronize(stream));
    CU(cudaEventElapsedTime(&result->kernel_ms, ctx->ev_k0, ctx->ev_k1));
    return KMC_OK;
}

extern "C" int32_t kmc_composition(kmc_ctx *ctx, const kmc_seqs *seqs, int32_t k, int32_t mode, uint32_t *table,
                                   kmc_result *result)
{
    int32_t st = check_common(ctx, seqs, k);
    if (st) return st;
    if (!table || !result) return fail(ctx, KMC_E_BAD_ARG, "table / result is NULL");
    if (mode != KMC_FW && mode != KMC_CANON) return fail(ctx, KMC_E_BAD_ARG, "mode must be KMC_FW or KMC_CANON");
    if (seqs->src_bits != 2) return fail(ctx, KMC_E_UNSUPPORTED, "the composition needs a 2-bit source");
    if (k > 14) return fail(ctx, KMC_E_UNSUPPORTED, "the dense composition table supports K <= 14 (4^K counters)");
    CU(cudaSetDevice(ctx->device));
    memset(result, 0, sizeof *result);
    const Geometry ge = geometry(k);
    cudaStream_t stream = ctx->stream;
    CU(cudaEventRecord(ctx->ev_k0, stream));
    st = ensure_scratch(ctx, layout_scratch_bytes(seqs) + 256);
    if (st) return st;
    Scratch scratch{static_cast<char *>(ctx->scratch), ctx->scratch_bytes, 0};
    Layout L;
    st = plan_layout(ctx, seqs, k, ge, stream, KnownTotals(), scratch, &L);
    if (st) return st;
    result->n_written = L.total;
    if (L.total == 0) return KMC_OK;
    ExtractParams p = base_params(seqs, k, ge, L, 0);
    ConsumeParams c{};
    c.table = table;
    c.table_entries = 1u << (2 * k);
    // few counters: every block counts in shared memory and adds its non-zero counters once (global
    // atomics on a handful of addresses serialise); many counters: direct increments, L2-resident up to K = 12
    const bool shared = c.table_entries <= static_cast<uint32_t>(kSharedCompositionMax);
    ConsumeLaunchFn fn = shared ? consume_launcher<OP_COMPOSITION_SHARED>(ge, !L.uniform_len, mode == KMC_CANON)
                                : consume_launcher<OP_COMPOSITION>(ge, !L.uniform_len, mode == KMC_CANON);
    if (!fn) return fail(ctx, KMC_E_UNSUPPORTED, "no kernel for this K");
    CU(fn(p, c, stream));
    CU(cudaEventRecord(ctx->ev_k1, stream));
    CU(cudaStreamSynchronize(stream));
    CU(cudaEventElapsedTime(&result->kernel_ms, ctx->ev_k0, ctx->ev_k1));
    return KMC_OK;
}
