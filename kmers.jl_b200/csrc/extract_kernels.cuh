// extract_kernels.cuh -- the batched replacement of the reference's per-symbol iterator loops
// (FwKmers.jl:88-94, CanonicalKmers.jl:94-105,220-225, UnambiguousKmers.jl:64-77).
//
// Work decomposition.  The output of a read set is one flat array of windows ("flat index"):
// read r owns flat indices [f0_r, f0_r + wcount_r).  The flat array is cut into aligned groups
// of G windows (G*N*8 bytes = a multiple of 32, so a full group is written with 256-bit stores).
// A work item is (read r, group slot gi): the gi-th aligned group that intersects read r.  An item
// therefore only ever touches ONE read (no divergent slow path at read boundaries); groups
// that straddle two reads are emitted as two partial items by two threads.
//   uniform sets  : item -> (r, gi) by one division per thread at kernel start, then an
//                   incremental update per step (no division in the loop)
//   ragged sets   : item -> r through a two-level narrowing of the exclusive scan of group slots:
//                   a per-tile first read (tile_first, one binary search per TILE in a tiny
//                   pre-kernel), a per-warp-bucket first read (65 short searches per block, kept
//                   in shared memory), and a 0-2 step search per item inside its bucket's range
#pragma once
#include <cstdlib>
#include <type_traits>
#include "kmer_core.cuh"

namespace kmc {

enum : int { MODE_FW = 0, MODE_FWRV = 1, MODE_CANON = 2 };
// where the k-mers go: the output streams, or (north_star extension) a hash-bucket count table
// SINK_BUCKETS increments table[bucket] directly; SINK_IDS writes the 32-bit bucket id of every window
// to a flat array instead (first pass of the binned count for tables that do not fit L2, buckets.cu)
enum : int { SINK_STREAMS = 0, SINK_BUCKETS = 1, SINK_IDS = 2 };

constexpr int kBlockThreads = 256;
constexpr int kTileIters = 8;                                // work items per thread per tile
constexpr int kTileItems = kBlockThreads * kTileIters;       // work items per block

// EXPERIMENT ("measure a TMA bulk store, do not argue it"): with -DKMC_TMA_STORE=1 the SoA streams
// of aligned uniform sets (C2) leave the kernel as 1-D bulk copies, cp.async.bulk.global.shared::cta: every warp stages
// the 256 elements of a step in shared memory (2 KB per stream, double-buffered) and one lane issues one bulk store per
// stream.  Off by default: the measured A/B is in profiles/r02_ab_tma_store.txt and DESIGN.md 3.1.
#ifndef KMC_TMA_STORE
#define KMC_TMA_STORE 0
#endif
constexpr int kTmaStageBytes = 2 /*buffers*/ * 2 /*streams*/ * 32 * 8 * 8; // per warp, one-limb k-mers (G = 8)

KMC_DEV void bulk_store(void *gdst, const void *ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
                 "r"(static_cast<uint32_t>(__cvta_generic_to_shared(ssrc))), "r"(bytes), "l"(0x12F0000000000000ull)
                 : "memory");
}

struct ExtractParams {
    const uint32_t *w32; // sequence stream viewed as 32-bit words
    int64_t nw32;        // addressable 32-bit words (loads are clamped into it)
    uint32_t unit_bits;  // bits per offset unit: 64 (LongSequence words) or 32 (recoded 4-bit stream)
    uint64_t unit_bias;  // subtracted from every sequence offset (chunk views of a larger set)
    uint32_t first;      // first_symbol_offset
    int32_t k;
    uint32_t s0;         // 32*NX - 2K - 2(G-1)
    uint64_t head_mask;
    uint64_t n_seqs;
    uint64_t items;      // total work items (group slots); the grid covers at least this many
    const uint64_t *items_dev; // if not NULL the exact total lives on the device (grid is an upper bound)
    // uniform locator
    uint64_t stride_units;
    uint64_t wpr;        // windows per read
    uint64_t gprm;       // group slots per read
    uint64_t it_dq, it_dr; // divmod(kBlockThreads, gprm): per-iteration advance of (r, gi)
    uint64_t read_bits;    // stride_units * unit_bits: stream bits from one read to the next (uniform offsets)
    uint32_t aligned;      // uniform set whose windows per read are a multiple of G: item i IS flat group i (set_iteration_strides)
    uint32_t al_magic;     // floor(2^32 / gprm) (2^32 - 1 for gprm = 1): extract_aligned_kernel divides by gprm with it
    uint32_t al_tail;      // aligned single sequence: windows of its last, partial group (0: the last group is whole too)
    // ragged locator
    const uint64_t *seq_unit_off; // [n_seqs] or NULL (then r * stride_units); used by both locators
    const uint64_t *win_off;      // [n_seqs+1] exclusive scan of window counts
    const uint64_t *item_off;     // [n_seqs+1] exclusive scan of group slots
    const uint64_t *tile_first;   // [tiles+1] read that owns the first item of each tile
    const uint64_t *seq_index_base; // [n_seqs] or NULL: added to the emitted index (runs of a larger read)
    // outputs
    uint64_t *out_a;
    uint64_t *out_b;
    uint64_t *out_hash;
    int64_t *out_index;
    int64_t index_base;
    uint32_t aos;    // Julia tuple layout for FWRV / index
    uint32_t vec_ok; // every output base pointer is 32-byte aligned
    // fused bucket count (MODE_CANON only): table[fx_hash >> bucket_shift] += 1
    uint32_t *bucket_table;
    uint32_t bucket_shift;
    // 4-bit sources (FourToTwo): the kernels run on the recoded 2-bit stream (unit_bits = 32, one
    // unit per source word) and consult vstart: bit P = "no uncertain symbol in [P, P+K)", P the
    // absolute symbol index in the stream.
    const uint32_t *vstart;
    unsigned long long *err_flat;  // strict modes: atomicMin of the first flat window with an uncertain symbol
    // source prefetch in bursts (see burst_prefetch): tiles per chunk; 0 = off.  Set by launch_extract.
    uint32_t pf_tiles;
    // DIGEST instantiations: xor / wrapping sum of the words written to out_a -> digest[0..1], of out_hash -> [2..3]
    unsigned long long *digest;
};

// Validity bits of the slots [jlo, jhi) of one item: bit j set <=> window j has no uncertain symbol.
// sym0 = absolute symbol index (in the recoded stream) of slot 0; sym0 + jlo >= 0.
KMC_DEV uint32_t valid_slots(const uint32_t *__restrict__ vstart, int64_t sym0, int jlo, int jhi)
{
    const int64_t q = sym0 + jlo;
    const uint32_t w0 = __ldg(vstart + (q >> 5)), w1 = __ldg(vstart + (q >> 5) + 1);
    const uint32_t bits = __funnelshift_r(w0, w1, static_cast<uint32_t>(q) & 31u);
    return (bits & ((1u << (jhi - jlo)) - 1u)) << jlo;
}

// G windows per thread for N limbs: G*N*8 must be a multiple of 32 bytes (kmer_core.cuh: group_of).
template <int N> struct GroupOf { static constexpr int G = group_of(N); };

// ---------------------------------------------------------------------------------------------
// TileCursor: maps the work items of one tile (kTileItems consecutive group slots) to
// (read, slot) and derives everything the kernels need about the item's windows.
// ---------------------------------------------------------------------------------------------
constexpr int kTileBuckets = kTileItems / 32; // one bucket = the 32 items a warp handles in one step
constexpr int kTileMetaCap = 512;             // sequences of a tile whose descriptors are staged in shared memory

template <bool RAGGED> struct TileShared;
template <> struct TileShared<false> {
    uint64_t r0, gi0; // (read, slot) of the tile's first item
};
// Ragged sets: the descriptors of the sequences [r_first, r_first + n_meta) that own the tile's items
// are staged once per tile with coalesced loads, so that locating an item costs shared-memory
// latency instead of a chain of dependent global loads.  Tiles that span more than kTileMetaCap
// sequences (very short reads) fall back to the global arrays.
template <> struct TileShared<true> {
    uint32_t fr[kTileBuckets + 1]; // sequence (relative to r_first) owning the first item of each bucket
    uint32_t staged;
    uint64_t item_off[kTileMetaCap + 1];
    uint64_t win_off[kTileMetaCap + 1];
    uint64_t unit_off[kTileMetaCap];
    uint64_t ibase[kTileMetaCap];
};

template <bool RAGGED, int G>
struct TileCursor {
    uint64_t r, gi;       // current item
    uint64_t r_first;     // ragged: read owning the tile's first item
    // derived for the current item
    uint64_t f0, wcount;  // flat index of the read's first window, its window count
    uint64_t unit_off;    // ragged: where the read starts in the stream (units of p.unit_bits)
    uint64_t ubit;        // the same in bits
    uint64_t seq_ibase;   // p.seq_index_base[r] (0 without it)
    uint64_t q;           // aligned flat group of this item
    int64_t wbase;        // window (within the read) of slot 0, in (-G, wcount)
    int jlo, jhi;         // slots [jlo, jhi) are windows of this read

    KMC_DEV static uint64_t unit_off_of(const ExtractParams &p, uint64_t rr)
    {
        return p.seq_unit_off ? __ldg(p.seq_unit_off + rr) - p.unit_bias : rr * p.stride_units;
    }

    // li0 = this thread's first local item of the tile (threadIdx.x for the strided item order);
    // tile = index of the tile (blockIdx.x, except in compact_kernel, which takes its tiles from a ticket counter)
    KMC_DEV void init(const ExtractParams &p, uint64_t tile_base, TileShared<RAGGED> &sh, uint32_t li0)
    {
        init(p, tile_base, sh, li0, blockIdx.x);
    }
    KMC_DEV void init(const ExtractParams &p, uint64_t tile_base, TileShared<RAGGED> &sh, uint32_t li0, uint32_t tile)
    {
        seq_ibase = 0;
        if constexpr (!RAGGED) {
            // (r, gi) = divmod(item, gprm): one 64-bit division per BLOCK, then a 32-bit one per thread
            if (threadIdx.x == 0) {
                sh.r0 = tile_base / p.gprm;
                sh.gi0 = tile_base - sh.r0 * p.gprm;
            }
            __syncthreads();
            r = sh.r0;
            gi = sh.gi0 + li0;
            if (p.gprm > 0xffffffffull - kTileItems) { // gi0 + li0 wraps at most once
                if (gi >= p.gprm) {
                    gi -= p.gprm;
                    ++r;
                }
            } else {
                const uint32_t d = static_cast<uint32_t>(gi) / static_cast<uint32_t>(p.gprm);
                gi -= static_cast<uint64_t>(d) * p.gprm;
                r += d;
            }
        } else {
            r_first = __ldg(p.tile_first + tile);
            const uint64_t r_last = __ldg(p.tile_first + tile + 1);
            const uint64_t n_meta = r_last - r_first + 1;
            const bool staged = n_meta <= kTileMetaCap;
            if (staged) {
                for (uint32_t t = threadIdx.x; t <= n_meta; t += kBlockThreads) {
                    sh.item_off[t] = __ldg(p.item_off + r_first + t);
                    sh.win_off[t] = __ldg(p.win_off + r_first + t);
                    if (t < n_meta) {
                        sh.unit_off[t] = unit_off_of(p, r_first + t);
                        sh.ibase[t] = p.seq_index_base ? __ldg(p.seq_index_base + r_first + t) : 0ull;
                    }
                }
                if (threadIdx.x == 0) sh.staged = 1;
                __syncthreads();
            } else if (threadIdx.x == 0) {
                sh.staged = 0;
            }
            if (threadIdx.x <= kTileBuckets) {
                const uint64_t item = tile_base + 32ull * threadIdx.x;
                uint64_t lo = 0, hi = n_meta; // relative to r_first: largest rr in [lo, hi) with item_off <= item
                while (hi - lo > 1) {
                    const uint64_t mid = (lo + hi) >> 1;
                    const uint64_t v = staged ? sh.item_off[mid] : __ldg(p.item_off + r_first + mid);
                    if (v <= item) lo = mid; else hi = mid;
                }
                sh.fr[threadIdx.x] = static_cast<uint32_t>(lo);
            }
            __syncthreads();
            r = gi = 0;
        }
    }

    // item = tile_base + li
    KMC_DEV void locate(const ExtractParams &p, uint64_t item, uint32_t li, const TileShared<RAGGED> &sh)
    {
        if constexpr (RAGGED) {
            const uint32_t b = li >> 5;
            uint32_t lo = sh.fr[b], hi = sh.fr[b + 1] + 1;
            if (sh.staged) {
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (sh.item_off[mid] <= item) lo = mid; else hi = mid;
                }
                r = r_first + lo;
                gi = item - sh.item_off[lo];
                f0 = sh.win_off[lo];
                wcount = sh.win_off[lo + 1] - f0;
                unit_off = sh.unit_off[lo];
                seq_ibase = sh.ibase[lo];
            } else {
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (__ldg(p.item_off + r_first + mid) <= item) lo = mid; else hi = mid;
                }
                r = r_first + lo;
                gi = item - __ldg(p.item_off + r);
                f0 = __ldg(p.win_off + r);
                wcount = __ldg(p.win_off + r + 1) - f0;
                unit_off = unit_off_of(p, r);
                seq_ibase = p.seq_index_base ? __ldg(p.seq_index_base + r) : 0ull;
            }
            ubit = unit_off * p.unit_bits;
        } else {
            if (p.aligned && !p.al_tail) {
                // windows per read are a multiple of G (C2: 120 = 15 x 8): every group lies wholly inside one read, item i
                // is flat group i and all G slots are windows.  One 32 x 32 -> 64-bit product instead of the two 64-bit
                // products and the divisions-by-G bookkeeping below (60 of the 314 instructions per item, ncu r01).
                q = item;
                wbase = static_cast<int64_t>(static_cast<uint32_t>(gi) * static_cast<uint32_t>(G));
                jlo = 0;
                jhi = G;
                ubit = static_cast<uint64_t>(static_cast<uint32_t>(r)) * static_cast<uint32_t>(p.read_bits);
                return;
            }
            // (recomputed per item: carrying f0 / ubit incrementally across the iterations measured 2 % slower
            // on the ALU-bound modes -- two more live 64-bit values per thread)
            f0 = r * p.wpr;
            wcount = p.wpr;
            ubit = p.seq_unit_off ? (__ldg(p.seq_unit_off + r) - p.unit_bias) * p.unit_bits : r * p.read_bits;
        }
        q = f0 / G + gi;
        wbase = static_cast<int64_t>(q * G - f0);
        const int64_t rem = static_cast<int64_t>(wcount) - wbase; // windows available from slot 0
        jlo = wbase < 0 ? static_cast<int>(-wbase) : 0;
        jhi = rem < G ? static_cast<int>(rem) : G;
    }

    // bit offset in the stream of slot 0's first symbol (BPS = bits per symbol of the stream)
    template <int BPS = 2>
    KMC_DEV int64_t bit(const ExtractParams &p) const
    {
        return static_cast<int64_t>(ubit) + BPS * (static_cast<int64_t>(p.first) + wbase);
    }

    // to the item kBlockThreads further on
    KMC_DEV void advance(const ExtractParams &p)
    {
        if (!RAGGED) {
            r += p.it_dq;
            gi += p.it_dr;
            if (gi >= p.gprm) {
                gi -= p.gprm;
                ++r;
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Source prefetch in bursts.  The kernels write 16-50 bytes for every byte they read, and HBM pays for the
// mix out of proportion: a pure stream of these stores runs at 7.48 TB/s, the same stores with the 2 % trickle
// of source reads among them at 6.5 TB/s (tools/bw_probe.py, KMC_STORE_PROBE_PATTERN=2 / 3) -- every read that
// reaches DRAM interrupts a run of writes.  So the reads are taken out of the trickle: the tiles are cut into
// chunks of pf_tiles tiles (about 20 MB of source), and kPfBlocks blocks shortly before the end of chunk c
// pull the source of chunk c + 1 into L2 with evict_last prefetches -- one short burst of reads per chunk --
// where the loads of the tiles find it; the stores are marked evict_first so that they do not push it out again.
// Probe (pattern 7): 7.1 TB/s, 95 % of the pure-write ceiling.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kPfBlocks = 32;   // blocks that share one chunk's prefetch
constexpr uint32_t kPfLead = 128;    // tiles before the chunk boundary at which they run (blocks are dispatched in order)
constexpr uint64_t kPfChunkBytes = 20ull << 20;

// bit offset in the source stream of the first item of `tile` (end of the stream for tiles beyond the last)
template <bool RAGGED, int G, int BPS>
KMC_DEV int64_t tile_start_bit(const ExtractParams &p, uint64_t tile, uint64_t n_items)
{
    const int64_t end = p.nw32 * 32;
    const uint64_t item = tile * kTileItems;
    if (item >= n_items) return end;
    uint64_t r, gi, f0, ubit;
    if (RAGGED) {
        r = __ldg(p.tile_first + tile);
        gi = item - __ldg(p.item_off + r);
        f0 = __ldg(p.win_off + r);
        ubit = (p.seq_unit_off ? __ldg(p.seq_unit_off + r) - p.unit_bias : r * p.stride_units) * p.unit_bits;
    } else {
        r = item / p.gprm;
        gi = item - r * p.gprm;
        f0 = r * p.wpr;
        ubit = p.seq_unit_off ? (__ldg(p.seq_unit_off + r) - p.unit_bias) * p.unit_bits : r * p.read_bits;
    }
    const int64_t wbase = static_cast<int64_t>((f0 / G + gi) * G - f0);
    const int64_t bit = static_cast<int64_t>(ubit) + BPS * (static_cast<int64_t>(p.first) + wbase);
    return bit < 0 ? 0 : (bit > end ? end : bit);
}

template <bool RAGGED, int G, int BPS>
KMC_DEV void burst_prefetch(const ExtractParams &p, uint64_t n_items)
{
    const uint32_t T = p.pf_tiles;
    const uint32_t c = blockIdx.x / T, k = blockIdx.x - c * T;
    const uint32_t lead = kPfLead < T ? kPfLead : T, k0 = T - lead;
    const bool first = blockIdx.x < kPfBlocks; // the first blocks of the grid pull chunk 0
    if (!first && (k < k0 || k >= k0 + kPfBlocks)) return;
    const uint64_t cc = first ? 0 : c + 1, kk = first ? blockIdx.x : k - k0;
    const int64_t b0 = (tile_start_bit<RAGGED, G, BPS>(p, cc * T, n_items) >> 3) & ~127ll;
    int64_t b1 = (tile_start_bit<RAGGED, G, BPS>(p, (cc + 1) * T, n_items) >> 3) + 256; // + the halo of the last item
    if (b1 > p.nw32 * 4) b1 = p.nw32 * 4;
    const char *base = reinterpret_cast<const char *>(p.w32);
    for (int64_t o = b0 + (static_cast<int64_t>(kk) * kBlockThreads + threadIdx.x) * 128; o < b1;
         o += static_cast<int64_t>(kPfBlocks) * kBlockThreads * 128)
        asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(base + o));
    if (BPS == 2 && p.vstart) {
        // strict iteration over a recoded source: the valid-start bits of the chunk (one bit per symbol = half the bytes
        // of the 2-bit stream) are a second trickle of reads; pull them in with the same burst
        const char *vb = reinterpret_cast<const char *>(p.vstart);
        const int64_t v0 = (b0 >> 1) & ~127ll, v1 = (b1 >> 1) + 128; // (the array has two spare words past the last symbol's)
        for (int64_t o = v0 + (static_cast<int64_t>(kk) * kBlockThreads + threadIdx.x) * 128; o < v1;
             o += static_cast<int64_t>(kPfBlocks) * kBlockThreads * 128)
            asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(vb + o));
    }
}

// KMC_PREFETCH=0 switches the burst prefetch off (A/B measurements)
inline bool prefetch_enabled()
{
    static const bool on = [] {
        const char *e = getenv("KMC_PREFETCH");
        return !(e && e[0] == '0');
    }();
    return on;
}

// per-iteration strides of the uniform locator (every launcher calls this before the launch)
// g = windows per work item of a uniform-layout launch (0: ragged, or a kernel that does not use the aligned form)
inline void set_iteration_strides(ExtractParams &p, int g = 0)
{
    p.it_dq = kBlockThreads / p.gprm;
    p.it_dr = kBlockThreads % p.gprm;
    p.read_bits = p.stride_units * p.unit_bits;
    // groups never straddle two reads when the windows per read are a multiple of g -- or when there is one read only (a
    // long sequence, a shard of one: C4), whose last group may then be partial (al_tail windows)
    const bool whole = g > 0 && p.wpr > 0 && p.wpr % static_cast<uint64_t>(g) == 0 && p.gprm == p.wpr / g;
    const bool single = g > 0 && p.wpr > 0 && p.n_seqs == 1 && p.gprm == (p.wpr + g - 1) / g;
    p.aligned = ((whole || single) && !p.seq_unit_off && p.n_seqs < 0xffffffffull && p.read_bits < 0xffffffffull && p.gprm < 0xffffffffull)
                    ? 1u
                    : 0u;
    p.al_tail = (p.aligned && !whole) ? static_cast<uint32_t>(p.wpr % static_cast<uint64_t>(g)) : 0u;
}



// DIGEST instantiations: the threads' shares of the fingerprint -> one set of atomics per block (same-address atomics
// serialise in L2).  Every thread of the block calls it.
KMC_DEV void digest_epilogue(unsigned long long *digest, uint64_t dg_xa, uint64_t dg_sa, uint64_t dg_xh, uint64_t dg_sh)
{
    __shared__ uint64_t s_dg[4][kBlockThreads / 32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        dg_xa ^= __shfl_xor_sync(0xffffffffu, dg_xa, d);
        dg_sa += __shfl_xor_sync(0xffffffffu, dg_sa, d);
        dg_xh ^= __shfl_xor_sync(0xffffffffu, dg_xh, d);
        dg_sh += __shfl_xor_sync(0xffffffffu, dg_sh, d);
    }
    if ((threadIdx.x & 31) == 0) {
        s_dg[0][threadIdx.x >> 5] = dg_xa;
        s_dg[1][threadIdx.x >> 5] = dg_sa;
        s_dg[2][threadIdx.x >> 5] = dg_xh;
        s_dg[3][threadIdx.x >> 5] = dg_sh;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        uint64_t v = 0;
#pragma unroll
        for (int w = 0; w < kBlockThreads / 32; ++w) v = (threadIdx.x & 1) ? v + s_dg[threadIdx.x][w] : v ^ s_dg[threadIdx.x][w];
        if (threadIdx.x & 1)
            atomicAdd(digest + threadIdx.x, static_cast<unsigned long long>(v));
        else
            atomicXor(digest + threadIdx.x, static_cast<unsigned long long>(v));
    }
}

template <int N, int NX, int MODE, bool HASH, bool RAGGED, int SINK = SINK_STREAMS, bool STRICT4 = false, int BPS = 2,
          bool DIGEST = false>
__global__ void __launch_bounds__(kBlockThreads) extract_kernel(const ExtractParams p)
{
    static_assert(!DIGEST || (SINK == SINK_STREAMS && MODE != MODE_FWRV), "the fused fingerprint covers out_a and out_hash");
    uint64_t dg_xa = 0, dg_sa = 0, dg_xh = 0, dg_sh = 0; // DIGEST: this thread's share of the fingerprint
    constexpr int G = GroupOf<N>::G;
    constexpr bool WANT_FW = true;
    constexpr bool WANT_RV = (MODE != MODE_FW);

    // One tile of kTileIters x 256 consecutive work items per block; blocks are scheduled by the
    // hardware as SMs drain, which balances the SMs (a static persistent partition left the
    // fastest SMs idle for 25 % of the kernel: profiles/r01_c2_canon31_hash_v1_persistent.txt).
    __shared__ TileShared<RAGGED> sh;
    const uint64_t tile_base = static_cast<uint64_t>(blockIdx.x) * kTileItems;
    const uint64_t n_items = p.items_dev ? __ldg(p.items_dev) : p.items;
    if (tile_base >= n_items) return; // block-uniform
    if (p.pf_tiles) burst_prefetch<RAGGED, G, BPS>(p, n_items);
    TileCursor<RAGGED, G> cur;
    cur.init(p, tile_base, sh, threadIdx.x);

#pragma unroll 1
    for (int it = 0; it < kTileIters; ++it) {
        const uint64_t item = tile_base + static_cast<uint64_t>(it) * kBlockThreads + threadIdx.x;
        if (item >= n_items) break;
        cur.locate(p, item, static_cast<uint32_t>(it) * kBlockThreads + threadIdx.x, sh);
        const uint64_t q = cur.q;
        const int64_t wbase = cur.wbase;
        const int jlo = cur.jlo, jhi = cur.jhi;

        if (jhi > jlo) {
            const int64_t bit = cur.template bit<BPS>(p);
            if (STRICT4) {
                // FourToTwo, strict (FwKmers.jl:104-115, CanonicalKmers.jl:131-144): an uncertain symbol
                // is an error.  Record the first offending window; the host resolves it to the symbol
                // the reference would have thrown on.
                const uint32_t ok = valid_slots(p.vstart, bit >> 1, jlo, jhi);
                const uint32_t want = ((1u << (jhi - jlo)) - 1u) << jlo;
                if (ok != want) atomicMin(p.err_flat, static_cast<unsigned long long>(q * G + (__ffs(ok ^ want) - 1)));
            }
            uint32_t x[NX];
            load_block<NX>(p.w32, p.nw32, bit, x);
            uint64_t fw[G][N], rv[G][N];
            block_kmers<N, NX, G, WANT_FW, WANT_RV, BPS>(x, p.s0, p.head_mask, fw, rv);

            // what lands in out_a, and its hash
            uint64_t a[G][N], h[G];
#pragma unroll
            for (int j = 0; j < G; ++j) {
                bool take_fw = true;
                if (MODE == MODE_CANON) take_fw = limbs_less<N>(fw[j], rv[j]); // fw < rv ? fw : rv
#pragma unroll
                for (int i = 0; i < N; ++i) a[j][i] = take_fw ? fw[j][i] : rv[j][i];
                if (HASH) h[j] = fx_hash<N>(a[j], 0);
                if (DIGEST && j >= jlo && j < jhi) {
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        dg_xa ^= a[j][i];
                        dg_sa += a[j][i];
                    }
                    if (HASH) {
                        dg_xh ^= h[j];
                        dg_sh += h[j];
                    }
                }
            }

            if (SINK == SINK_IDS) {
                uint32_t *ids = reinterpret_cast<uint32_t *>(p.out_a) + q * G;
                if (G == 8 && jlo == 0 && jhi == G && p.vec_ok) {
                    uint64_t w[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        w[t] = (h[(2 * t) % G] >> p.bucket_shift) | ((h[(2 * t + 1) % G] >> p.bucket_shift) << 32);
                    st_v4(reinterpret_cast<uint64_t *>(ids), w[0], w[1], w[2], w[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < G; ++j)
                        if (j >= jlo && j < jhi) ids[j] = static_cast<uint32_t>(h[j] >> p.bucket_shift);
                }
                goto next_item;
            }
            if (SINK == SINK_BUCKETS) {
                // canonical k-mer -> fx_hash -> bucket -> counter (no k-mer stream is written)
#pragma unroll
                for (int j = 0; j < G; ++j)
                    if (j >= jlo && j < jhi) atomicAdd(p.bucket_table + (h[j] >> p.bucket_shift), 1u);
                goto next_item;
            }

            const uint64_t fbase = q * G; // flat index of slot 0
#if KMC_TMA_STORE
            // (a warp whose last lane is past the end takes the ordinary stores: its lanes have left the loop)
            const bool warp_full = tile_base + static_cast<uint64_t>(it) * kBlockThreads + (threadIdx.x | 31u) < n_items;
            if (N == 1 && SINK == SINK_STREAMS && MODE != MODE_FWRV && p.aligned && !p.al_tail && p.vec_ok && !p.out_index && warp_full) {
                // the warp's 32 items are 256 consecutive elements of every stream: lane l's 8 go to bytes [64 l, 64 l + 64)
                extern __shared__ __align__(128) unsigned char s_tma[];
                const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
                unsigned char *buf = s_tma + warp * kTmaStageBytes + (it & 1) * (kTmaStageBytes / 2);
                if (it >= 2) { // the bulk copies of step it - 2 have finished reading this buffer
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
                }
                uint64_t *sa = reinterpret_cast<uint64_t *>(buf) + lane * G, *sh = sa + 32 * G;
#pragma unroll
                for (int j = 0; j < G; j += 2) {
                    *reinterpret_cast<ulonglong2 *>(sa + j) = make_ulonglong2(a[j][0], a[j + 1][0]);
                    if (HASH) *reinterpret_cast<ulonglong2 *>(sh + j) = make_ulonglong2(h[j], h[j + 1]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy writes -> visible to the bulk copy
                __syncwarp();
                if (lane == 0) {
                    bulk_store(p.out_a + fbase, buf, 32 * G * 8); // lane 0's fbase is the warp's first element
                    if (HASH) bulk_store(p.out_hash + fbase, buf + 32 * G * 8, 32 * G * 8);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                goto next_item;
            }
#endif
            // 1-based start (within its sequence) of slot 0's window -- only the index-emitting branches evaluate it
            auto ibase_of = [&]() -> int64_t { return wbase + 1 + p.index_base + static_cast<int64_t>(cur.seq_ibase); };
            const bool full = (jlo == 0) && (jhi == G);
            const bool tuple_rv = (MODE == MODE_FWRV) && p.aos;
            const bool tuple_ix = (p.out_index != nullptr) && p.aos;
            const bool fast = full && p.vec_ok;
            // Every stream is staged as one run of words per item and written with the widest stores
            // its alignment allows: whole 256-bit stores for a full group, and for a partial group
            // (read / run boundary) 256- and 128-bit stores over the aligned parts of [jlo, jhi).
            if (tuple_rv) {
                uint64_t buf[2 * G * N];
#pragma unroll
                for (int j = 0; j < G; ++j)
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        buf[j * 2 * N + i] = fw[j][i];
                        buf[j * 2 * N + N + i] = rv[j][i];
                    }
                store_words<2 * G * N>(p.out_a + fbase * (2 * N), buf, jlo * 2 * N, jhi * 2 * N, fast, p.vec_ok);
            } else if (tuple_ix) {
                // Tuple{Kmer,Int} = {u64[N]; i64} elements
                uint64_t buf[G * (N + 1)];
                const int64_t ibase = ibase_of();
#pragma unroll
                for (int j = 0; j < G; ++j) {
#pragma unroll
                    for (int i = 0; i < N; ++i) buf[j * (N + 1) + i] = a[j][i];
                    buf[j * (N + 1) + N] = static_cast<uint64_t>(ibase + j);
                }
                store_words<G * (N + 1)>(p.out_a + fbase * (N + 1), buf, jlo * (N + 1), jhi * (N + 1), fast, p.vec_ok);
            } else {
                uint64_t buf[G * N];
#pragma unroll
                for (int j = 0; j < G; ++j)
#pragma unroll
                    for (int i = 0; i < N; ++i) buf[j * N + i] = a[j][i];
                store_words<G * N>(p.out_a + fbase * N, buf, jlo * N, jhi * N, fast, p.vec_ok);
                if (MODE == MODE_FWRV) {
#pragma unroll
                    for (int j = 0; j < G; ++j)
#pragma unroll
                        for (int i = 0; i < N; ++i) buf[j * N + i] = rv[j][i];
                    store_words<G * N>(p.out_b + fbase * N, buf, jlo * N, jhi * N, fast, p.vec_ok);
                }
                if (p.out_index) {
                    uint64_t ib[G];
                    const int64_t ibase = ibase_of();
#pragma unroll
                    for (int j = 0; j < G; ++j) ib[j] = static_cast<uint64_t>(ibase + j);
                    store_words<G>(reinterpret_cast<uint64_t *>(p.out_index) + fbase, ib, jlo, jhi, fast, p.vec_ok);
                }
            }
            if (HASH) store_words<G>(p.out_hash + fbase, h, jlo, jhi, fast, p.vec_ok);
        }

    next_item:
        cur.advance(p);
    }
#if KMC_TMA_STORE
    if ((threadIdx.x & 31) == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // shared memory stays valid until read
#endif
    if (DIGEST) digest_epilogue(p.digest, dg_xa, dg_sa, dg_xh, dg_sh);
}

// The work items of an aligned uniform set (ExtractParams::aligned): item i is flat group i of G windows, wholly inside
// read i / gprm (kmer_core.cuh: AlignedLocator, which the CPU tests also drive).
template <int G, int BPS> struct AlignedItems : AlignedLocator<G, BPS> {
    KMC_DEV explicit AlignedItems(const ExtractParams &p)
        : AlignedLocator<G, BPS>(static_cast<uint32_t>(p.gprm), p.al_magic, static_cast<uint32_t>(p.read_bits), p.first)
    {
    }
    // Offsets grow with the item, so the last item of a tile bounds the block loads of all of them: true when NX + 1
    // words from there lie inside the buffer (all tiles but the one or two that reach the end of it).
    template <int NX> KMC_DEV bool loads_inside(const ExtractParams &p, uint32_t last_item) const
    {
        return static_cast<int64_t>(this->bit_of(last_item) >> 5) + NX < p.nw32;
    }
};

// the aligned x-stream of the block at `bit`; inside = loads_inside() of the tile (no clamping needed)
template <int NX> KMC_DEV void load_block_at(const ExtractParams &p, uint64_t bit, bool inside, uint32_t (&x)[NX])
{
    if (inside) {
        const uint32_t *w = p.w32 + (bit >> 5);
        uint32_t a[NX + 1];
#pragma unroll
        for (int i = 0; i <= NX; ++i) a[i] = __ldg(w + i);
        const uint32_t sh = static_cast<uint32_t>(bit) & 31u;
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = __funnelshift_r(a[i], a[i + 1], sh);
    } else {
        load_block<NX>(p.w32, p.nw32, static_cast<int64_t>(bit), x);
    }
}

// ---------------------------------------------------------------------------------------------
// extract_aligned_kernel: the same work items for the case every headline configuration is in -- a uniform read set
// (or one sequence) whose windows per read are a multiple of G (C2: 120 = 15 x 8), SoA streams with 32-byte aligned
// bases, fewer than 2^32 items.  Then item i IS flat group i, all G slots are windows, and nothing of the general
// kernel's bookkeeping is left: no (read, slot) cursor carried through the loop (one multiply-high by a precomputed
// reciprocal per item instead), no partial groups, no run-time choice of an output layout, no bounds test on the
// block load (one test per tile: only a tile that can reach the last words of the buffer takes the clamped loads).
// The k-mer arithmetic is the shared block_kmers / limbs_less / fx_hash.  SINK_IDS writes the 32-bit bucket id of
// every window (first pass of the binned count, buckets.cu).
// ---------------------------------------------------------------------------------------------
// AOS: the Julia tuple layouts.  A thread's windows must then fill whole 32-byte sectors that neighbouring lanes continue, or
// a warp's store instruction touches 32 different lines: the group size GG is chosen so that one item IS 32 bytes of output
// (one-limb k-mers: GG = 2 for both tuples) -- a quarter of the windows per block load, but one 256-bit store per lane and
// every line written whole by four neighbouring lanes.
enum : int { AOS_NONE = 0, AOS_FWRV = 1 /* Tuple{Kmer,Kmer} */, AOS_INDEX = 2 /* Tuple{Kmer,Int} */ };

// INDEX: an SoA stream of the windows' 1-based starts beside the k-mers (UnambiguousKmers over a 2-bit source).
template <int N, int NX, int MODE, bool HASH, int SINK = SINK_STREAMS, int BPS = 2, bool DIGEST = false, bool STRICT4 = false, int GG = 0,
          int AOS = AOS_NONE, bool INDEX = false>
__global__ void __launch_bounds__(kBlockThreads) extract_aligned_kernel(const ExtractParams p)
{
    static_assert(!INDEX || (AOS == AOS_NONE && SINK == SINK_STREAMS), "the index stream of the SoA form");
    static_assert(SINK == SINK_STREAMS || (SINK == SINK_IDS && MODE == MODE_CANON && HASH), "bucket ids are hashes of canonical k-mers");
    static_assert(AOS == AOS_NONE || (SINK == SINK_STREAMS && !DIGEST), "tuple layouts are output streams");
    static_assert(AOS != AOS_FWRV || MODE == MODE_FWRV, "Tuple{Kmer,Kmer} is the FwRvIterator's element");
    constexpr int G = GG > 0 ? GG : GroupOf<N>::G;
    constexpr bool WANT_RV = (MODE != MODE_FW);
    uint64_t dg_xa = 0, dg_sa = 0, dg_xh = 0, dg_sh = 0;
    const uint32_t n_items = static_cast<uint32_t>(p.items);
    const uint32_t tile_base = blockIdx.x * static_cast<uint32_t>(kTileItems); // the launcher keeps items below 2^32 - kTileItems
    if (p.pf_tiles) burst_prefetch<false, G, BPS>(p, n_items);
    const AlignedItems<G, BPS> items(p);
    const uint32_t tile_last = (n_items - tile_base > static_cast<uint32_t>(kTileItems) ? tile_base + kTileItems : n_items) - 1u;
    const bool safe = items.template loads_inside<NX>(p, tile_last); // block-uniform

#pragma unroll 1
    for (int it = 0; it < kTileIters; ++it) {
        const uint32_t item = tile_base + static_cast<uint32_t>(it) * kBlockThreads + threadIdx.x;
        if (item >= n_items) break;
        uint32_t gi;
        const uint64_t bit = items.bit_of(item, gi);
        // all G slots are windows, except in the last group of a single sequence whose window count is not a multiple of G
        const bool partial = p.al_tail != 0 && item == n_items - 1u;
        const int jhi = partial ? static_cast<int>(p.al_tail) : G;
        if (STRICT4) {
            // recoded 4-bit / ASCII source, strict iteration: the first window with a symbol that cannot be encoded is the
            // error (extract_kernel has the same test; the host resolves it to the symbol the reference throws on)
            const uint32_t ok = valid_slots(p.vstart, static_cast<int64_t>(bit >> 1), 0, jhi);
            const uint32_t want = (1u << jhi) - 1u;
            if (ok != want)
                atomicMin(p.err_flat, static_cast<unsigned long long>(static_cast<uint64_t>(item) * G + (__ffs(ok ^ want) - 1)));
        }
        uint32_t x[NX];
        load_block_at<NX>(p, bit, safe, x);
        uint64_t fw[G][N], rv[G][N];
        block_kmers<N, NX, G, true, WANT_RV, BPS>(x, p.s0, p.head_mask, fw, rv);

        uint64_t a[G][N], h[G];
#pragma unroll
        for (int j = 0; j < G; ++j) {
            bool take_fw = true;
            if (MODE == MODE_CANON) take_fw = limbs_less<N>(fw[j], rv[j]);
#pragma unroll
            for (int i = 0; i < N; ++i) a[j][i] = take_fw ? fw[j][i] : rv[j][i];
            if (HASH) h[j] = fx_hash<N>(a[j], 0);
            if (DIGEST && j < jhi) {
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    dg_xa ^= a[j][i];
                    dg_sa += a[j][i];
                }
                if (HASH) {
                    dg_xh ^= h[j];
                    dg_sh += h[j];
                }
            }
        }

        if constexpr (SINK == SINK_IDS) {
            // G ids = G / 2 words; the group is 4 G bytes and the base is 32-byte aligned
            uint64_t w[G / 2];
#pragma unroll
            for (int t = 0; t < G / 2; ++t) w[t] = (h[2 * t] >> p.bucket_shift) | ((h[2 * t + 1] >> p.bucket_shift) << 32);
            uint64_t *dst = p.out_a + static_cast<uint64_t>(item) * (G / 2);
            if (partial) {
                uint32_t *ids = reinterpret_cast<uint32_t *>(dst);
#pragma unroll
                for (int j = 0; j < G; ++j)
                    if (j < jhi) ids[j] = static_cast<uint32_t>(h[j] >> p.bucket_shift);
            } else if (G == 8) st_v4(dst, w[0], w[1 % (G / 2)], w[2 % (G / 2)], w[3 % (G / 2)]);
            else if (G == 4) st_v2(dst, w[0], w[1 % (G / 2)]);
            else st_u64(dst, w[0]);
        } else if constexpr (AOS != AOS_NONE) {
            constexpr int E = AOS == AOS_FWRV ? 2 * N : N + 1; // words per element
            uint64_t buf[G * E];
            const int64_t ibase = static_cast<int64_t>(gi * static_cast<uint32_t>(G)) + 1 + p.index_base; // 1-based start of slot 0
#pragma unroll
            for (int j = 0; j < G; ++j) {
#pragma unroll
                for (int i = 0; i < N; ++i) buf[j * E + i] = AOS == AOS_FWRV ? fw[j][i] : a[j][i];
                if (AOS == AOS_FWRV) {
#pragma unroll
                    for (int i = 0; i < N; ++i) buf[j * E + N + i] = rv[j][i];
                } else {
                    buf[j * E + N] = static_cast<uint64_t>(ibase + j);
                }
            }
            if (partial) {
                store_words<G * E>(p.out_a + static_cast<uint64_t>(item) * (G * E), buf, 0, jhi * E, false, true);
                if (HASH) store_words<G>(p.out_hash + static_cast<uint64_t>(item) * G, h, 0, jhi, false, true);
                continue;
            }
            store_run<G * E>(p.out_a + static_cast<uint64_t>(item) * (G * E), buf, true);
            if (HASH) store_run<G>(p.out_hash + static_cast<uint64_t>(item) * G, h, false);
        } else {
            uint64_t buf[G * N];
#pragma unroll
            for (int j = 0; j < G; ++j)
#pragma unroll
                for (int i = 0; i < N; ++i) buf[j * N + i] = a[j][i];
            if (partial) { // (one thread of the launch)
                store_words<G * N>(p.out_a + static_cast<uint64_t>(item) * (G * N), buf, 0, jhi * N, false, true);
                if (MODE == MODE_FWRV) {
#pragma unroll
                    for (int j = 0; j < G; ++j)
#pragma unroll
                        for (int i = 0; i < N; ++i) buf[j * N + i] = rv[j][i];
                    store_words<G * N>(p.out_b + static_cast<uint64_t>(item) * (G * N), buf, 0, jhi * N, false, true);
                }
                if (HASH) store_words<G>(p.out_hash + static_cast<uint64_t>(item) * G, h, 0, jhi, false, true);
                if (INDEX) {
                    uint64_t ib[G];
#pragma unroll
                    for (int j = 0; j < G; ++j) ib[j] = static_cast<uint64_t>(static_cast<int64_t>(gi * static_cast<uint32_t>(G)) + 1 + j + p.index_base);
                    store_words<G>(reinterpret_cast<uint64_t *>(p.out_index) + static_cast<uint64_t>(item) * G, ib, 0, jhi, false, true);
                }
                continue;
            }
            store_run<G * N>(p.out_a + static_cast<uint64_t>(item) * (G * N), buf, true);
            if (MODE == MODE_FWRV) {
#pragma unroll
                for (int j = 0; j < G; ++j)
#pragma unroll
                    for (int i = 0; i < N; ++i) buf[j * N + i] = rv[j][i];
                store_run<G * N>(p.out_b + static_cast<uint64_t>(item) * (G * N), buf, true);
            }
            if (HASH) store_run<G>(p.out_hash + static_cast<uint64_t>(item) * G, h, true);
            if (INDEX) { // UnambiguousKmers over a 2-bit source: every window, with its 1-based start (SoA index stream)
                uint64_t ib[G];
#pragma unroll
                for (int j = 0; j < G; ++j) ib[j] = static_cast<uint64_t>(static_cast<int64_t>(gi * static_cast<uint32_t>(G)) + 1 + j + p.index_base);
                store_run<G>(reinterpret_cast<uint64_t *>(p.out_index) + static_cast<uint64_t>(item) * G, ib, true);
            }
        }
    }
    if (DIGEST) digest_epilogue(p.digest, dg_xa, dg_sa, dg_xh, dg_sh);
}

// KMC_ALIGNED_KERNEL=0 sends aligned uniform sets through the general kernel (A/B measurements, and the tests of both)
inline bool aligned_kernel_enabled()
{
    static const bool on = [] {
        const char *e = getenv("KMC_ALIGNED_KERNEL");
        return !(e && e[0] == '0');
    }();
    return on;
}

// Host-side launcher: one block per tile of kTileItems work items.  Defined per N in
// extract_n*.cu so the instantiations compile in parallel.
using ExtractLaunchFn = cudaError_t (*)(ExtractParams, int sm_count, cudaStream_t);

template <int N, int NX, int MODE, bool HASH, bool RAGGED, int SINK = SINK_STREAMS, bool STRICT4 = false, int BPS = 2,
          bool DIGEST = false>
cudaError_t launch_extract(ExtractParams p, int /*sm_count*/, cudaStream_t stream)
{
    const uint64_t tiles = (p.items + kTileItems - 1) / kTileItems;
    if (tiles == 0) return cudaSuccess;
    if (tiles > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    set_iteration_strides(p, RAGGED ? 0 : GroupOf<N>::G);
    p.pf_tiles = 0;
    if (SINK != SINK_BUCKETS && prefetch_enabled()) {
        // chunks of about kPfChunkBytes of source: tiles per chunk from the average source bytes per tile
        const uint64_t per_tile = static_cast<uint64_t>(p.nw32) * 4 / tiles + 1;
        const uint64_t t = kPfChunkBytes / per_tile;
        p.pf_tiles = static_cast<uint32_t>(t < 2 * kPfLead ? 2 * kPfLead : (t > (1u << 20) ? (1u << 20) : t));
    }
    if (DIGEST && !p.digest) return cudaErrorInvalidValue;
    size_t smem = 0;
#if KMC_TMA_STORE
    if (N == 1 && SINK == SINK_STREAMS && MODE != MODE_FWRV) {
        smem = static_cast<size_t>(kBlockThreads / 32) * kTmaStageBytes;
        cudaError_t e = cudaFuncSetAttribute(extract_kernel<N, NX, MODE, HASH, RAGGED, SINK, STRICT4, BPS, DIGEST>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
    }
#endif
    if constexpr (!RAGGED && SINK != SINK_BUCKETS && !KMC_TMA_STORE) {
        // the index stream exists for the plain forward form only (UnambiguousKmers over a 2-bit source)
        constexpr bool kCanIndex = MODE == MODE_FW && SINK == SINK_STREAMS && !DIGEST && !STRICT4 && BPS == 2;
        const bool plain_soa = SINK == SINK_IDS || (!p.aos && (kCanIndex || !p.out_index));
        if (p.aligned && p.vec_ok && plain_soa && !p.items_dev && p.gprm < 0x80000000ull &&
            p.items < 0xffffffffull - kTileItems && aligned_kernel_enabled()) {
            p.al_magic = aligned_magic(p.gprm);
            // Three resident blocks per SM, not the six its 38 registers allow: the kernel is bound by the HBM write stream, and
            // more blocks only mean more write streams in flight at once (measured, profiles/r02_ab_aligned_occupancy.txt:
            // FwRvIterator SoA 2.73 ms with 6 blocks per SM, 2.68 with 5, 2.66 with 4, 2.64 with 3; the ALU-bound canonical
            // stream without hash 1.41 ms with 3 to 6 and 1.6 ms with 2).  The cap is an unused dynamic shared-memory
            // request; KMC_ALIGNED_SMEM overrides its size in bytes (0 = no cap).
            static const int pad = [] {
                const char *e = getenv("KMC_ALIGNED_SMEM");
                const int v = e ? atoi(e) : (SINK == SINK_STREAMS ? 72 * 1024 : 0);
                return v < 0 ? 0 : (v > 200 * 1024 ? 200 * 1024 : v);
            }();
            auto go = [&](auto index_tag) -> cudaError_t {
                constexpr bool IDX = decltype(index_tag)::value;
                auto kernel = extract_aligned_kernel<N, NX, MODE, HASH, SINK, BPS, DIGEST, STRICT4, 0, AOS_NONE, IDX>;
                if (pad > 0) {
                    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pad);
                    if (e != cudaSuccess) return e;
                }
                kernel<<<static_cast<unsigned>(tiles), kBlockThreads, pad, stream>>>(p);
                return cudaGetLastError();
            };
            if constexpr (kCanIndex) {
                if (p.out_index) return go(std::true_type());
            }
            return go(std::false_type());
        }
    }
    extract_kernel<N, NX, MODE, HASH, RAGGED, SINK, STRICT4, BPS, DIGEST>
        <<<static_cast<unsigned>(tiles), kBlockThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

// The tuple layouts of one-limb k-mers over an aligned uniform set (or a single sequence) whose layout was planned with groups
// of two windows (geometry(k, 2, 2)): Tuple{Kmer,Kmer} (fwrv) or Tuple{Kmer,Int} (every window with its index).  Returns
// cudaErrorNotSupported when the set does not qualify; the caller then plans the ordinary layout and takes extract_kernel.
using AosLaunchFn = cudaError_t (*)(ExtractParams, cudaStream_t);
constexpr int kAosGroup = 2;

template <int NX, int AOS, bool HASH>
cudaError_t launch_extract_aos(ExtractParams p, cudaStream_t stream)
{
    constexpr int MODE = AOS == AOS_FWRV ? MODE_FWRV : MODE_FW;
    const uint64_t tiles = (p.items + kTileItems - 1) / kTileItems;
    if (tiles == 0) return cudaSuccess;
    set_iteration_strides(p, kAosGroup);
    if (!p.aligned || !p.vec_ok || p.items_dev || p.gprm >= 0x80000000ull || p.items >= 0xffffffffull - kTileItems || tiles > 0x7fffffffull ||
        !aligned_kernel_enabled())
        return cudaErrorNotSupported;
    p.al_magic = aligned_magic(p.gprm);
    p.pf_tiles = 0;
    if (prefetch_enabled()) {
        const uint64_t per_tile = static_cast<uint64_t>(p.nw32) * 4 / tiles + 1, t = kPfChunkBytes / per_tile;
        p.pf_tiles = static_cast<uint32_t>(t < 2 * kPfLead ? 2 * kPfLead : (t > (1u << 20) ? (1u << 20) : t));
    }
    extract_aligned_kernel<1, NX, MODE, HASH, SINK_STREAMS, 2, false, false, kAosGroup, AOS>
        <<<static_cast<unsigned>(tiles), kBlockThreads, 0, stream>>>(p);
    return cudaGetLastError();
}
AosLaunchFn get_aos_launcher_n1(int nx, bool fwrv, bool hash); // extract_n1.cu

// table lookup implemented in extract_n{1,2,3,4}.cu
ExtractLaunchFn get_extract_launcher_n1(int nx, int mode, bool hash, bool ragged);
ExtractLaunchFn get_extract_launcher_n2(int nx, int mode, bool hash, bool ragged);
ExtractLaunchFn get_extract_launcher_n3(int nx, int mode, bool hash, bool ragged);
ExtractLaunchFn get_extract_launcher_n4(int nx, int mode, bool hash, bool ragged);
// mode == -1 selects the bucket-count sink (canonical + fx_hash -> table)
constexpr int MODE_BUCKETS = -1;
constexpr int MODE_BUCKET_IDS = -2; // canonical + fx_hash -> flat array of 32-bit bucket ids

// One translation unit per N instantiates its 3 NX variants x 3 modes x hash x locator.
#define KMC_DEFINE_LAUNCHER_TABLE(FN, N)                                                            \
    template <int NX, int MODE, bool HASH>                                                          \
    static ExtractLaunchFn pick_loc_##N(bool ragged)                                                \
    {                                                                                               \
        return ragged ? &launch_extract<N, NX, MODE, HASH, true> : &launch_extract<N, NX, MODE, HASH, false>; \
    }                                                                                               \
    template <int NX, int MODE>                                                                     \
    static ExtractLaunchFn pick_hash_##N(bool hash, bool ragged)                                    \
    {                                                                                               \
        return hash ? pick_loc_##N<NX, MODE, true>(ragged) : pick_loc_##N<NX, MODE, false>(ragged); \
    }                                                                                               \
    template <int NX>                                                                               \
    static ExtractLaunchFn pick_mode_##N(int mode, bool hash, bool ragged)                          \
    {                                                                                               \
        switch (mode) {                                                                             \
        case MODE_BUCKETS:                                                                          \
            return ragged ? &launch_extract<N, NX, MODE_CANON, true, true, SINK_BUCKETS>            \
                          : &launch_extract<N, NX, MODE_CANON, true, false, SINK_BUCKETS>;          \
        case MODE_BUCKET_IDS:                                                                       \
            return ragged ? &launch_extract<N, NX, MODE_CANON, true, true, SINK_IDS>                \
                          : &launch_extract<N, NX, MODE_CANON, true, false, SINK_IDS>;              \
        case MODE_FW: return pick_hash_##N<NX, MODE_FW>(hash, ragged);                              \
        case MODE_FWRV: return pick_hash_##N<NX, MODE_FWRV>(hash, ragged);                          \
        case MODE_CANON: return pick_hash_##N<NX, MODE_CANON>(hash, ragged);                        \
        }                                                                                           \
        return nullptr;                                                                             \
    }                                                                                               \
    ExtractLaunchFn FN(int nx, int mode, bool hash, bool ragged)                                    \
    {                                                                                               \
        constexpr int NXMAX = (64 * N + 2 * GroupOf<N>::G - 2 + 31) / 32;                           \
        if (nx == NXMAX) return pick_mode_##N<NXMAX>(mode, hash, ragged);                           \
        if (nx == NXMAX - 1) return pick_mode_##N<(NXMAX - 1 > 0 ? NXMAX - 1 : 1)>(mode, hash, ragged); \
        if (nx == NXMAX - 2) return pick_mode_##N<(NXMAX - 2 > 0 ? NXMAX - 2 : 1)>(mode, hash, ragged); \
        return nullptr;                                                                             \
    }

// The same kernels with the fingerprint of their output fused in (KMC_DIGEST of the host pipeline): FwKmers and
// CanonicalKmers, SoA, N <= 2 (extract_n{1,2}.cu).
ExtractLaunchFn get_digest_launcher_n1(int nx, int mode, bool hash, bool ragged);
ExtractLaunchFn get_digest_launcher_n2(int nx, int mode, bool hash, bool ragged);

#define KMC_DEFINE_DIGEST_TABLE(FN, N)                                                              \
    template <int NX, int MODE>                                                                     \
    static ExtractLaunchFn pickdg_##N(bool hash, bool ragged)                                       \
    {                                                                                               \
        if (hash)                                                                                   \
            return ragged ? &launch_extract<N, NX, MODE, true, true, SINK_STREAMS, false, 2, true>  \
                          : &launch_extract<N, NX, MODE, true, false, SINK_STREAMS, false, 2, true>; \
        return ragged ? &launch_extract<N, NX, MODE, false, true, SINK_STREAMS, false, 2, true>     \
                      : &launch_extract<N, NX, MODE, false, false, SINK_STREAMS, false, 2, true>;   \
    }                                                                                               \
    template <int NX>                                                                               \
    static ExtractLaunchFn pickdg_mode_##N(int mode, bool hash, bool ragged)                        \
    {                                                                                               \
        switch (mode) {                                                                             \
        case MODE_FW: return pickdg_##N<NX, MODE_FW>(hash, ragged);                                 \
        case MODE_CANON: return pickdg_##N<NX, MODE_CANON>(hash, ragged);                           \
        }                                                                                           \
        return nullptr;                                                                             \
    }                                                                                               \
    ExtractLaunchFn FN(int nx, int mode, bool hash, bool ragged)                                    \
    {                                                                                               \
        constexpr int NXMAX = (64 * N + 2 * GroupOf<N>::G - 2 + 31) / 32;                           \
        if (nx == NXMAX) return pickdg_mode_##N<NXMAX>(mode, hash, ragged);                         \
        if (nx == NXMAX - 1) return pickdg_mode_##N<(NXMAX - 1 > 0 ? NXMAX - 1 : 1)>(mode, hash, ragged); \
        if (nx == NXMAX - 2) return pickdg_mode_##N<(NXMAX - 2 > 0 ? NXMAX - 2 : 1)>(mode, hash, ragged); \
        return nullptr;                                                                             \
    }

// Kmer{<:NucleicAcidAlphabet{4}} (BPS = 4: the Copyable 4 -> 4 scheme, and TwoToFour over an expanded
// stream), one translation unit per N (extract_b4_n{1,2,3,4}.cu): K <= 16 N, the three streaming modes.
ExtractLaunchFn get_kmer4_launcher_n1(int nx, int mode, bool hash, bool ragged);
ExtractLaunchFn get_kmer4_launcher_n2(int nx, int mode, bool hash, bool ragged);
ExtractLaunchFn get_kmer4_launcher_n3(int nx, int mode, bool hash, bool ragged);
ExtractLaunchFn get_kmer4_launcher_n4(int nx, int mode, bool hash, bool ragged);

#define KMC_DEFINE_KMER4_TABLE(FN, N)                                                               \
    template <int NX, int MODE>                                                                     \
    static ExtractLaunchFn pickb4_##N(bool hash, bool ragged)                                       \
    {                                                                                               \
        if (hash)                                                                                   \
            return ragged ? &launch_extract<N, NX, MODE, true, true, SINK_STREAMS, false, 4>        \
                          : &launch_extract<N, NX, MODE, true, false, SINK_STREAMS, false, 4>;      \
        return ragged ? &launch_extract<N, NX, MODE, false, true, SINK_STREAMS, false, 4>           \
                      : &launch_extract<N, NX, MODE, false, false, SINK_STREAMS, false, 4>;         \
    }                                                                                               \
    template <int NX>                                                                               \
    static ExtractLaunchFn pickb4_mode_##N(int mode, bool hash, bool ragged)                        \
    {                                                                                               \
        switch (mode) {                                                                             \
        case MODE_FW: return pickb4_##N<NX, MODE_FW>(hash, ragged);                                 \
        case MODE_FWRV: return pickb4_##N<NX, MODE_FWRV>(hash, ragged);                             \
        case MODE_CANON: return pickb4_##N<NX, MODE_CANON>(hash, ragged);                           \
        }                                                                                           \
        return nullptr;                                                                             \
    }                                                                                               \
    ExtractLaunchFn FN(int nx, int mode, bool hash, bool ragged)                                    \
    {                                                                                               \
        constexpr int NXMAX = (64 * N + 4 * GroupOf<N>::G - 4 + 31) / 32;                           \
        if (nx == NXMAX) return pickb4_mode_##N<NXMAX>(mode, hash, ragged);                         \
        if (nx == NXMAX - 1) return pickb4_mode_##N<(NXMAX - 1 > 0 ? NXMAX - 1 : 1)>(mode, hash, ragged); \
        if (nx == NXMAX - 2) return pickb4_mode_##N<(NXMAX - 2 > 0 ? NXMAX - 2 : 1)>(mode, hash, ragged); \
        return nullptr;                                                                             \
    }

// 4-bit (FourToTwo) launchers, one translation unit per N (extract4_n{1,2,3,4}.cu):
//   strict FwKmers / FwRvIterator / CanonicalKmers with the uncertain-symbol check.
// (UnambiguousKmers over a 4-bit source runs the ordinary ragged kernels over a run list, fourbit.cu.)
ExtractLaunchFn get_strict4_launcher_n1(int nx, int mode, bool hash, bool ragged);
ExtractLaunchFn get_strict4_launcher_n2(int nx, int mode, bool hash, bool ragged);
ExtractLaunchFn get_strict4_launcher_n3(int nx, int mode, bool hash, bool ragged);
ExtractLaunchFn get_strict4_launcher_n4(int nx, int mode, bool hash, bool ragged);

#define KMC_DEFINE_FOURBIT_TABLES(FN_STRICT, N)                                         \
    template <int NX, int MODE>                                                                     \
    static ExtractLaunchFn pick4_##N(bool hash, bool ragged)                                        \
    {                                                                                               \
        if (hash)                                                                                   \
            return ragged ? &launch_extract<N, NX, MODE, true, true, SINK_STREAMS, true>            \
                          : &launch_extract<N, NX, MODE, true, false, SINK_STREAMS, true>;          \
        return ragged ? &launch_extract<N, NX, MODE, false, true, SINK_STREAMS, true>               \
                      : &launch_extract<N, NX, MODE, false, false, SINK_STREAMS, true>;             \
    }                                                                                               \
    template <int NX>                                                                               \
    static ExtractLaunchFn pick4_mode_##N(int mode, bool hash, bool ragged)                         \
    {                                                                                               \
        switch (mode) {                                                                             \
        case MODE_FW: return pick4_##N<NX, MODE_FW>(hash, ragged);                                  \
        case MODE_FWRV: return pick4_##N<NX, MODE_FWRV>(hash, ragged);                              \
        case MODE_CANON: return pick4_##N<NX, MODE_CANON>(hash, ragged);                            \
        }                                                                                           \
        return nullptr;                                                                             \
    }                                                                                               \
    ExtractLaunchFn FN_STRICT(int nx, int mode, bool hash, bool ragged)                             \
    {                                                                                               \
        constexpr int NXMAX = (64 * N + 2 * GroupOf<N>::G - 2 + 31) / 32;                           \
        if (nx == NXMAX) return pick4_mode_##N<NXMAX>(mode, hash, ragged);                          \
        if (nx == NXMAX - 1) return pick4_mode_##N<(NXMAX - 1 > 0 ? NXMAX - 1 : 1)>(mode, hash, ragged); \
        if (nx == NXMAX - 2) return pick4_mode_##N<(NXMAX - 2 > 0 ? NXMAX - 2 : 1)>(mode, hash, ragged); \
        return nullptr;                                                                             \
    }

} // namespace kmc
