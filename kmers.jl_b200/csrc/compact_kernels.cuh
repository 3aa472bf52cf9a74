// compact_kernels.cuh -- UnambiguousKmers over a recoded (4-bit or ASCII) source
// (UnambiguousKmers.jl:109-148) as ONE ordered stream compaction over the FLAT WINDOW order of the set.
// This is the general path: it handles any layout of the sequences in the buffer (overlapping views, any
// order).  Sets whose sequences are ascending and disjoint in the buffer -- all the host mirrors build -- take
// the cheaper source-order compaction of lincompact.cuh instead.
//
// The iterator emits, in order, every window whose K symbols are all certain, with its 1-based
// start.  The recoding pass (fourbit.cu / ascii.cu) has left a 2-bit stream and one "valid start"
// bit per symbol; this kernel walks the same (read, group slot) work items as extract_kernel, and
//
//   phase 1  reads only the valid-start bits of the tile's items: survivors per (iteration, warp),
//            one block scan, and the tile's place in the output by a decoupled look-back over the
//            tiles before it (single pass: no count kernel, no host round trip, no run list);
//   phase 2  computes the G windows of every item in registers exactly as extract_kernel does,
//            and each warp compacts the survivors of its 32 items through a shared-memory staging
//            buffer, so that the stream goes out as ALIGNED 256-bit stores whatever the survivor
//            pattern is (a direct scatter would write 8 bytes per 32-byte sector).
//
// Staging layout: the warp's words of one stream, in output order, at word index a + w where a is
// the misalignment (in words) of the warp's first output word against 32 bytes, plus one pad word
// per 16 (w + (w >> 4)): lanes write runs that start G*E words apart and lanes read aligned quads,
// and the pad makes both 2-way (= minimal for 64-bit accesses) instead of 16-way bank conflicts.
#pragma once
#include "extract_kernels.cuh"

namespace kmc {

struct CompactParams {
    unsigned long long *tile_state; // [tiles + 1], zeroed before the launch: flag (2 bits) | count (62 bits); the last
                                    // entry is the ticket counter the blocks draw their tile from
    uint64_t *total_out;            // device: number of k-mers the set emits (written by the last tile)
    uint64_t capacity;              // elements the output buffers hold; nothing is written beyond
};

// Resident blocks per SM the register allocation must allow.  The kernel waits on its two barriers around the
// look-back and on its loads, so it wants warps: 5 blocks (<= 51 registers, no spills) measured 4-6 % faster than
// the 4 the compiler picks on its own (56 registers); 6 needs spills and is slower than 4.
#ifndef KMC_COMPACT_MIN_BLOCKS
#define KMC_COMPACT_MIN_BLOCKS 5
#endif

constexpr uint64_t kTileAggregate = 1ull << 62, kTilePrefix = 2ull << 62, kTileValue = (1ull << 62) - 1;
constexpr int kWarpsPerBlock = kBlockThreads / 32;

KMC_DEV uint64_t ld_relaxed_u64(const unsigned long long *p)
{
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
KMC_DEV void st_relaxed_u64(unsigned long long *p, uint64_t v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

KMC_DEV uint32_t stage_slot(uint32_t w) { return w + (w >> 4); }

// predicated 64-bit store to shared memory (keeps the per-slot staging branch-free)
KMC_DEV void sts64_if(uint32_t saddr, uint64_t v, uint32_t on)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u64 [%0], %1;\n\t}" ::"r"(saddr), "l"(v), "r"(on)
                 : "memory");
}

// words of staging one warp needs for N limbs (the widest stream is the Tuple{Kmer,Int} element)
constexpr int stage_words(int n)
{
    const int w = 32 * group_of(n) * (n + 1) + 4;
    return w + (w >> 4) + 1;
}

template <typename T> KMC_DEV T warp_sum(T v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// One stream of one warp step: E words per element; the warp's survivors are the elements
// [o, o + c) of the stream; this lane's survivors (bits of m, G slots) start at element o + x.
// v holds the lane's G elements.  Any 8-byte aligned stream base works: the quads are aligned to
// the ADDRESS of the warp's first word, not to its index.
template <int E, int G>
KMC_DEV void warp_emit(uint64_t *__restrict__ gbase, uint64_t o, uint32_t c, uint32_t x, uint32_t m,
                       const uint64_t (&v)[G * E], uint64_t *__restrict__ stage, int lane)
{
    const uint64_t w0 = o * E;
    const uint32_t a = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(gbase + w0) >> 3) & 3u;
    const uint32_t end = a + c * E;
    uint32_t w = a + x * E;
    {
        const uint32_t sbase = static_cast<uint32_t>(__cvta_generic_to_shared(stage));
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const uint32_t on = (m >> j) & 1u;
#pragma unroll
            for (int i = 0; i < E; ++i) sts64_if(sbase + 8u * stage_slot(w + i), v[j * E + i], on);
            w += on * E;
        }
    }
    __syncwarp();
    uint64_t *g0 = gbase + (w0 - a);
    for (uint32_t t = 4u * lane; t < end; t += 128u) {
        const uint32_t s = stage_slot(t); // t is a multiple of 4: the quad shares one pad offset
        const uint64_t q0 = stage[s], q1 = stage[s + 1], q2 = stage[s + 2], q3 = stage[s + 3];
        if (t >= a && t + 4 <= end) {
            st_v4(g0 + t, q0, q1, q2, q3);
        } else {
            if (t >= a && t < end) st_u64(g0 + t, q0);
            if (t + 1 >= a && t + 1 < end) st_u64(g0 + t + 1, q1);
            if (t + 2 >= a && t + 2 < end) st_u64(g0 + t + 2, q2);
            if (t + 3 >= a && t + 3 < end) st_u64(g0 + t + 3, q3);
        }
    }
    __syncwarp();
}

template <int N, int NX, bool HASH, bool RAGGED>
__global__ void __launch_bounds__(kBlockThreads, KMC_COMPACT_MIN_BLOCKS) compact_kernel(const ExtractParams p, const CompactParams cp)
{
    constexpr int G = GroupOf<N>::G;
    constexpr int kCells = kTileIters * kWarpsPerBlock; // (iteration, warp) steps of a tile, in output order
    static_assert(kCells == 64, "the block scan below handles two cells per lane of warp 0");
    extern __shared__ uint64_t s_stage_all[];
    __shared__ TileShared<RAGGED> sh;
    __shared__ uint32_t s_cnt[kCells], s_base[kCells];
    __shared__ uint64_t s_tile_out;

    __shared__ uint32_t s_tile;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t *stage = s_stage_all + static_cast<size_t>(warp) * stage_words(N);
    // The tile comes from a ticket, not from blockIdx.x: the look-back below waits for the tiles before this one,
    // and a ticket guarantees that they were STARTED before it whatever order the hardware dispatches blocks in
    // (a block that has drawn its ticket publishes its aggregate without waiting for anything).
    if (threadIdx.x == 0) s_tile = static_cast<uint32_t>(atomicAdd(cp.tile_state + gridDim.x, 1ull));
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t tile_base = static_cast<uint64_t>(tile) * kTileItems;
    TileCursor<RAGGED, G> cur;
    cur.init(p, tile_base, sh, threadIdx.x, tile);
    const TileCursor<RAGGED, G> cur0 = cur;

    // ---- phase 1: survivors of every item (G <= 8 bits each, kept for phase 2) ------------------
    uint64_t masks = 0;
#pragma unroll 1
    for (int it = 0; it < kTileIters; ++it) {
        const uint32_t li = static_cast<uint32_t>(it) * kBlockThreads + threadIdx.x;
        const uint64_t item = tile_base + li;
        uint32_t m = 0;
        if (item < p.items) {
            cur.locate(p, item, li, sh);
            if (cur.jhi > cur.jlo) m = valid_slots(p.vstart, cur.bit(p) >> 1, cur.jlo, cur.jhi);
            cur.advance(p);
        }
        masks |= static_cast<uint64_t>(m) << (8 * it);
        const uint32_t n = __reduce_add_sync(0xffffffffu, __popc(m));
        if (lane == 0) s_cnt[it * kWarpsPerBlock + warp] = n;
    }
    __syncthreads();
    if (warp == 0) {
        const uint32_t v0 = s_cnt[lane], v1 = s_cnt[lane + 32];
        uint32_t i0 = v0, i1 = v1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t0 = __shfl_up_sync(0xffffffffu, i0, d), t1 = __shfl_up_sync(0xffffffffu, i1, d);
            if (lane >= d) {
                i0 += t0;
                i1 += t1;
            }
        }
        i1 += __shfl_sync(0xffffffffu, i0, 31);
        s_base[lane] = i0 - v0;
        s_base[lane + 32] = i1 - v1;
        const uint64_t tile_total = __shfl_sync(0xffffffffu, i1, 31);
        // decoupled look-back: every tile before this one has drawn its ticket earlier, so it is running or done
        // and publishes its aggregate without waiting for anything after it
        uint64_t excl = 0;
        if (tile > 0) {
            if (lane == 0) st_relaxed_u64(cp.tile_state + tile, kTileAggregate | tile_total);
            int64_t top = static_cast<int64_t>(tile) - 1;
            for (;;) {
                const int64_t idx = top - lane;
                uint64_t s;
                do {
                    s = idx >= 0 ? ld_relaxed_u64(cp.tile_state + idx) : kTilePrefix;
                } while (__any_sync(0xffffffffu, (s >> 62) == 0));
                const uint32_t pref = __ballot_sync(0xffffffffu, (s >> 62) == 2);
                const int stop = pref ? __ffs(pref) - 1 : 31; // nearest tile with a full prefix
                excl += warp_sum<uint64_t>(lane <= stop ? (s & kTileValue) : 0ull);
                if (pref) break;
                top -= 32;
            }
        }
        if (lane == 0) {
            st_relaxed_u64(cp.tile_state + tile, kTilePrefix | (excl + tile_total));
            s_tile_out = excl;
            if (tile == gridDim.x - 1) *cp.total_out = excl + tile_total;
        }
    }
    __syncthreads();
    const uint64_t tile_out = s_tile_out;

    // ---- phase 2: the windows, compacted warp by warp ------------------------------------------
    // The source words of item it+1 are requested before item it is processed, so their latency
    // hides behind a whole step of staging and stores.
    cur = cur0;
    const bool tuple_ix = p.aos != 0;
    uint32_t raw_next[NX + 1], sh_next = 0;
    int64_t ibase_next = 0;
    auto prepare = [&](int it) {
        const uint32_t li = static_cast<uint32_t>(it) * kBlockThreads + threadIdx.x;
        if ((masks >> (8 * it)) & 0xffull) {
            cur.locate(p, tile_base + li, li, sh);
            const int64_t bit = cur.bit(p);
            load_raw<NX>(p.w32, p.nw32, bit, raw_next);
            sh_next = static_cast<uint32_t>(bit) & 31u;
            ibase_next = cur.wbase + 1 + p.index_base + static_cast<int64_t>(cur.seq_ibase);
        }
        cur.advance(p);
    };
    prepare(0);
#pragma unroll 1
    for (int it = 0; it < kTileIters; ++it) {
        uint32_t raw[NX + 1];
#pragma unroll
        for (int i = 0; i <= NX; ++i) raw[i] = raw_next[i];
        const uint32_t shift = sh_next;
        const int64_t ibase = ibase_next;
        if (it + 1 < kTileIters) prepare(it + 1);

        const int cell = it * kWarpsPerBlock + warp;
        uint32_t c = s_cnt[cell];
        if (c == 0) continue; // warp-uniform
        const uint64_t o = tile_out + s_base[cell];
        if (o + c > cp.capacity) c = o < cp.capacity ? static_cast<uint32_t>(cp.capacity - o) : 0u;
        const uint32_t m = static_cast<uint32_t>(masks >> (8 * it)) & 0xffu;
        uint32_t x = __popc(m); // -> exclusive prefix of the lanes' survivor counts
        {
            uint32_t incl = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            x = incl - x;
        }
        uint32_t xw[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) xw[i] = __funnelshift_r(raw[i], raw[i + 1], shift);
        uint64_t fw[G][N], rv[G][N];
        block_kmers<N, NX, G, true, false>(xw, p.s0, p.head_mask, fw, rv); // lanes without survivors: never staged
        if (tuple_ix) {
            // Vector{Tuple{Kmer,Int}}: {u64[N]; i64} elements
            uint64_t buf[G * (N + 1)];
#pragma unroll
            for (int j = 0; j < G; ++j) {
#pragma unroll
                for (int i = 0; i < N; ++i) buf[j * (N + 1) + i] = fw[j][i];
                buf[j * (N + 1) + N] = static_cast<uint64_t>(ibase + j);
            }
            warp_emit<N + 1, G>(p.out_a, o, c, x, m, buf, stage, lane);
        } else {
            uint64_t buf[G * N];
#pragma unroll
            for (int j = 0; j < G; ++j)
#pragma unroll
                for (int i = 0; i < N; ++i) buf[j * N + i] = fw[j][i];
            warp_emit<N, G>(p.out_a, o, c, x, m, buf, stage, lane);
            uint64_t ib[G];
#pragma unroll
            for (int j = 0; j < G; ++j) ib[j] = static_cast<uint64_t>(ibase + j);
            warp_emit<1, G>(reinterpret_cast<uint64_t *>(p.out_index), o, c, x, m, ib, stage, lane);
        }
        if (HASH) {
            uint64_t h[G];
#pragma unroll
            for (int j = 0; j < G; ++j) h[j] = fx_hash<N>(fw[j], 0);
            warp_emit<1, G>(p.out_hash, o, c, x, m, h, stage, lane);
        }
    }
}

using CompactLaunchFn = cudaError_t (*)(ExtractParams, CompactParams, cudaStream_t);

template <int N, int NX, bool HASH, bool RAGGED>
cudaError_t launch_compact(ExtractParams p, CompactParams cp, cudaStream_t stream)
{
    const uint64_t tiles = (p.items + kTileItems - 1) / kTileItems;
    if (tiles == 0) return cudaSuccess;
    if (tiles > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    set_iteration_strides(p);
    p.pf_tiles = 0;
    constexpr size_t smem = static_cast<size_t>(kWarpsPerBlock) * stage_words(N) * sizeof(uint64_t);
    // (set on every launch: the attribute belongs to the function on the CURRENT device, and a process may hold
    // contexts on several devices; the call costs a few microseconds)
    cudaError_t e = cudaFuncSetAttribute(compact_kernel<N, NX, HASH, RAGGED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    compact_kernel<N, NX, HASH, RAGGED><<<static_cast<unsigned>(tiles), kBlockThreads, smem, stream>>>(p, cp);
    return cudaGetLastError();
}

CompactLaunchFn get_compact_launcher_n1(int nx, bool hash, bool ragged);
CompactLaunchFn get_compact_launcher_n2(int nx, bool hash, bool ragged);
CompactLaunchFn get_compact_launcher_n3(int nx, bool hash, bool ragged);
CompactLaunchFn get_compact_launcher_n4(int nx, bool hash, bool ragged);

#define KMC_DEFINE_COMPACT_TABLE(FN, N)                                                             \
    template <int NX>                                                                               \
    static CompactLaunchFn pickc_##N(bool hash, bool ragged)                                        \
    {                                                                                               \
        if (hash) return ragged ? &launch_compact<N, NX, true, true> : &launch_compact<N, NX, true, false>; \
        return ragged ? &launch_compact<N, NX, false, true> : &launch_compact<N, NX, false, false>; \
    }                                                                                               \
    CompactLaunchFn FN(int nx, bool hash, bool ragged)                                              \
    {                                                                                               \
        constexpr int NXMAX = (64 * N + 2 * GroupOf<N>::G - 2 + 31) / 32;                           \
        if (nx == NXMAX) return pickc_##N<NXMAX>(hash, ragged);                                     \
        if (nx == NXMAX - 1) return pickc_##N<(NXMAX - 1 > 0 ? NXMAX - 1 : 1)>(hash, ragged);       \
        if (nx == NXMAX - 2) return pickc_##N<(NXMAX - 2 > 0 ? NXMAX - 2 : 1)>(hash, ragged);       \
        return nullptr;                                                                             \
    }

} // namespace kmc
