// lincompact.cuh -- UnambiguousKmers over a recoded (4-bit or ASCII) source (UnambiguousKmers.jl:109-148)
// as a stream compaction in SOURCE order.
//
// The iterator emits, in order, every window whose K symbols are all certain, with its 1-based start.  When
// the sequences of a set lie in the source buffer in ascending, non-overlapping order (every set the host
// mirrors build; checked, see lin_prepare in valid_count.cu) the output order IS the order of the window
// starts in the buffer, and the whole job is "compact the set bits of one bit array":
//
//   lin_prepare   lays the candidate window starts out as POSITIONS p = 0, 1, 2, ... and leaves one bit per
//                 position: "a window of a sequence starts here and its K symbols are certain".
//                   uniform sets   position p = (sequence r, offset u) with u < W8 = windows per sequence
//                                  rounded up to the item width G: the tails (last K-1 symbols) and the padding
//                                  between sequences are not positions at all, and an item of G positions never
//                                  straddles two sequences;
//                   sets with      position = symbol index in the recoded stream; the valid-start bits are
//                   offsets        masked in place down to the window starts of the sequences.
//                 The positions are cut into chunks of 2048; one popcount per chunk and ONE small exclusive scan
//                 give every chunk its place in the output -- and the total, before a single k-mer exists;
//   this kernel   one warp per chunk, no cooperation between warps or blocks at all: G positions per lane and
//                 step, k-mers from the closed form of extract_kernel (one block load, static funnel shifts),
//                 survivors compacted through a per-warp staging buffer so that the streams go out as ALIGNED
//                 256-bit stores whatever the survivor pattern is.
//
// It replaces compact_kernel's flat-window walk (two item locates per work item, a block scan between two
// barriers and a decoupled look-back that spins on its predecessors), with no forward-progress assumption about
// the block scheduler.  compact_kernel stays for sets whose sequences overlap or are out of order in the buffer.
#pragma once
#include "compact_kernels.cuh"

namespace kmc {

constexpr int kLinChunkPos = 2048;                 // positions per warp chunk
constexpr int kLinChunkWords = kLinChunkPos / 32;  // u32 words of position bits per chunk
constexpr int kLinWarps = kBlockThreads / 32;      // chunks per block

#ifndef KMC_LIN_MIN_BLOCKS
#define KMC_LIN_MIN_BLOCKS 4
#endif

struct LinParams {
    const uint32_t *bits;        // one bit per position (uniform sets: packed by lin_prepare; offsets: the masked valid-start bits)
    const uint64_t *chunk_off;   // [n_chunks + 1] exclusive scan of the survivors per chunk
    const uint64_t *chunk_first; // offsets given: [n_chunks + 1] sequence that owns the chunk's first symbol (0 if none yet)
    uint64_t n_chunks;
    uint64_t capacity;           // elements the output buffers hold; nothing is written beyond
    uint64_t stride_syms;        // uniform sets: symbols from one sequence's start to the next
    uint32_t w8;                 // uniform sets: positions per sequence (windows rounded up to G)
    float inv_w8;                // a little below 1 / w8 (used when w8 < 4096)
    uint32_t spu;                // symbols per offset unit (16: 4-bit source words, 1: ASCII bytes)
};

// staging index: two pad words per 16 keep pairs 16-byte aligned (128-bit shared-memory accesses) and spread
// both the lanes' runs (G*E words apart) and the aligned quads of the read-out over all banks
KMC_DEV uint32_t lin_slot(uint32_t w) { return w + 2u * (w >> 4); }

constexpr int lin_stage_words(int n)
{
    const int w = 32 * group_of(n) * (n + 1) + 8; // + the misalignment of the first word and the read-out's last quad
    return ((w + 2 * (w >> 4) + 4) + 1) & ~1;
}

KMC_DEV void sts128_if(uint32_t saddr, uint64_t a, uint64_t b, uint32_t on)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p st.shared.v2.u64 [%0], {%1, %2};\n\t}" ::"r"(saddr), "l"(a), "l"(b), "r"(on)
                 : "memory");
}

// Where this lane's survivors go in the staging buffer of a stream with E words per element whose first output
// word is misaligned by `a` words against 32 bytes: slot j (if it survives) -> shared-memory byte address addr[j].
// Computed once per step and used by every stream of the same E (the k-mer, index and hash streams of the SoA
// layout share them: their base pointers are 32-byte aligned, so they share the misalignment too).
template <int E, int G>
KMC_DEV void lin_stage_addrs(uint32_t sbase, uint32_t a, uint32_t x, uint32_t m, uint32_t (&addr)[G])
{
    uint32_t w = a + x * E;
#pragma unroll
    for (int j = 0; j < G; ++j) {
        addr[j] = sbase + 8u * lin_slot(w);
        w += ((m >> j) & 1u) * E;
    }
}

// One stream of one warp step: the lane's G elements v (E words each) to the staging buffer at addr[], then the
// warp's c * E words -- elements [o, o + c) of the stream -- to global memory: aligned quads as 256-bit stores, the
// at most three words before the first and after the last aligned quad one by one.
template <int E, int G>
KMC_DEV void lin_emit(uint64_t *__restrict__ gbase, uint64_t o, uint32_t c, uint32_t a, uint32_t m, const uint32_t (&addr)[G],
                      const uint64_t (&v)[G * E], const uint64_t *__restrict__ stage, uint32_t sbase, int lane)
{
    const bool wide = (E % 2 == 0) && (a & 1u) == 0; // warp-uniform: elements are 16-byte aligned in the staging buffer
#pragma unroll
    for (int j = 0; j < G; ++j) {
        const uint32_t on = (m >> j) & 1u;
        if (E == 1) {
            sts64_if(addr[j], v[j], on);
        } else if (E == 2 && wide) {
            sts128_if(addr[j], v[2 * j], v[2 * j + 1], on); // an aligned pair never straddles a 16-word run
        } else {
            // word i of the element lies i words on, plus the two pad words if it has crossed into the next 16-word
            // run: a padded index is 18 * (w >> 4) + (w & 15), so the element's place in its run is (index mod 18)
            const uint32_t r15 = ((addr[j] - sbase) >> 3) % 18u;
            if (wide) {
#pragma unroll
                for (int i = 0; i + 1 < E; i += 2)
                    sts128_if(addr[j] + 8u * (i + 2u * ((r15 + i) >> 4)), v[j * E + i], v[j * E + i + 1], on);
            } else {
#pragma unroll
                for (int i = 0; i < E; ++i) sts64_if(addr[j] + 8u * (i + 2u * ((r15 + i) >> 4)), v[j * E + i], on);
            }
        }
    }
    __syncwarp();
    const uint32_t end = a + c * E;
    uint64_t *g0 = gbase + (o * E - a); // 32-byte aligned
    const uint32_t q_lo = (a + 3u) & ~3u, q_hi = end & ~3u;
    for (uint32_t t = q_lo + 4u * lane; t < q_hi; t += 128u) {
        const uint32_t s = lin_slot(t); // t is a multiple of 4: the quad lies inside one 16-word run, 16-byte aligned
        const ulonglong2 lo = *reinterpret_cast<const ulonglong2 *>(stage + s);
        const ulonglong2 hi = *reinterpret_cast<const ulonglong2 *>(stage + s + 2);
        st_v4(g0 + t, lo.x, lo.y, hi.x, hi.y);
    }
    if (lane < 6) { // head [a, min(q_lo, end)) by lanes 0-2, tail [max(q_hi, q_lo), end) by lanes 3-5
        const uint32_t t = lane < 3 ? a + lane : (q_hi > q_lo ? q_hi : q_lo) + (lane - 3);
        const uint32_t lim = lane < 3 ? (q_lo < end ? q_lo : end) : end;
        if (t < lim) st_u64(g0 + t, stage[lin_slot(t)]);
    }
    __syncwarp();
}

// OFFSETS = the set gives per-sequence offsets (seq_unit_off) and positions are symbols of the stream; otherwise
// position p = (sequence p / w8, offset p % w8) and sequence r starts at symbol r * stride_syms + first
template <int N, int NX, bool HASH, bool OFFSETS>
__global__ void __launch_bounds__(kBlockThreads, KMC_LIN_MIN_BLOCKS) lin_compact_kernel(const ExtractParams p, const LinParams lp)
{
    constexpr int G = GroupOf<N>::G;
    constexpr int LPW = 32 / G;                       // lanes that share one word of position bits
    constexpr int ITERS = kLinChunkPos / (32 * G);    // warp steps per chunk
    extern __shared__ __align__(16) uint64_t s_lin_stage[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t *stage = s_lin_stage + static_cast<size_t>(warp) * lin_stage_words(N);
    const uint32_t sbase = static_cast<uint32_t>(__cvta_generic_to_shared(stage));

    const uint64_t c = static_cast<uint64_t>(blockIdx.x) * kLinWarps + warp;
    if (c >= lp.n_chunks) return;
    uint64_t o = __ldg(lp.chunk_off + c);
    if (__ldg(lp.chunk_off + c + 1) == o) return; // nothing survives in this chunk (warp-uniform)
    const uint32_t *__restrict__ vs = lp.bits + c * kLinChunkWords;
    const uint64_t pos_c = c * kLinChunkPos;
    const bool tuple_ix = p.aos != 0;
    // misalignment (in words, against 32 bytes) of the streams' base pointers
    const uint32_t a_a = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(p.out_a) >> 3) & 3u;
    const uint32_t a_i = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(p.out_index) >> 3) & 3u;
    const uint32_t a_h = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(p.out_hash) >> 3) & 3u;

    // where the chunk lies among the sequences
    uint64_t r_lo = 0, r_hi = 0; // OFFSETS: the sequences that can own symbols of this chunk
    uint32_t u_c = 0;            // uniform: offset of the chunk's first position in its sequence
    uint64_t sym_rc = 0;         // uniform: first symbol (stream index) of that sequence
    if (OFFSETS) {
        r_lo = __ldg(lp.chunk_first + c);
        r_hi = __ldg(lp.chunk_first + c + 1);
    } else {
        const uint64_t r_c = (pos_c >> 32) == 0 ? static_cast<uint32_t>(pos_c) / lp.w8 : pos_c / lp.w8;
        u_c = static_cast<uint32_t>(pos_c - r_c * lp.w8);
        sym_rc = r_c * lp.stride_syms + p.first;
    }
    auto seq_start = [&](uint64_t r) -> uint64_t { // first symbol of sequence r in the stream
        return (__ldg(p.seq_unit_off + r) - p.unit_bias) * lp.spu + p.first;
    };
    // stream symbol of slot 0 of step `it` (uniform: and the slot's offset u in its sequence)
    auto item_symbol = [&](int it, uint32_t &u) -> uint64_t {
        const uint32_t off_it = static_cast<uint32_t>(it) * 32 * G + static_cast<uint32_t>(lane) * G;
        if (OFFSETS) {
            u = 0;
            return pos_c + off_it;
        }
        const uint32_t off = u_c + off_it; // < w8 + 2048
        uint32_t q;
        if (lp.w8 >= 4096) {
            q = off >= lp.w8 ? 1u : 0u;
        } else {
            q = static_cast<uint32_t>(static_cast<float>(off) * lp.inv_w8); // <= the quotient, at most 1 below it
            if (off - q * lp.w8 >= lp.w8) ++q;
        }
        u = off - q * lp.w8;
        return sym_rc + static_cast<uint64_t>(q) * lp.stride_syms + u;
    };

    // software pipeline: the position bits and the source words of the next step are requested before this step's
    // k-mers are computed
    uint32_t v_next = __ldg(vs + lane / LPW);
    uint32_t raw_next[NX + 1], u_next;
    uint64_t sym_next = item_symbol(0, u_next);
    load_raw<NX>(p.w32, p.nw32, static_cast<int64_t>(2 * sym_next), raw_next);
    const uint32_t sub = static_cast<uint32_t>(lane) % LPW;

#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        const uint32_t v = v_next, u0 = u_next;
        const uint64_t sym0 = sym_next;
        uint32_t raw[NX + 1];
#pragma unroll
        for (int i = 0; i <= NX; ++i) raw[i] = raw_next[i];
        if (it + 1 < ITERS) {
            v_next = __ldg(vs + (it + 1) * G + lane / LPW);
            sym_next = item_symbol(it + 1, u_next);
            load_raw<NX>(p.w32, p.nw32, static_cast<int64_t>(2 * sym_next), raw_next);
        }
        // survivors of the step, and of the lanes before this one: the LPW lanes of a word hold the same popcount
        const uint32_t pw = __popc(v);
        uint32_t incl = pw;
#pragma unroll
        for (int d = LPW; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const uint32_t cnt = __shfl_sync(0xffffffffu, incl, 31);
        if (cnt == 0) continue; // warp-uniform
        const uint32_t m = (v >> (G * sub)) & ((1u << G) - 1u);
        const uint32_t x = incl - pw + __popc(v & ((1u << (G * sub)) - 1u));
        uint32_t cc = cnt;
        if (o + cnt > lp.capacity) cc = o < lp.capacity ? static_cast<uint32_t>(lp.capacity - o) : 0u;
        const uint64_t o_it = o;
        o += cnt;

        const uint32_t shift = (2u * static_cast<uint32_t>(sym0)) & 31u;
        uint32_t xw[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) xw[i] = __funnelshift_r(raw[i], raw[i + 1], shift);
        uint64_t fw[G][N], rv[G][N];
        block_kmers<N, NX, G, true, false>(xw, p.s0, p.head_mask, fw, rv); // lanes without survivors: never staged

        // 1-based start of every slot inside its sequence
        int64_t ib[G];
        if (OFFSETS) {
            if (m) {
                uint64_t lo = r_lo, hi = r_hi + 1; // largest r in [lo, hi) whose first symbol is <= sym0 (r_lo if none)
                while (hi - lo > 1) {
                    const uint64_t mid = (lo + hi) >> 1;
                    if (seq_start(mid) <= sym0) lo = mid; else hi = mid;
                }
                uint64_t r = lo, s_r = seq_start(r);
                uint64_t next = r + 1 < p.n_seqs ? seq_start(r + 1) : ~0ull;
                if (sym0 + G <= next) {
                    const int64_t b = static_cast<int64_t>(sym0 - s_r) + 1 + p.index_base;
#pragma unroll
                    for (int j = 0; j < G; ++j) ib[j] = b + j;
                } else { // the slots straddle sequences
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        while (sym0 + j >= next) {
                            ++r;
                            s_r = next;
                            next = r + 1 < p.n_seqs ? seq_start(r + 1) : ~0ull;
                        }
                        ib[j] = static_cast<int64_t>(sym0 + j - s_r) + 1 + p.index_base;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < G; ++j) ib[j] = 0;
            }
        } else {
            const int64_t b = static_cast<int64_t>(u0) + 1 + p.index_base; // an item never straddles two sequences
#pragma unroll
            for (int j = 0; j < G; ++j) ib[j] = b + j;
        }

        if (tuple_ix) {
            // Vector{Tuple{Kmer,Int}}: {u64[N]; i64} elements
            uint64_t buf[G * (N + 1)];
#pragma unroll
            for (int j = 0; j < G; ++j) {
#pragma unroll
                for (int i = 0; i < N; ++i) buf[j * (N + 1) + i] = fw[j][i];
                buf[j * (N + 1) + N] = static_cast<uint64_t>(ib[j]);
            }
            const uint32_t a = (a_a + static_cast<uint32_t>(o_it) * (N + 1)) & 3u;
            uint32_t addr[G];
            lin_stage_addrs<N + 1, G>(sbase, a, x, m, addr);
            lin_emit<N + 1, G>(p.out_a, o_it, cc, a, m, addr, buf, stage, sbase, lane);
            if (HASH) {
                uint64_t h[G];
#pragma unroll
                for (int j = 0; j < G; ++j) h[j] = fx_hash<N>(fw[j], 0);
                const uint32_t ah = (a_h + static_cast<uint32_t>(o_it)) & 3u;
                lin_stage_addrs<1, G>(sbase, ah, x, m, addr);
                lin_emit<1, G>(p.out_hash, o_it, cc, ah, m, addr, h, stage, sbase, lane);
            }
        } else {
            uint64_t buf[G * N];
#pragma unroll
            for (int j = 0; j < G; ++j)
#pragma unroll
                for (int i = 0; i < N; ++i) buf[j * N + i] = fw[j][i];
            const uint32_t a = (a_a + static_cast<uint32_t>(o_it) * N) & 3u;
            uint32_t addr[G];
            lin_stage_addrs<N, G>(sbase, a, x, m, addr);
            lin_emit<N, G>(p.out_a, o_it, cc, a, m, addr, buf, stage, sbase, lane);
            // the one-word streams: the index, and the hash
            uint64_t iw[G];
#pragma unroll
            for (int j = 0; j < G; ++j) iw[j] = static_cast<uint64_t>(ib[j]);
            const uint32_t ai = (a_i + static_cast<uint32_t>(o_it)) & 3u;
            if (N != 1 || ai != a) lin_stage_addrs<1, G>(sbase, ai, x, m, addr);
            lin_emit<1, G>(reinterpret_cast<uint64_t *>(p.out_index), o_it, cc, ai, m, addr, iw, stage, sbase, lane);
            if (HASH) {
                uint64_t h[G];
#pragma unroll
                for (int j = 0; j < G; ++j) h[j] = fx_hash<N>(fw[j], 0);
                const uint32_t ah = (a_h + static_cast<uint32_t>(o_it)) & 3u;
                if (ah != ai) lin_stage_addrs<1, G>(sbase, ah, x, m, addr);
                lin_emit<1, G>(p.out_hash, o_it, cc, ah, m, addr, h, stage, sbase, lane);
            }
        }
    }
}

using LinLaunchFn = cudaError_t (*)(ExtractParams, LinParams, cudaStream_t);

template <int N, int NX, bool HASH, bool OFFSETS>
cudaError_t launch_lin_compact(ExtractParams p, LinParams lp, cudaStream_t stream)
{
    if (lp.n_chunks == 0) return cudaSuccess;
    const uint64_t blocks = (lp.n_chunks + kLinWarps - 1) / kLinWarps;
    if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    constexpr size_t smem = static_cast<size_t>(kLinWarps) * lin_stage_words(N) * sizeof(uint64_t);
    cudaError_t e = cudaFuncSetAttribute(lin_compact_kernel<N, NX, HASH, OFFSETS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    lin_compact_kernel<N, NX, HASH, OFFSETS><<<static_cast<unsigned>(blocks), kBlockThreads, smem, stream>>>(p, lp);
    return cudaGetLastError();
}

LinLaunchFn get_lin_launcher_n1(int nx, bool hash, bool offsets);
LinLaunchFn get_lin_launcher_n2(int nx, bool hash, bool offsets);
LinLaunchFn get_lin_launcher_n3(int nx, bool hash, bool offsets);
LinLaunchFn get_lin_launcher_n4(int nx, bool hash, bool offsets);

#define KMC_DEFINE_LIN_TABLE(FN, N)                                                                 \
    template <int NX>                                                                               \
    static LinLaunchFn pickl_##N(bool hash, bool offsets)                                           \
    {                                                                                               \
        if (hash) return offsets ? &launch_lin_compact<N, NX, true, true> : &launch_lin_compact<N, NX, true, false>; \
        return offsets ? &launch_lin_compact<N, NX, false, true> : &launch_lin_compact<N, NX, false, false>; \
    }                                                                                               \
    LinLaunchFn FN(int nx, bool hash, bool offsets)                                                 \
    {                                                                                               \
        constexpr int NXMAX = (64 * N + 2 * GroupOf<N>::G - 2 + 31) / 32;                           \
        if (nx == NXMAX) return pickl_##N<NXMAX>(hash, offsets);                                    \
        if (nx == NXMAX - 1) return pickl_##N<(NXMAX - 1 > 0 ? NXMAX - 1 : 1)>(hash, offsets);      \
        if (nx == NXMAX - 2) return pickl_##N<(NXMAX - 2 > 0 ? NXMAX - 2 : 1)>(hash, offsets);      \
        return nullptr;                                                                             \
    }

} // namespace kmc
