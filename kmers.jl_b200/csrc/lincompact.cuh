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
//                   uniform sets   position p = (sequence, window) = divmod(p, windows per sequence): the tails
//                                  (last K-1 symbols) and the padding between sequences are not positions at all;
//                   sets with      position = symbol index in the recoded stream; the valid-start bits are
//                   offsets        masked in place down to the window starts of the sequences.
//                 The positions are cut into chunks of 2048; one popcount per chunk and ONE small exclusive scan
//                 give every chunk its place in the output -- and the total, before a single k-mer exists;
//   this kernel   one warp per chunk, no cooperation between warps or blocks and NO shared memory: a step takes
//                 the 32 positions of one word of bits, lane l the l-th.  The survivors of a step are consecutive
//                 elements of the output, so a lane's rank among them (a popcount of the bits below it) is its
//                 place, and the warp's stores of a step are one contiguous run per stream.
//                 The k-mer of a position comes out of the REVERSED stream the recoding pass writes beside the
//                 forward one: symbol i at position T - 1 - i, so that the K symbols of a window, first symbol
//                 highest -- the Kmer layout (kmer.jl:32-51) -- are a plain run of 2K bits: three overlapping
//                 words (the lanes of a step read the same three or four words: one L1 wavefront each), two
//                 funnel shifts and a mask per limb.  No bit reversal, no per-window block to amortise.
//
// History (profiles/r02_c3_*): the flat-window look-back kernel (compact_kernels.cuh) ran 2.5 G warp instructions per
// 10 M reads with a block scan between two barriers; a first source-order kernel with G positions per lane and a
// staging buffer in shared memory got that to 2.3 G but moved the bound to the L1 data pipe (87-90 % busy: every
// survivor crosses shared memory once in each direction, 480 M wavefronts against 220 M ideal).  This form has no
// staging to pay for.  compact_kernel stays for sets whose sequences overlap or are out of order in the buffer.
#pragma once
#include "extract_kernels.cuh"

namespace kmc {

#ifndef KMC_LIN_MIN_BLOCKS
#define KMC_LIN_MIN_BLOCKS 1 // resident blocks per SM the compaction kernel's registers are capped for (A/B: -DKMC_LIN_MIN_BLOCKS=5)
#endif
constexpr int kLinChunkPos = 2048;                 // positions per warp chunk
constexpr int kLinChunkWords = kLinChunkPos / 32;  // u32 words of position bits per chunk
constexpr int kLinWarps = kBlockThreads / 32;      // chunks per block

struct LinParams {
    const uint32_t *bits;        // one bit per position (uniform sets: gathered by lin_prepare; offsets: the masked valid-start bits)
    const uint64_t *chunk_off;   // [n_chunks + 1] exclusive scan of the survivors per chunk
    const uint64_t *chunk_first; // offsets given: [n_chunks + 1] sequence that owns the chunk's first symbol (0 if none yet)
    const uint32_t *rev32;       // the 2-bit codes in reversed symbol order: symbol i at symbol position t_syms - 1 - i
    uint64_t t_syms;             // symbols in the reversed stream (a multiple of 32)
    uint64_t n_chunks;
    uint64_t capacity;           // elements the output buffers hold; nothing is written beyond
    uint64_t stride_syms;        // uniform sets: symbols from one sequence's start to the next
    uint32_t w8;                 // uniform sets: positions (= windows) per sequence
    uint32_t jump;               // uniform sets: stride_syms - w8, the symbols between two sequences' windows
    uint32_t spu;                // symbols per offset unit (16: 4-bit source words, 1: ASCII bytes)
};

// Predicated loads (plain asm: a pure function of its operands, which the compiler may schedule freely -- the kernel
// below issues the loads of the NEXT two steps before it finishes the current two) and predicated stores; one
// predicate per group of accesses.
template <int NP> KMC_DEV void ldg2_if(const uint2 *p, uint32_t on, uint2 (&v)[NP])
{
    static_assert(NP >= 2 && NP <= 5, "one to four limbs");
    if constexpr (NP == 2) {
        asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\tmov.u32 %0, 0;\n\tmov.u32 %1, 0;\n\tmov.u32 %2, 0;\n\tmov.u32 %3, 0;\n\t"
            "@q ld.global.nc.v2.u32 {%0,%1}, [%4];\n\t@q ld.global.nc.v2.u32 {%2,%3}, [%4+8];\n\t}"
            : "=r"(v[0].x), "=r"(v[0].y), "=r"(v[1].x), "=r"(v[1].y)
            : "l"(p), "r"(on));
    } else {
#pragma unroll
        for (int i = 0; i < NP; ++i)
            asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\tmov.u32 %0, 0;\n\tmov.u32 %1, 0;\n\t@q ld.global.nc.v2.u32 {%0,%1}, [%2];\n\t}"
                : "=r"(v[i].x), "=r"(v[i].y)
                : "l"(p + i), "r"(on));
    }
}
KMC_DEV void stg64_if(uint64_t *p, uint64_t v, uint32_t on)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.L1::no_allocate.L2::cache_hint.u64 [%0], %1, %3;\n\t}" ::"l"(p), "l"(v),
                 "r"(on), "l"(kEvictFirst));
}
// one word to each of two streams
KMC_DEV void stg64x2_if(uint64_t *p0, uint64_t v0, uint64_t *p1, uint64_t v1, uint32_t on)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %4, 0;\n\t@q st.global.L1::no_allocate.L2::cache_hint.u64 [%0], %1, %5;\n\t"
                 "@q st.global.L1::no_allocate.L2::cache_hint.u64 [%2], %3, %5;\n\t}" ::"l"(p0), "l"(v0), "l"(p1), "l"(v1), "r"(on), "l"(kEvictFirst));
}
KMC_DEV void stg128_if(uint64_t *p, uint64_t a, uint64_t b, uint32_t on)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q st.global.L1::no_allocate.L2::cache_hint.v2.u64 [%0], {%1,%2}, %4;\n\t}" ::"l"(p),
                 "l"(a), "l"(b), "r"(on), "l"(kEvictFirst));
}

#ifndef KMC_LIN_PREFETCH
#define KMC_LIN_PREFETCH 2 // 0: none, 1: into L1 (a load whose result nobody reads), 2: into L2
#endif
KMC_DEV void lin_prefetch(const char *p)
{
#if KMC_LIN_PREFETCH == 2
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#elif KMC_LIN_PREFETCH == 1
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(reinterpret_cast<uintptr_t>(p) & ~static_cast<uintptr_t>(3)));
#else
    (void)p;
#endif
}

constexpr uint32_t kLinGone = 0xffffffffu; // LinStep::at of a position that does not survive

// One step of one lane between the two halves of the pipeline: the words of its window are on their way
template <int N> struct LinStep {
    uint32_t at;      // its element of the chunk's output, or kLinGone
    uint32_t pos;     // uniform sets: its window inside its sequence; offsets: its symbol against the chunk's first
    int32_t d;        // bit offset of its window from w0 (bits 0..4: inside the first word; bit 5: that word is the odd one of its pair)
    uint32_t any;     // offsets: the step's 32 bits (warp-uniform)
    uint2 w[N + 1];   // 2N + 2 words of the reversed stream from an 8-byte boundary: the 2K bits of the window lie inside
};

// requests the words of the window that starts d bits (may be negative) from the 8-byte aligned address w0
template <int N> KMC_DEV void lin_fetch(const char *__restrict__ w0, int32_t d, uint32_t on, LinStep<N> &s)
{
    s.d = d;
    const uint2 *q = reinterpret_cast<const uint2 *>(w0 + static_cast<int64_t>(d >> 6) * 8); // arithmetic shift: floor
    ldg2_if<N + 1>(q, on, s.w); // (the pair past the last needed word is readable: the stream is padded)
}

// N limbs (head first) of the window: 2K bits of the reversed stream -- overlapping words, two funnel shifts per limb
template <int N> KMC_DEV void lin_limbs(const LinStep<N> &s, uint64_t head_mask, uint64_t (&limb)[N])
{
    const bool odd = (s.d & 32) != 0;
    const uint32_t sh = static_cast<uint32_t>(s.d) & 31u;
    uint32_t W[2 * N + 2], x[2 * N + 1];
#pragma unroll
    for (int i = 0; i <= N; ++i) {
        W[2 * i] = s.w[i].x;
        W[2 * i + 1] = s.w[i].y;
    }
#pragma unroll
    for (int i = 0; i <= 2 * N; ++i) x[i] = odd ? W[i + 1] : W[i];
#pragma unroll
    for (int m = 0; m < N; ++m) { // m = 0 is the least significant limb
        uint64_t v = pack64(__funnelshift_r(x[2 * m], x[2 * m + 1], sh), __funnelshift_r(x[2 * m + 1], x[2 * m + 2], sh));
        if (m == N - 1) v &= head_mask;
        limb[N - 1 - m] = v;
    }
}

// CNT consecutive words to p, if `on`.  wide = p is 16-byte aligned (known per kernel, not per store: elements of an
// even number of words in a 16-byte aligned buffer)
template <int CNT> KMC_DEV void lin_store(uint64_t *p, const uint64_t (&v)[CNT], uint32_t on, bool wide)
{
    if (CNT % 2 == 0 && wide) {
#pragma unroll
        for (int i = 0; i + 1 < CNT; i += 2) stg128_if(p + i, v[i], v[i + 1], on);
    } else {
#pragma unroll
        for (int i = 0; i < CNT; ++i) stg64_if(p + i, v[i], on);
    }
}

// OFFSETS = the set gives per-sequence offsets (seq_unit_off) and positions are symbols of the stream; otherwise
// position p = (sequence p / w8, window p % w8) and sequence r starts at symbol r * stride_syms + first (w8 >= 32:
// lin_uniform_ok; a step of 32 positions then crosses at most one sequence boundary).  AOS = Vector{Tuple{Kmer,Int}}.
// Everything inside a chunk is 32-bit arithmetic relative to the chunk's first position: the output element (against the
// chunk's first element, whose pointers are formed once), the stream symbol (against the chunk's first symbol, whose
// place in the reversed stream is formed once), the capacity left.  lin_prepare guarantees that a chunk's symbols span
// less than 2^30 (lin_uniform_ok).
// The chunk's 2048 position bits are read ONCE, 64 per lane, and so are the survivors before every pair of steps (a warp
// scan of the lanes' popcounts); a step gets its word and its running count by shuffle.  The loop is a two-stage
// software pipeline over "trips" of two steps: the words of trip t + 1 are requested (64-bit loads: two per window of
// up to 32 symbols) before trip t is shifted out and stored, so that a warp always has two to four steps' loads in
// flight.  (One step at a time, every shift waited for its own loads -- 42 % of the stall samples -- and the kernel ran
// at the pace of the load/store unit's queue; with the pipeline but the bits still loaded a trip ahead, a third of the
// samples waited for those.)
template <int N, bool HASH, bool OFFSETS, bool AOS>
__global__ void __launch_bounds__(kBlockThreads, KMC_LIN_MIN_BLOCKS) lin_compact_kernel(const ExtractParams p, const LinParams lp)
{
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t c = static_cast<uint64_t>(blockIdx.x) * kLinWarps + warp;
    if (c >= lp.n_chunks) return;
    const uint64_t o_c = __ldg(lp.chunk_off + c);
    if (__ldg(lp.chunk_off + c + 1) == o_c) return; // nothing survives in this chunk (warp-uniform)
    constexpr int kTrips = kLinChunkWords / 2;
    static_assert(kTrips == 32, "one trip's bits per lane");
    // lane t: the 64 position bits of trip t and the survivors of the trips before it
    const uint2 my_bits = __ldg(reinterpret_cast<const uint2 *>(lp.bits + c * kLinChunkWords) + lane);
    uint32_t my_run;
    {
        const uint32_t n = __popc(my_bits.x) + __popc(my_bits.y);
        uint32_t incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= static_cast<uint32_t>(d)) incl += o;
        }
        my_run = incl - n;
    }
    const uint64_t pos_c = c * kLinChunkPos;
    uint32_t lt_mask = (1u << lane) - 1u, lane_bit = 1u << lane;
    uint32_t cap = o_c >= lp.capacity ? 0u : (lp.capacity - o_c > 0xffffffffull ? 0xffffffffu : static_cast<uint32_t>(lp.capacity - o_c));
    uint64_t *out_a = p.out_a + o_c * (AOS ? N + 1 : N);
    uint64_t *out_i = AOS ? nullptr : reinterpret_cast<uint64_t *>(p.out_index) + o_c;
    uint64_t *out_h = HASH ? p.out_hash + o_c : nullptr;
    // 16-byte stores: every element of the k-mer stream starts on a 16-byte boundary
    const bool wide = (reinterpret_cast<uintptr_t>(out_a) & 15) == 0 && ((AOS ? N + 1 : N) % 2) == 0;
    uint64_t head_mask = p.head_mask;
    int64_t index_base = p.index_base + 1;
    // (the chunk's constants are formed ONCE: without the barriers the compiler re-derives them from the parameter block
    // inside the loop -- a dozen issue slots per step of a kernel that is bound by its issue slots)
    asm volatile("" : "+r"(lt_mask), "+r"(lane_bit), "+r"(cap), "+l"(out_a), "+l"(out_i), "+l"(out_h), "+l"(head_mask), "+l"(index_base));

    // this lane's position: its symbol relative to the chunk's first one and (uniform sets) its window inside its sequence
    uint32_t rel = lane, u = 0;
    uint32_t w8 = lp.w8, jump = lp.jump; // jump = stride_syms - w8: over a sequence's tail and the padding behind it
    asm volatile("" : "+r"(w8), "+r"(jump));
    uint64_t sym_c; // stream symbol of the chunk's first position
    // offsets given: the sequence that owns the step's first symbol (warp-uniform), its first symbol and the next one's
    uint64_t r_it = 0, s_cur = 0, s_next = ~0ull;
    auto seq_start = [&](uint64_t r) -> uint64_t { return (__ldg(p.seq_unit_off + r) - p.unit_bias) * lp.spu + p.first; };
    if (OFFSETS) {
        sym_c = pos_c;
        r_it = __ldg(lp.chunk_first + c);
        s_cur = seq_start(r_it);
        s_next = r_it + 1 < p.n_seqs ? seq_start(r_it + 1) : ~0ull;
    } else {
        const uint64_t r_c = (pos_c >> 32) == 0 ? static_cast<uint32_t>(pos_c) / w8 : pos_c / w8;
        const uint32_t u_c = static_cast<uint32_t>(pos_c - r_c * w8);
        sym_c = r_c * lp.stride_syms + p.first + u_c;
        u = u_c + lane;
        if (u >= w8) { // into the next sequence (w8 >= 32: at most one)
            u -= w8;
            rel += jump;
        }
    }
    // the window of relative symbol `rel` starts at bit b0 - 2 rel of the reversed stream, counted from the 8-byte
    // aligned address w0
    const int64_t bit_c = 2 * (static_cast<int64_t>(lp.t_syms) - static_cast<int64_t>(sym_c) - p.k);
    const char *w0 = reinterpret_cast<const char *>(lp.rev32 + ((bit_c >> 6) << 1));
    int32_t b0 = static_cast<int32_t>(bit_c & 63);
    asm volatile("" : "+l"(w0), "+r"(b0));
    // The chunk reads a few hundred consecutive bytes of the reversed stream, downwards, 8 bytes per step: every fourth
    // step would meet a sector that is still in DRAM.  One request per 128-byte line, all at once, instead.
    {
        const uint32_t span_syms = OFFSETS ? kLinChunkPos : kLinChunkPos + (kLinChunkPos / w8 + 1) * jump; // < 2^30 (lin_uniform_ok)
        const char *hi = w0 + ((b0 + 2 * p.k + 63) >> 3);
        const char *lo = w0 + ((static_cast<int64_t>(b0) - 2 * static_cast<int64_t>(span_syms)) >> 3);
        const char *base = reinterpret_cast<const char *>(lp.rev32);
        if (lo < base) lo = base;
        const char *line = hi - 128 * static_cast<int64_t>(lane);
        if (line >= lo) lin_prefetch(line);
    }

    // first half of a step: the 32 positions of bit word v, `run` survivors of the chunk before them
    auto issue = [&](uint32_t v, uint32_t run, LinStep<N> &s) {
        const uint32_t at = run + __popc(v & lt_mask);
        const bool on = (v & lane_bit) && at < cap;
        s.at = on ? at : kLinGone;
        lin_fetch<N>(w0, b0 - 2 * static_cast<int32_t>(rel), on ? 1u : 0u, s);
        s.pos = OFFSETS ? rel : u;
        s.any = v;
        rel += 32; // the next step's 32 positions
        if (!OFFSETS) {
            u += 32;
            const bool wrap = u >= w8;
            u -= wrap ? w8 : 0u;
            rel += wrap ? jump : 0u;
        }
    };
    auto issue_trip = [&](int t, LinStep<N> &s0, LinStep<N> &s1) {
        const uint32_t vx = __shfl_sync(0xffffffffu, my_bits.x, t), vy = __shfl_sync(0xffffffffu, my_bits.y, t);
        const uint32_t run = __shfl_sync(0xffffffffu, my_run, t);
        issue(vx, run, s0);
        issue(vy, run + __popc(vx), s1);
    };
    // second half: the k-mer out of the words, its index, the stores.  `it` = the step's number in the chunk
    auto finish = [&](const LinStep<N> &s, int it) {
        const uint32_t on = s.at != kLinGone ? 1u : 0u;
        uint64_t limb[N];
        lin_limbs<N>(s, head_mask, limb);
        int64_t index;
        if (OFFSETS) {
            index = 0;
            if (s.any) { // warp-uniform: the sequence of the step's first symbol (symbols only ascend)
                const uint64_t sym0 = pos_c + 32ull * it;
                while (s_next <= sym0) {
                    ++r_it;
                    s_cur = s_next;
                    s_next = r_it + 1 < p.n_seqs ? seq_start(r_it + 1) : ~0ull;
                }
            }
            if (on) {
                const uint64_t sym = sym_c + s.pos;
                uint64_t s_r = s_cur;
                if (sym >= s_next) { // a later sequence than the step's first
                    uint64_t r = r_it + 1, nx;
                    s_r = s_next;
                    while (r + 1 < p.n_seqs && (nx = seq_start(r + 1)) <= sym) {
                        ++r;
                        s_r = nx;
                    }
                }
                index = static_cast<int64_t>(sym - s_r) + index_base;
            }
        } else {
            index = static_cast<int64_t>(s.pos) + index_base;
        }
        if (AOS) { // {u64[N]; i64} elements
            uint64_t e[N + 1];
#pragma unroll
            for (int i = 0; i < N; ++i) e[i] = limb[i];
            e[N] = static_cast<uint64_t>(index);
            lin_store<N + 1>(out_a + static_cast<uint64_t>(s.at) * (N + 1), e, on, wide);
        } else if (N == 1) {
            stg64x2_if(out_a + s.at, limb[0], out_i + s.at, static_cast<uint64_t>(index), on);
        } else {
            lin_store<N>(out_a + static_cast<uint64_t>(s.at) * N, limb, on, wide);
            stg64_if(out_i + s.at, static_cast<uint64_t>(index), on);
        }
        if (HASH) stg64_if(out_h + s.at, fx_hash<N>(limb, 0), on);
    };

    LinStep<N> a0, a1, b0s, b1s;
    issue_trip(0, a0, a1);
#pragma unroll 1
    for (int t = 0; t < kTrips; t += 2) { // trip t is in (a0, a1)
        issue_trip(t + 1, b0s, b1s);
        finish(a0, 2 * t);
        finish(a1, 2 * t + 1);
        if (t + 2 < kTrips) issue_trip(t + 2, a0, a1);
        finish(b0s, 2 * t + 2);
        finish(b1s, 2 * t + 3);
    }
}

using LinLaunchFn = cudaError_t (*)(ExtractParams, LinParams, cudaStream_t);

template <int N, bool HASH, bool OFFSETS>
cudaError_t launch_lin_compact(ExtractParams p, LinParams lp, cudaStream_t stream)
{
    if (lp.n_chunks == 0) return cudaSuccess;
    const uint64_t blocks = (lp.n_chunks + kLinWarps - 1) / kLinWarps;
    if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    if (p.aos)
        lin_compact_kernel<N, HASH, OFFSETS, true><<<static_cast<unsigned>(blocks), kBlockThreads, 0, stream>>>(p, lp);
    else
        lin_compact_kernel<N, HASH, OFFSETS, false><<<static_cast<unsigned>(blocks), kBlockThreads, 0, stream>>>(p, lp);
    return cudaGetLastError();
}

LinLaunchFn get_lin_launcher_n1(bool hash, bool offsets);
LinLaunchFn get_lin_launcher_n2(bool hash, bool offsets);
LinLaunchFn get_lin_launcher_n3(bool hash, bool offsets);
LinLaunchFn get_lin_launcher_n4(bool hash, bool offsets);

#define KMC_DEFINE_LIN_TABLE(FN, N)                                                                                  \
    LinLaunchFn FN(bool hash, bool offsets)                                                                          \
    {                                                                                                                \
        if (hash) return offsets ? &launch_lin_compact<N, true, true> : &launch_lin_compact<N, true, false>;         \
        return offsets ? &launch_lin_compact<N, false, true> : &launch_lin_compact<N, false, false>;                 \
    }

} // namespace kmc
