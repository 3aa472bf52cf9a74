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

// Predicated loads / stores without a memory clobber: the kernel below is straight-line code for two steps at a time, and
// the compiler is free to move the loads of the second step above the stores of the first (they never alias: the loads
// are read-only source data).
KMC_DEV uint32_t ldg_if(const uint32_t *p, uint32_t on)
{
    uint32_t v;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\tmov.u32 %0, 0;\n\t@q ld.global.nc.u32 %0, [%1];\n\t}" : "=r"(v) : "l"(p), "r"(on));
    return v;
}
KMC_DEV void stg64_if(uint64_t *p, uint64_t v, uint32_t on)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.L1::no_allocate.L2::cache_hint.u64 [%0], %1, %3;\n\t}" ::"l"(p), "l"(v),
                 "r"(on), "l"(kEvictFirst));
}
KMC_DEV void stg128_if(uint64_t *p, uint64_t a, uint64_t b, uint32_t on)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q st.global.L1::no_allocate.L2::cache_hint.v2.u64 [%0], {%1,%2}, %4;\n\t}" ::"l"(p),
                 "l"(a), "l"(b), "r"(on), "l"(kEvictFirst));
}

// N limbs (head first) of a window: 2K bits of the reversed stream from bit offset d (may be negative) relative to
// the byte address w0 on
template <int N>
KMC_DEV void lin_kmer(const char *__restrict__ w0, int32_t d, uint64_t head_mask, uint32_t on, uint64_t (&limb)[N])
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(w0 + static_cast<int64_t>(d >> 5) * 4); // arithmetic shift: floor
    const uint32_t sh = static_cast<uint32_t>(d) & 31u;
    uint32_t x[2 * N + 1];
#pragma unroll
    for (int i = 0; i <= 2 * N; ++i) x[i] = ldg_if(w + i, on); // (the word past the last needed one is readable: the stream is padded)
#pragma unroll
    for (int m = 0; m < N; ++m) { // m = 0 is the least significant limb
        uint64_t v = pack64(__funnelshift_r(x[2 * m], x[2 * m + 1], sh), __funnelshift_r(x[2 * m + 1], x[2 * m + 2], sh));
        if (m == N - 1) v &= head_mask;
        limb[N - 1 - m] = v;
    }
}

// CNT consecutive words to p (8-byte aligned; 16-byte stores where the address allows), if `on`
template <int CNT> KMC_DEV void lin_store(uint64_t *p, const uint64_t (&v)[CNT], uint32_t on)
{
    if (CNT % 2 == 0) {
        const uint32_t al = (reinterpret_cast<uintptr_t>(p) & 15) == 0 ? on : 0u, un = on & ~al;
#pragma unroll
        for (int i = 0; i + 1 < CNT; i += 2) stg128_if(p + i, v[i], v[i + 1], al);
        if (un) { // an output buffer that is only 8-byte aligned: rare
#pragma unroll
            for (int i = 0; i < CNT; ++i) stg64_if(p + i, v[i], 1u);
        }
    } else {
#pragma unroll
        for (int i = 0; i < CNT; ++i) stg64_if(p + i, v[i], on);
    }
}

// OFFSETS = the set gives per-sequence offsets (seq_unit_off) and positions are symbols of the stream; otherwise
// position p = (sequence p / w8, window p % w8) and sequence r starts at symbol r * stride_syms + first.
// Everything inside a chunk is 32-bit arithmetic relative to the chunk's first position: the output element (against the
// chunk's first element, whose pointers are formed once), the stream symbol (against the chunk's first symbol, whose
// place in the reversed stream is formed once), the capacity left.  lin_prepare guarantees that a chunk's symbols span
// less than 2^30 (lin_uniform_ok).  The loop body is branch-free for uniform sets (predicated loads and stores) and takes
// two steps per trip, so that the loads of the second step are in flight while the first is finished.
template <int N, bool HASH, bool OFFSETS>
__global__ void __launch_bounds__(kBlockThreads) lin_compact_kernel(const ExtractParams p, const LinParams lp)
{
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t c = static_cast<uint64_t>(blockIdx.x) * kLinWarps + warp;
    if (c >= lp.n_chunks) return;
    const uint64_t o_c = __ldg(lp.chunk_off + c);
    if (__ldg(lp.chunk_off + c + 1) == o_c) return; // nothing survives in this chunk (warp-uniform)
    const uint32_t *bits = lp.bits + c * kLinChunkWords;
    asm volatile("" : "+l"(bits));
    const uint64_t pos_c = c * kLinChunkPos;
    const bool tuple_ix = p.aos != 0;
    uint32_t lt_mask = (1u << lane) - 1u, lane_bit = 1u << lane;
    uint32_t cap = o_c >= lp.capacity ? 0u : (lp.capacity - o_c > 0xffffffffull ? 0xffffffffu : static_cast<uint32_t>(lp.capacity - o_c));
    uint64_t *out_a = p.out_a + o_c * (tuple_ix ? N + 1 : N);
    uint64_t *out_i = tuple_ix ? nullptr : reinterpret_cast<uint64_t *>(p.out_index) + o_c;
    uint64_t *out_h = HASH ? p.out_hash + o_c : nullptr;
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
        while (u >= w8) { // into the next sequence(s)
            u -= w8;
            rel += jump;
        }
    }
    // the window of relative symbol `rel` starts at bit b0 - 2 rel of the reversed stream, counted from byte address w0
    const int64_t bit_c = 2 * (static_cast<int64_t>(lp.t_syms) - static_cast<int64_t>(sym_c) - p.k);
    const char *w0 = reinterpret_cast<const char *>(lp.rev32 + (bit_c >> 5));
    int32_t b0 = static_cast<int32_t>(bit_c & 31);
    asm volatile("" : "+l"(w0), "+r"(b0));

    // one step: the 32 positions of bit word v; `run` survivors of the chunk came before it
    auto step = [&](uint32_t v, uint32_t run, int it) {
        const uint32_t at = run + __popc(v & lt_mask); // this lane's element of the chunk's output, if it survives
        const uint32_t on = ((v & lane_bit) && at < cap) ? 1u : 0u;
        uint64_t limb[N];
        lin_kmer<N>(w0, b0 - 2 * static_cast<int32_t>(rel), head_mask, on, limb);
        int64_t index;
        if (OFFSETS) {
            index = 0;
            if (v) { // warp-uniform: the sequence of the step's first symbol (symbols only ascend)
                const uint64_t sym0 = pos_c + 32ull * it;
                while (s_next <= sym0) {
                    ++r_it;
                    s_cur = s_next;
                    s_next = r_it + 1 < p.n_seqs ? seq_start(r_it + 1) : ~0ull;
                }
            }
            if (on) {
                const uint64_t sym = sym_c + rel;
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
            index = static_cast<int64_t>(u) + index_base;
        }
        if (tuple_ix) { // Vector{Tuple{Kmer,Int}}: {u64[N]; i64} elements
            uint64_t e[N + 1];
#pragma unroll
            for (int i = 0; i < N; ++i) e[i] = limb[i];
            e[N] = static_cast<uint64_t>(index);
            lin_store<N + 1>(out_a + static_cast<uint64_t>(at) * (N + 1), e, on);
        } else {
            lin_store<N>(out_a + static_cast<uint64_t>(at) * N, limb, on);
            stg64_if(out_i + at, static_cast<uint64_t>(index), on);
        }
        if (HASH) stg64_if(out_h + at, fx_hash<N>(limb, 0), on);
        // the next step's 32 positions
        rel += 32;
        if (!OFFSETS) {
            u += 32;
            if (u >= w8) {
                u -= w8;
                rel += jump;
                while (u >= w8) { // sequences of fewer than 32 windows
                    u -= w8;
                    rel += jump;
                }
            }
        }
    };

    uint32_t run = 0; // survivors of the chunk's earlier steps
    uint2 v_next = __ldg(reinterpret_cast<const uint2 *>(bits));
#pragma unroll 1
    for (int it = 0; it < kLinChunkWords; it += 2) {
        const uint2 v = v_next;
        v_next = __ldg(reinterpret_cast<const uint2 *>(bits + it + 2)); // (the bit array is padded)
        step(v.x, run, it);
        step(v.y, run + __popc(v.x), it + 1);
        run += __popc(v.x) + __popc(v.y);
    }
}

using LinLaunchFn = cudaError_t (*)(ExtractParams, LinParams, cudaStream_t);

template <int N, bool HASH, bool OFFSETS>
cudaError_t launch_lin_compact(ExtractParams p, LinParams lp, cudaStream_t stream)
{
    if (lp.n_chunks == 0) return cudaSuccess;
    const uint64_t blocks = (lp.n_chunks + kLinWarps - 1) / kLinWarps;
    if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    lin_compact_kernel<N, HASH, OFFSETS><<<static_cast<unsigned>(blocks), kBlockThreads, 0, stream>>>(p, lp);
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
