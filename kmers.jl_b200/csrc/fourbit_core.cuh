// fourbit_core.cuh -- register-level primitives of the 4-bit source path (FourToTwo, src/construction.jl:85-86):
// the bit-parallel recoding of one source word and the valid-start word of one group of 32 symbols.  Device
// inline functions only (no launches, no runtime API), so that the CPU test of the primitives
// (tests/host_core) can compile exactly this code for the host.
#pragma once
#include <cstdint>

namespace kmc {

// ---- bit-parallel recoding of one LongSequence{<:NucleicAcidAlphabet{4}} word (16 nibbles) ------
// Everything is done on 32-bit halves (8 nibbles): the integer pipe is 32 bits wide, a 64-bit
// formulation costs two instructions per operation.
//   2-bit code of a one-hot nibble = trailing_zeros (A=1,C=2,G=4,T=8 -> 0,1,2,3;
//   construction_utils.jl:51): bit0 = x1|x3, bit1 = x2|x3.
//   flag <=> the nibble is not one-hot (count_ones(enc) != 1: the reference's uncertainty test,
//   FwKmers.jl:112, UnambiguousKmers.jl:145; covers IUPAC ambiguity codes, N and gap): with the four
//   bit planes a..d aligned at bit 0 of each nibble, exactly one is set iff
//   ((a^b) ^ (c^d)) & ~(a&b) & ~(c&d).
__device__ __forceinline__ void recode_half(uint32_t x, uint32_t &codes16, uint32_t &flags8)
{
    const uint32_t M = 0x11111111u;
    const uint32_t t1 = x >> 1, t2 = x >> 2, t3 = x >> 3;
    uint32_t c = ((t1 | t3) & M) | (((t2 | t3) & M) << 1); // nibble i holds its code in its low 2 bits
    c = (c | (c >> 2)) & 0x0f0f0f0fu;
    c = (c | (c >> 4)) & 0x00ff00ffu;
    codes16 = __byte_perm(c, 0u, 0x4420); // bytes 0 and 2
    const uint32_t one = ((x ^ t1) ^ (t2 ^ t3)) & ~(x & t1) & ~(t2 & t3);
    uint32_t n = ~one & M;
    n = (n | (n >> 3)) & 0x03030303u;
    n = (n | (n >> 6)) & 0x000f000fu;
    flags8 = (n | (n >> 12)) & 0xffu;
}

// one source word: 32 bits of 2-bit codes, 16 flags
__device__ __forceinline__ void recode_word(uint64_t w, uint32_t &codes, uint32_t &flags)
{
    uint32_t c0, c1, f0, f1;
    recode_half(static_cast<uint32_t>(w), c0, f0);
    recode_half(static_cast<uint32_t>(w >> 32), c1, f1);
    codes = c0 | (c1 << 16);
    flags = f0 | (f1 << 8);
}

// Valid-start bits of one group of 32 symbols from the flag words a[0..4] of this and the next four
// groups (a[5] = 0): bit t set <=> no flagged symbol in [t, t + K).  Sliding-window OR of length K by
// doubling -- A_1 = flags, A_2L = A_L | A_L >> L while 2L <= K, then two windows of length L cover
// [P, P+K): A_L | A_L >> (K - L).  Branch-free in the data (K is uniform).  Only the first
// NW = (30 + K) / 32 + 1 words can reach the result (a window starting at bit 31 ends at bit 30 + K), so
// the doubling runs on NW words: 2 for K <= 33 instead of 5.
template <int NW>
__device__ __forceinline__ uint32_t valid_start_word_n(const uint32_t (&a6)[6], int k)
{
    uint32_t a[NW + 1];
#pragma unroll
    for (int w = 0; w < NW; ++w) a[w] = a6[w];
    a[NW] = 0;
    int L = 1;
#pragma unroll
    for (int step = 0; step < 7; ++step) {
        const int s = 1 << step; // current window length
        if (2 * s <= k) {
            if (s < 32) {
#pragma unroll
                for (int w = 0; w < NW; ++w) a[w] |= __funnelshift_r(a[w], a[w + 1], s);
            } else if (s == 32) {
#pragma unroll
                for (int w = 0; w < NW; ++w) a[w] |= a[w + 1];
            } else {
#pragma unroll
                for (int w = 0; w + 1 < NW; ++w) a[w] |= a[w + 2 <= NW ? w + 2 : NW];
            }
            L = 2 * s;
        }
    }
    const int r = k - L; // 0 <= r < L, r < 64
    uint32_t v = a[0];
    if (r) {
        const int b = r & 31;
        if (r < 32) v |= __funnelshift_r(a[0], a[NW >= 1 ? 1 : 0], b);
        else v |= b ? __funnelshift_r(a[1 <= NW ? 1 : NW], a[2 <= NW ? 2 : NW], b) : a[1 <= NW ? 1 : NW];
    }
    return ~v;
}

__device__ __forceinline__ uint32_t valid_start_word(const uint32_t (&a)[6], int k)
{
    switch ((30 + k) / 32) { // warp-uniform
    case 0: return valid_start_word_n<1>(a, k);
    case 1: return valid_start_word_n<2>(a, k);
    case 2: return valid_start_word_n<3>(a, k);
    case 3: return valid_start_word_n<4>(a, k);
    }
    return valid_start_word_n<5>(a, k);
}
// TwoToFour (src/construction_utils.jl:35: enc4 = 1 << enc2)
// 8 two-bit codes (16 bits) -> 8 one-hot nibbles
__device__ __forceinline__ uint32_t onehot8(uint32_t s)
{
    s = (s | (s << 8)) & 0x00ff00ffu;
    s = (s | (s << 4)) & 0x0f0f0f0fu;
    s = (s | (s << 2)) & 0x33333333u; // nibble i holds code i in its low 2 bits
    const uint32_t M = 0x11111111u;
    const uint32_t b0 = s & M, b1 = (s >> 1) & M;
    return (~b1 & ~b0 & M) | ((~b1 & b0) << 1) | ((b1 & ~b0) << 2) | ((b1 & b0) << 3);
}

} // namespace kmc
