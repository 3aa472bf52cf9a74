// core_host.cpp -- TEST INFRASTRUCTURE: kmers.jl_b200/csrc/kmer_core.cuh compiled for the host, so that the bit logic of
// the device primitives (block load and alignment, block_kmers, limbs_less, fx_hash) can be checked against the
// oracle without a GPU.  The CUDA intrinsics the header uses are given portable definitions here; nothing in the
// product links or calls this file (the library has no CPU path).
#include <cstdint>

#define __device__
#define __forceinline__ inline __attribute__((always_inline))

static inline uint32_t __brev(uint32_t x)
{
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
    return (x >> 16) | (x << 16);
}
// PRMT, default mode: byte i of the result is byte (selector nibble i) of the 8 bytes {b, a}
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s)
{
    const uint64_t v = (static_cast<uint64_t>(b) << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= static_cast<uint32_t>((v >> (8 * ((s >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
}
// SHF.R.WRAP: the low 32 bits of {hi, lo} >> (s & 31)
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s)
{
    s &= 31u;
    return s ? (lo >> s) | (hi << (32u - s)) : lo;
}
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32); }

#include "../../kmers.jl_b200/csrc/ascii_luts.h"
#include "../../kmers.jl_b200/csrc/fourbit_core.cuh"
#include "../../kmers.jl_b200/csrc/kmer_core.cuh"

namespace {

using namespace kmc;

template <int N, int NX, int BPS>
void windows(const uint32_t *w32, int64_t nw32, int64_t bit, const Geometry &ge, uint64_t *fw, uint64_t *rv, uint64_t *canon,
             uint64_t *hash)
{
    constexpr int G = group_of(N);
    uint32_t x[NX];
    load_block<NX>(w32, nw32, bit, x);
    uint64_t f[G][N], r[G][N];
    block_kmers<N, NX, G, true, true, BPS>(x, ge.s0, ge.head_mask, f, r);
    for (int j = 0; j < G; ++j) {
        const bool take_fw = limbs_less<N>(f[j], r[j]);
        uint64_t a[N];
        for (int i = 0; i < N; ++i) {
            fw[j * N + i] = f[j][i];
            rv[j * N + i] = r[j][i];
            a[i] = take_fw ? f[j][i] : r[j][i];
            canon[j * N + i] = a[i];
        }
        hash[j] = fx_hash<N>(a, 0);
    }
}

template <int N, int BPS>
int dispatch_nx(const uint32_t *w32, int64_t nw32, int64_t bit, const Geometry &ge, uint64_t *fw, uint64_t *rv, uint64_t *canon,
                uint64_t *hash)
{
    constexpr int G = group_of(N);
    constexpr int NXMAX = (64 * N + BPS * G - BPS + 31) / 32; // the launcher tables' range: NXMAX - 2 .. NXMAX
    if (ge.nx == NXMAX) return windows<N, NXMAX, BPS>(w32, nw32, bit, ge, fw, rv, canon, hash), 0;
    if (ge.nx == NXMAX - 1) return windows<N, (NXMAX - 1 > 0 ? NXMAX - 1 : 1), BPS>(w32, nw32, bit, ge, fw, rv, canon, hash), 0;
    if (ge.nx == NXMAX - 2) return windows<N, (NXMAX - 2 > 0 ? NXMAX - 2 : 1), BPS>(w32, nw32, bit, ge, fw, rv, canon, hash), 0;
    if (ge.nx == NXMAX - 3) return windows<N, (NXMAX - 3 > 0 ? NXMAX - 3 : 1), BPS>(w32, nw32, bit, ge, fw, rv, canon, hash), 0;
    return -1;
}

template <int BPS>
int dispatch_n(const uint32_t *w32, int64_t nw32, int64_t bit, const Geometry &ge, uint64_t *fw, uint64_t *rv, uint64_t *canon,
               uint64_t *hash)
{
    switch (ge.n_limbs) {
    case 1: return dispatch_nx<1, BPS>(w32, nw32, bit, ge, fw, rv, canon, hash);
    case 2: return dispatch_nx<2, BPS>(w32, nw32, bit, ge, fw, rv, canon, hash);
    case 3: return dispatch_nx<3, BPS>(w32, nw32, bit, ge, fw, rv, canon, hash);
    case 4: return dispatch_nx<4, BPS>(w32, nw32, bit, ge, fw, rv, canon, hash);
    }
    return -1;
}

} // namespace

// The G windows of the work item whose first window starts at stream bit `bit` (bps * symbol index), exactly as a
// kernel thread computes them: fw / rv / canon are [G][N] limbs (head first), hash is [G].  Returns G, or -1.
extern "C" int core_item_windows(const uint32_t *w32, int64_t nw32, int64_t bit, int k, int bps, uint64_t *fw, uint64_t *rv,
                                 uint64_t *canon, uint64_t *hash, int *n_limbs)
{
    const kmc::Geometry ge = kmc::geometry(k, bps);
    *n_limbs = ge.n_limbs;
    const int rc = bps == 2 ? dispatch_n<2>(w32, nw32, bit, ge, fw, rv, canon, hash)
                            : dispatch_n<4>(w32, nw32, bit, ge, fw, rv, canon, hash);
    return rc < 0 ? rc : ge.g;
}

// The same for the tuple-layout kernels (extract_kernels.cuh: launch_extract_aos): one-limb 2-bit k-mers in groups of TWO
// windows, geometry(k, 2, 2), block widths 1..3.  fw / rv are [2] words; returns 2, or -1.
template <int NX> static void windows_g2(const uint32_t *w32, int64_t nw32, int64_t bit, const kmc::Geometry &ge, uint64_t *fw, uint64_t *rv)
{
    uint32_t x[NX];
    kmc::load_block<NX>(w32, nw32, bit, x);
    uint64_t f[2][1], r[2][1];
    kmc::block_kmers<1, NX, 2, true, true, 2>(x, ge.s0, ge.head_mask, f, r);
    for (int j = 0; j < 2; ++j) {
        fw[j] = f[j][0];
        rv[j] = r[j][0];
    }
}
extern "C" int core_item_windows_g2(const uint32_t *w32, int64_t nw32, int64_t bit, int k, uint64_t *fw, uint64_t *rv)
{
    if (k < 1 || k > 32) return -1;
    const kmc::Geometry ge = kmc::geometry(k, 2, 2);
    switch (ge.nx) {
    case 1: windows_g2<1>(w32, nw32, bit, ge, fw, rv); return 2;
    case 2: windows_g2<2>(w32, nw32, bit, ge, fw, rv); return 2;
    case 3: windows_g2<3>(w32, nw32, bit, ge, fw, rv); return 2;
    }
    return -1;
}

// The locator of the lean kernels (kmer_core.cuh: AlignedLocator, G = 8, 2-bit symbols): stream bit and group of a work item
extern "C" uint64_t core_aligned_bit(uint32_t item, uint32_t gprm, uint32_t read_bits, uint32_t first, uint32_t *gi)
{
    const kmc::AlignedLocator<8, 2> loc(gprm, kmc::aligned_magic(gprm), read_bits, first);
    return loc.bit_of(item, *gi);
}

// FourToTwo primitives (fourbit_core.cuh): one source word of 16 nibbles -> 32 bits of 2-bit codes + 16 flags
extern "C" void core_recode_word(uint64_t w, uint32_t *codes, uint32_t *flags) { kmc::recode_word(w, *codes, *flags); }

// valid-start bits of one group of 32 symbols from the flag words of this and the next four groups (a[5] = 0)
extern "C" uint32_t core_valid_start_word(const uint32_t *a, int k)
{
    uint32_t b[6];
    for (int i = 0; i < 6; ++i) b[i] = a[i];
    return kmc::valid_start_word(b, k);
}

// Base.hash of one k-mer (n limbs, head first) with h already xor-ed with K, as base_hash_kernel folds it
extern "C" uint64_t core_base_hash(const uint64_t *limbs, int n, uint64_t h)
{
    uint64_t acc = kmc::base_hash_seed(h);
    for (int j = n - 1; j >= 0; --j) acc = kmc::base_hash_fold(limbs[j], acc);
    return acc;
}

// fx_hash of one k-mer (n limbs, head first), src/kmer.jl:255-261
extern "C" uint64_t core_fx_hash(const uint64_t *limbs, int n, uint64_t h)
{
    switch (n) {
    case 0: return h;
    case 1: return kmc::fx_hash<1>(*reinterpret_cast<const uint64_t(*)[1]>(limbs), h);
    case 2: return kmc::fx_hash<2>(*reinterpret_cast<const uint64_t(*)[2]>(limbs), h);
    case 3: return kmc::fx_hash<3>(*reinterpret_cast<const uint64_t(*)[3]>(limbs), h);
    case 4: return kmc::fx_hash<4>(*reinterpret_cast<const uint64_t(*)[4]>(limbs), h);
    }
    return 0;
}

// TwoToFour: 8 two-bit codes (16 bits) -> 8 one-hot nibbles
extern "C" uint32_t core_onehot8(uint32_t s) { return kmc::onehot8(s); }

// the three byte tables of the ASCII sources: strict DNA, strict RNA, the UnambiguousKmers skipping table
extern "C" void core_ascii_luts(uint8_t *out)
{
    const kmc::AsciiLuts l = kmc::make_luts();
    for (int i = 0; i < 256; ++i) {
        out[i] = l.strict_dna[i];
        out[256 + i] = l.strict_rna[i];
        out[512 + i] = l.skipping[i];
    }
}

// The positioned tables of ascii_recode_kernel applied to 32 bytes on the host, exactly as the kernel combines them: the OR of
// the eight entries of every byte pair, then the byte permutes.  which: 0 strict DNA, 1 strict RNA, 2 skipping.
// out: codes (2 words), "not a base" flags, error flags.
extern "C" void core_ascii_group_positioned(int which, const uint8_t *bytes32, uint32_t *out)
{
    const kmc::AsciiLuts l = kmc::make_luts();
    static uint32_t pos[8][256];
    kmc::make_positioned(which == 0 ? l.strict_dna : which == 1 ? l.strict_rna : l.skipping, pos);
    uint32_t f[4];
    for (int m = 0; m < 4; ++m) {
        f[m] = 0;
        for (int par = 0; par < 2; ++par)
            for (int p = 0; p < 4; ++p) f[m] |= pos[4 * par + p][bytes32[8 * m + 4 * par + p]];
    }
    out[0] = __byte_perm(f[0], f[1], 0x7430);
    out[1] = __byte_perm(f[2], f[3], 0x7430);
    out[2] = __byte_perm(__byte_perm(f[0], f[1], 0x0051), __byte_perm(f[2], f[3], 0x0051), 0x5410);
    out[3] = __byte_perm(__byte_perm(f[0], f[1], 0x0062), __byte_perm(f[2], f[3], 0x0062), 0x5410);
}

// the two 4-bit tables (k-mers over DNAAlphabet{4} / RNAAlphabet{4} from ASCII sources)
extern "C" void core_ascii_luts4(uint8_t *out)
{
    const kmc::AsciiLuts l = kmc::make_luts();
    for (int i = 0; i < 256; ++i) {
        out[i] = l.dna4[i];
        out[256 + i] = l.rna4[i];
    }
}
