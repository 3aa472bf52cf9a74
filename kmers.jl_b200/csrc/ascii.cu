// ascii.cu -- ASCII sources (the AsciiEncode recoding scheme, src/construction.jl:95-96): byte
// strings (String, SubString, codeunits, Vector{UInt8}, StringView -- what FASTA/FASTQ parsers hand
// over) are recoded ON THE DEVICE into the same three streams a 4-bit source produces (2-bit codes,
// "cannot be part of a k-mer" flags, hard-error flags), so the bytes cross PCIe once and the host
// never packs anything.
//
//   strict FwKmers / FwRvIterator / CanonicalKmers (FwKmers.jl:117-129, CanonicalKmers.jl:146-174,
//     construction_utils.jl:71-88): BioSequences.ascii_encode(A, byte) > 0x7f is an EncodeError.
//     For a 2-bit alphabet the valid bytes are the alphabet's own four symbols in either case:
//     ACGT/acgt for DNAAlphabet{2}, ACGU/acgu for RNAAlphabet{2} (BioSequences v3 builds the table
//     from symbols(A) and their lowercase forms; restated, BioSequences is not vendored).
//   UnambiguousKmers (UnambiguousKmers.jl:109-132): ASCII_SKIPPING_LUT (iterators/common.jl:22-32):
//     Aa Cc Gg TtUu -> 0..3 for both alphabets, "-MRSVWYHKDBN" in either case -> skip and
//     restart, every other byte -> EncodeError.
#include <mutex>

#include "ascii_luts.h"
#include "fourbit.h"

namespace kmc {

namespace {

// (LUT entry format, AsciiLuts, make_luts: ascii_luts.h)

__constant__ uint8_t c_luts[5][256]; // strict DNA, strict RNA, skipping, 4-bit DNA, 4-bit RNA (ascii_luts.h)

__device__ uint32_t g_pos[3][8][256]; // the positioned forms of the three 2-bit tables (ascii_luts.h: make_positioned)

constexpr int kAsciiTiles = 4; // tiles of 256 groups a block recodes with one copy of the tables in shared memory

// 32 bytes -> 64 bits of 2-bit codes, 32 "not a base" flags, 32 error flags.  s_pos: the positioned tables of the LUT in
// shared memory, [8][256] words (a warp's look-ups fall into different banks unless two lanes hold bytes 32 apart).
__device__ __forceinline__ void ascii_group(const uint8_t *__restrict__ bytes, uint64_t n_bytes, bool aligned, uint64_t g,
                                            const uint32_t *s_pos, uint64_t &codes, uint32_t &fb, uint32_t &fe)
{
    const uint64_t b0 = 32 * g;
    uint32_t v[8];
    if (b0 + 32 <= n_bytes && aligned) {
        const uint4 *p = reinterpret_cast<const uint4 *>(bytes + b0);
        const uint4 x = __ldg(p), y = __ldg(p + 1);
        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
        v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
    } else {
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            uint32_t t = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint64_t b = b0 + 4 * w + c;
                t |= static_cast<uint32_t>(b < n_bytes ? bytes[b] : 'A') << (8 * c); // past the end: never part of a window
            }
            v[w] = t;
        }
    }
    uint32_t f[4]; // one result word per pair of words (eight bytes): see make_positioned
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        uint32_t acc = 0;
#pragma unroll
        for (int par = 0; par < 2; ++par) {
            const uint32_t x = v[2 * m + par];
#pragma unroll
            for (int pos = 0; pos < 4; ++pos) {
                // byte offset of the entry: 4 * ((x >> 8 pos) & 0xff), as one shift and one mask
                const uint32_t off = (pos == 0 ? (x << 2) : (x >> (8 * pos - 2))) & 0x3fcu;
                acc |= *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(s_pos + (4 * par + pos) * 256) + off);
            }
        }
        f[m] = acc;
    }
    codes = (static_cast<uint64_t>(__byte_perm(f[2], f[3], 0x7430)) << 32) | __byte_perm(f[0], f[1], 0x7430);
    fb = __byte_perm(__byte_perm(f[0], f[1], 0x0051), __byte_perm(f[2], f[3], 0x0051), 0x5410);
    fe = __byte_perm(__byte_perm(f[0], f[1], 0x0062), __byte_perm(f[2], f[3], 0x0062), 0x5410);
}

// One thread per group of 32 bytes: 2 x u32 of 2-bit codes, 1 x u32 of "not a base" flags, 1 x u32 of
// error flags and, fused through shared memory (+ a recomputed 5-group halo), the valid-start word.
// (rev: the codes once more in reversed symbol order, see recode_vstart_kernel in fourbit.cu)
// A block recodes kAsciiTiles consecutive tiles of 256 groups.
__global__ void __launch_bounds__(256) ascii_recode_kernel(const uint8_t *__restrict__ bytes, uint64_t n_bytes, int lut, int k,
                                                           uint32_t *__restrict__ rec, uint32_t *__restrict__ bad,
                                                           uint32_t *__restrict__ err, uint32_t *__restrict__ vstart,
                                                           uint64_t n_groups, uint64_t n_vstart, uint32_t *__restrict__ rev,
                                                           unsigned long long *__restrict__ any_err)
{
    __shared__ uint32_t s_pos[8 * 256];
    __shared__ uint32_t s_bad[256 + 8];
    {
        const uint32_t *src = &g_pos[lut][0][0];
#pragma unroll
        for (int i = 0; i < 8; ++i) s_pos[i * 256 + threadIdx.x] = src[i * 256 + threadIdx.x];
    }
    __syncthreads();
    const bool aligned = (reinterpret_cast<uintptr_t>(bytes) & 15) == 0;
    for (int tile = 0; tile < kAsciiTiles; ++tile) {
        const uint64_t g0 = (static_cast<uint64_t>(blockIdx.x) * kAsciiTiles + tile) * 256;
        if (g0 >= n_vstart) break; // block-uniform (n_vstart >= n_groups)
        {
            const uint64_t g = g0 + threadIdx.x;
            uint32_t fb = 0, fe = 0;
            if (g < n_groups) {
                uint64_t codes;
                ascii_group(bytes, n_bytes, aligned, g, s_pos, codes, fb, fe);
                if (rec) reinterpret_cast<uint2 *>(rec)[g] = make_uint2(static_cast<uint32_t>(codes), static_cast<uint32_t>(codes >> 32));
                if (rev)
                    reinterpret_cast<uint2 *>(rev)[n_groups - 1 - g] =
                        make_uint2(rev2_32(static_cast<uint32_t>(codes >> 32)), rev2_32(static_cast<uint32_t>(codes)));
                if (bad) bad[g] = fb;
                if (err) err[g] = fe;
                if (fe && any_err) atomicOr(any_err, 1ull); // (rare: lets the per-sequence search for the first error return at once)
            }
            s_bad[threadIdx.x] = fb;
        }
        if (threadIdx.x < kRecodeHalo) {
            const uint64_t g = g0 + 256 + threadIdx.x;
            uint32_t fb = 0, fe = 0;
            uint64_t codes;
            if (g < n_groups) ascii_group(bytes, n_bytes, aligned, g, s_pos, codes, fb, fe);
            s_bad[256 + threadIdx.x] = fb;
        }
        __syncthreads();
        const uint64_t g = g0 + threadIdx.x;
        if (g < n_vstart) {
            uint32_t a[6];
#pragma unroll
            for (int d = 0; d < 5; ++d) a[d] = s_bad[threadIdx.x + d];
            a[5] = 0;
            vstart[g] = valid_start_word(a, k);
        }
        __syncthreads(); // s_bad is rewritten by the next tile
    }
}

// k-mers over a 4-bit alphabet from ASCII bytes (FwKmers.jl:117-129, CanonicalKmers.jl:146-174 with the FourBit branch
// :160-162, SpacedKmers.jl:110-119): one thread per group of 32 bytes writes 2 x u64 of nibbles -- the layout of a
// LongSequence{DNAAlphabet{4}}, which the Copyable 4 -> 4 kernels then read -- 32 error flags and, fused through shared
// memory as above, the "K encodable symbols from here on" word.
__global__ void __launch_bounds__(256) ascii4_recode_kernel(const uint8_t *__restrict__ bytes, uint64_t n_bytes, int lut, int k,
                                                            uint64_t *__restrict__ nib, uint32_t *__restrict__ bad,
                                                            uint32_t *__restrict__ vstart, uint64_t n_groups, uint64_t n_vstart,
                                                            unsigned long long *__restrict__ any_err)
{
    __shared__ uint8_t s_lut[256];
    __shared__ uint32_t s_bad[256 + 8];
    s_lut[threadIdx.x] = c_luts[lut][threadIdx.x];
    __syncthreads();
    const uint64_t g0 = static_cast<uint64_t>(blockIdx.x) * 256;
    auto group = [&](uint64_t g, uint64_t &w0, uint64_t &w1) -> uint32_t {
        uint32_t fe = 0;
        w0 = w1 = 0;
#pragma unroll 4
        for (int t = 0; t < 32; ++t) {
            const uint64_t b = 32 * g + t;
            const uint32_t e = s_lut[b < n_bytes ? bytes[b] : 'A']; // past the end: never part of a window
            const uint64_t code = e & 15u;
            if (t < 16) w0 |= code << (4 * t); else w1 |= code << (4 * (t - 16));
            fe |= (e >> 7) << t;
        }
        return fe;
    };
    {
        const uint64_t g = g0 + threadIdx.x;
        uint32_t fe = 0;
        if (g < n_groups) {
            uint64_t w0, w1;
            fe = group(g, w0, w1);
            nib[2 * g] = w0;
            nib[2 * g + 1] = w1;
            bad[g] = fe;
            if (fe && any_err) atomicOr(any_err, 1ull);
        }
        s_bad[threadIdx.x] = fe;
    }
    if (threadIdx.x < kRecodeHalo) {
        const uint64_t g = g0 + 256 + threadIdx.x;
        uint64_t w0, w1;
        s_bad[256 + threadIdx.x] = g < n_groups ? group(g, w0, w1) : 0u;
    }
    __syncthreads();
    const uint64_t g = g0 + threadIdx.x;
    if (g < n_vstart) {
        uint32_t a[6];
#pragma unroll
        for (int d = 0; d < 5; ++d) a[d] = s_bad[threadIdx.x + d];
        a[5] = 0;
        vstart[g] = valid_start_word(a, k);
    }
}

// UnambiguousKmers over ASCII reads EVERY byte of every sequence (even sequences shorter than K),
// so any error byte fails the call: the first sequence that holds one.  One warp per sequence.
// (min_len: the strict iterators never touch a sequence shorter than K, FwKmers.jl:62-66)
__global__ void __launch_bounds__(256) seq_first_error_kernel(ExtractParams p, const uint32_t *__restrict__ err,
                                                              const uint64_t *__restrict__ seq_len, uint64_t uniform_len,
                                                              uint64_t min_len, unsigned long long *__restrict__ err_seq,
                                                              const unsigned long long *__restrict__ any_err)
{
    // the recoding pass has seen no error byte anywhere in the buffer: nothing to look for (a warp per sequence and a
    // handful of dependent loads each took 0.9 ms per 10 M reads -- a fifth of an UnambiguousKmers call over ASCII reads)
    if (any_err && *any_err == 0) return;
    const uint64_t warps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p.n_seqs == 1) { // one long sequence: the whole grid strides over its flag words
        const uint64_t len = seq_len ? seq_len[0] : uniform_len;
        if (len == 0 || len < min_len) return;
        const uint64_t unit_off = p.seq_unit_off ? p.seq_unit_off[0] - p.unit_bias : 0;
        const uint64_t a = unit_off * (p.unit_bits >> 1) + p.first, b = a + len;
        const uint64_t threads = static_cast<uint64_t>(gridDim.x) * blockDim.x;
        for (uint64_t w = (a >> 5) + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; w <= ((b - 1) >> 5); w += threads) {
            uint32_t v = __ldg(err + w);
            if (w == (a >> 5)) v &= 0xffffffffu << (a & 31);
            if (w == ((b - 1) >> 5)) v &= 0xffffffffu >> (31 - ((b - 1) & 31));
            if (v) {
                atomicMin(err_seq, 0ull);
                return;
            }
        }
        return;
    }
    for (uint64_t r = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < p.n_seqs; r += warps) {
        const uint64_t len = seq_len ? seq_len[r] : uniform_len;
        if (len == 0 || len < min_len) continue;
        const uint64_t unit_off = p.seq_unit_off ? p.seq_unit_off[r] - p.unit_bias : r * p.stride_units;
        const uint64_t a = unit_off * (p.unit_bits >> 1) + p.first, b = a + len; // error bits [a, b)
        bool found = false;
        for (uint64_t w = (a >> 5) + lane; w <= ((b - 1) >> 5) && !found; w += 32) {
            uint32_t v = __ldg(err + w);
            if (w == (a >> 5)) v &= 0xffffffffu << (a & 31);
            if (w == ((b - 1) >> 5)) v &= 0xffffffffu >> (31 - ((b - 1) & 31));
            found = v != 0;
        }
        if (__any_sync(0xffffffffu, found)) {
            if (lane == 0) atomicMin(err_seq, static_cast<unsigned long long>(r));
            return; // later sequences of this warp cannot lower the minimum
        }
    }
}

// position (1-based) and byte of the first error byte of sequence r.  One block.
__global__ void __launch_bounds__(256) resolve_ascii_error_kernel(ExtractParams p, const uint8_t *__restrict__ bytes,
                                                                  const uint32_t *__restrict__ err,
                                                                  const uint64_t *__restrict__ seq_len, uint64_t uniform_len,
                                                                  uint64_t r, uint64_t *__restrict__ err_out)
{
    __shared__ unsigned long long s_min;
    if (threadIdx.x == 0) s_min = ~0ull;
    __syncthreads();
    const uint64_t len = seq_len ? seq_len[r] : uniform_len;
    const uint64_t unit_off = p.seq_unit_off ? p.seq_unit_off[r] - p.unit_bias : r * p.stride_units;
    const uint64_t a = unit_off * (p.unit_bits >> 1) + p.first;
    for (uint64_t j = threadIdx.x; j < len; j += blockDim.x) {
        const uint64_t s = a + j;
        if ((err[s >> 5] >> (s & 31)) & 1u) {
            atomicMin(&s_min, static_cast<unsigned long long>(j));
            break;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        err_out[0] = r;
        err_out[1] = s_min + 1;
        err_out[2] = bytes[a + s_min];
    }
}

} // namespace

static cudaError_t upload_luts()
{
    static bool uploaded[64] = {};
    static std::mutex mu; // the device group calls in from one host thread per device
    std::lock_guard<std::mutex> lock(mu);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && !uploaded[dev]) {
        const AsciiLuts l = make_luts();
        static_assert(sizeof(AsciiLuts) == sizeof(c_luts), "the constant-memory copy holds every table");
        e = cudaMemcpyToSymbol(c_luts, &l, sizeof l);
        if (e != cudaSuccess) return e;
        static uint32_t pos[3][8][256]; // (guarded by `uploaded`: filled before the flag of the first device is set)
        make_positioned(l.strict_dna, pos[0]);
        make_positioned(l.strict_rna, pos[1]);
        make_positioned(l.skipping, pos[2]);
        e = cudaMemcpyToSymbol(g_pos, pos, sizeof pos);
        if (e != cudaSuccess) return e;
        uploaded[dev] = true;
    }
    return cudaSuccess;
}

// rna: U (not T) is the fourth base.  nib: 2 u64 per group of 32 bytes; bad / vstart as in ascii_recode.
cudaError_t ascii4_recode(const uint8_t *bytes, uint64_t n_bytes, bool rna, int k, uint64_t *nib, uint32_t *bad, uint32_t *vstart,
                          uint64_t n_groups, uint64_t n_vstart, cudaStream_t stream, unsigned long long *any_err)
{
    cudaError_t e = upload_luts();
    if (e != cudaSuccess) return e;
    if (any_err) {
        e = cudaMemsetAsync(any_err, 0, 8, stream);
        if (e != cudaSuccess) return e;
    }
    ascii4_recode_kernel<<<static_cast<unsigned>((n_vstart + 255) / 256), 256, 0, stream>>>(bytes, n_bytes, rna ? 4 : 3, k, nib, bad,
                                                                                           vstart, n_groups, n_vstart, any_err);
    return cudaGetLastError();
}

cudaError_t ascii_recode(const uint8_t *bytes, uint64_t n_bytes, int lut, int k, uint32_t *rec, uint32_t *bad, uint32_t *err,
                         uint32_t *vstart, uint64_t n_groups, uint64_t n_vstart, cudaStream_t stream, uint32_t *rev,
                         unsigned long long *any_err)
{
    cudaError_t e = upload_luts();
    if (e != cudaSuccess) return e;
    if (any_err) {
        e = cudaMemsetAsync(any_err, 0, 8, stream);
        if (e != cudaSuccess) return e;
    }
    ascii_recode_kernel<<<static_cast<unsigned>((n_vstart + 256 * kAsciiTiles - 1) / (256 * kAsciiTiles)), 256, 0, stream>>>(bytes, n_bytes, lut, k, rec, bad, err,
                                                                                          vstart, n_groups, n_vstart, rev, any_err);
    return cudaGetLastError();
}

cudaError_t ascii_first_error_seq(const ExtractParams &p, const uint32_t *err, const uint64_t *seq_len, uint64_t uniform_len,
                                  unsigned long long *err_seq, int sm_count, cudaStream_t stream, uint64_t min_len,
                                  const unsigned long long *any_err)
{
    if (p.n_seqs == 0) return cudaSuccess;
    const uint64_t want = p.n_seqs == 1 ? static_cast<uint64_t>(sm_count) * 32 : (p.n_seqs * 32 + 255) / 256;
    const unsigned grid = static_cast<unsigned>(want < static_cast<uint64_t>(sm_count) * 32 ? want : static_cast<uint64_t>(sm_count) * 32);
    seq_first_error_kernel<<<grid, 256, 0, stream>>>(p, err, seq_len, uniform_len, min_len, err_seq, any_err);
    return cudaGetLastError();
}

cudaError_t ascii_resolve_error(const ExtractParams &p, const uint8_t *bytes, const uint32_t *err, const uint64_t *seq_len,
                                uint64_t uniform_len, uint64_t r, uint64_t *err_out, cudaStream_t stream)
{
    resolve_ascii_error_kernel<<<1, 256, 0, stream>>>(p, bytes, err, seq_len, uniform_len, r, err_out);
    return cudaGetLastError();
}

} // namespace kmc
