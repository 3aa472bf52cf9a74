// spaced.cu -- SpacedKmers{A,K,J} (src/iterators/SpacedKmers.jl:22-139; each_codon = SpacedKmers{A,3,3}, :78-82):
// the k-mers at the 1-based starts 1, 1+J, 1+2J, ... of every sequence -- div(L - K, J) + 1 of them (:36-40).
//
// The reference walks the sequence (unsafe_extract for J >= K, unsafe_shift_from for J < K, :92-139); here every
// element is computed on its own from the closed form of kmer_core.cuh (one thread per element, G = 1: with a step
// between the windows there is no shared block to amortise), for every recoding scheme the nucleotide alphabets
// have (construction.jl:75-100):
//   Copyable 2 -> 2 and 4 -> 4      the LongSequence words as they are
//   TwoToFour                       the one-hot expansion of kmer4.cu
//   FourToTwo                       the recoding pass of fourbit.cu; an uncertain symbol inside a sampled window is
//                                   the reference's EncodeError (construction.jl:108-110)
//   AsciiEncode -> 2-bit / 4-bit    ascii.cu; a byte that is no symbol of the alphabet inside a sampled window is the
//                                   EncodeError of construction_utils.jl:71-88
// Symbols BETWEEN the windows of a step J > K are never read by the reference and are not checked here either.
#include <algorithm>

#include "fourbit.h"
#include "plan.h"

namespace kmc {

namespace {

struct SpacedParams {
    const uint32_t *w32; // the stream, in the k-mer alphabet's bits per symbol
    int64_t nw32;
    uint32_t unit_bits;  // stream bits per offset unit
    uint64_t unit_bias;
    uint32_t first, s0;
    uint64_t head_mask;
    uint32_t k, step;
    uint64_t n_seqs, total;
    uint64_t stride_units;
    const uint64_t *seq_unit_off;
    uint64_t cpr;              // uniform lengths: elements per sequence
    const uint64_t *out_off;   // ragged: [n_seqs + 1] exclusive scan of the elements per sequence
    uint64_t *out_a, *out_hash;
    const uint32_t *ok_bits;   // or NULL: bit P = the K symbols from stream symbol P on can all be encoded
    unsigned long long *err_flat; // first element whose window cannot be encoded
};

// element e -> (sequence, first symbol of its window relative to the sequence's offset unit)
__device__ __forceinline__ void spaced_locate(const SpacedParams &p, uint64_t e, uint64_t &r, uint64_t &sym)
{
    uint64_t j;
    if (p.out_off) {
        uint64_t lo = 0, hi = p.n_seqs; // largest r with out_off[r] <= e < out_off[r + 1]
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) >> 1;
            if (__ldg(p.out_off + mid) <= e) lo = mid; else hi = mid;
        }
        r = lo;
        j = e - __ldg(p.out_off + r);
    } else {
        r = e / p.cpr;
        j = e - r * p.cpr;
    }
    sym = p.first + j * p.step;
}

__device__ __forceinline__ uint64_t spaced_unit(const SpacedParams &p, uint64_t r)
{
    return p.seq_unit_off ? __ldg(p.seq_unit_off + r) - p.unit_bias : r * p.stride_units;
}

template <int N, int NX, int BPS, bool HASH>
__global__ void __launch_bounds__(256) spaced_kernel(const SpacedParams p)
{
    const uint64_t e = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= p.total) return;
    uint64_t r, sym;
    spaced_locate(p, e, r, sym);
    const int64_t bit = static_cast<int64_t>(spaced_unit(p, r) * p.unit_bits + static_cast<uint64_t>(BPS) * sym);
    if (p.ok_bits) {
        const uint64_t P = static_cast<uint64_t>(bit) / BPS;
        if (!((__ldg(p.ok_bits + (P >> 5)) >> (P & 31)) & 1u)) atomicMin(p.err_flat, static_cast<unsigned long long>(e));
    }
    uint32_t x[NX];
    load_block<NX>(p.w32, p.nw32, bit, x);
    uint64_t fw[1][N], rv[1][N];
    block_kmers<N, NX, 1, true, false, BPS>(x, p.s0, p.head_mask, fw, rv);
#pragma unroll
    for (int i = 0; i < N; ++i) st_u64(p.out_a + e * N + i, fw[0][i]);
    if (HASH) st_u64(p.out_hash + e, fx_hash<N>(fw[0], 0));
}

using SpacedLaunch = void (*)(const SpacedParams &, cudaStream_t);

template <int N, int NX, int BPS, bool HASH>
void launch_spaced(const SpacedParams &p, cudaStream_t stream)
{
    spaced_kernel<N, NX, BPS, HASH><<<static_cast<unsigned>((p.total + 255) / 256), 256, 0, stream>>>(p);
}

template <int N, int BPS>
SpacedLaunch pick_spaced(int nx, bool hash)
{
    // K * BPS bits in (64 (N - 1), 64 N]: 2N - 1 or 2N words of 32 bits
    if (nx == 2 * N) return hash ? &launch_spaced<N, 2 * N, BPS, true> : &launch_spaced<N, 2 * N, BPS, false>;
    if (nx == 2 * N - 1) return hash ? &launch_spaced<N, 2 * N - 1, BPS, true> : &launch_spaced<N, 2 * N - 1, BPS, false>;
    return nullptr;
}

template <int BPS>
SpacedLaunch spaced_launcher(int n_limbs, int nx, bool hash)
{
    switch (n_limbs) {
    case 1: return pick_spaced<1, BPS>(nx, hash);
    case 2: return pick_spaced<2, BPS>(nx, hash);
    case 3: return pick_spaced<3, BPS>(nx, hash);
    case 4: return pick_spaced<4, BPS>(nx, hash);
    }
    return nullptr;
}

__global__ void spaced_counts_kernel(const uint64_t *__restrict__ seq_len, uint64_t n, uint64_t k, uint64_t step, uint64_t *__restrict__ cnt)
{
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint64_t len = seq_len[i];
        cnt[i] = len >= k ? (len - k) / step + 1 : 0; // SpacedKmers.jl:36-40
    }
}

// The element `e` is the first whose window cannot be encoded: the first flagged symbol of that window is what the
// reference throws on (every window before it was clean).  One thread.
__global__ void spaced_resolve_kernel(const SpacedParams p, int bps, uint64_t e, const uint32_t *__restrict__ bad, const void *src,
                                      int src_bits, uint64_t *__restrict__ err_out)
{
    uint64_t r, sym;
    spaced_locate(p, e, r, sym);
    const uint64_t P = (spaced_unit(p, r) * p.unit_bits) / bps + sym; // stream symbol of the window's first symbol
    uint64_t t = 0;
    for (; t + 1 < p.k; ++t)
        if ((bad[(P + t) >> 5] >> ((P + t) & 31)) & 1u) break;
    const uint64_t a = P + t; // source symbol index = stream symbol index
    err_out[0] = r;
    err_out[1] = sym - p.first + t + 1;
    err_out[2] = src_bits == 8 ? static_cast<const uint8_t *>(src)[a] : (static_cast<const uint64_t *>(src)[a >> 4] >> (4 * (a & 15))) & 15u;
}

} // namespace

} // namespace kmc

using namespace kmc;

extern "C" int32_t kmc_extract_spaced(kmc_ctx *ctx, const kmc_seqs *s, int32_t k, int32_t step, uint32_t flags, const kmc_out *out,
                                      kmc_result *res)
{
    int32_t st = check_common(ctx, s, k);
    if (st) return st;
    if (!out || !res) return fail(ctx, KMC_E_BAD_ARG, "kmc_out / kmc_result is NULL");
    if (step < 1) return fail(ctx, KMC_E_BAD_K, "J must be at least 1"); // SpacedKmers.jl:30
    const bool kmer4 = (flags & KMC_KMER4) != 0;
    if (kmer4 && k > KMC_MAX_K4) return fail(ctx, KMC_E_BAD_K, "K exceeds KMC_MAX_K4 (64) for k-mers over a 4-bit alphabet");
    const bool hash = (flags & KMC_HASH_FX) != 0;
    if (hash && !out->hash) return fail(ctx, KMC_E_BAD_ARG, "out.hash is NULL (KMC_HASH_FX)");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t stream = ctx->stream;
    res->n_written = 0;
    res->err_seq = res->err_pos = 0;
    res->err_sym = 0;
    res->kernel_ms = 0.f;
    const int bps = kmer4 ? 4 : 2;
    const uint64_t K = static_cast<uint64_t>(k), J = static_cast<uint64_t>(step), n = s->n_seqs;
    const bool ascii = s->src_bits == 8;
    const uint64_t nb = ascii ? (s->n_words + 31) / 32 : (s->n_words + 1) / 2; // groups of 32 symbols of a recoded source

    uint64_t need = 4 * round_up(8 * (n + 2), 256) + round_up(8 * (scan_tmp_elems(n) + 1), 256) + 4096;
    need += round_up(16 * (std::max(s->n_words, nb) + 4), 256) + 2 * round_up(4 * (nb + 8), 256); // recoded / expanded stream, flags
    st = ensure_scratch(ctx, need);
    if (st) return st;
    st = ensure_host_small(ctx);
    if (st) return st;
    Scratch scratch{static_cast<char *>(ctx->scratch), ctx->scratch_bytes, 0};
    CU(cudaEventRecord(ctx->ev_k0, stream));

    SpacedParams p{};
    p.k = static_cast<uint32_t>(k);
    p.step = static_cast<uint32_t>(step);
    p.n_seqs = n;
    p.first = s->first_symbol_offset;
    p.stride_units = s->uniform_stride_words;
    p.seq_unit_off = s->seq_word_offset;
    const int n_limbs = (bps * k + 63) / 64, nx = (bps * k + 31) / 32;
    p.s0 = static_cast<uint32_t>(32 * nx - bps * k);
    const int used = bps * k - 64 * (n_limbs - 1);
    p.head_mask = used >= 64 ? ~0ull : ((1ull << used) - 1);

    // how many elements, and where each sequence's elements start
    if (s->seq_len == nullptr) {
        p.cpr = s->uniform_len >= K ? (s->uniform_len - K) / J + 1 : 0;
        p.total = p.cpr * n;
        if (out->seq_out_offset) CU(fill_uniform_offsets(out->seq_out_offset, n + 1, p.cpr, stream));
    } else {
        uint64_t *cnt = static_cast<uint64_t *>(scratch.take(8 * (n + 2)));
        uint64_t *off = static_cast<uint64_t *>(scratch.take(8 * (n + 2)));
        uint64_t *tmp = static_cast<uint64_t *>(scratch.take(8 * (scan_tmp_elems(n) + 1)));
        if (!cnt || !off || !tmp) return fail(ctx, KMC_E_BAD_ARG, "internal: scratch window too small");
        if (n) {
            spaced_counts_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(s->seq_len, n, K, J, cnt);
            CU(cudaGetLastError());
        }
        CU(inclusive_offsets_u64(cnt, off, n, tmp, stream));
        uint64_t *h = ctx->host_small + 56;
        CU(cudaMemcpyAsync(h, off + n, 8, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        p.total = *h;
        p.out_off = off;
        if (out->seq_out_offset) CU(cudaMemcpyAsync(out->seq_out_offset, off, 8 * (n + 1), cudaMemcpyDeviceToDevice, stream));
    }
    res->n_written = p.total;
    if (p.total == 0) return KMC_OK;
    if (p.total > out->capacity) return fail(ctx, KMC_E_OUT_TOO_SMALL, "output capacity smaller than the number of k-mers");
    if (!out->a) return fail(ctx, KMC_E_BAD_ARG, "out.a is NULL");
    p.out_a = out->a;
    p.out_hash = hash ? out->hash : nullptr;

    // the stream in the k-mer alphabet's bits per symbol
    uint32_t *bad = nullptr;
    p.w32 = reinterpret_cast<const uint32_t *>(s->words);
    p.nw32 = static_cast<int64_t>(s->n_words) * 2;
    p.unit_bits = 64;
    if (s->src_bits == 2 && kmer4) { // TwoToFour
        void *wide = scratch.take(16 * (s->n_words + 2));
        if (!wide) return fail(ctx, KMC_E_BAD_ARG, "internal: scratch window too small");
        CU(expand_two_to_four(s->words, s->n_words, wide, stream));
        p.w32 = static_cast<const uint32_t *>(wide);
        p.nw32 = static_cast<int64_t>(s->n_words) * 4;
        p.unit_bits = 128;
    } else if ((s->src_bits == 4 && !kmer4) || ascii) { // FourToTwo, AsciiEncode: recode, flag what cannot be encoded
        void *rec = scratch.take(16 * (nb + 2));
        bad = static_cast<uint32_t *>(scratch.take(4 * (nb + 8)));
        uint32_t *vstart = static_cast<uint32_t *>(scratch.take(4 * (nb + 8)));
        p.err_flat = static_cast<unsigned long long *>(scratch.take(8));
        if (!rec || !bad || !vstart || !p.err_flat) return fail(ctx, KMC_E_BAD_ARG, "internal: scratch window too small");
        if (!ascii) {
            CU(fourbit_recode(s->words, s->n_words, k, static_cast<uint32_t *>(rec), bad, vstart, nb, nb + 2, stream));
            p.nw32 = static_cast<int64_t>(nb) * 2;
            p.unit_bits = 32; // one source word = 16 symbols = 32 bits of the recoded stream
        } else if (!kmer4) {
            CU(ascii_recode(reinterpret_cast<const uint8_t *>(s->words), s->n_words, (flags & KMC_RNA) ? 1 : 0, k,
                            static_cast<uint32_t *>(rec), bad, nullptr, vstart, nb, nb + 2, stream));
            p.nw32 = static_cast<int64_t>(nb) * 2;
            p.unit_bits = 2; // offsets count bytes = symbols
        } else {
            CU(ascii4_recode(reinterpret_cast<const uint8_t *>(s->words), s->n_words, (flags & KMC_RNA) != 0, k,
                             static_cast<uint64_t *>(rec), bad, vstart, nb, nb + 2, stream));
            p.nw32 = static_cast<int64_t>(nb) * 4;
            p.unit_bits = 4;
        }
        p.w32 = static_cast<const uint32_t *>(rec);
        p.ok_bits = vstart;
        CU(cudaMemsetAsync(p.err_flat, 0xff, 8, stream));
    }

    SpacedLaunch fn = kmer4 ? spaced_launcher<4>(n_limbs, nx, hash) : spaced_launcher<2>(n_limbs, nx, hash);
    if (!fn) return fail(ctx, KMC_E_UNSUPPORTED, "no kernel for this K");
    fn(p, stream);
    CU(cudaGetLastError());
    if (p.ok_bits) {
        uint64_t *h = ctx->host_small + 56;
        CU(cudaMemcpyAsync(h, p.err_flat, 8, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        if (*h != ~0ull) {
            uint64_t *err_out = static_cast<uint64_t *>(scratch.take(24));
            if (!err_out) return fail(ctx, KMC_E_BAD_ARG, "internal: scratch window too small");
            spaced_resolve_kernel<<<1, 1, 0, stream>>>(p, bps, *h, bad, s->words, static_cast<int>(s->src_bits), err_out);
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(h + 1, err_out, 24, cudaMemcpyDeviceToHost, stream));
            CU(cudaStreamSynchronize(stream));
            res->n_written = 0;
            res->err_seq = h[1];
            res->err_pos = h[2];
            res->err_sym = static_cast<uint32_t>(h[3]);
            return fail(ctx, KMC_E_AMBIGUOUS, "cannot encode this symbol in the k-mer alphabet");
        }
    }
    CU(cudaEventRecord(ctx->ev_k1, stream));
    CU(cudaStreamSynchronize(stream));
    CU(cudaEventElapsedTime(&res->kernel_ms, ctx->ev_k0, ctx->ev_k1));
    return KMC_OK;
}
