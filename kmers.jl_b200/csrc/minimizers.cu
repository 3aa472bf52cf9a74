// minimizers.cu -- minimizers on the device: the step AFTER the k-mer stream.
//
// "Minimizers ... are defined as the minimum of W consecutive kmers, as ordered by some ordering O"
// (docs/src/replacements.md:28-30), with fx_hash as the ordering (replacements.md:32-58,
// test/benchmark.jl:96-119).  For every window start i = 1, 1+step, ... with W k-mers available,
// emit the k-mer with the smallest fx_hash among the k-mers starting at i .. i+W-1 (ties: the
// first).  The full k-mer stream is never written: the kernel reads 2 bits per symbol and writes one
// k-mer per window start.
//
// NOTE: the reference's *example* loop shifts the next symbol into the best-so-far k-mer rather than
// into the previous window's k-mer, so whenever a later k-mer is not smaller it compares k-mers that
// do not occur in the sequence.  This kernel implements the prose definition (the true W
// consecutive k-mers); tests/ check it against a naive per-window definition.
//
// Dense windows (step 1, W <= 31, uniform sets / single sequences: minimizer_dense_kernel): a thread
// owns G consecutive window starts, so it hashes each of their G+W-1 k-mers ONCE (register-blocked
// exactly like the extraction kernels: one block load, static funnel shifts) and takes the G
// sliding minima by doubling: m_1 = h, m_2s[i] = min(m_s[i], m_s[i+s]), and a window of W =
// 2^L + d k-mers is min(m_2^L[i], m_2^L[i+d]).  Ties keep the leftmost k-mer at every merge, so
// the result is the first smallest, as the definition asks.  The winning k-mer is recovered from
// its hash (fx_hash of one limb is a multiplication by an odd constant: invertible mod 2^64).
//
// Everything else (step > 1, ragged sets, W >= 32): one thread per window start.  The K+W-1 symbols
// of the window (<= 64 symbols = 128 bits) are fetched once; k-mer j is a shift of the
// 2-bit-reversed block (forward) or of the complemented raw block (reverse complement).
#include "kmer_core.cuh"
#include "plan.h"

namespace kmc {

namespace {

struct MinimizerParams {
    const uint64_t *words;
    int64_t n_words;
    uint64_t n_seqs;
    const uint64_t *seq_word_off; // or NULL
    uint64_t stride_words;
    uint32_t first;
    uint64_t uniform_cnt;         // window starts per sequence (uniform sets)
    const uint64_t *cnt_off;      // ragged: exclusive scan of the per-sequence counts [n_seqs+1]
    uint64_t total;
    int k, w, step;
    int canon;
    uint64_t *out_kmer, *out_hash;
    int64_t *out_index;
    int64_t index_base;
};

KMC_DEV uint64_t rev2_64(uint64_t x)
{
    x = __brevll(x);
    return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}

__global__ void __launch_bounds__(256) minimizer_kernel(const MinimizerParams p)
{
    const uint64_t e = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= p.total) return;
    uint64_t r, t;
    if (p.cnt_off) {
        uint64_t lo = 0, hi = p.n_seqs;
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) >> 1;
            if (__ldg(p.cnt_off + mid) <= e) lo = mid; else hi = mid;
        }
        r = lo;
        t = e - __ldg(p.cnt_off + r);
    } else {
        r = e / p.uniform_cnt;
        t = e - r * p.uniform_cnt;
    }
    const uint64_t word_off = p.seq_word_off ? __ldg(p.seq_word_off + r) : r * p.stride_words;
    const uint64_t sym = t * static_cast<uint64_t>(p.step);              // 0-based window start within the sequence
    const uint64_t bit = word_off * 64 + 2 * (p.first + sym);
    const int64_t idx = static_cast<int64_t>(bit >> 6);
    const unsigned sh = static_cast<unsigned>(bit & 63);
    auto W = [&](int64_t i) -> uint64_t { return i < p.n_words ? __ldg(p.words + i) : 0ull; };
    const uint64_t w0 = W(idx), w1 = W(idx + 1), w2 = W(idx + 2);
    const uint64_t lo = sh ? (w0 >> sh) | (w1 << (64 - sh)) : w0;
    const uint64_t hi = sh ? (w1 >> sh) | (w2 << (64 - sh)) : w1;
    const int L = p.k + p.w - 1; // symbols in the window, <= 64
    unsigned __int128 S = (static_cast<unsigned __int128>(hi) << 64) | lo;
    if (L < 64) S &= (static_cast<unsigned __int128>(1) << (2 * L)) - 1;
    // rev2 of the whole 128-bit register pair: symbol i -> 63-i, so k-mer j sits in the top 2K bits
    // after a left shift by 2j -- both streams advance by a constant 2 bits per k-mer
    unsigned __int128 T = (static_cast<unsigned __int128>(rev2_64(lo)) << 64) | rev2_64(hi);
    const uint64_t mask = p.k == 32 ? ~0ull : ((1ull << (2 * p.k)) - 1);
    const int top = 64 - 2 * p.k; // 0..62
    uint64_t best_k = 0, best_h = 0;
    int best_j = 0;
    for (int j = 0; j < p.w; ++j) {
        uint64_t km = static_cast<uint64_t>(T >> 64) >> top;
        if (p.canon) {
            const uint64_t rv = ~static_cast<uint64_t>(S) & mask;
            km = km < rv ? km : rv;
        }
        const uint64_t h = km * FX_CONSTANT; // fx_hash of a one-limb k-mer, h0 = 0 (src/kmer.jl:255-261)
        if (j == 0 || h < best_h) {
            best_h = h;
            best_k = km;
            best_j = j;
        }
        T <<= 2;
        S >>= 2;
    }
    p.out_kmer[e] = best_k;
    if (p.out_hash) p.out_hash[e] = best_h;
    if (p.out_index) p.out_index[e] = static_cast<int64_t>(sym) + best_j + 1 + p.index_base;
}

constexpr uint64_t FX_INVERSE = 0x2040003d780970bdull; // FX_CONSTANT * FX_INVERSE == 1 (mod 2^64)
static_assert(FX_CONSTANT * FX_INVERSE == 1ull, "inverse of the fx_hash multiplier");

struct DenseParams {
    const uint32_t *w32;
    int64_t nw32;
    uint64_t n_seqs;
    const uint64_t *seq_word_off; // or NULL
    uint64_t stride_words;
    uint32_t first;
    uint64_t cnt;   // window starts per sequence
    uint64_t ipr;   // work items (groups of G window starts) per sequence
    uint64_t items; // ipr * n_seqs
    int k, w;
    uint64_t mask;
    uint64_t *out_kmer, *out_hash;
    int64_t *out_index;
    int64_t index_base;
};

// windows per thread for the class 2^L <= W < 2^(L+1)
#ifndef KMC_MIN_G
#define KMC_MIN_G 16
#endif
constexpr int dense_group(int l) { return l <= 3 ? KMC_MIN_G : 8; }

template <int CNT>
KMC_DEV void store_dense(uint64_t *p, const uint64_t (&v)[CNT], int n)
{
    if (n == CNT && (reinterpret_cast<uintptr_t>(p) & 31) == 0) {
#pragma unroll
        for (int i = 0; i < CNT; i += 4) st_v4(p + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else if (n == CNT && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
        for (int i = 0; i < CNT; i += 2) st_v2(p + i, v[i], v[i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < CNT; ++i)
            if (i < n) st_u64(p + i, v[i]);
    }
}

constexpr int log2_floor(int w) { return w < 2 ? 0 : 1 + log2_floor(w / 2); }

// W is a template parameter (31 instantiations per form): a thread builds and hashes exactly the G + W - 1 k-mers its G
// windows can see, and the last merge has a compile-time distance.  (Templated on the class 2^L <= W < 2^(L+1) only, the
// kernel hashed G + 2^(L+1) - 2 k-mers -- 30 instead of 25 for W = 10 -- and carried them through every doubling step.)
template <int W, bool CANON, bool INDEX>
__global__ void __launch_bounds__(128) minimizer_dense_kernel(const DenseParams p)
{
    constexpr int L = log2_floor(W);
    constexpr int G = dense_group(L);
    constexpr int MM = G + W - 1;                     // k-mers a thread hashes
    constexpr int NX = (64 + 2 * (MM - 1) + 31) / 32; // block words for K = 32
    const uint64_t item = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (item >= p.items) return;
    const uint64_t r = item / p.ipr;
    const uint64_t t0 = (item - r * p.ipr) * G; // first window start of this thread within its sequence
    const uint64_t left = p.cnt - t0;
    const int nwin = left < static_cast<uint64_t>(G) ? static_cast<int>(left) : G;
    const uint64_t word_off = p.seq_word_off ? __ldg(p.seq_word_off + r) : r * p.stride_words;
    const int64_t bit = static_cast<int64_t>(word_off) * 64 + 2 * static_cast<int64_t>(p.first + t0);
    uint32_t x[NX];
    load_block<NX>(p.w32, p.nw32, bit, x);

    // forward k-mer j = rev2(block) >> 2(MM-1-j) once the block is right-aligned to k-mer MM-1:
    // drop s0 = 32 NX - 2K - 2(MM-1) bits; s0 = 32 q + sr with a uniform q in 0..2 (NX is sized for K = 32)
    uint32_t t[NX];
    {
        uint32_t y[NX + 3];
#pragma unroll
        for (int i = 0; i < NX; ++i) y[i] = rev2_32(x[NX - 1 - i]);
        y[NX] = y[NX + 1] = y[NX + 2] = 0u;
        const uint32_t s0 = 32u * NX - 2u * p.k - 2u * (MM - 1);
        const uint32_t q = s0 >> 5, sr = s0 & 31u;
        if (q == 0) {
#pragma unroll
            for (int i = 0; i < NX; ++i) t[i] = __funnelshift_r(y[i], y[i + 1], sr);
        } else if (q == 1) {
#pragma unroll
            for (int i = 0; i < NX; ++i) t[i] = __funnelshift_r(y[i + 1], y[i + 2], sr);
        } else {
#pragma unroll
            for (int i = 0; i < NX; ++i) t[i] = __funnelshift_r(y[i + 2], y[i + 3], sr);
        }
    }
    uint32_t nx[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) nx[i] = ~x[i];

    uint64_t h[MM];
    uint32_t pos[INDEX ? MM : 1];
#pragma unroll
    for (int j = 0; j < MM; ++j) {
        uint64_t km = stream64<NX>(t, 2 * (MM - 1 - j)) & p.mask;
        if (CANON) {
            const uint64_t rv = stream64<NX>(nx, 2 * j) & p.mask; // ~W: the reverse complement (kmer_core.cuh)
            km = km < rv ? km : rv;
        }
        h[j] = mul_fx(km); // fx_hash of a one-limb k-mer, h0 = 0 (src/kmer.jl:255-261)
        if (INDEX) pos[j] = j;
    }
    // sliding minima by doubling; ties keep the left (earlier) k-mer
#pragma unroll
    for (int l = 0; l < L; ++l) {
        const int s = 1 << l;
#pragma unroll
        for (int i = 0; i + s < MM; ++i) {
            const bool take = h[i + s] < h[i];
            h[i] = take ? h[i + s] : h[i];
            if (INDEX) pos[i] = take ? pos[i + s] : pos[i];
        }
    }
    // the window of W = 2^L + d k-mers: two blocks of 2^L, d apart (in place: h[i + d] is still a block minimum
    // when h[i] is overwritten, for every d >= 0)
    constexpr int d = W - (1 << L); // 0 <= d < 2^L
    if (d > 0) {
#pragma unroll
        for (int i = 0; i < G; ++i) {
            const bool take = h[i + d] < h[i];
            h[i] = take ? h[i + d] : h[i];
            if (INDEX) pos[i] = take ? pos[i + d] : pos[i];
        }
    }
    const uint64_t e0 = r * p.cnt + t0;
    uint64_t o[G];
    if (p.out_hash) {
#pragma unroll
        for (int i = 0; i < G; ++i) o[i] = h[i];
        store_dense<G>(p.out_hash + e0, o, nwin);
    }
    if (INDEX) {
        const int64_t ib = static_cast<int64_t>(t0) + 1 + p.index_base;
#pragma unroll
        for (int i = 0; i < G; ++i) o[i] = static_cast<uint64_t>(ib + pos[i]);
        store_dense<G>(reinterpret_cast<uint64_t *>(p.out_index) + e0, o, nwin);
    }
#pragma unroll
    for (int i = 0; i < G; ++i) o[i] = mul_c64<FX_INVERSE>(h[i]); // the k-mer itself
    store_dense<G>(p.out_kmer + e0, o, nwin);
}

using DenseLaunchFn = void (*)(const DenseParams &, unsigned, cudaStream_t);

template <int W, bool CANON, bool INDEX>
void launch_dense(const DenseParams &p, unsigned blocks, cudaStream_t stream)
{
    minimizer_dense_kernel<W, CANON, INDEX><<<blocks, 128, 0, stream>>>(p);
}

template <int W>
DenseLaunchFn pick_dense(bool canon, bool index)
{
    if (canon) return index ? &launch_dense<W, true, true> : &launch_dense<W, true, false>;
    return index ? &launch_dense<W, false, true> : &launch_dense<W, false, false>;
}

template <int W = 1>
DenseLaunchFn dense_launcher(int w, bool canon, bool index)
{
    if (w == W) return pick_dense<W>(canon, index);
    if constexpr (W < 31) return dense_launcher<W + 1>(w, canon, index);
    return nullptr;
}

__global__ void minimizer_counts_kernel(const uint64_t *__restrict__ seq_len, uint64_t n, uint64_t span, uint64_t step,
                                        uint64_t *__restrict__ cnt)
{
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint64_t len = seq_len[i];
        cnt[i] = len >= span ? (len - span) / step + 1 : 0;
    }
}

} // namespace

} // namespace kmc

using namespace kmc;

extern "C" int32_t kmc_minimizers(kmc_ctx *ctx, const kmc_seqs *s, int32_t k, int32_t w, int32_t step, int32_t mode,
                                  uint32_t flags, const kmc_out *out, kmc_result *res)
{
    int32_t st = check_common(ctx, s, k);
    if (st) return st;
    if (!out || !res) return fail(ctx, KMC_E_BAD_ARG, "kmc_out / kmc_result is NULL");
    if (s->src_bits != 2) return fail(ctx, KMC_E_UNSUPPORTED, "minimizers need a 2-bit source");
    if (mode != KMC_FW && mode != KMC_CANON) return fail(ctx, KMC_E_BAD_ARG, "mode must be KMC_FW or KMC_CANON");
    if (k > 32) return fail(ctx, KMC_E_UNSUPPORTED, "minimizers support K <= 32");
    if (w < 1 || step < 1 || k + w - 1 > 64) return fail(ctx, KMC_E_BAD_ARG, "need W >= 1, step >= 1 and K + W - 1 <= 64");
    CU(cudaSetDevice(ctx->device));
    memset(res, 0, sizeof *res);
    cudaStream_t stream = ctx->stream;
    MinimizerParams p{};
    p.words = s->words;
    p.n_words = static_cast<int64_t>(s->n_words);
    p.n_seqs = s->n_seqs;
    p.seq_word_off = s->seq_word_offset;
    p.stride_words = s->uniform_stride_words;
    p.first = s->first_symbol_offset;
    p.k = k;
    p.w = w;
    p.step = step;
    p.canon = mode == KMC_CANON;
    const uint64_t span = static_cast<uint64_t>(k + w - 1);
    CU(cudaEventRecord(ctx->ev_k0, stream));
    if (s->seq_len == nullptr) {
        p.uniform_cnt = s->uniform_len >= span ? (s->uniform_len - span) / step + 1 : 0;
        p.total = p.uniform_cnt * s->n_seqs;
    } else {
        const uint64_t n = s->n_seqs;
        st = ensure_scratch(ctx, 2 * round_up((n + 1) * 8, 256) + round_up(scan_tmp_elems(n) * 8, 256) + 256);
        if (st) return st;
        Scratch sc{static_cast<char *>(ctx->scratch), ctx->scratch_bytes, 0};
        uint64_t *cnt = static_cast<uint64_t *>(sc.take((n + 1) * 8));
        uint64_t *off = static_cast<uint64_t *>(sc.take((n + 1) * 8));
        uint64_t *tmp = static_cast<uint64_t *>(sc.take(scan_tmp_elems(n) * 8));
        if (n) {
            minimizer_counts_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(s->seq_len, n, span, step, cnt);
            CU(cudaGetLastError());
        }
        CU(inclusive_offsets_u64(cnt, off, n, tmp, stream));
        CU(cudaMemcpyAsync(&p.total, off + n, 8, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        p.cnt_off = off;
        if (out->seq_out_offset) CU(cudaMemcpyAsync(out->seq_out_offset, off, (n + 1) * 8, cudaMemcpyDeviceToDevice, stream));
    }
    if (s->seq_len == nullptr && out->seq_out_offset)
        CU(fill_uniform_offsets(out->seq_out_offset, s->n_seqs + 1, p.uniform_cnt, stream));
    res->n_written = p.total;
    if (p.total == 0) return KMC_OK;
    if (p.total > out->capacity) return fail(ctx, KMC_E_OUT_TOO_SMALL, "output capacity smaller than the number of minimizers");
    if (!out->a) return fail(ctx, KMC_E_BAD_ARG, "out.a is NULL");
    if ((flags & KMC_HASH_FX) && !out->hash) return fail(ctx, KMC_E_BAD_ARG, "out.hash is NULL (KMC_HASH_FX)");
    p.out_kmer = out->a;
    p.out_hash = (flags & KMC_HASH_FX) ? out->hash : nullptr;
    p.out_index = out->index;
    p.index_base = out->index_base;
    if (p.total > 0x7fffffffull * 256) return fail(ctx, KMC_E_UNSUPPORTED, "too many minimizers for one launch");
    if (step == 1 && w <= 31 && s->seq_len == nullptr) {
        int l = 0;
        while ((2 << l) <= w) ++l; // 2^l <= w < 2^(l+1)
        const uint64_t g = static_cast<uint64_t>(dense_group(l));
        DenseParams d{};
        d.w32 = reinterpret_cast<const uint32_t *>(s->words);
        d.nw32 = static_cast<int64_t>(s->n_words) * 2;
        d.n_seqs = s->n_seqs;
        d.seq_word_off = s->seq_word_offset;
        d.stride_words = s->uniform_stride_words;
        d.first = s->first_symbol_offset;
        d.cnt = p.uniform_cnt;
        d.ipr = (p.uniform_cnt + g - 1) / g;
        d.items = d.ipr * s->n_seqs;
        d.k = k;
        d.w = w;
        d.mask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1);
        d.out_kmer = p.out_kmer;
        d.out_hash = p.out_hash;
        d.out_index = p.out_index;
        d.index_base = p.index_base;
        DenseLaunchFn fn = dense_launcher<>(w, p.canon != 0, p.out_index != nullptr);
        if (!fn) return fail(ctx, KMC_E_UNSUPPORTED, "no dense minimizer kernel for this W");
        fn(d, static_cast<unsigned>((d.items + 127) / 128), stream);
    } else {
        minimizer_kernel<<<static_cast<unsigned>((p.total + 255) / 256), 256, 0, stream>>>(p);
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev_k1, stream));
    CU(cudaStreamSynchronize(stream));
    CU(cudaEventElapsedTime(&res->kernel_ms, ctx->ev_k0, ctx->ev_k1));
    return KMC_OK;
}
