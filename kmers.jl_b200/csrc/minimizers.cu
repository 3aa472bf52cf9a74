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
// One thread per window start.  The K+W-1 symbols of the window (<= 64 symbols = 128 bits) are
// fetched once; k-mer j is a shift of the 2-bit-reversed block (forward) or of the complemented
// raw block (reverse complement), exactly as in kmer_core.cuh.
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
    // rev2 of the whole block: symbol i -> L-1-i
    unsigned __int128 T = (static_cast<unsigned __int128>(rev2_64(lo)) << 64) | rev2_64(hi);
    T >>= (128 - 2 * L);
    const uint64_t mask = p.k == 32 ? ~0ull : ((1ull << (2 * p.k)) - 1);
    uint64_t best_k = 0, best_h = 0;
    int best_j = 0;
    for (int j = 0; j < p.w; ++j) {
        uint64_t km = static_cast<uint64_t>(T >> (2 * (p.w - 1 - j))) & mask;
        if (p.canon) {
            const uint64_t rv = ~static_cast<uint64_t>(S >> (2 * j)) & mask;
            km = km < rv ? km : rv;
        }
        const uint64_t h = km * FX_CONSTANT; // fx_hash of a one-limb k-mer, h0 = 0 (src/kmer.jl:255-261)
        if (j == 0 || h < best_h) {
            best_h = h;
            best_k = km;
            best_j = j;
        }
    }
    p.out_kmer[e] = best_k;
    if (p.out_hash) p.out_hash[e] = best_h;
    if (p.out_index) p.out_index[e] = static_cast<int64_t>(sym) + best_j + 1 + p.index_base;
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
    minimizer_kernel<<<static_cast<unsigned>((p.total + 255) / 256), 256, 0, stream>>>(p);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev_k1, stream));
    CU(cudaStreamSynchronize(stream));
    CU(cudaEventElapsedTime(&res->kernel_ms, ctx->ev_k0, ctx->ev_k1));
    return KMC_OK;
}
