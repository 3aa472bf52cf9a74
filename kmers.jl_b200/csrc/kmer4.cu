// kmer4.cu -- k-mers over the 4-bit alphabets, Kmer{DNAAlphabet{4},K,N} / Kmer{RNAAlphabet{4},K,N} with
// N = cld(4K, 64), for FwKmers / FwRvIterator / CanonicalKmers:
//
//   Copyable, 4-bit source   (FwKmers.jl:88-94, CanonicalKmers.jl:107-120): the extraction kernels run
//       on the LongSequence words as they are, with nibbles for bit pairs (kmer_core.cuh: rev4_32 /
//       comp4_32).  Every symbol is allowed (IUPAC sets, N, gap): nothing to check, nothing to skip.
//   TwoToFour, 2-bit source  (construction.jl:87-88; FwKmers.jl:96-102, CanonicalKmers.jl:122-129):
//       enc4 = 1 << enc2.  expand_kernel turns every source word into 128 bits of one-hot nibbles
//       in scratch memory, then the same kernels run on that stream (one source word = one unit
//       of 128 bits).
//
// UnambiguousKmers only exists for 2-bit k-mers (UnambiguousKmers{A<:TwoBit}, UnambiguousKmers.jl:29).
#include "fourbit_core.cuh"
#include "plan.h"

namespace kmc {

namespace {

// (onehot8: fourbit_core.cuh)

__global__ void __launch_bounds__(256) expand_kernel(const uint64_t *__restrict__ words, uint64_t n_words,
                                                     uint4 *__restrict__ out)
{
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_words) return;
    const uint64_t w = __ldg(words + i);
    const uint32_t lo = static_cast<uint32_t>(w), hi = static_cast<uint32_t>(w >> 32);
    out[i] = make_uint4(onehot8(lo & 0xffffu), onehot8(lo >> 16), onehot8(hi & 0xffffu), onehot8(hi >> 16));
}

ExtractLaunchFn kmer4_launcher(const Geometry &ge, int mode, bool hash, bool ragged)
{
    switch (ge.n_limbs) {
    case 1: return get_kmer4_launcher_n1(ge.nx, mode, hash, ragged);
    case 2: return get_kmer4_launcher_n2(ge.nx, mode, hash, ragged);
    case 3: return get_kmer4_launcher_n3(ge.nx, mode, hash, ragged);
    case 4: return get_kmer4_launcher_n4(ge.nx, mode, hash, ragged);
    }
    return nullptr;
}

} // namespace

int32_t check_kmer4(kmc_ctx *ctx, const kmc_seqs *s, int32_t k, int32_t mode)
{
    if (k > KMC_MAX_K4) return fail(ctx, KMC_E_BAD_K, "K exceeds KMC_MAX_K4 (64) for k-mers over a 4-bit alphabet");
    if (mode == KMC_UNAMBIG)
        return fail(ctx, KMC_E_BAD_ARG, "UnambiguousKmers yields k-mers over a 2-bit alphabet only (UnambiguousKmers{A<:TwoBit})");
    if (s->src_bits == 8) return fail(ctx, KMC_E_UNSUPPORTED, "ASCII sources are recoded to 2-bit k-mers only");
    return KMC_OK;
}

uint64_t kmer4_scratch_bytes(const kmc_seqs *s)
{
    return layout_scratch_bytes(s) + (s->src_bits == 2 ? round_up(16 * (s->n_words + 2), 256) : 0) + 512;
}

int32_t extract_device_kmer4(kmc_ctx *ctx, const kmc_seqs *s, int32_t k, int32_t mode, uint32_t flags, const kmc_out *out,
                             kmc_result *res, cudaStream_t stream, const KnownTotals &known, uint64_t unit_bias, bool sync,
                             Scratch &scratch)
{
    int32_t st = check_kmer4(ctx, s, k, mode);
    if (st) return st;
    const Geometry ge = geometry(k, 4);
    const bool hash = (flags & KMC_HASH_FX) != 0;
    if (sync) CU(cudaEventRecord(ctx->ev_k0, stream));
    Layout L;
    st = plan_layout(ctx, s, k, ge, stream, known, scratch, &L);
    if (st) return st;
    res->n_written = L.total;
    if (out->seq_out_offset) {
        if (L.uniform_len)
            CU(fill_uniform_offsets(out->seq_out_offset, s->n_seqs + 1, L.wpr, stream));
        else
            CU(cudaMemcpyAsync(out->seq_out_offset, L.win_off, (s->n_seqs + 1) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, stream));
    }
    if (L.total == 0) return KMC_OK;
    if (L.total > out->capacity) return fail(ctx, KMC_E_OUT_TOO_SMALL, "output capacity smaller than the number of k-mers");
    ExtractParams p = base_params(s, k, ge, L, unit_bias);
    if (s->src_bits == 2) {
        uint4 *wide = static_cast<uint4 *>(scratch.take(16 * (s->n_words + 2)));
        if (!wide) return fail(ctx, KMC_E_BAD_ARG, "internal: scratch window too small");
        if (s->n_words) {
            expand_kernel<<<static_cast<unsigned>((s->n_words + 255) / 256), 256, 0, stream>>>(s->words, s->n_words, wide);
            CU(cudaGetLastError());
        }
        p.w32 = reinterpret_cast<const uint32_t *>(wide);
        p.nw32 = static_cast<int64_t>(s->n_words) * 4;
        p.unit_bits = 128; // one source word (32 symbols) = 128 bits of the expanded stream
    }
    st = bind_outputs(ctx, out, mode, flags, &p);
    if (st) return st;
    ExtractLaunchFn fn = kmer4_launcher(ge, mode, hash, !L.uniform_len);
    if (!fn) return fail(ctx, KMC_E_UNSUPPORTED, "no kernel for this K");
    CU(fn(p, ctx->sm_count, stream));
    if (sync) {
        CU(cudaEventRecord(ctx->ev_k1, stream));
        CU(cudaStreamSynchronize(stream));
        CU(cudaEventElapsedTime(&res->kernel_ms, ctx->ev_k0, ctx->ev_k1));
    }
    return KMC_OK;
}

} // namespace kmc
