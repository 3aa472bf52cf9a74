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
//   AsciiEncode, ASCII bytes (construction.jl:95-96; FwKmers.jl:117-129, CanonicalKmers.jl:146-174 with the
//       FourBit branch :160-162): ascii4_recode (ascii.cu) writes the bytes as nibbles -- every IUPAC letter
//       and the gap in either case, BioSequences.ascii_encode of a 4-bit alphabet -- and flags the bytes that
//       are no symbol of the alphabet; the first flagged byte of a sequence of at least K symbols is the
//       EncodeError the reference throws; then the Copyable kernels run on the nibbles.
//
// UnambiguousKmers only exists for 2-bit k-mers (UnambiguousKmers{A<:TwoBit}, UnambiguousKmers.jl:29).
#include "fourbit.h"
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

cudaError_t expand_two_to_four(const uint64_t *words, uint64_t n_words, void *wide, cudaStream_t stream)
{
    if (n_words == 0) return cudaSuccess;
    expand_kernel<<<static_cast<unsigned>((n_words + 255) / 256), 256, 0, stream>>>(words, n_words, static_cast<uint4 *>(wide));
    return cudaGetLastError();
}

int32_t check_kmer4(kmc_ctx *ctx, const kmc_seqs *s, int32_t k, int32_t mode)
{
    if (k > KMC_MAX_K4) return fail(ctx, KMC_E_BAD_K, "K exceeds KMC_MAX_K4 (64) for k-mers over a 4-bit alphabet");
    if (mode == KMC_UNAMBIG)
        return fail(ctx, KMC_E_BAD_ARG, "UnambiguousKmers yields k-mers over a 2-bit alphabet only (UnambiguousKmers{A<:TwoBit})");
    (void)s;
    return KMC_OK;
}

uint64_t kmer4_scratch_bytes(const kmc_seqs *s)
{
    uint64_t need = layout_scratch_bytes(s) + 512;
    if (s->src_bits == 2) need += round_up(16 * (s->n_words + 2), 256);
    if (s->src_bits == 8) {
        const uint64_t nb = (s->n_words + 31) / 32; // groups of 32 bytes
        need += round_up(16 * (nb + 2), 256) + 2 * round_up(4 * (nb + 8), 256) + 1024;
    }
    return need;
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
        CU(expand_two_to_four(s->words, s->n_words, wide, stream));
        p.w32 = reinterpret_cast<const uint32_t *>(wide);
        p.nw32 = static_cast<int64_t>(s->n_words) * 4;
        p.unit_bits = 128; // one source word (32 symbols) = 128 bits of the expanded stream
    } else if (s->src_bits == 8) {
        const uint64_t nb = (s->n_words + 31) / 32;
        const uint8_t *bytes = reinterpret_cast<const uint8_t *>(s->words);
        uint64_t *nib = static_cast<uint64_t *>(scratch.take(16 * (nb + 2)));
        uint32_t *bad = static_cast<uint32_t *>(scratch.take(4 * (nb + 8)));
        uint32_t *vstart = static_cast<uint32_t *>(scratch.take(4 * (nb + 8)));
        unsigned long long *err_seq = static_cast<unsigned long long *>(scratch.take(16)); // [1]: "the recoding pass flagged a byte"
        uint64_t *err_out = static_cast<uint64_t *>(scratch.take(24));
        if (!nib || !bad || !vstart || !err_seq || !err_out) return fail(ctx, KMC_E_BAD_ARG, "internal: scratch window too small");
        st = ensure_host_small(ctx);
        if (st) return st;
        CU(ascii4_recode(bytes, s->n_words, (flags & KMC_RNA) != 0, k, nib, bad, vstart, nb, nb + 2, stream, err_seq + 1));
        p.w32 = reinterpret_cast<const uint32_t *>(nib);
        p.nw32 = static_cast<int64_t>(nb) * 4;
        p.unit_bits = 4; // offsets count bytes = symbols = nibbles of the recoded stream
        // the strict iterators read every symbol of every sequence of at least K symbols, in order: the first byte
        // that is no symbol of the alphabet is the error (FwKmers.jl:124-126).  (The error helpers index the flag
        // bits by symbol: unit_bits = 2 is "one symbol per offset unit" to them.)
        ExtractParams pe = p;
        pe.unit_bits = 2;
        uint64_t *flag = ctx->host_small + 60;
        CU(cudaMemsetAsync(err_seq, 0xff, 8, stream));
        CU(ascii_first_error_seq(pe, bad, s->seq_len, s->uniform_len, err_seq, ctx->sm_count, stream, static_cast<uint64_t>(k), err_seq + 1));
        CU(cudaMemcpyAsync(flag, err_seq, 8, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        if (*flag != ~0ull) {
            CU(ascii_resolve_error(pe, bytes, bad, s->seq_len, s->uniform_len, *flag, err_out, stream));
            CU(cudaMemcpyAsync(flag + 1, err_out, 24, cudaMemcpyDeviceToHost, stream));
            CU(cudaStreamSynchronize(stream));
            res->n_written = 0;
            res->err_seq = flag[1];
            res->err_pos = flag[2];
            res->err_sym = static_cast<uint32_t>(flag[3]);
            return fail(ctx, KMC_E_AMBIGUOUS, "cannot encode this byte in a 4-bit alphabet");
        }
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
