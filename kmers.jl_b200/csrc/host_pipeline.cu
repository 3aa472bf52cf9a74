// host_pipeline.cu -- kmc_extract_host: the entry point with HOST buffers (what a Julia
// `collect(it)` replacement calls).  The set is cut into chunks of ~4 M windows; each chunk is
// uploaded, extracted and downloaded on one of three streams so that H2D, kernels and D2H of
// neighbouring chunks overlap.  Works for 2-bit and 4-bit sources; for UnambiguousKmers over a
// 4-bit source the number of k-mers of a chunk is only known after its count pass, so a chunk
// runs in two phases (fourbit.h) and the download goes to a running host offset.
#include <algorithm>
#include <cstring>
#include <vector>

#include "fourbit.h"
#include "plan.h"

using namespace kmc;

namespace {

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_a = nullptr;
    uint64_t *words = nullptr, *off = nullptr, *len = nullptr, *seq_out = nullptr;
    uint64_t *a = nullptr, *b = nullptr, *hash = nullptr;
    int64_t *index = nullptr;
    Scratch scratch;
    uint64_t *host_small = nullptr;
};

struct Chunk {
    uint64_t seq0 = 0, nseq = 0;     // reads [seq0, seq0+nseq)   (n_seqs > 1)
    uint64_t word0 = 0, nwords = 0;  // words to upload
    uint64_t out0 = 0, nout = 0;     // flat window range (upper bound of the output for UnambiguousKmers/4-bit)
    uint64_t items = 0;              // group slots (ragged)
    uint32_t first = 0;              // first_symbol_offset of the chunk view
    int64_t index_base = 0;
    uint64_t len = 0;                // single-sequence chunks: symbols in view
    // run state
    kmc_seqs ds{};
    kmc_out dout{};
    FourBitState fb;
};

} // namespace

static int32_t run_pipeline(kmc_ctx *ctx, const kmc_seqs *hs, int32_t k, int32_t mode, uint32_t flags, const kmc_out *ho,
                            kmc_result *result);

extern "C" int32_t kmc_extract_host(kmc_ctx *ctx, const kmc_seqs *hs, int32_t k, int32_t mode, uint32_t flags,
                                    const kmc_out *ho, kmc_result *result)
{
    const int32_t st = run_pipeline(ctx, hs, k, mode, flags, ho, result);
    if (st != KMC_OK && ctx) {
        // never return while copies into the caller's buffers are still in flight
        const std::string keep = ctx->last_error;
        for (int i = 0; i < 3; ++i)
            if (ctx->pipe_streams[i]) cudaStreamSynchronize(ctx->pipe_streams[i]);
        ctx->last_error = keep;
    }
    return st;
}

static int32_t run_pipeline(kmc_ctx *ctx, const kmc_seqs *hs, int32_t k, int32_t mode, uint32_t flags, const kmc_out *ho,
                            kmc_result *result)
{
    int32_t st = check_common(ctx, hs, k);
    if (st) return st;
    if (!ho || !result) return fail(ctx, KMC_E_BAD_ARG, "kmc_out / kmc_result is NULL");
    if (mode < KMC_FW || mode > KMC_UNAMBIG) return fail(ctx, KMC_E_BAD_ARG, "unknown mode");
    CU(cudaSetDevice(ctx->device));
    result->n_written = 0;
    result->err_seq = result->err_pos = 0;
    result->err_sym = 0;
    result->kernel_ms = 0.f;
    memset(result->digest, 0, sizeof result->digest);

    const bool kmer4 = (flags & KMC_KMER4) != 0; // k-mers over a 4-bit alphabet (kmer4.cu): no recoding, no checks
    if (kmer4) {
        st = check_kmer4(ctx, hs, k, mode);
        if (st) return st;
    }
    const Geometry ge = geometry(k, kmer4 ? 4 : 2);
    const uint64_t N = static_cast<uint64_t>(ge.n_limbs);
    const bool four = hs->src_bits != 2 && !kmer4; // recoded on the device first: 4-bit words or ASCII bytes (fourbit.h)
    const bool ascii = hs->src_bits == 8;
    const uint64_t spw = ascii ? 1 : hs->src_bits == 4 ? 16 : 32; // symbols per source unit (word, or byte for ASCII)
    const uint64_t unit_bytes = ascii ? 1 : 8;
    const bool hash = (flags & KMC_HASH_FX) != 0;
    const bool aos = (flags & KMC_AOS) != 0;
    const bool dev_out = (flags & KMC_OUT_DEVICE) != 0; // outputs stay in (caller-provided) device memory
    if (dev_out && ho->seq_out_offset) return fail(ctx, KMC_E_UNSUPPORTED, "seq_out_offset is not available with KMC_OUT_DEVICE");
    const bool want_digest = (flags & KMC_DIGEST) != 0;
    const bool want_index = (mode == KMC_UNAMBIG);
    const bool compacting = four && mode == KMC_UNAMBIG; // variable-length output
    // windows per work item of the layout the device planner will build
    const uint64_t G = static_cast<uint64_t>(ge.g);
    const bool two = (mode == KMC_FWRV);
    const bool ragged_len = hs->seq_len != nullptr;
    const bool ragged_off = hs->seq_word_offset != nullptr;
    const uint64_t K = static_cast<uint64_t>(k);
    // u64 per element of out.a
    const uint64_t a_elems = (aos && two) ? 2 * N : (aos && want_index) ? N + 1 : N;

    // ---- chunk plan (host side; lengths are host arrays here) --------------------------------
    std::vector<Chunk> chunks;
    // windows per chunk: 64-128 MB of staged output at 16-32 B / window; with KMC_OUT_DEVICE nothing is
    // staged, so larger chunks (fewer, larger H2D transfers) are free
    const uint64_t target = dev_out ? (32ull << 20) : (4ull << 20);
    uint64_t total = 0;
    const bool single = (hs->n_seqs == 1);
    if (single) {
        const uint64_t len = ragged_len ? hs->seq_len[0] : hs->uniform_len;
        const uint64_t w0 = ragged_off ? hs->seq_word_offset[0] : 0;
        const uint64_t wcount = len >= K ? len - K + 1 : 0;
        total = wcount;
        for (uint64_t s = 0; s < wcount; s += target) {
            Chunk c;
            c.nout = std::min(target, wcount - s);
            c.out0 = s;
            const uint64_t sym0 = hs->first_symbol_offset + s; // absolute first symbol
            const uint64_t sym1 = sym0 + c.nout + K - 1;       // one past last symbol
            c.word0 = w0 + sym0 / spw;
            c.nwords = (sym1 + spw - 1) / spw - sym0 / spw;
            c.first = static_cast<uint32_t>(sym0 % spw);
            c.index_base = static_cast<int64_t>(s);
            c.len = c.nout + K - 1;
            c.seq0 = 0;
            c.nseq = 1;
            chunks.push_back(c);
        }
    } else {
        uint64_t r = 0;
        const uint64_t n = hs->n_seqs;
        const uint64_t wpr_u = (!ragged_len && hs->uniform_len >= K) ? hs->uniform_len - K + 1 : 0;
        while (r < n) {
            Chunk c;
            c.seq0 = r;
            c.out0 = total;
            c.first = hs->first_symbol_offset;
            if (!ragged_len) {
                uint64_t take = wpr_u ? std::max<uint64_t>(1, target / wpr_u) : n;
                if (take > 4) take &= ~3ull; // chunk outputs start on a 32-byte boundary (vector stores, KMC_OUT_DEVICE)
                take = std::min(take, n - r);
                c.nseq = take;
                c.nout = take * wpr_u;
                r += take;
            } else {
                uint64_t acc = 0, items = 0;
                uint64_t extra = 0;
                while (r < n && (acc < target || c.nseq == 0 || ((acc & 3) && extra < 64))) {
                    uint64_t len = hs->seq_len[r];
                    acc += len >= K ? len - K + 1 : 0;
                    ++r;
                    ++c.nseq;
                    if (acc >= target) ++extra; // a few more reads so that the next chunk starts 32-byte aligned
                }
                // group slots are relative to the chunk's own flat origin (0); the host only needs
                // their total so that the device planner does not have to be read back
                uint64_t fo = 0;
                for (uint64_t q = c.seq0; q < c.seq0 + c.nseq; ++q) {
                    uint64_t len = hs->seq_len[q];
                    uint64_t wc = len >= K ? len - K + 1 : 0;
                    if (wc) items += (fo + wc + G - 1) / G - fo / G;
                    fo += wc;
                }
                c.nout = acc;
                c.items = items;
            }
            if (ragged_off) {
                c.word0 = hs->seq_word_offset[c.seq0];
                uint64_t last = c.seq0 + c.nseq - 1;
                uint64_t last_len = ragged_len ? hs->seq_len[last] : hs->uniform_len;
                uint64_t end = hs->seq_word_offset[last] + (hs->first_symbol_offset + last_len + spw - 1) / spw;
                c.nwords = end > c.word0 ? end - c.word0 : 0;
            } else {
                c.word0 = c.seq0 * hs->uniform_stride_words;
                c.nwords = c.nseq * hs->uniform_stride_words;
            }
            c.nwords = std::min(c.nwords, hs->n_words > c.word0 ? hs->n_words - c.word0 : 0);
            total += c.nout;
            chunks.push_back(c);
        }
    }
    if (!compacting) {
        result->n_written = total;
        if (ho->seq_out_offset) {
            // per-sequence output offsets are a host-side prefix sum here
            uint64_t acc = 0;
            for (uint64_t r = 0; r < hs->n_seqs; ++r) {
                ho->seq_out_offset[r] = acc;
                uint64_t len = ragged_len ? hs->seq_len[r] : hs->uniform_len;
                acc += len >= K ? len - K + 1 : 0;
            }
            ho->seq_out_offset[hs->n_seqs] = acc;
        }
        if (total == 0) return KMC_OK;
        if (total > ho->capacity) return fail(ctx, KMC_E_OUT_TOO_SMALL, "output capacity smaller than the number of k-mers");
    } else if (ho->seq_out_offset) {
        for (uint64_t r = 0; r <= hs->n_seqs; ++r) ho->seq_out_offset[r] = 0;
    }
    if (total == 0) return KMC_OK;
    if (!ho->a) return fail(ctx, KMC_E_BAD_ARG, "out.a is NULL");
    if (two && !aos && !ho->b) return fail(ctx, KMC_E_BAD_ARG, "out.b is NULL (FWRV, SoA)");
    if (hash && !ho->hash) return fail(ctx, KMC_E_BAD_ARG, "out.hash is NULL (KMC_HASH_FX)");
    if (want_index && !aos && !ho->index) return fail(ctx, KMC_E_BAD_ARG, "out.index is NULL (UNAMBIG, SoA)");

    // UnambiguousKmers over a recoded source: are the sequences ascending and disjoint in the buffer?  (The device
    // path would have to check its device-resident offsets and synchronise; here the offsets are host arrays.)
    int host_linear = -1;
    if (compacting && ragged_off && !single) {
        host_linear = 1;
        uint64_t prev_end = 0;
        for (uint64_t r = 0; r < hs->n_seqs && host_linear; ++r) {
            const uint64_t len = ragged_len ? hs->seq_len[r] : hs->uniform_len;
            const uint64_t start = hs->seq_word_offset[r] * spw;
            if (r && start < prev_end) host_linear = 0;
            prev_end = start + (len >= K ? len - K + 1 : 0);
        }
    }

    // ---- slot buffers --------------------------------------------------------------------------
    uint64_t max_words = 0, max_out = 0, max_seq = 0;
    for (const Chunk &c : chunks) {
        max_words = std::max(max_words, c.nwords);
        max_out = std::max(max_out, c.nout);
        max_seq = std::max(max_seq, c.nseq);
    }
    const bool want_seq_out = compacting && ho->seq_out_offset && !single;
    const uint64_t b_words = round_up((max_words + 32) * unit_bytes, 256);
    const uint64_t b_meta = round_up((max_seq + 1) * 8, 256);
    const uint64_t b_a = dev_out ? 0 : round_up(max_out * a_elems * 8, 256);
    const uint64_t b_b = (two && !aos && !dev_out) ? round_up(max_out * N * 8, 256) : 0;
    const uint64_t b_h = (hash && !dev_out) ? round_up(max_out * 8, 256) : 0;
    const uint64_t b_i = (want_index && !aos && !dev_out) ? round_up(max_out * 8, 256) : 0;
    const uint64_t b_so = want_seq_out ? b_meta : 0;
    const uint64_t per_slot = b_words + 2 * b_meta + b_so + b_a + b_b + b_h + b_i;
    const int n_slots = chunks.size() >= 3 ? 3 : static_cast<int>(chunks.size());

    kmc_seqs worst = *hs; // the largest chunk, for scratch sizing
    worst.n_words = max_words;
    worst.n_seqs = single ? 1 : max_seq;
    if (single) worst.seq_len = nullptr;
    const uint64_t scratch_per_slot = round_up(kmer4 ? kmer4_scratch_bytes(&worst) : extract_scratch_bytes(&worst, k, mode), 256);
    st = ensure_scratch(ctx, scratch_per_slot * n_slots);
    if (st) return st;
    st = ensure_host_small(ctx);
    if (st) return st;
    if (want_digest) {
        // the three slot streams accumulate into it with atomics; ordered after this memset by an event
        CU(cudaMemsetAsync(ctx->dev_small, 0, 32, ctx->pipe_streams[0]));
        if (!ctx->ev_k0) return KMC_E_BAD_ARG;
        CU(cudaEventRecord(ctx->ev_k0, ctx->pipe_streams[0]));
        for (int i = 1; i < 3; ++i) CU(cudaStreamWaitEvent(ctx->pipe_streams[i], ctx->ev_k0, 0));
    }

    Slot slots[3];
    for (int i = 0; i < n_slots; ++i) {
        if (ctx->pipe_bytes[i] < per_slot) {
            CU(cudaStreamSynchronize(ctx->pipe_streams[i]));
            if (ctx->pipe_buf[i]) CU(cudaFree(ctx->pipe_buf[i]));
            ctx->pipe_buf[i] = nullptr;
            ctx->pipe_bytes[i] = 0;
            CU(cudaMalloc(&ctx->pipe_buf[i], per_slot));
            ctx->pipe_bytes[i] = per_slot;
        }
        if (!ctx->pipe_events[i]) CU(cudaEventCreateWithFlags(&ctx->pipe_events[i], cudaEventDisableTiming));
        char *base = static_cast<char *>(ctx->pipe_buf[i]);
        Slot &sl = slots[i];
        sl.stream = ctx->pipe_streams[i];
        sl.ev_a = ctx->pipe_events[i];
        sl.words = reinterpret_cast<uint64_t *>(base); base += b_words;
        sl.off = reinterpret_cast<uint64_t *>(base); base += b_meta;
        sl.len = reinterpret_cast<uint64_t *>(base); base += b_meta;
        sl.seq_out = b_so ? reinterpret_cast<uint64_t *>(base) : nullptr; base += b_so;
        sl.a = b_a ? reinterpret_cast<uint64_t *>(base) : nullptr; base += b_a;
        sl.b = b_b ? reinterpret_cast<uint64_t *>(base) : nullptr; base += b_b;
        sl.hash = b_h ? reinterpret_cast<uint64_t *>(base) : nullptr; base += b_h;
        sl.index = b_i ? reinterpret_cast<int64_t *>(base) : nullptr;
        sl.scratch.base = static_cast<char *>(ctx->scratch) + scratch_per_slot * i;
        sl.scratch.bytes = scratch_per_slot;
        sl.host_small = ctx->host_small + 16 * i;
    }

    // ---- pipeline ------------------------------------------------------------------------------
    uint64_t emitted = 0; // compacting: elements written to the host so far
    struct Fixup { uint64_t seq0, nseq, base; };
    std::vector<Fixup> fixups;

    auto download = [&](Slot &sl, uint64_t host_elem0, uint64_t n) -> int32_t {
        if (!n || dev_out) return KMC_OK;
        cudaStream_t sm = sl.stream;
        CU(cudaMemcpyAsync(ho->a + host_elem0 * a_elems, sl.a, n * a_elems * 8, cudaMemcpyDeviceToHost, sm));
        if (sl.b) CU(cudaMemcpyAsync(ho->b + host_elem0 * N, sl.b, n * N * 8, cudaMemcpyDeviceToHost, sm));
        if (sl.hash) CU(cudaMemcpyAsync(ho->hash + host_elem0, sl.hash, n * 8, cudaMemcpyDeviceToHost, sm));
        if (sl.index) CU(cudaMemcpyAsync(ho->index + host_elem0, sl.index, n * 8, cudaMemcpyDeviceToHost, sm));
        return KMC_OK;
    };

    // where a chunk's k-mers go: the slot's staging buffers, or (KMC_OUT_DEVICE) the caller's device
    // buffers at the chunk's element offset
    auto bind_chunk_out = [&](Chunk &c, Slot &sl, uint64_t elem0, uint64_t cap) {
        kmc_out &d = c.dout;
        if (dev_out) {
            d.a = ho->a + elem0 * a_elems;
            d.b = (two && !aos) ? ho->b + elem0 * N : nullptr;
            d.hash = hash ? ho->hash + elem0 : nullptr;
            d.index = (want_index && !aos) ? ho->index + elem0 : nullptr;
        } else {
            d.a = sl.a;
            d.b = sl.b;
            d.hash = sl.hash;
            d.index = sl.index;
        }
        d.capacity = cap;
    };

    // KMC_DIGEST: fingerprint the chunk's k-mers and hashes right behind the kernel that wrote them
    // (the 2-bit SoA forms of FwKmers / CanonicalKmers fingerprint their output inside the extraction kernel)
    const bool fused_digest = want_digest && digest_fusable(hs, ge.n_limbs, mode, flags);
    auto digest_chunk = [&](Chunk &c, Slot &sl, uint64_t n) -> int32_t {
        if (!want_digest || !n || fused_digest) return KMC_OK;
        CU(launch_digest(c.dout.a, n * a_elems, ctx->dev_small, ctx->sm_count, sl.stream, false));
        if (hash) CU(launch_digest(c.dout.hash, n, ctx->dev_small + 2, ctx->sm_count, sl.stream, false));
        return KMC_OK;
    };

    // phase B of a 4-bit chunk: runs once the chunk's phase A has completed on the device
    auto finish = [&](size_t ci) -> int32_t {
        Chunk &c = chunks[ci];
        Slot &sl = slots[ci % n_slots];
        CU(cudaEventSynchronize(sl.ev_a));
        kmc_result r{};
        if (compacting) bind_chunk_out(c, sl, emitted, dev_out ? ho->capacity - emitted : c.nout);
        int32_t rc = fourbit_phase_b(ctx, &c.fb, &c.dout, sl.stream, &r);
        if (rc == KMC_E_AMBIGUOUS) {
            result->n_written = 0;
            result->err_seq = single ? 0 : c.seq0 + r.err_seq;
            result->err_pos = r.err_pos + (single ? static_cast<uint64_t>(c.index_base) : 0);
            result->err_sym = r.err_sym;
            for (int i = 0; i < n_slots; ++i) cudaStreamSynchronize(slots[i].stream);
            return rc;
        }
        if (rc) return rc;
        if (compacting) {
            if (emitted + r.n_written > ho->capacity)
                return fail(ctx, KMC_E_OUT_TOO_SMALL, "output capacity smaller than the number of k-mers");
            rc = digest_chunk(c, sl, r.n_written);
            if (rc) return rc;
            rc = download(sl, emitted, r.n_written);
            if (rc) return rc;
            if (want_seq_out) {
                CU(cudaMemcpyAsync(ho->seq_out_offset + c.seq0, sl.seq_out, c.nseq * 8, cudaMemcpyDeviceToHost, sl.stream));
                fixups.push_back({c.seq0, c.nseq, emitted});
            }
            emitted += r.n_written;
        }
        return KMC_OK;
    };

    for (size_t ci = 0; ci < chunks.size(); ++ci) {
        Chunk &c = chunks[ci];
        Slot &sl = slots[ci % n_slots];
        cudaStream_t sm = sl.stream;
        if (c.nwords)
            CU(cudaMemcpyAsync(sl.words, reinterpret_cast<const char *>(hs->words) + c.word0 * unit_bytes, c.nwords * unit_bytes,
                               cudaMemcpyHostToDevice, sm));
        kmc_seqs &ds = c.ds;
        ds = *hs;
        ds.words = sl.words;
        ds.n_words = c.nwords;
        ds.n_seqs = c.nseq;
        ds.first_symbol_offset = c.first;
        ds.seq_word_offset = nullptr;
        ds.seq_len = nullptr;
        uint64_t bias = 0;
        KnownTotals known;
        known.linear = host_linear;
        if (single) {
            ds.uniform_len = c.len;
            ds.uniform_stride_words = c.nwords;
        } else {
            if (ragged_off) {
                CU(cudaMemcpyAsync(sl.off, hs->seq_word_offset + c.seq0, c.nseq * 8, cudaMemcpyHostToDevice, sm));
                ds.seq_word_offset = sl.off;
                bias = c.word0;
            }
            if (ragged_len) {
                CU(cudaMemcpyAsync(sl.len, hs->seq_len + c.seq0, c.nseq * 8, cudaMemcpyHostToDevice, sm));
                ds.seq_len = sl.len;
                known.valid = true;
                known.windows = c.nout;
                known.items = c.items;
            }
        }
        kmc_out &dout = c.dout;
        dout = kmc_out{};
        bind_chunk_out(c, sl, c.out0, c.nout);
        dout.seq_out_offset = want_seq_out ? sl.seq_out : nullptr;
        dout.index_base = c.index_base + ho->index_base;
        Scratch scratch = sl.scratch;
        scratch.used = 0;
        if (!four) {
            kmc_result r{};
            if (fused_digest) known.digest = reinterpret_cast<unsigned long long *>(ctx->dev_small);
            st = kmer4 ? extract_device_kmer4(ctx, &ds, k, mode, flags, &dout, &r, sm, known, bias, false, scratch)
                       : extract_device(ctx, &ds, k, mode, flags, &dout, &r, sm, known, bias, false, scratch);
            if (st == KMC_E_AMBIGUOUS) { // 4-bit k-mers from ASCII bytes: a byte that is no symbol of the alphabet
                result->n_written = 0;
                result->err_seq = single ? 0 : c.seq0 + r.err_seq;
                result->err_pos = r.err_pos + (single ? static_cast<uint64_t>(c.index_base) : 0);
                result->err_sym = r.err_sym;
            }
            if (st) return st;
            st = digest_chunk(c, sl, c.nout);
            if (st) return st;
            st = download(sl, c.out0, c.nout);
            if (st) return st;
        } else {
            // count first: the next chunk's place in the caller's buffer depends on this chunk's survivors
            st = fourbit_phase_a(ctx, &ds, k, mode, flags, &dout, sm, known, bias, scratch, sl.host_small, true, &c.fb);
            if (st) return st;
            if (!compacting) {
                st = digest_chunk(c, sl, c.nout);
                if (st) return st;
                st = download(sl, c.out0, c.nout);
                if (st) return st;
            }
            CU(cudaEventRecord(sl.ev_a, sm));
            if (ci >= 1) {
                st = finish(ci - 1);
                if (st) return st;
            }
        }
    }
    if (four) {
        st = finish(chunks.size() - 1);
        if (st) return st;
    }
    for (int i = 0; i < n_slots; ++i) CU(cudaStreamSynchronize(slots[i].stream));
    if (want_digest) CU(cudaMemcpy(result->digest, ctx->dev_small, 32, cudaMemcpyDeviceToHost));
    if (compacting) {
        result->n_written = emitted;
        if (ho->seq_out_offset) {
            if (single) {
                ho->seq_out_offset[0] = 0;
                ho->seq_out_offset[1] = emitted;
            } else {
                for (const Fixup &f : fixups)
                    for (uint64_t r = f.seq0; r < f.seq0 + f.nseq; ++r) ho->seq_out_offset[r] += f.base;
                ho->seq_out_offset[hs->n_seqs] = emitted;
            }
        }
    }
    return KMC_OK;
}
