// valid_count.cu -- how many k-mers UnambiguousKmers will emit over a recoded (4-bit / ASCII) source
// (UnambiguousKmers.jl:109-148), from the valid-start BIT stream alone (1 bit per window).
//
// Used where the number is needed BEFORE the k-mers are written: kmc_count, and the host pipeline,
// which has to know where a chunk's k-mers go in the caller's buffer before it enqueues the chunk.
// The device-resident extraction does not run it: compact_kernel finds its tiles' places itself.
//
// One thread per work item (read, group slot) of the set's layout, kTileItems items per block like
// the extraction kernels; every block adds its survivors to one device counter.
#include <cstdlib>

#include "fourbit.h"
#include "lincompact.cuh"

namespace kmc {

namespace {

template <bool RAGGED, int G>
__global__ void __launch_bounds__(kBlockThreads) count_valid_kernel(const ExtractParams p, unsigned long long *__restrict__ total)
{
    __shared__ TileShared<RAGGED> sh;
    __shared__ uint32_t s_v[kBlockThreads / 32];
    const uint64_t tile_base = static_cast<uint64_t>(blockIdx.x) * kTileItems;
    TileCursor<RAGGED, G> cur;
    cur.init(p, tile_base, sh, threadIdx.x);
    uint32_t nv = 0;
#pragma unroll 1
    for (int it = 0; it < kTileIters; ++it) {
        const uint32_t li = static_cast<uint32_t>(it) * kBlockThreads + threadIdx.x;
        const uint64_t item = tile_base + li;
        if (item >= p.items) break;
        cur.locate(p, item, li, sh);
        if (cur.jhi > cur.jlo) {
            // valid-start bits of the slots [jlo, jhi): up to 32 of them
            const int64_t q = (cur.bit(p) >> 1) + cur.jlo;
            const int width = cur.jhi - cur.jlo;
            const uint32_t w0 = __ldg(p.vstart + (q >> 5)), w1 = __ldg(p.vstart + (q >> 5) + 1);
            const uint32_t bits = __funnelshift_r(w0, w1, static_cast<uint32_t>(q) & 31u);
            nv += __popc(bits & (width == 32 ? 0xffffffffu : ((1u << width) - 1u)));
        }
        cur.advance(p);
    }
    nv = __reduce_add_sync(0xffffffffu, nv);
    if ((threadIdx.x & 31) == 0) s_v[threadIdx.x >> 5] = nv;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tv = 0;
#pragma unroll
        for (int w = 0; w < kBlockThreads / 32; ++w) tv += s_v[w];
        if (tv) atomicAdd(total, static_cast<unsigned long long>(tv));
    }
}

template <bool RAGGED, int G>
cudaError_t launch_count(ExtractParams p, unsigned long long *total, cudaStream_t stream)
{
    const uint64_t tiles = (p.items + kTileItems - 1) / kTileItems;
    if (tiles == 0) return cudaSuccess;
    if (tiles > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    set_iteration_strides(p);
    count_valid_kernel<RAGGED, G><<<static_cast<unsigned>(tiles), kBlockThreads, 0, stream>>>(p, total);
    return cudaGetLastError();
}

template <int G>
cudaError_t launch_count_g(const ExtractParams &p, bool ragged, unsigned long long *total, cudaStream_t stream)
{
    return ragged ? launch_count<true, G>(p, total, stream) : launch_count<false, G>(p, total, stream);
}

} // namespace

// *total must be zero (it is accumulated into); g = windows per work item of the layout p describes
cudaError_t count_valid(const ExtractParams &p, bool ragged, int g, unsigned long long *total, cudaStream_t stream)
{
    switch (g) {
    case 2: return launch_count_g<2>(p, ragged, total, stream);
    case 4: return launch_count_g<4>(p, ragged, total, stream);
    case 8: return launch_count_g<8>(p, ragged, total, stream);
    case 32: return launch_count_g<32>(p, ragged, total, stream);
    }
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------------------
// lin_prepare: the source-order compaction's preparation (lincompact.cuh).  Everything here is proportional to
// the valid-start BITS (1 bit per symbol) or to the number of sequences -- a few dozen microseconds beside the
// milliseconds of the compaction itself.
// ---------------------------------------------------------------------------------------------------------
namespace {

struct LinGeom {
    uint32_t *vstart;   // bit P = no uncertain symbol in [P, P + K) of the recoded stream
    uint32_t *packed;   // uniform sets: the position bits written here (positions (r, u < w8))
    uint64_t n_chunks;
    uint64_t n_seqs;
    const uint64_t *seq_unit_off; // or NULL: sequence r starts at unit r * stride_units
    uint64_t unit_bias, stride_syms;
    uint32_t spu, first;
    const uint64_t *seq_len; // or NULL: uniform_len
    uint64_t uniform_len;
    uint64_t k;
    uint64_t wpr; // uniform sets: windows per sequence
    uint32_t w8;  // uniform sets: positions per sequence
};

__device__ __forceinline__ uint64_t lin_start(const LinGeom &g, uint64_t r)
{
    return (g.seq_unit_off ? (__ldg(g.seq_unit_off + r) - g.unit_bias) * g.spu : r * g.stride_syms) + g.first;
}
__device__ __forceinline__ uint64_t lin_windows(const LinGeom &g, uint64_t r)
{
    const uint64_t len = g.seq_len ? __ldg(g.seq_len + r) : g.uniform_len;
    return len >= g.k ? len - g.k + 1 : 0;
}
__device__ __forceinline__ uint32_t bit_range(uint32_t lo, uint32_t hi) // bits [lo, hi), 0 <= lo < hi <= 32
{
    return (hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
}

// offsets given: the window starts of consecutive sequences must be ascending and disjoint in the stream
__global__ void __launch_bounds__(256) lin_check_kernel(const LinGeom g, unsigned long long *__restrict__ bad)
{
    const uint64_t r = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r == 0 || r >= g.n_seqs) return;
    if (lin_start(g, r) < lin_start(g, r - 1) + lin_windows(g, r - 1)) *bad = 1ull;
}

// offsets given: thread r clears the bits between the last window start of sequence r - 1 and the first of
// sequence r (thread n_seqs: from the last sequence to the end of the bit array).  The gaps are disjoint, so a
// word that lies wholly inside one gap is touched by nobody else; the two edge words are shared with the
// neighbouring sequences' bits and are cleared with atomicAnd.
__global__ void __launch_bounds__(256) lin_mask_offsets_kernel(const LinGeom g)
{
    const uint64_t r = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r > g.n_seqs) return;
    const uint64_t total = g.n_chunks * kLinChunkPos;
    uint64_t a = r == 0 ? 0 : lin_start(g, r - 1) + lin_windows(g, r - 1);
    uint64_t b = r == g.n_seqs ? total : lin_start(g, r);
    if (b > total) b = total;
    if (a >= b) return;
    const uint64_t wa = a >> 5, wb = (b - 1) >> 5;
    const uint32_t la = static_cast<uint32_t>(a) & 31u, lb = (static_cast<uint32_t>(b - 1) & 31u) + 1u;
    if (wa == wb) {
        atomicAnd(g.vstart + wa, ~bit_range(la, lb));
        return;
    }
    atomicAnd(g.vstart + wa, ~bit_range(la, 32));
    for (uint64_t w = wa + 1; w < wb; ++w) g.vstart[w] = 0;
    atomicAnd(g.vstart + wb, ~bit_range(0, lb));
}

// n <= 32 valid-start bits from bit `sym` on
__device__ __forceinline__ uint32_t vstart_bits(const uint32_t *__restrict__ vs, uint64_t sym, uint32_t n)
{
    const uint64_t w = sym >> 5;
    const uint32_t b = static_cast<uint32_t>(sym) & 31u;
    const uint32_t lo = __ldg(vs + w), hi = b + n > 32 ? __ldg(vs + w + 1) : 0u;
    const uint32_t v = __funnelshift_r(lo, hi, b);
    return n >= 32 ? v : v & ((1u << n) - 1u);
}

// One warp per chunk of 64 words of position bits: count the survivors of the chunk.
//   uniform sets   the position bits are gathered first: position p = (r, u) = divmod(p, w8) is a window start iff
//                  u < wpr, and survives iff the valid-start bit of symbol r * stride + first + u is set;
//   offsets given  the bits are the (already masked) valid-start bits; also the sequence that owns the chunk's
//                  first symbol is looked up.
template <bool OFFSETS>
__global__ void __launch_bounds__(256) lin_chunk_kernel(const LinGeom g, uint64_t *__restrict__ cnt, uint64_t *__restrict__ chunk_first)
{
    const uint64_t c = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= g.n_chunks) return;
    uint32_t n = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint64_t w = c * kLinChunkWords + lane + 32 * h;
        uint32_t v;
        if (OFFSETS) {
            v = g.vstart[w];
        } else {
            // positions [32 w, 32 w + 32): runs of consecutive windows of one sequence each; the run of sequence r starts
            // at its window u and ends with the sequence's windows.  With at least 32 windows per sequence a word holds at
            // most two runs, and both gathers are issued before either is used (the kernel is a chain of dependent loads).
            v = 0;
            const uint64_t p0 = w * 32;
            uint64_t r;
            uint32_t u;
            if (g.n_seqs == 1) {
                r = p0 < g.w8 ? 0 : 1;
                u = static_cast<uint32_t>(p0 < g.w8 ? p0 : 0);
            } else {
                r = (p0 >> 32) == 0 ? static_cast<uint32_t>(p0) / g.w8 : p0 / g.w8;
                u = static_cast<uint32_t>(p0 - r * g.w8);
            }
            {
                const uint32_t n0 = min(32u, g.w8 - u);
                const uint32_t b0 = r < g.n_seqs ? vstart_bits(g.vstart, r * g.stride_syms + g.first + u, n0) : 0u;
                const uint32_t b1 = (n0 < 32 && r + 1 < g.n_seqs) ? vstart_bits(g.vstart, (r + 1) * g.stride_syms + g.first, 32 - n0) : 0u;
                v = b0 | (n0 < 32 ? b1 << n0 : 0u);
            }
            g.packed[w] = v;
        }
        n += __popc(v);
    }
    n = __reduce_add_sync(0xffffffffu, n);
    if (lane == 0) {
        cnt[c] = n;
        if (OFFSETS) {
            const uint64_t sym = c * kLinChunkPos;
            uint64_t lo = 0, hi = g.n_seqs; // largest r whose first symbol is <= sym (0 if none)
            while (hi - lo > 1) {
                const uint64_t mid = (lo + hi) >> 1;
                if (lin_start(g, mid) <= sym) lo = mid; else hi = mid;
            }
            chunk_first[c] = lo;
            if (c + 1 == g.n_chunks) chunk_first[c + 1] = g.n_seqs - 1;
        }
    }
}

} // namespace

// KMC_LINEAR=0 keeps every set on compact_kernel (A/B measurements, and the tests of that path)
bool lin_enabled()
{
    static const bool on = [] {
        const char *e = getenv("KMC_LINEAR");
        return !(e && e[0] == '0');
    }();
    return on;
}

// A set without offsets and with host-known lengths takes the source-order compaction iff this holds (lin_prepare
// applies the same test; phase A of fourbit.cu asks before it decides which streams the recoding pass writes)
bool lin_uniform_ok(const kmc_seqs *s, int k)
{
    if (s->n_seqs == 0 || s->seq_word_offset != nullptr || s->seq_len != nullptr) return false;
    const uint64_t K = static_cast<uint64_t>(k), spu = s->src_bits == 8 ? 1 : 16;
    const uint64_t wpr = s->uniform_len >= K ? s->uniform_len - K + 1 : 0, stride = s->uniform_stride_words * spu;
    if (wpr < 32 || wpr > 0x7fffffffull) return false; // (fewer than 32 windows per sequence: compact_kernel; the kernel's steps cross at most one boundary)
    if (s->n_seqs > 1 && stride < wpr) return false; // overlapping sequences
    const uint64_t jump = s->n_seqs > 1 ? stride - wpr : 0;
    return jump < (1ull << 30) && (kLinChunkPos / wpr + 2) * jump + kLinChunkPos < (1ull << 30);
}

uint64_t lin_chunks(uint64_t n_positions) { return (n_positions + kLinChunkPos - 1) / kLinChunkPos; }

// scratch lin_prepare can take for a set of n_seqs sequences over n_symbols symbols (any K, any layout)
uint64_t lin_scratch_bytes(uint64_t n_symbols, uint64_t n_seqs)
{
    const uint64_t nc = lin_chunks(n_symbols + 64) + 1;
    return 3 * round_up(8 * (nc + 2), 256) + round_up(8 * (scan_tmp_elems(nc) + 1), 256) + round_up(4 * nc * kLinChunkWords, 256) + 1024;
}

// Decides whether the set can be compacted in source order and, if so, prepares the position bits and lays the
// chunks out.  plan->linear is the answer.  n_vstart_words: words of valid-start bits the recoding pass has written
// (the array has room for whole chunks beyond).  `host_flag`: a pinned u64 the device check is copied to; the call
// synchronises the stream only when the set gives offsets and the caller does not know the answer (known_linear < 0).
int32_t lin_prepare(kmc_ctx *ctx, const ExtractParams &p, const kmc_seqs *s, int k, uint64_t n_vstart_words,
                    int known_linear, uint64_t *host_flag, cudaStream_t stream, Scratch &scratch, LinPlan *plan)
{
    plan->linear = false;
    if (!lin_enabled() || known_linear == 0 || s->n_seqs == 0) return KMC_OK;
    const bool offsets = s->seq_word_offset != nullptr;
    if (!offsets && s->seq_len != nullptr) return KMC_OK; // uniform stride with device-resident lengths: compact_kernel
    LinGeom g;
    g.vstart = const_cast<uint32_t *>(p.vstart);
    g.packed = nullptr;
    g.n_seqs = s->n_seqs;
    g.seq_unit_off = offsets ? s->seq_word_offset : nullptr;
    g.unit_bias = p.unit_bias;
    g.spu = p.unit_bits >> 1;
    g.stride_syms = s->uniform_stride_words * g.spu;
    g.first = s->first_symbol_offset;
    g.seq_len = s->seq_len;
    g.uniform_len = s->uniform_len;
    g.k = static_cast<uint64_t>(k);
    g.wpr = 0;
    g.w8 = 0;
    uint64_t n_pos;
    if (offsets) {
        n_pos = n_vstart_words * 32; // positions are the symbols of the stream
    } else {
        // (no overlapping sequences; and the kernel's 32-bit arithmetic: the symbols of one chunk of 2048 positions must
        // span less than 2^30)
        if (!lin_uniform_ok(s, k)) return KMC_OK;
        g.wpr = s->uniform_len - g.k + 1;
        const uint64_t w8 = g.wpr; // positions per sequence = its windows: the tails and the padding are no positions
        g.w8 = static_cast<uint32_t>(w8);
        n_pos = s->n_seqs * w8;
    }
    const uint64_t nc = lin_chunks(n_pos);
    g.n_chunks = nc;
    uint64_t *cnt = static_cast<uint64_t *>(scratch.take(8 * (nc + 2)));
    uint64_t *chunk_off = static_cast<uint64_t *>(scratch.take(8 * (nc + 2)));
    uint64_t *chunk_first = static_cast<uint64_t *>(scratch.take(8 * (nc + 2)));
    uint64_t *tmp = static_cast<uint64_t *>(scratch.take(8 * (scan_tmp_elems(nc) + 1)));
    unsigned long long *flag = static_cast<unsigned long long *>(scratch.take(8));
    if (!offsets) g.packed = static_cast<uint32_t *>(scratch.take(4 * (nc * kLinChunkWords + 8)));
    if (!cnt || !chunk_off || !chunk_first || !tmp || !flag || (!offsets && !g.packed))
        return fail(ctx, KMC_E_BAD_ARG, "internal: scratch window too small");
    if (offsets) {
        if (known_linear < 0 && s->n_seqs > 1) {
            CU(cudaMemsetAsync(flag, 0, 8, stream));
            lin_check_kernel<<<static_cast<unsigned>((s->n_seqs + 255) / 256), 256, 0, stream>>>(g, flag);
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(host_flag, flag, 8, cudaMemcpyDeviceToHost, stream));
            CU(cudaStreamSynchronize(stream));
            if (*host_flag) return KMC_OK; // out of order or overlapping: compact_kernel handles any layout
        }
        lin_mask_offsets_kernel<<<static_cast<unsigned>((s->n_seqs + 1 + 255) / 256), 256, 0, stream>>>(g);
        CU(cudaGetLastError());
        lin_chunk_kernel<true><<<static_cast<unsigned>((nc * 32 + 255) / 256), 256, 0, stream>>>(g, cnt, chunk_first);
    } else {
        lin_chunk_kernel<false><<<static_cast<unsigned>((nc * 32 + 255) / 256), 256, 0, stream>>>(g, cnt, nullptr);
    }
    CU(cudaGetLastError());
    CU(inclusive_offsets_u64(cnt, chunk_off, nc, tmp, stream));
    plan->linear = true;
    plan->lp.bits = offsets ? g.vstart : g.packed;
    plan->lp.chunk_off = chunk_off;
    plan->lp.chunk_first = offsets ? chunk_first : nullptr;
    plan->lp.n_chunks = nc;
    plan->lp.capacity = 0;
    plan->lp.stride_syms = g.stride_syms;
    plan->lp.w8 = g.w8;
    plan->lp.jump = (!offsets && s->n_seqs > 1) ? static_cast<uint32_t>(g.stride_syms - g.w8) : 0u;
    plan->lp.spu = g.spu;
    plan->offsets = offsets;
    plan->total_dev = chunk_off + nc;
    return KMC_OK;
}

} // namespace kmc
