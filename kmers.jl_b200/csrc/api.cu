// api.cu -- the C ABI of libkmerscuda.so (include/kmerscuda.h): contexts, memory, and the host-side
// planning of every launch.  No compute happens on the host: a missing device is an error, never a
// fallback.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "extract_kernels.cuh"
#include "fourbit.h"
#include "kmc_internal.h"

using namespace kmc;

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
namespace {

int32_t fail_cuda(kmc_ctx *ctx, cudaError_t e, const char *what)
{
    if (ctx) {
        char buf[512];
        snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
        ctx->last_error = buf;
    }
    return -static_cast<int32_t>(e);
}

int32_t fail(kmc_ctx *ctx, int32_t code, const char *msg)
{
    if (ctx) ctx->last_error = msg;
    return code;
}

#define CU(call)                                                  \
    do {                                                          \
        cudaError_t e__ = (call);                                 \
        if (e__ != cudaSuccess) return fail_cuda(ctx, e__, #call); \
    } while (0)

struct Geometry {
    int n_limbs, g, nx;
    uint32_t s0;
    uint64_t head_mask;
};

// src/kmer.jl:117-137 (N = cld(2K, 64)) and :603-605 (get_mask)
Geometry geometry(int k)
{
    Geometry ge;
    ge.n_limbs = (2 * k + 63) / 64;
    ge.g = ge.n_limbs == 1 ? 4 : ge.n_limbs == 2 ? 2 : ge.n_limbs == 3 ? 4 : 1;
    ge.nx = (2 * k + 2 * ge.g - 2 + 31) / 32;
    ge.s0 = static_cast<uint32_t>(32 * ge.nx - 2 * k - 2 * (ge.g - 1));
    int used = 2 * k - 64 * (ge.n_limbs - 1); // bits used in the head limb, 2..64
    ge.head_mask = used >= 64 ? ~0ull : ((1ull << used) - 1);
    return ge;
}

ExtractLaunchFn get_launcher(const Geometry &ge, int mode, bool hash, bool ragged)
{
    switch (ge.n_limbs) {
    case 1: return get_extract_launcher_n1(ge.nx, mode, hash, ragged);
    case 2: return get_extract_launcher_n2(ge.nx, mode, hash, ragged);
    case 3: return get_extract_launcher_n3(ge.nx, mode, hash, ragged);
    case 4: return get_extract_launcher_n4(ge.nx, mode, hash, ragged);
    }
    return nullptr;
}

bool aligned32(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; }

__global__ void fill_uniform_offsets_kernel(uint64_t *out, uint64_t n_plus_1, uint64_t step)
{
    uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n_plus_1) out[i] = i * step;
}

// Totals a caller may already know (the host pipeline computes them from host-side lengths), so
// the device path does not have to synchronise to read them back.
struct KnownTotals {
    bool valid = false;
    uint64_t windows = 0;
    uint64_t items = 0;
};

struct Layout {
    bool uniform_len, uniform_off;
    uint64_t wpr;     // uniform_len only
    uint64_t total;   // windows
    uint64_t items;
    uint64_t gprm;
    const uint64_t *win_off = nullptr;  // device, ragged
    const uint64_t *item_off = nullptr; // device, ragged
};

int32_t check_common(kmc_ctx *ctx, const kmc_seqs *s, int32_t k)
{
    if (!ctx) return KMC_E_BAD_ARG;
    if (!s) return fail(ctx, KMC_E_BAD_ARG, "kmc_seqs is NULL");
    if (k < 1) return fail(ctx, KMC_E_BAD_K, "K must be at least 1");
    if (k > KMC_MAX_K) return fail(ctx, KMC_E_BAD_K, "K exceeds KMC_MAX_K (128)");
    if (s->src_bits != 2 && s->src_bits != 4) return fail(ctx, KMC_E_BAD_ARG, "src_bits must be 2 or 4");
    if (s->n_seqs > 0 && s->words == nullptr && s->n_words > 0) return fail(ctx, KMC_E_BAD_ARG, "words is NULL");
    return KMC_OK;
}

// Window / group-slot layout of a set whose descriptor arrays live on the DEVICE.
int32_t plan_layout(kmc_ctx *ctx, const kmc_seqs *s, int k, const Geometry &ge, cudaStream_t stream,
                    const KnownTotals &known, Layout *L)
{
    L->uniform_len = (s->seq_len == nullptr);
    L->uniform_off = (s->seq_word_offset == nullptr);
    const uint64_t G = static_cast<uint64_t>(ge.g);
    if (L->uniform_len) {
        L->wpr = s->uniform_len >= static_cast<uint64_t>(k) ? s->uniform_len - k + 1 : 0;
        L->total = L->wpr * s->n_seqs;
        L->gprm = (L->wpr % G == 0) ? L->wpr / G : (L->wpr + G - 1) / G + 1;
        if (s->n_seqs == 1) L->gprm = (L->wpr + G - 1) / G;
        L->items = L->gprm * s->n_seqs;
        return KMC_OK;
    }
    // ragged: cnt -> win_off ; slots -> item_off   (all in scratch)
    const uint64_t n = s->n_seqs;
    const uint64_t tmp_elems = scan_tmp_elems(n);
    const uint64_t need = ((n + 1) * 4 + tmp_elems) * sizeof(uint64_t);
    int32_t st = ensure_scratch(ctx, need);
    if (st) return st;
    uint64_t *base = static_cast<uint64_t *>(ctx->scratch);
    uint64_t *cnt = base, *win_off = base + (n + 1), *slots = base + 2 * (n + 1), *item_off = base + 3 * (n + 1),
             *tmp = base + 4 * (n + 1);
    CU(window_counts(s->seq_len, n, k, cnt, stream));
    CU(inclusive_offsets_u64(cnt, win_off, n, tmp, stream));
    CU(group_slots(win_off, n, ge.g, slots, stream));
    CU(inclusive_offsets_u64(slots, item_off, n, tmp, stream));
    L->win_off = win_off;
    L->item_off = item_off;
    L->wpr = 0;
    L->gprm = 1;
    if (known.valid) {
        L->total = known.windows;
        L->items = known.items;
    } else {
        uint64_t h[2] = {0, 0};
        CU(cudaMemcpyAsync(&h[0], win_off + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(&h[1], item_off + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        L->total = h[0];
        L->items = h[1];
    }
    return KMC_OK;
}

// Source-side fields of the kernel parameter block (2-bit LongSequence words).
ExtractParams base_params(const kmc_seqs *s, int k, const Geometry &ge, const Layout &L, uint64_t unit_bias)
{
    ExtractParams p;
    memset(&p, 0, sizeof p);
    p.w32 = reinterpret_cast<const uint32_t *>(s->words);
    p.nw32 = static_cast<int64_t>(s->n_words) * 2;
    p.unit_bits = 64;
    p.unit_bias = unit_bias;
    p.first = s->first_symbol_offset;
    p.k = k;
    p.s0 = ge.s0;
    p.head_mask = ge.head_mask;
    p.n_seqs = s->n_seqs;
    p.items = L.items;
    p.stride_units = s->uniform_stride_words;
    p.wpr = L.wpr;
    p.gprm = L.gprm;
    p.seq_unit_off = s->seq_word_offset;
    p.win_off = L.win_off;
    p.item_off = L.item_off;
    return p;
}

// The device-resident extraction; everything is enqueued on `stream`.
int32_t extract_device(kmc_ctx *ctx, const kmc_seqs *s, int32_t k, int32_t mode, uint32_t flags, const kmc_out *out,
                       kmc_result *res, cudaStream_t stream, const KnownTotals &known, uint64_t unit_bias, bool sync)
{
    int32_t st = check_common(ctx, s, k);
    if (st) return st;
    if (!out || !res) return fail(ctx, KMC_E_BAD_ARG, "kmc_out / kmc_result is NULL");
    if (mode < KMC_FW || mode > KMC_UNAMBIG) return fail(ctx, KMC_E_BAD_ARG, "unknown mode");
    res->n_written = 0;
    res->err_seq = res->err_pos = 0;
    res->err_sym = 0;
    res->kernel_ms = 0.f;
    if (s->src_bits == 4)
        return extract_device_4bit(ctx, s, k, mode, flags, out, res, stream, sync);

    const Geometry ge = geometry(k);
    const bool hash = (flags & KMC_HASH_FX) != 0;
    const bool aos = (flags & KMC_AOS) != 0;
    const bool want_index = (mode == KMC_UNAMBIG); // 2-bit source: every window + its index (UnambiguousKmers.jl:64-77)
    const int kmode = (mode == KMC_UNAMBIG) ? MODE_FW : mode;

    if (sync) CU(cudaEventRecord(ctx->ev_k0, stream));
    Layout L;
    st = plan_layout(ctx, s, k, ge, stream, known, &L);
    if (st) return st;
    res->n_written = L.total;
    if (out->seq_out_offset) {
        if (L.uniform_len) {
            uint64_t n1 = s->n_seqs + 1;
            fill_uniform_offsets_kernel<<<static_cast<unsigned>((n1 + 255) / 256), 256, 0, stream>>>(out->seq_out_offset, n1, L.wpr);
            CU(cudaGetLastError());
        } else {
            CU(cudaMemcpyAsync(out->seq_out_offset, L.win_off, (s->n_seqs + 1) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, stream));
        }
    }
    if (L.total == 0) return KMC_OK;
    if (L.total > out->capacity) return fail(ctx, KMC_E_OUT_TOO_SMALL, "output capacity smaller than the number of k-mers");
    if (!out->a) return fail(ctx, KMC_E_BAD_ARG, "out.a is NULL");
    if (kmode == MODE_FWRV && !aos && !out->b) return fail(ctx, KMC_E_BAD_ARG, "out.b is NULL (FWRV, SoA)");
    if (hash && !out->hash) return fail(ctx, KMC_E_BAD_ARG, "out.hash is NULL (KMC_HASH_FX)");
    if (want_index && !aos && !out->index) return fail(ctx, KMC_E_BAD_ARG, "out.index is NULL (UNAMBIG, SoA)");

    ExtractParams p = base_params(s, k, ge, L, unit_bias);
    p.out_a = out->a;
    p.out_b = out->b;
    p.out_hash = hash ? out->hash : nullptr;
    // AoS index lives inside out.a; the kernel keys "emit an index" on out_index != NULL
    p.out_index = want_index ? (aos ? reinterpret_cast<int64_t *>(out->a) : out->index) : nullptr;
    p.index_base = out->index_base;
    p.aos = aos ? 1u : 0u;
    p.vec_ok = aligned32(out->a) && (!p.out_b || aligned32(p.out_b)) && (!p.out_hash || aligned32(p.out_hash)) &&
               (!p.out_index || aligned32(p.out_index));

    const bool ragged = !L.uniform_len;
    ExtractLaunchFn fn = get_launcher(ge, kmode, hash, ragged);
    if (!fn) return fail(ctx, KMC_E_UNSUPPORTED, "no kernel for this K");
    CU(fn(p, ctx->sm_count, stream));
    if (sync) {
        CU(cudaEventRecord(ctx->ev_k1, stream));
        CU(cudaStreamSynchronize(stream));
        CU(cudaEventElapsedTime(&res->kernel_ms, ctx->ev_k0, ctx->ev_k1));
    }
    return KMC_OK;
}

} // namespace

namespace kmc {

int32_t ensure_scratch(kmc_ctx *ctx, uint64_t bytes)
{
    if (bytes <= ctx->scratch_bytes) return KMC_OK;
    if (ctx->scratch) {
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaFree(ctx->scratch));
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
    }
    uint64_t want = bytes + bytes / 4 + 4096;
    CU(cudaMalloc(&ctx->scratch, want));
    ctx->scratch_bytes = want;
    return KMC_OK;
}

} // namespace kmc

// ---------------------------------------------------------------------------------------------
// lifecycle
// ---------------------------------------------------------------------------------------------
extern "C" {

int32_t kmc_version(void) { return KMC_VERSION; }

int32_t kmc_device_count(int32_t *n)
{
    if (!n) return KMC_E_BAD_ARG;
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        *n = 0;
        return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? KMC_E_NO_DEVICE : -static_cast<int32_t>(e);
    }
    *n = c;
    return KMC_OK;
}

int32_t kmc_ctx_create(int32_t device, kmc_ctx **out)
{
    if (!out) return KMC_E_BAD_ARG;
    *out = nullptr;
    int32_t n = 0;
    int32_t st = kmc_device_count(&n);
    if (st) return st;
    if (n == 0 || device < 0 || device >= n) return KMC_E_NO_DEVICE;
    kmc_ctx *ctx = new (std::nothrow) kmc_ctx();
    if (!ctx) return KMC_E_BAD_ARG;
    ctx->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&ctx->pipe_streams[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_begin);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_end);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_k0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_k1);
    if (e != cudaSuccess) {
        delete ctx;
        return -static_cast<int32_t>(e);
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return KMC_OK;
}

int32_t kmc_ctx_destroy(kmc_ctx *ctx)
{
    if (!ctx) return KMC_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->scratch) cudaFree(ctx->scratch);
    for (int i = 0; i < 3; ++i) {
        if (ctx->pipe_buf[i]) cudaFree(ctx->pipe_buf[i]);
        if (ctx->pipe_streams[i]) cudaStreamDestroy(ctx->pipe_streams[i]);
    }
    if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
    if (ctx->ev_end) cudaEventDestroy(ctx->ev_end);
    if (ctx->ev_k0) cudaEventDestroy(ctx->ev_k0);
    if (ctx->ev_k1) cudaEventDestroy(ctx->ev_k1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return KMC_OK;
}

int32_t kmc_ctx_set_stream(kmc_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return KMC_E_BAD_ARG;
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return KMC_OK;
}

int32_t kmc_sync(kmc_ctx *ctx)
{
    if (!ctx) return KMC_E_BAD_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return KMC_OK;
}

const char *kmc_last_error(kmc_ctx *ctx) { return ctx ? ctx->last_error.c_str() : "kmc_ctx is NULL"; }

const char *kmc_status_string(int32_t status)
{
    switch (status) {
    case KMC_OK: return "ok";
    case KMC_E_BAD_K: return "K must be at least 1 (and at most KMC_MAX_K)";
    case KMC_E_BAD_ARG: return "bad argument";
    case KMC_E_AMBIGUOUS: return "cannot encode an uncertain symbol in a 2-bit alphabet";
    case KMC_E_OUT_TOO_SMALL: return "output buffer too small";
    case KMC_E_NO_DEVICE: return "no CUDA device";
    case KMC_E_UNSUPPORTED: return "unsupported configuration";
    }
    if (status < 0) return cudaGetErrorString(static_cast<cudaError_t>(-status));
    return "unknown status";
}

int32_t kmc_device_info(kmc_ctx *ctx, int32_t *sm_count, uint64_t *total_mem, char *name, int32_t name_len)
{
    if (!ctx) return KMC_E_BAD_ARG;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, ctx->device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (total_mem) *total_mem = prop.totalGlobalMem;
    if (name && name_len > 0) {
        strncpy(name, prop.name, static_cast<size_t>(name_len) - 1);
        name[name_len - 1] = 0;
    }
    return KMC_OK;
}

// ---------------------------------------------------------------------------------------------
// memory
// ---------------------------------------------------------------------------------------------
int32_t kmc_malloc(kmc_ctx *ctx, uint64_t bytes, void **dptr)
{
    if (!ctx || !dptr) return KMC_E_BAD_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMalloc(dptr, bytes ? bytes : 1));
    return KMC_OK;
}

int32_t kmc_free(kmc_ctx *ctx, void *dptr)
{
    if (!ctx) return KMC_E_BAD_ARG;
    if (!dptr) return KMC_OK;
    CU(cudaSetDevice(ctx->device));
    CU(cudaFree(dptr));
    return KMC_OK;
}

int32_t kmc_memset(kmc_ctx *ctx, void *dptr, int32_t value, uint64_t bytes)
{
    if (!ctx) return KMC_E_BAD_ARG;
    if (bytes == 0) return KMC_OK;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemsetAsync(dptr, value, bytes, ctx->stream));
    return KMC_OK;
}

int32_t kmc_upload(kmc_ctx *ctx, void *dptr, const void *host, uint64_t bytes)
{
    if (!ctx) return KMC_E_BAD_ARG;
    if (bytes == 0) return KMC_OK;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(dptr, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return KMC_OK;
}

int32_t kmc_download(kmc_ctx *ctx, void *host, const void *dptr, uint64_t bytes)
{
    if (!ctx) return KMC_E_BAD_ARG;
    CU(cudaSetDevice(ctx->device));
    if (bytes) CU(cudaMemcpyAsync(host, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KMC_OK;
}

int32_t kmc_host_alloc(kmc_ctx *ctx, uint64_t bytes, void **hptr)
{
    if (!ctx || !hptr) return KMC_E_BAD_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return KMC_OK;
}

int32_t kmc_host_free(kmc_ctx *ctx, void *hptr)
{
    if (!ctx) return KMC_E_BAD_ARG;
    if (!hptr) return KMC_OK;
    CU(cudaFreeHost(hptr));
    return KMC_OK;
}

int32_t kmc_host_register(kmc_ctx *ctx, void *hptr, uint64_t bytes)
{
    if (!ctx || !hptr) return KMC_E_BAD_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaHostRegister(hptr, bytes, cudaHostRegisterDefault));
    return KMC_OK;
}

int32_t kmc_host_unregister(kmc_ctx *ctx, void *hptr)
{
    if (!ctx || !hptr) return KMC_E_BAD_ARG;
    CU(cudaHostUnregister(hptr));
    return KMC_OK;
}

// ---------------------------------------------------------------------------------------------
// hot path
// ---------------------------------------------------------------------------------------------
int32_t kmc_count(kmc_ctx *ctx, const kmc_seqs *seqs, int32_t k, int32_t mode, uint64_t *n_out)
{
    int32_t st = check_common(ctx, seqs, k);
    if (st) return st;
    if (!n_out) return fail(ctx, KMC_E_BAD_ARG, "n_out is NULL");
    if (mode < KMC_FW || mode > KMC_UNAMBIG) return fail(ctx, KMC_E_BAD_ARG, "unknown mode");
    CU(cudaSetDevice(ctx->device));
    if (mode == KMC_UNAMBIG && seqs->src_bits == 4) return count_unambiguous_4bit(ctx, seqs, k, n_out, ctx->stream);
    const Geometry ge = geometry(k);
    Layout L;
    st = plan_layout(ctx, seqs, k, ge, ctx->stream, KnownTotals(), &L);
    if (st) return st;
    *n_out = L.total;
    return KMC_OK;
}

int32_t kmc_extract(kmc_ctx *ctx, const kmc_seqs *seqs, int32_t k, int32_t mode, uint32_t flags, const kmc_out *out,
                    kmc_result *result)
{
    if (!ctx) return KMC_E_BAD_ARG;
    CU(cudaSetDevice(ctx->device));
    return extract_device(ctx, seqs, k, mode, flags, out, result, ctx->stream, KnownTotals(), 0,
                          (flags & KMC_NO_SYNC) == 0);
}

int32_t kmc_fx_hash(kmc_ctx *ctx, const uint64_t *kmers, uint64_t n, int32_t n_limbs, uint64_t h0, uint64_t *out)
{
    if (!ctx) return KMC_E_BAD_ARG;
    if (n_limbs < 0 || n_limbs > 4) return fail(ctx, KMC_E_BAD_ARG, "n_limbs must be in 0..4");
    if (n && (!out || (n_limbs && !kmers))) return fail(ctx, KMC_E_BAD_ARG, "NULL buffer");
    CU(cudaSetDevice(ctx->device));
    CU(launch_fx_hash(kmers, n, n_limbs, h0, out, ctx->sm_count, ctx->stream));
    return KMC_OK;
}

int32_t kmc_bucket_count(kmc_ctx *ctx, const kmc_seqs *seqs, int32_t k, int32_t bucket_bits, uint32_t *table,
                         kmc_result *result)
{
    int32_t st = check_common(ctx, seqs, k);
    if (st) return st;
    if (!table || !result) return fail(ctx, KMC_E_BAD_ARG, "table / result is NULL");
    if (bucket_bits < 1 || bucket_bits > 32) return fail(ctx, KMC_E_BAD_ARG, "bucket_bits must be in 1..32");
    if (seqs->src_bits != 2) return fail(ctx, KMC_E_UNSUPPORTED, "bucket count needs a 2-bit source");
    CU(cudaSetDevice(ctx->device));
    result->n_written = 0;
    result->err_seq = result->err_pos = 0;
    result->err_sym = 0;
    result->kernel_ms = 0.f;
    const Geometry ge = geometry(k);
    cudaStream_t stream = ctx->stream;
    CU(cudaEventRecord(ctx->ev_k0, stream));
    Layout L;
    st = plan_layout(ctx, seqs, k, ge, stream, KnownTotals(), &L);
    if (st) return st;
    result->n_written = L.total;
    if (L.total == 0) return KMC_OK;
    ExtractParams p = base_params(seqs, k, ge, L, 0);
    p.bucket_table = table;
    p.bucket_shift = static_cast<uint32_t>(64 - bucket_bits);
    ExtractLaunchFn fn = get_launcher(ge, MODE_BUCKETS, true, !L.uniform_len);
    if (!fn) return fail(ctx, KMC_E_UNSUPPORTED, "no kernel for this K");
    CU(fn(p, ctx->sm_count, stream));
    CU(cudaEventRecord(ctx->ev_k1, stream));
    CU(cudaStreamSynchronize(stream));
    CU(cudaEventElapsedTime(&result->kernel_ms, ctx->ev_k0, ctx->ev_k1));
    return KMC_OK;
}

int32_t kmc_timer_begin(kmc_ctx *ctx)
{
    if (!ctx) return KMC_E_BAD_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->ev_begin, ctx->stream));
    return KMC_OK;
}

int32_t kmc_timer_end(kmc_ctx *ctx, float *ms)
{
    if (!ctx || !ms) return KMC_E_BAD_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->ev_end, ctx->stream));
    CU(cudaEventSynchronize(ctx->ev_end));
    CU(cudaEventElapsedTime(ms, ctx->ev_begin, ctx->ev_end));
    return KMC_OK;
}

int32_t kmc_store_probe(kmc_ctx *ctx, void *dptr, uint64_t bytes)
{
    if (!ctx || !dptr) return KMC_E_BAD_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(launch_store_probe(dptr, bytes, ctx->sm_count, ctx->stream));
    return KMC_OK;
}

// ---------------------------------------------------------------------------------------------
// host-buffer entry point: upload -> extract -> download in pipelined chunks
// ---------------------------------------------------------------------------------------------
namespace {

struct Slot {
    cudaStream_t stream;
    uint64_t *words = nullptr, *off = nullptr, *len = nullptr;
    uint64_t *a = nullptr, *b = nullptr, *hash = nullptr;
    int64_t *index = nullptr;
};

uint64_t round_up(uint64_t x, uint64_t m) { return (x + m - 1) / m * m; }

} // namespace

int32_t kmc_extract_host(kmc_ctx *ctx, const kmc_seqs *hs, int32_t k, int32_t mode, uint32_t flags,
                         const kmc_out *ho, kmc_result *result)
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
    if (hs->src_bits == 4) return extract_host_4bit(ctx, hs, k, mode, flags, ho, result);

    const Geometry ge = geometry(k);
    const uint64_t N = static_cast<uint64_t>(ge.n_limbs), G = static_cast<uint64_t>(ge.g);
    const bool hash = (flags & KMC_HASH_FX) != 0;
    const bool aos = (flags & KMC_AOS) != 0;
    const bool want_index = (mode == KMC_UNAMBIG);
    const bool two = (mode == KMC_FWRV);
    const bool ragged_len = hs->seq_len != nullptr;
    const bool ragged_off = hs->seq_word_offset != nullptr;
    const uint64_t K = static_cast<uint64_t>(k);
    // u64 per element of out.a
    const uint64_t a_elems = (aos && two) ? 2 * N : (aos && want_index) ? N + 1 : N;

    // ---- chunk plan (host side; lengths are host arrays here) --------------------------------
    struct Chunk {
        uint64_t seq0, nseq;       // reads [seq0, seq0+nseq)   (n_seqs > 1)
        uint64_t word0, nwords;    // words to upload
        uint64_t out0, nout;       // flat output range
        uint64_t items;            // group slots (ragged)
        uint32_t first;            // first_symbol_offset of the chunk view
        int64_t index_base;
        uint64_t len;              // single-sequence chunks: symbols in view
    };
    std::vector<Chunk> chunks;
    const uint64_t target = 4ull << 20; // windows per chunk (64-128 MB of output at 16-32 B / window)
    uint64_t total = 0;
    const bool single = (hs->n_seqs == 1);
    if (single) {
        const uint64_t len = ragged_len ? hs->seq_len[0] : hs->uniform_len;
        const uint64_t w0 = ragged_off ? hs->seq_word_offset[0] : 0;
        const uint64_t wcount = len >= K ? len - K + 1 : 0;
        total = wcount;
        for (uint64_t s = 0; s < wcount; s += target) {
            Chunk c{};
            c.nout = std::min(target, wcount - s);
            c.out0 = s;
            const uint64_t sym0 = hs->first_symbol_offset + s;      // absolute first symbol
            const uint64_t sym1 = sym0 + c.nout + K - 1;            // one past last symbol
            c.word0 = w0 + sym0 / 32;
            c.nwords = (sym1 + 31) / 32 - sym0 / 32;
            c.first = static_cast<uint32_t>(sym0 % 32);
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
            Chunk c{};
            c.seq0 = r;
            c.out0 = total;
            c.first = hs->first_symbol_offset;
            if (!ragged_len) {
                uint64_t take = wpr_u ? std::max<uint64_t>(1, target / wpr_u) : n;
                take = std::min(take, n - r);
                c.nseq = take;
                c.nout = take * wpr_u;
                r += take;
            } else {
                uint64_t acc = 0, items = 0;
                while (r < n && (acc < target || c.nseq == 0)) {
                    uint64_t len = hs->seq_len[r];
                    uint64_t wc = len >= K ? len - K + 1 : 0;
                    acc += wc;
                    ++r;
                    ++c.nseq;
                }
                // group slots are relative to the chunk's own flat origin (0), recomputed on device;
                // the host only needs the total to skip the read-back
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
                uint64_t end = hs->seq_word_offset[last] + (hs->first_symbol_offset + last_len + 31) / 32;
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
    if (!ho->a) return fail(ctx, KMC_E_BAD_ARG, "out.a is NULL");
    if (two && !aos && !ho->b) return fail(ctx, KMC_E_BAD_ARG, "out.b is NULL (FWRV, SoA)");
    if (hash && !ho->hash) return fail(ctx, KMC_E_BAD_ARG, "out.hash is NULL (KMC_HASH_FX)");
    if (want_index && !aos && !ho->index) return fail(ctx, KMC_E_BAD_ARG, "out.index is NULL (UNAMBIG, SoA)");

    // ---- slot buffers --------------------------------------------------------------------------
    uint64_t max_words = 0, max_out = 0, max_seq = 0;
    for (const Chunk &c : chunks) {
        max_words = std::max(max_words, c.nwords);
        max_out = std::max(max_out, c.nout);
        max_seq = std::max(max_seq, c.nseq);
    }
    const uint64_t b_words = round_up((max_words + 4) * 8, 256);
    const uint64_t b_meta = round_up((max_seq + 1) * 8, 256);
    const uint64_t b_a = round_up(max_out * a_elems * 8, 256);
    const uint64_t b_b = (two && !aos) ? round_up(max_out * N * 8, 256) : 0;
    const uint64_t b_h = hash ? round_up(max_out * 8, 256) : 0;
    const uint64_t b_i = (want_index && !aos) ? round_up(max_out * 8, 256) : 0;
    const uint64_t per_slot = b_words + 2 * b_meta + b_a + b_b + b_h + b_i;
    const int n_slots = chunks.size() >= 3 ? 3 : static_cast<int>(chunks.size());
    Slot slots[3];
    for (int i = 0; i < n_slots; ++i) {
        if (ctx->pipe_bytes[i] < per_slot) {
            if (ctx->pipe_buf[i]) CU(cudaFree(ctx->pipe_buf[i]));
            ctx->pipe_buf[i] = nullptr;
            ctx->pipe_bytes[i] = 0;
            CU(cudaMalloc(&ctx->pipe_buf[i], per_slot));
            ctx->pipe_bytes[i] = per_slot;
        }
        char *base = static_cast<char *>(ctx->pipe_buf[i]);
        Slot &sl = slots[i];
        sl.stream = ctx->pipe_streams[i];
        sl.words = reinterpret_cast<uint64_t *>(base); base += b_words;
        sl.off = reinterpret_cast<uint64_t *>(base); base += b_meta;
        sl.len = reinterpret_cast<uint64_t *>(base); base += b_meta;
        sl.a = reinterpret_cast<uint64_t *>(base); base += b_a;
        sl.b = b_b ? reinterpret_cast<uint64_t *>(base) : nullptr; base += b_b;
        sl.hash = b_h ? reinterpret_cast<uint64_t *>(base) : nullptr; base += b_h;
        sl.index = b_i ? reinterpret_cast<int64_t *>(base) : nullptr;
    }
    // the ragged layout scans use ctx->scratch, one region per slot would be needed for full overlap;
    // keep it simple and safe: ragged chunks serialise their planning on the slot stream through a
    // private scratch carve per slot.
    uint64_t scratch_per_slot = 0;
    if (ragged_len && !single) {
        scratch_per_slot = round_up(((max_seq + 1) * 4 + scan_tmp_elems(max_seq)) * 8, 256);
        st = ensure_scratch(ctx, scratch_per_slot * n_slots);
        if (st) return st;
    }

    // ---- pipeline ------------------------------------------------------------------------------
    for (size_t ci = 0; ci < chunks.size(); ++ci) {
        const Chunk &c = chunks[ci];
        Slot &sl = slots[ci % n_slots];
        cudaStream_t sm = sl.stream;
        if (c.nwords) CU(cudaMemcpyAsync(sl.words, hs->words + c.word0, c.nwords * 8, cudaMemcpyHostToDevice, sm));
        kmc_seqs ds = *hs;
        ds.words = sl.words;
        ds.n_words = c.nwords;
        ds.n_seqs = c.nseq;
        ds.first_symbol_offset = c.first;
        ds.seq_word_offset = nullptr;
        ds.seq_len = nullptr;
        uint64_t bias = 0;
        KnownTotals known;
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
        kmc_out dout{};
        dout.a = sl.a;
        dout.b = sl.b;
        dout.hash = sl.hash;
        dout.index = sl.index;
        dout.capacity = c.nout;
        dout.index_base = c.index_base + ho->index_base;
        kmc_result r{};
        // per-slot scratch window for the ragged planner
        void *saved = ctx->scratch;
        uint64_t saved_bytes = ctx->scratch_bytes;
        if (scratch_per_slot) {
            ctx->scratch = static_cast<char *>(saved) + scratch_per_slot * (ci % n_slots);
            ctx->scratch_bytes = scratch_per_slot;
        }
        st = extract_device(ctx, &ds, k, mode, flags, &dout, &r, sm, known, bias, false);
        ctx->scratch = saved;
        ctx->scratch_bytes = saved_bytes;
        if (st) return st;
        if (c.nout) {
            CU(cudaMemcpyAsync(ho->a + c.out0 * a_elems, sl.a, c.nout * a_elems * 8, cudaMemcpyDeviceToHost, sm));
            if (sl.b) CU(cudaMemcpyAsync(ho->b + c.out0 * N, sl.b, c.nout * N * 8, cudaMemcpyDeviceToHost, sm));
            if (sl.hash) CU(cudaMemcpyAsync(ho->hash + c.out0, sl.hash, c.nout * 8, cudaMemcpyDeviceToHost, sm));
            if (sl.index) CU(cudaMemcpyAsync(ho->index + c.out0, sl.index, c.nout * 8, cudaMemcpyDeviceToHost, sm));
        }
    }
    for (int i = 0; i < n_slots; ++i) CU(cudaStreamSynchronize(slots[i].stream));
    return KMC_OK;
}

} // extern "C"
