// misc_kernels.cu -- standalone fx_hash over an existing device k-mer array (src/kmer.jl:255-261)
// and the pure-store probe used to measure the write roofline of the device.
#include <cstdlib>
#include "kmc_internal.h"
#include "kmer_core.cuh"

namespace kmc {

template <int N>
__global__ void __launch_bounds__(256) fx_hash_kernel(const uint64_t *__restrict__ kmers, uint64_t n, uint64_t h0,
                                                      uint64_t *__restrict__ out)
{
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t d[N];
#pragma unroll
        for (int j = 0; j < N; ++j) d[j] = __ldg(kmers + i * N + j);
        st_u64(out + i, fx_hash<N>(d, h0));
    }
}

__global__ void __launch_bounds__(256) fx_hash_empty_kernel(uint64_t n, uint64_t h0, uint64_t *__restrict__ out)
{
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = h0;
}

cudaError_t launch_fx_hash(const uint64_t *kmers, uint64_t n, int n_limbs, uint64_t h0, uint64_t *out, int sm_count,
                           cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    uint64_t want = (n + 255) / 256;
    unsigned grid = static_cast<unsigned>(want < static_cast<uint64_t>(sm_count) * 16 ? want : static_cast<uint64_t>(sm_count) * 16);
    switch (n_limbs) {
    case 0: fx_hash_empty_kernel<<<grid, 256, 0, stream>>>(n, h0, out); break; // empty k-mer hashes to h0 (0 by default)
    case 1: fx_hash_kernel<1><<<grid, 256, 0, stream>>>(kmers, n, h0, out); break;
    case 2: fx_hash_kernel<2><<<grid, 256, 0, stream>>>(kmers, n, h0, out); break;
    case 3: fx_hash_kernel<3><<<grid, 256, 0, stream>>>(kmers, n, h0, out); break;
    case 4: fx_hash_kernel<4><<<grid, 256, 0, stream>>>(kmers, n, h0, out); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// Base.hash(x::Kmer, h) = hash(x.data, h ⊻ K) (src/kmer.jl:206) with Julia 1.10 / 1.11's Base:
//   hash(::Tuple{}, h) = h + 0x77cfa1eef01bca90 ; hash(t::Tuple, h) = hash(t[1], hash(tail(t), h))   (tuple.jl)
//   hash(x::UInt64, h) = hash_64_64(x) - 3h                                                          (hashing.jl)
// Pinned by the reference's documented value hash(mer"UGCUGUAC"r) == 0xe5057d38c8907b22
// (docs/src/hashing.md:18-20).  Julia >= 1.12 hashes integers differently; the Julia binding checks
// that value at load time before it trusts this kernel.
// (hash_64_64, base_hash_seed, base_hash_fold: kmer_core.cuh)

template <int N>
__global__ void __launch_bounds__(256) base_hash_kernel(const uint64_t *__restrict__ kmers, uint64_t n, uint64_t h,
                                                        uint64_t *__restrict__ out)
{
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t acc = base_hash_seed(h); // the empty tail
#pragma unroll
        for (int j = N - 1; j >= 0; --j) acc = base_hash_fold(__ldg(kmers + i * N + j), acc);
        st_u64(out + i, acc);
    }
}

cudaError_t launch_base_hash(const uint64_t *kmers, uint64_t n, int n_limbs, uint64_t h, uint64_t *out, int sm_count,
                             cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    uint64_t want = (n + 255) / 256;
    unsigned grid = static_cast<unsigned>(want < static_cast<uint64_t>(sm_count) * 16 ? want : static_cast<uint64_t>(sm_count) * 16);
    switch (n_limbs) {
    case 0: base_hash_kernel<0><<<grid, 256, 0, stream>>>(kmers, n, h, out); break;
    case 1: base_hash_kernel<1><<<grid, 256, 0, stream>>>(kmers, n, h, out); break;
    case 2: base_hash_kernel<2><<<grid, 256, 0, stream>>>(kmers, n, h, out); break;
    case 3: base_hash_kernel<3><<<grid, 256, 0, stream>>>(kmers, n, h, out); break;
    case 4: base_hash_kernel<4><<<grid, 256, 0, stream>>>(kmers, n, h, out); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// XOR / wrapping-sum fingerprint of a u64 stream (kmc_digest): 128-bit loads, warp shuffle +
// shared-memory reduction, one pair of atomics per block.
__global__ void __launch_bounds__(256) digest_kernel(const uint64_t *__restrict__ p, uint64_t n,
                                                     unsigned long long *__restrict__ acc)
{
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    uint64_t x = 0, s = 0;
    const uint64_t head = (16 - (reinterpret_cast<uintptr_t>(p) & 15)) / 8 % 2; // words before 16-byte alignment
    const uint64_t tid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (tid < head && tid < n) {
        x ^= p[tid];
        s += p[tid];
    }
    const uint64_t n2 = n > head ? (n - head) / 2 : 0;
    const ulonglong2 *p2 = reinterpret_cast<const ulonglong2 *>(p + head);
    for (uint64_t i = tid; i < n2; i += stride) {
        const ulonglong2 v = __ldg(p2 + i);
        x ^= v.x ^ v.y;
        s += v.x + v.y;
    }
    const uint64_t tail = head + 2 * n2;
    if (tid == 0 && tail < n) {
        x ^= p[tail];
        s += p[tail];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        x ^= __shfl_xor_sync(0xffffffffu, x, d);
        s += __shfl_xor_sync(0xffffffffu, s, d);
    }
    // one pair of atomics per BLOCK: same-address atomics serialise in L2
    __shared__ uint64_t sx[8], ss[8];
    if ((threadIdx.x & 31) == 0) {
        sx[threadIdx.x >> 5] = x;
        ss[threadIdx.x >> 5] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < 8; ++w) {
            x ^= sx[w];
            s += ss[w];
        }
        atomicXor(acc, static_cast<unsigned long long>(x));
        atomicAdd(acc + 1, static_cast<unsigned long long>(s));
    }
}

cudaError_t launch_digest(const uint64_t *p, uint64_t n, uint64_t *acc, int sm_count, cudaStream_t stream, bool zero)
{
    cudaError_t e = zero ? cudaMemsetAsync(acc, 0, 16, stream) : cudaSuccess;
    if (e != cudaSuccess || n == 0) return e;
    const uint64_t want = (n / 2 + 255) / 256 + 1;
    const unsigned grid = static_cast<unsigned>(want < static_cast<uint64_t>(sm_count) * 8 ? want : static_cast<uint64_t>(sm_count) * 8);
    digest_kernel<<<grid, 256, 0, stream>>>(p, n, reinterpret_cast<unsigned long long *>(acc));
    return cudaGetLastError();
}

// Every thread writes 32 bytes per step with one 256-bit store; nothing is read.  Same launch shape
// as the extraction kernels (one tile of 8 x 256 stores per block, hardware-scheduled), so it is the
// write ceiling for exactly that store pattern.
__global__ void __launch_bounds__(256) store_probe_kernel(uint64_t *__restrict__ p, uint64_t n_vec)
{
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * 2048 + threadIdx.x;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const uint64_t i = base + static_cast<uint64_t>(it) * 256;
        if (i < n_vec) st_v4(p + 4 * i, i, i + 1, i + 2, i + 3);
    }
}

// The extraction kernels' own pattern: a thread owns 64 contiguous bytes of each of STREAMS output streams and
// writes them as two 256-bit stores, so one warp-wide store instruction covers 2 KB with 32-byte holes that the
// next instruction fills.  (KMC_STORE_PROBE_PATTERN = 1 or 2 selects it; experiment only.)
template <int STREAMS>
__global__ void __launch_bounds__(256) store_probe_items_kernel(uint64_t *__restrict__ p, uint64_t n_items)
{
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * 2048 + threadIdx.x;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const uint64_t i = base + static_cast<uint64_t>(it) * 256;
        if (i < n_items) {
#pragma unroll
            for (int s = 0; s < STREAMS; ++s) {
                uint64_t *q = p + static_cast<uint64_t>(s) * n_items * 8 + 8 * i;
                st_v4(q, i, i + 1, i + 2, i + 3);
                st_v4(q + 4, i + 4, i + 5, i + 6, i + 7);
            }
        }
    }
}

// Pattern 3: pattern 2 plus the extraction kernels' input side -- every item first loads four 32-bit words from a
// packed source (2.67 bytes per item, neighbours overlap, so L1 absorbs most of it) and its stores depend on them.
// Pattern 4: the same bytes, but the block's slice of the source is loaded once, coalesced, into shared memory.
__device__ __forceinline__ uint32_t ldg_l2_256(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.global.nc.L2::256B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

template <bool STAGED, bool WIDE = false>
__global__ void __launch_bounds__(256) store_probe_rw_kernel(uint64_t *__restrict__ p, uint64_t n_items,
                                                             const uint32_t *__restrict__ src, uint64_t n_src)
{
    __shared__ uint32_t s_src[2048 * 8 / 12 + 16];
    const uint64_t tile = static_cast<uint64_t>(blockIdx.x) * 2048;
    const uint64_t base = tile + threadIdx.x;
    if (STAGED) {
        const uint64_t w0 = tile * 8 / 12;
        for (uint32_t t = threadIdx.x; t < 2048 * 8 / 12 + 16; t += 256) s_src[t] = w0 + t < n_src ? __ldg(src + w0 + t) : 0u;
        __syncthreads();
    }
#pragma unroll 1
    for (int it = 0; it < 8; ++it) {
        const uint64_t i = base + static_cast<uint64_t>(it) * 256;
        if (i < n_items) {
            uint64_t w = i * 8 / 12; // 32-bit word of the source this item starts in (2.67 bytes per item)
            uint32_t a[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                a[k] = STAGED ? s_src[w - tile * 8 / 12 + k] : (w + k < n_src ? (WIDE ? ldg_l2_256(src + w + k) : __ldg(src + w + k)) : 0u);
            const uint64_t x = (static_cast<uint64_t>(a[1]) << 32 | a[0]) ^ i, y = (static_cast<uint64_t>(a[3]) << 32 | a[2]) + i;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                uint64_t *q = p + static_cast<uint64_t>(s) * n_items * 8 + 8 * i;
                st_v4(q, x, y, x + 2, y + 3);
                st_v4(q + 4, x + 4, y + 5, x + 6, y + 7);
            }
        }
    }
}

// Pattern 5: pattern 3 in chunks, each chunk's slice of the source pulled into L2 (evict_last) by a prefetch kernel
// first, so that DRAM sees bursts of reads between long runs of writes instead of a 2 % trickle of reads among them.
__global__ void __launch_bounds__(256) l2_prefetch_kernel(const char *__restrict__ p, uint64_t bytes)
{
    for (uint64_t o = (static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x) * 128; o < bytes;
         o += static_cast<uint64_t>(gridDim.x) * 256 * 128)
        asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p + o));
}

__device__ __forceinline__ void st_v4_evict_first(uint64_t *p, uint64_t a, uint64_t b, uint64_t c, uint64_t d, uint64_t pol)
{
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u64 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d),
                 "l"(pol)
                 : "memory");
}

template <bool EVF>
__global__ void __launch_bounds__(256) store_probe_rw_chunk_kernel(uint64_t *__restrict__ p, uint64_t n_items, uint64_t item0,
                                                                   uint64_t item1, const uint32_t *__restrict__ src, uint64_t n_src)
{
    const uint64_t base = item0 + static_cast<uint64_t>(blockIdx.x) * 2048 + threadIdx.x;
    uint64_t pol = 0;
    if (EVF) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
#pragma unroll 1
    for (int it = 0; it < 8; ++it) {
        const uint64_t i = base + static_cast<uint64_t>(it) * 256;
        if (i < item1) {
            uint64_t w = i * 8 / 12;
            uint32_t a[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) a[k] = w + k < n_src ? __ldg(src + w + k) : 0u;
            const uint64_t x = (static_cast<uint64_t>(a[1]) << 32 | a[0]) ^ i, y = (static_cast<uint64_t>(a[3]) << 32 | a[2]) + i;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                uint64_t *q = p + static_cast<uint64_t>(s) * n_items * 8 + 8 * i;
                if (EVF) {
                    st_v4_evict_first(q, x, y, x + 2, y + 3, pol);
                    st_v4_evict_first(q + 4, x + 4, y + 5, x + 6, y + 7, pol);
                } else {
                    st_v4(q, x, y, x + 2, y + 3);
                    st_v4(q + 4, x + 4, y + 5, x + 6, y + 7);
                }
            }
        }
    }
}

// Pattern 7: pattern 3 as ONE launch; the tiles are cut into chunks of T tiles and the first P tiles of a chunk pull
// the NEXT chunk's slice of the source into L2 (evict_last) -- a short burst of reads once per chunk.
__global__ void __launch_bounds__(256) store_probe_rw_burst_kernel(uint64_t *__restrict__ p, uint64_t n_items,
                                                                   const uint32_t *__restrict__ src, uint64_t n_src, uint32_t T,
                                                                   uint32_t P, uint32_t lead)
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    {
        // the tiles [T - lead, T - lead + P) of chunk c pull chunk c + 1; the first P tiles of the grid pull chunk 0
        const uint32_t c = blockIdx.x / T, k = blockIdx.x - c * T;
        const uint32_t k0 = T - lead;
        const bool first = blockIdx.x < P;
        if (first || (k >= k0 && k < k0 + P)) {
            const uint32_t cc = first ? 0 : c + 1, kk = first ? blockIdx.x : k - k0;
            const uint64_t i0 = static_cast<uint64_t>(cc) * T * 2048, i1 = i0 + static_cast<uint64_t>(T) * 2048;
            uint64_t b0 = (i0 * 8 / 12 * 4) & ~127ull, b1 = (i1 * 8 / 12 + 4) * 4;
            if (b1 > n_src * 4) b1 = n_src * 4;
            const char *sb = reinterpret_cast<const char *>(src);
            for (uint64_t o = b0 + (static_cast<uint64_t>(kk) * 256 + threadIdx.x) * 128; o < b1; o += static_cast<uint64_t>(P) * 256 * 128)
                asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(sb + o));
        }
    }
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * 2048 + threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < 8; ++it) {
        const uint64_t i = base + static_cast<uint64_t>(it) * 256;
        if (i < n_items) {
            uint64_t w = i * 8 / 12;
            uint32_t a[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) a[k] = w + k < n_src ? __ldg(src + w + k) : 0u;
            const uint64_t x = (static_cast<uint64_t>(a[1]) << 32 | a[0]) ^ i, y = (static_cast<uint64_t>(a[3]) << 32 | a[2]) + i;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                uint64_t *q = p + static_cast<uint64_t>(s) * n_items * 8 + 8 * i;
                st_v4_evict_first(q, x, y, x + 2, y + 3, pol);
                st_v4_evict_first(q + 4, x + 4, y + 5, x + 6, y + 7, pol);
            }
        }
    }
}

cudaError_t launch_store_probe(void *dptr, uint64_t bytes, int sm_count, cudaStream_t stream)
{
    static const int pattern = [] {
        const char *e = getenv("KMC_STORE_PROBE_PATTERN");
        return e ? atoi(e) : 0;
    }();
    uint64_t n_vec = bytes / 32;
    if (n_vec == 0) return cudaSuccess;
    if (pattern == 1 || pattern == 2) {
        const uint64_t n_items = bytes / (64ull * pattern);
        if (n_items == 0) return cudaSuccess;
        const unsigned grid = static_cast<unsigned>((n_items + 2047) / 2048);
        if (pattern == 1)
            store_probe_items_kernel<1><<<grid, 256, 0, stream>>>(static_cast<uint64_t *>(dptr), n_items);
        else
            store_probe_items_kernel<2><<<grid, 256, 0, stream>>>(static_cast<uint64_t *>(dptr), n_items);
        return cudaGetLastError();
    }
    if (pattern == 7) {
        static const int T = [] { const char *e = getenv("KMC_PROBE_T"); return e ? atoi(e) : 2048; }();
        static const int P = [] { const char *e = getenv("KMC_PROBE_P"); return e ? atoi(e) : 32; }();
        static const int lead = [] { const char *e = getenv("KMC_PROBE_LEAD"); return e ? atoi(e) : 600; }();
        const uint64_t n_items = bytes / (128 + 3);
        if (n_items == 0) return cudaSuccess;
        const char *src = static_cast<char *>(dptr) + n_items * 128;
        const uint64_t n_src = (bytes - n_items * 128) / 4;
        store_probe_rw_burst_kernel<<<static_cast<unsigned>((n_items + 2047) / 2048), 256, 0, stream>>>(
            static_cast<uint64_t *>(dptr), n_items, reinterpret_cast<const uint32_t *>(src), n_src, T, P, lead < T ? lead : T);
        return cudaGetLastError();
    }
    if (pattern == 5 || pattern == 6) {
        static const int chunks = [] {
            const char *e = getenv("KMC_PROBE_CHUNKS");
            const int v = e ? atoi(e) : 8;
            return v < 1 ? 1 : v;
        }();
        const uint64_t n_items = bytes / (128 + 3);
        if (n_items == 0) return cudaSuccess;
        const char *src = static_cast<char *>(dptr) + n_items * 128;
        const uint64_t n_src = (bytes - n_items * 128) / 4;
        const uint64_t per = ((n_items + chunks - 1) / chunks + 2047) / 2048 * 2048;
        for (uint64_t i0 = 0; i0 < n_items; i0 += per) {
            const uint64_t i1 = i0 + per < n_items ? i0 + per : n_items;
            const uint64_t b0 = (i0 * 8 / 12 * 4) & ~127ull, b1 = (i1 * 8 / 12 + 4) * 4;
            l2_prefetch_kernel<<<sm_count * 4, 256, 0, stream>>>(src + b0, (b1 < n_src * 4 ? b1 : n_src * 4) - b0);
            const unsigned grid = static_cast<unsigned>((i1 - i0 + 2047) / 2048);
            if (pattern == 5)
                store_probe_rw_chunk_kernel<false><<<grid, 256, 0, stream>>>(static_cast<uint64_t *>(dptr), n_items, i0, i1,
                                                                             reinterpret_cast<const uint32_t *>(src), n_src);
            else
                store_probe_rw_chunk_kernel<true><<<grid, 256, 0, stream>>>(static_cast<uint64_t *>(dptr), n_items, i0, i1,
                                                                            reinterpret_cast<const uint32_t *>(src), n_src);
        }
        return cudaGetLastError();
    }
    if (pattern == 3 || pattern == 4 || pattern == 8) {
        // the last 1/48 of the buffer is the source (2.67 of every 128 + 2.67 bytes), the rest the two streams
        const uint64_t n_items = bytes / (128 + 3);
        if (n_items == 0) return cudaSuccess;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(static_cast<char *>(dptr) + n_items * 128);
        const uint64_t n_src = (bytes - n_items * 128) / 4;
        const unsigned grid = static_cast<unsigned>((n_items + 2047) / 2048);
        if (pattern == 3)
            store_probe_rw_kernel<false><<<grid, 256, 0, stream>>>(static_cast<uint64_t *>(dptr), n_items, src, n_src);
        else if (pattern == 8)
            store_probe_rw_kernel<false, true><<<grid, 256, 0, stream>>>(static_cast<uint64_t *>(dptr), n_items, src, n_src);
        else
            store_probe_rw_kernel<true><<<grid, 256, 0, stream>>>(static_cast<uint64_t *>(dptr), n_items, src, n_src);
        return cudaGetLastError();
    }
    store_probe_kernel<<<static_cast<unsigned>((n_vec + 2047) / 2048), 256, 0, stream>>>(static_cast<uint64_t *>(dptr), n_vec);
    return cudaGetLastError();
}

} // namespace kmc
