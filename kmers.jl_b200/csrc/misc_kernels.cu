// misc_kernels.cu -- standalone fx_hash over an existing device k-mer array (src/kmer.jl:255-261)
// and the pure-store probe used to measure the write roofline of the device.
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
__device__ __forceinline__ uint64_t hash_64_64(uint64_t a)
{
    a = ~a + (a << 21);
    a = a ^ (a >> 24);
    a = a + (a << 3) + (a << 8);
    a = a ^ (a >> 14);
    a = a + (a << 2) + (a << 4);
    a = a ^ (a >> 28);
    a = a + (a << 31);
    return a;
}

template <int N>
__global__ void __launch_bounds__(256) base_hash_kernel(const uint64_t *__restrict__ kmers, uint64_t n, uint64_t h,
                                                        uint64_t *__restrict__ out)
{
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t acc = h + 0x77cfa1eef01bca90ull; // the empty tail
#pragma unroll
        for (int j = N - 1; j >= 0; --j) acc = hash_64_64(__ldg(kmers + i * N + j)) - 3 * acc;
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

cudaError_t launch_store_probe(void *dptr, uint64_t bytes, int /*sm_count*/, cudaStream_t stream)
{
    uint64_t n_vec = bytes / 32;
    if (n_vec == 0) return cudaSuccess;
    store_probe_kernel<<<static_cast<unsigned>((n_vec + 2047) / 2048), 256, 0, stream>>>(static_cast<uint64_t *>(dptr), n_vec);
    return cudaGetLastError();
}

} // namespace kmc
