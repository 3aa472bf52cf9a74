// valid_count.cu -- how many k-mers UnambiguousKmers will emit over a recoded (4-bit / ASCII) source
// (UnambiguousKmers.jl:109-148), from the valid-start BIT stream alone (1 bit per window).
//
// Used where the number is needed BEFORE the k-mers are written: kmc_count, and the host pipeline,
// which has to know where a chunk's k-mers go in the caller's buffer before it enqueues the chunk.
// The device-resident extraction does not run it: compact_kernel finds its tiles' places itself.
//
// One thread per work item (read, group slot) of the set's layout, kTileItems items per block like
// the extraction kernels; every block adds its survivors to one device counter.
#include "fourbit.h"

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

} // namespace kmc
