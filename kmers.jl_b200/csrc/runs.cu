// runs.cu -- UnambiguousKmers over a 4-bit source (UnambiguousKmers.jl:134-148) as a RUN LIST.
//
// The iterator emits every window whose K symbols are all certain, in order.  Inside one read the
// surviving windows form maximal runs of consecutive windows; the concatenated output of the set
// is the concatenation of the runs.  A run is exactly what the 2-bit extraction kernels call a
// sequence: a start symbol in the stream, a number of windows, and (new) the index of its first
// window inside its read.  So instead of compacting k-mers (a scatter of 8-byte stores), the run
// list is built from the valid-start BIT stream (1 bit per window) and the ordinary ragged
// extraction kernel does all the heavy lifting with its aligned 256-bit stores.
//
//   mark_runs_kernel   per tile of the G=32 decomposition of the read set: survivors and run starts
//   (two small scans over the per-tile counts)
//   emit_runs_kernel   per tile again: block scan, then one descriptor per run start:
//                      run_sym[i]   absolute symbol index (recoded stream) of the run's first window
//                      run_woff[i]  flat output index of the run's first k-mer (= the "win_off" of the run set)
//                      run_ibase[i] 0-based window index of the run's first window inside its read
#include "fourbit.h"

namespace kmc {

namespace {

constexpr int kRunG = 32;

// Validity of the windows of one G=32 item, relative to its first in-read slot jlo:
//   m      bit t = window (jlo + t) has no uncertain symbol
//   starts bit t = that window starts a run (valid and its predecessor in the SAME read is not)
//   sym    absolute symbol index of slot jlo
template <bool RAGGED>
__device__ __forceinline__ void item_masks(const ExtractParams &p, const TileCursor<RAGGED, kRunG> &cur, uint32_t &m,
                                           uint32_t &starts, int64_t &sym)
{
    m = starts = 0;
    sym = 0;
    if (cur.jhi <= cur.jlo) return;
    sym = (cur.bit(p) >> 1) + cur.jlo;
    const int width = cur.jhi - cur.jlo; // 1..32
    const uint32_t w0 = __ldg(p.vstart + (sym >> 5)), w1 = __ldg(p.vstart + (sym >> 5) + 1);
    const uint32_t sh = static_cast<uint32_t>(sym) & 31u;
    const uint32_t bits = __funnelshift_r(w0, w1, sh);
    m = bits & (width == 32 ? 0xffffffffu : ((1u << width) - 1u));
    uint32_t prev = 0;
    if (cur.wbase + cur.jlo > 0) // the window before slot jlo belongs to the same read
        prev = sh ? ((w0 >> (sh - 1)) & 1u) : (__ldg(p.vstart + (sym >> 5) - 1) >> 31);
    starts = m & ~((m << 1) | prev);
}

template <bool RAGGED>
__global__ void __launch_bounds__(kBlockThreads) mark_runs_kernel(const ExtractParams p, uint64_t *__restrict__ tile_valid,
                                                                  uint64_t *__restrict__ tile_runs)
{
    __shared__ TileShared<RAGGED> sh;
    __shared__ uint32_t s_v[kBlockThreads / 32], s_r[kBlockThreads / 32];
    const uint64_t tile_base = static_cast<uint64_t>(blockIdx.x) * kTileItems;
    TileCursor<RAGGED, kRunG> cur;
    cur.init(p, tile_base, sh, threadIdx.x);
    uint32_t nv = 0, nr = 0;
#pragma unroll 1
    for (int it = 0; it < kTileIters; ++it) {
        const uint32_t li = static_cast<uint32_t>(it) * kBlockThreads + threadIdx.x;
        const uint64_t item = tile_base + li;
        if (item >= p.items) break;
        cur.locate(p, item, li, sh);
        uint32_t m, st;
        int64_t sym;
        item_masks<RAGGED>(p, cur, m, st, sym);
        nv += __popc(m);
        nr += __popc(st);
        cur.advance(p);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        nv += __shfl_xor_sync(0xffffffffu, nv, d);
        nr += __shfl_xor_sync(0xffffffffu, nr, d);
    }
    if ((threadIdx.x & 31) == 0) {
        s_v[threadIdx.x >> 5] = nv;
        s_r[threadIdx.x >> 5] = nr;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tv = 0, tr = 0;
#pragma unroll
        for (int w = 0; w < kBlockThreads / 32; ++w) {
            tv += s_v[w];
            tr += s_r[w];
        }
        tile_valid[blockIdx.x] = tv;
        tile_runs[blockIdx.x] = tr;
    }
}

// Thread t owns the kTileIters CONSECUTIVE items [tile_base + 8t, +8), so that one block-wide scan
// per tile orders everything.
template <bool RAGGED>
__global__ void __launch_bounds__(kBlockThreads) emit_runs_kernel(const ExtractParams p,
                                                                  const uint64_t *__restrict__ tile_valid_off,
                                                                  const uint64_t *__restrict__ tile_runs_off,
                                                                  uint64_t *__restrict__ run_sym,
                                                                  uint64_t *__restrict__ run_woff,
                                                                  uint64_t *__restrict__ run_ibase)
{
    __shared__ TileShared<RAGGED> sh;
    __shared__ uint64_t s_w[kBlockThreads / 32];
    const uint64_t tile_base = static_cast<uint64_t>(blockIdx.x) * kTileItems;
    const uint32_t li0 = threadIdx.x * kTileIters;
    TileCursor<RAGGED, kRunG> cur;
    cur.init(p, tile_base, sh, li0);
    TileCursor<RAGGED, kRunG> cur2 = cur;

    // pass 1: this thread's survivors (low 32 bits) and run starts (high 32 bits)
    uint64_t mine = 0;
#pragma unroll 1
    for (int i = 0; i < kTileIters; ++i) {
        const uint64_t item = tile_base + li0 + i;
        if (item >= p.items) break;
        cur.locate(p, item, li0 + i, sh);
        uint32_t m, st;
        int64_t sym;
        item_masks<RAGGED>(p, cur, m, st, sym);
        mine += static_cast<uint64_t>(__popc(m)) | (static_cast<uint64_t>(__popc(st)) << 32);
        cur.advance1(p);
    }
    // block-wide exclusive scan (both halves at once; no carry between them: totals < 2^32)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint64_t before = 0;
#pragma unroll
    for (int w = 0; w < kBlockThreads / 32; ++w) before += (w < warp) ? s_w[w] : 0ull;
    const uint64_t excl = before + incl - mine;
    uint64_t out_pos = __ldg(tile_valid_off + blockIdx.x) + (excl & 0xffffffffull);
    uint64_t run_id = __ldg(tile_runs_off + blockIdx.x) + (excl >> 32);

    // pass 2: one descriptor per run start
#pragma unroll 1
    for (int i = 0; i < kTileIters; ++i) {
        const uint64_t item = tile_base + li0 + i;
        if (item >= p.items) break;
        cur2.locate(p, item, li0 + i, sh);
        uint32_t m, st;
        int64_t sym;
        item_masks<RAGGED>(p, cur2, m, st, sym);
        const int64_t w_first = cur2.wbase + cur2.jlo; // window (in its read) of mask bit 0
        uint32_t s = st;
        while (s) {
            const int t = __ffs(s) - 1;
            s &= s - 1;
            run_sym[run_id] = static_cast<uint64_t>(sym + t);
            run_woff[run_id] = out_pos + __popc(m & ((1u << t) - 1u));
            run_ibase[run_id] = static_cast<uint64_t>(w_first + t);
            ++run_id;
        }
        out_pos += __popc(m);
        cur2.advance1(p);
    }
}

template <bool RAGGED>
cudaError_t launch_mark(ExtractParams p, uint64_t *tile_valid, uint64_t *tile_runs, cudaStream_t stream)
{
    const uint64_t tiles = (p.items + kTileItems - 1) / kTileItems;
    if (tiles == 0) return cudaSuccess;
    if (tiles > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    p.it_dq = kBlockThreads / p.gprm;
    p.it_dr = kBlockThreads % p.gprm;
    mark_runs_kernel<RAGGED><<<static_cast<unsigned>(tiles), kBlockThreads, 0, stream>>>(p, tile_valid, tile_runs);
    return cudaGetLastError();
}

template <bool RAGGED>
cudaError_t launch_emit(ExtractParams p, const uint64_t *tvo, const uint64_t *tro, uint64_t *run_sym, uint64_t *run_woff,
                        uint64_t *run_ibase, cudaStream_t stream)
{
    const uint64_t tiles = (p.items + kTileItems - 1) / kTileItems;
    if (tiles == 0) return cudaSuccess;
    emit_runs_kernel<RAGGED><<<static_cast<unsigned>(tiles), kBlockThreads, 0, stream>>>(p, tvo, tro, run_sym, run_woff, run_ibase);
    return cudaGetLastError();
}

} // namespace

cudaError_t mark_runs(const ExtractParams &p, bool ragged, uint64_t *tile_valid, uint64_t *tile_runs, cudaStream_t stream)
{
    return ragged ? launch_mark<true>(p, tile_valid, tile_runs, stream) : launch_mark<false>(p, tile_valid, tile_runs, stream);
}

cudaError_t emit_runs(const ExtractParams &p, bool ragged, const uint64_t *tile_valid_off, const uint64_t *tile_runs_off,
                      uint64_t *run_sym, uint64_t *run_woff, uint64_t *run_ibase, cudaStream_t stream)
{
    return ragged ? launch_emit<true>(p, tile_valid_off, tile_runs_off, run_sym, run_woff, run_ibase, stream)
                  : launch_emit<false>(p, tile_valid_off, tile_runs_off, run_sym, run_woff, run_ibase, stream);
}

} // namespace kmc
