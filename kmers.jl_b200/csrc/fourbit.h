// fourbit.h -- sources that are recoded on the device before extraction: 4-bit (DNAAlphabet{4} /
// RNAAlphabet{4}) LongSequences, the FourToTwo scheme (src/construction.jl:85-86), and ASCII byte
// strings, the AsciiEncode scheme (construction.jl:95-96, ascii.cu).
//
// A 4-bit source is first recoded on the device into (i) a 2-bit stream (trailing_zeros of each
// one-hot nibble, construction_utils.jl:41-54) and (ii) a bit stream that flags every uncertain
// symbol (count_ones != 1: IUPAC ambiguity codes, N and gap).  From (ii) a "valid start" bit
// stream is derived: bit P = no uncertain symbol in [P, P+K).  The 2-bit extraction kernels then
// run unchanged on (i) and consult the valid-start bits:
//   strict FwKmers / FwRvIterator / CanonicalKmers (FwKmers.jl:104-115, CanonicalKmers.jl:131-144):
//     the first window with an uncertain symbol is reported; the call fails with KMC_E_AMBIGUOUS
//     and the position / encoding the reference's throw_uncertain (construction.jl:108-110) names;
//   UnambiguousKmers (UnambiguousKmers.jl:134-148): windows with an uncertain symbol are skipped.
//     The survivors form runs of consecutive windows; a run list is built from the valid-start
//     bits (runs.cu) and the ordinary ragged extraction kernel emits the runs in order, each
//     k-mer with its 1-based start inside its read.
#pragma once
#include "plan.h"

namespace kmc {

// What one 4-bit extraction needs between its two phases.  Phase A enqueues everything whose size
// is known up front and leaves two words in `host_small` (pinned): [0] = number of k-mers
// UnambiguousKmers will emit, [1] = first offending flat window of a strict mode (~0 = none).
// Phase B runs after the caller has synchronised the stream.
struct FourBitState {
    const kmc_seqs *seqs = nullptr; // device descriptor (must outlive phase B)
    const uint64_t *words4 = nullptr;
    int k = 0, mode = 0;
    uint32_t flags = 0;
    Geometry ge{};
    Layout L{};      // strict: layout for the kernel's G; UnambiguousKmers: the G = 32 run-marking layout
    ExtractParams p{};
    uint32_t *bad = nullptr;
    uint32_t *err = nullptr; // ASCII UnambiguousKmers: hard-error flags (bytes outside the skipping table)
    uint64_t *tile_valid_off = nullptr, *tile_runs_off = nullptr; // UnambiguousKmers: scans of the per-tile counts
    uint64_t *err_out = nullptr; // device u64[3]: seq, 1-based pos, encoding
    uint64_t *host_small = nullptr;
    uint64_t unit_bias = 0;
    bool unambig = false;
};

// bytes phase B of UnambiguousKmers needs for a run list of n_runs runs and n_valid k-mers
uint64_t run_scratch_bytes(uint64_t n_runs, uint64_t n_valid, int g);

// runs.cu
cudaError_t mark_runs(const ExtractParams &p, bool ragged, uint64_t *tile_valid, uint64_t *tile_runs, cudaStream_t stream);
cudaError_t emit_runs(const ExtractParams &p, bool ragged, const uint64_t *tile_valid_off, const uint64_t *tile_runs_off,
                      uint64_t *run_sym, uint64_t *run_woff, uint64_t *run_ibase, cudaStream_t stream);

// Valid-start bits of one group of 32 symbols from the flag words a[0..4] of this and the next four
// groups (a[5] = 0): bit t set <=> no flagged symbol in [t, t + K).  Sliding-window OR of length K by
// doubling -- A_1 = flags, A_2L = A_L | A_L >> L while 2L <= K, then two windows of length L cover
// [P, P+K): A_L | A_L >> (K - L).  Branch-free in the data (K is uniform), ~40 funnel shifts.
__device__ __forceinline__ uint32_t valid_start_word(uint32_t (&a)[6], int k)
{
    int L = 1;
#pragma unroll
    for (int step = 0; step < 7; ++step) {
        const int s = 1 << step; // current window length
        if (2 * s <= k) {
            if (s < 32) {
#pragma unroll
                for (int w = 0; w < 5; ++w) a[w] |= __funnelshift_r(a[w], a[w + 1], s);
            } else if (s == 32) {
#pragma unroll
                for (int w = 0; w < 5; ++w) a[w] |= a[w + 1];
            } else {
#pragma unroll
                for (int w = 0; w < 4; ++w) a[w] |= a[w + 2];
            }
            L = 2 * s;
        }
    }
    const int r = k - L; // 0 <= r < L, r < 64
    uint32_t v = a[0];
    if (r) {
        const int b = r & 31;
        if (r < 32) v |= __funnelshift_r(a[0], a[1], b);
        else v |= b ? __funnelshift_r(a[1], a[2], b) : a[1];
    }
    return ~v;
}
constexpr int kRecodeHalo = 5; // flag words beyond a group that its windows can reach (31 + K - 1 <= 158 bits)

// ascii.cu: byte sources (AsciiEncode).  lut: 0 = strict DNAAlphabet{2}, 1 = strict RNAAlphabet{2},
// 2 = the UnambiguousKmers skipping table.  n_groups = groups of 32 bytes.
// Writes rec / bad / err for the groups [0, n_groups) and the valid-start words [0, n_vstart).
cudaError_t ascii_recode(const uint8_t *bytes, uint64_t n_bytes, int lut, int k, uint32_t *rec, uint32_t *bad, uint32_t *err,
                         uint32_t *vstart, uint64_t n_groups, uint64_t n_vstart, cudaStream_t stream);
cudaError_t ascii_first_error_seq(const ExtractParams &p, const uint32_t *err, const uint64_t *seq_len, uint64_t uniform_len,
                                  unsigned long long *err_seq, int sm_count, cudaStream_t stream);
cudaError_t ascii_resolve_error(const ExtractParams &p, const uint8_t *bytes, const uint32_t *err, const uint64_t *seq_len,
                                uint64_t uniform_len, uint64_t r, uint64_t *err_out, cudaStream_t stream);

uint64_t fourbit_scratch_bytes(const kmc_seqs *s, int k, int mode);

int32_t fourbit_phase_a(kmc_ctx *ctx, const kmc_seqs *s, int32_t k, int32_t mode, uint32_t flags, const kmc_out *out,
                        cudaStream_t stream, const KnownTotals &known, uint64_t unit_bias, Scratch &scratch,
                        uint64_t *host_small, FourBitState *st);
// Fills res->n_written (and err_* with KMC_E_AMBIGUOUS).  For UnambiguousKmers: builds the run list in
// `runs` (at least run_scratch_bytes(host_small[2], host_small[0], G) bytes) and enqueues the extraction.
int32_t fourbit_phase_b(kmc_ctx *ctx, FourBitState *st, const kmc_out *out, cudaStream_t stream, Scratch &runs,
                        kmc_result *res);

// kmc_extract on device buffers: phase A, sync, phase B, sync.
int32_t extract_device_4bit(kmc_ctx *ctx, const kmc_seqs *s, int32_t k, int32_t mode, uint32_t flags,
                            const kmc_out *out, kmc_result *res, cudaStream_t stream);
int32_t count_unambiguous_4bit(kmc_ctx *ctx, const kmc_seqs *s, int32_t k, uint64_t *n_out, cudaStream_t stream);

} // namespace kmc
