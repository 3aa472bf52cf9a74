/*
 * kmers_oracle.c -- CPU restatement of the Kmers.jl k-mer extraction hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA
 * library (libkmerscuda.so).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product
 * path never calls into it and there is no CPU fallback.
 *
 * It restates, literally (same recurrences, same state machines, same order
 * of operations), the following reference code (paths relative to
 * /root/reference, Kmers.jl v1.2.0):
 *
 *   src/tuple_bitflipping.jl:3-50      limb shifts with carry
 *   src/kmer.jl:117-137, 603-605       geometry, get_mask
 *   src/kmer.jl:176-201                cmp / isless (limb-lexicographic)
 *   src/kmer.jl:255-261                fx_hash
 *   src/kmer.jl:511-518                shift_first_encoding
 *   src/construction_utils.jl:41-69    unsafe_extract (FourToTwo, Copyable)
 *   src/construction_utils.jl:129-134  shift_encoding
 *   src/construction.jl:108-110        throw_uncertain (reported as an error code)
 *   src/transformations.jl:1-10,21-25,32-41  reverse, complement, RC, canonical
 *   src/iterators/FwKmers.jl:40-43,62-66,88-94,104-115
 *   src/iterators/CanonicalKmers.jl:61-66,94-105,131-144,220-225
 *   src/iterators/UnambiguousKmers.jl:64-86,134-148
 *
 * Third-party semantics restated from BioSequences.jl v3 (not vendored in the
 * reference; compat "~3.4.1, 3.5", Project.toml:19): LongSequence bit layout
 * (symbol i, 1-based, at bits [(i-1)*bps mod 64, +bps) of word (i-1)*bps div 64),
 * reversebits(x, BitsPerSymbol{2}) and complement_bitpar(x, 2-bit) = ~x.
 *
 * Parity pin: the reference cannot run here (no Julia).  The oracle is pinned
 * against every hot-path known-answer value the reference's own tests and
 * doctests hold (tests/golden/reference_kats.json, checked by
 * tests/test_oracle_golden.py).  Base.hash (Julia Base tuple hash) is NOT
 * restated: parity unpinned for that one function.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define KO_MAX_LIMBS 8
#define KO_OK 0
#define KO_E_BAD_K 1
#define KO_E_AMBIGUOUS 3

typedef uint64_t u64;
#define INL static inline __attribute__((always_inline))

/* ---- src/tuple_bitflipping.jl:3-19 ------------------------------------ */
INL u64 left_shift(u64 x, unsigned n) { return x << (n & 63u); }
INL u64 right_shift(u64 x, unsigned n) { return x >> (n & 63u); }
INL u64 left_carry(u64 x, unsigned n) { return right_shift(x, 64u - n); }
INL u64 right_carry(u64 x, unsigned n) { return left_shift(x, 64u - n); }

/* src/tuple_bitflipping.jl:24-33.  x[0] is the head (most significant) limb.
 * The reference recurses to the tail first, so the carry enters at the last
 * limb and ripples toward the head.  Returns the carry out of the head. */
INL u64 leftshift_carry(u64 *x, const int N, unsigned nbits, u64 carry)
{
    for (int i = N - 1; i >= 0; --i) {
        u64 out = left_carry(x[i], nbits);
        x[i] = left_shift(x[i], nbits) | carry;
        carry = out;
    }
    return carry;
}

/* src/tuple_bitflipping.jl:35-46.  Carry enters at the head. */
INL u64 rightshift_carry(u64 *x, const int N, unsigned nbits, u64 carry)
{
    for (int i = 0; i < N; ++i) {
        u64 new_head = right_shift(x[i], nbits) | right_carry(carry, nbits);
        u64 mask = left_shift(1, nbits) - 1;
        carry = x[i] & mask;
        x[i] = new_head;
    }
    return carry;
}

/* ---- geometry: src/kmer.jl:117-137 (2-bit output alphabet: bps = 2) ---- */
INL int n_limbs(int K, int bps) { return (K * bps + 63) / 64; }
INL int per_word_capacity(int bps) { return 64 / bps; }
INL int n_unused(int K, int N, int bps) { return per_word_capacity(bps) * N - K; }
INL int bits_unused(int K, int N, int bps) { return n_unused(K, N, bps) * bps; }
INL int elements_in_head(int K, int N, int bps) { return per_word_capacity(bps) - n_unused(K, N, bps); }
/* src/kmer.jl:603-605.  Julia's `UInt(1) << 64` is 0, so the mask is all ones
 * when no bits are unused. */
INL u64 get_mask(int K, int N, int bps)
{
    int s = 64 - bits_unused(K, N, bps);
    return (s >= 64 ? 0 : ((u64)1 << s)) - 1;
}

/* BioSequences.extract_encoded_element(::LongSequence, i), i 1-based. */
INL u64 extract_encoded_element(const u64 *words, u64 i, int bps)
{
    u64 bit = (i - 1) * (u64)bps;
    return (words[bit >> 6] >> (bit & 63)) & (((u64)1 << bps) - 1);
}

INL int count_ones(u64 x) { return __builtin_popcountll(x); }
INL u64 trailing_zeros(u64 x) { return x ? (u64)__builtin_ctzll(x) : 64; }

/* src/construction_utils.jl:129-134 */
INL void shift_encoding(u64 *d, int K, const int N, u64 enc)
{
    leftshift_carry(d, N, 2, enc);
    d[0] &= get_mask(K, N, 2);
}

/* src/kmer.jl:511-518 */
INL void shift_first_encoding(u64 *d, int K, const int N, u64 enc)
{
    rightshift_carry(d, N, 2, 0);
    d[0] |= left_shift(enc, (unsigned)((elements_in_head(K, N, 2) - 1) * 2));
}

/* BioSequences.reversebits(x, BitsPerSymbol{2}()): reverse the order of the
 * 32 two-bit groups of a word. */
INL u64 reversebits2(u64 x)
{
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0f0f0f0f0f0f0f0full) | ((x & 0x0f0f0f0f0f0f0f0full) << 4);
    return __builtin_bswap64(x);
}

/* src/transformations.jl:21-25 (2-bit complement, head masked) */
INL void complement2(u64 *d, int K, const int N)
{
    for (int i = 0; i < N; ++i) d[i] = ~d[i];
    d[0] &= get_mask(K, N, 2);
}

/* src/transformations.jl:1-10 */
INL void reverse2(u64 *d, int K, const int N)
{
    u64 t[KO_MAX_LIMBS];
    for (int i = 0; i < N; ++i) t[i] = reversebits2(d[N - 1 - i]);
    rightshift_carry(t, N, (unsigned)bits_unused(K, N, 2), 0);
    for (int i = 0; i < N; ++i) d[i] = t[i];
}

/* src/transformations.jl:32-34 */
INL void reverse_complement2(u64 *d, int K, const int N)
{
    complement2(d, K, N);
    reverse2(d, K, N);
}

/* src/kmer.jl:176-201: cmp(x.data, y.data), tuples compare head first. */
INL int cmp_limbs(const u64 *a, const u64 *b, const int N)
{
    for (int i = 0; i < N; ++i) {
        if (a[i] < b[i]) return -1;
        if (a[i] > b[i]) return 1;
    }
    return 0;
}

/* src/kmer.jl:218,255-261 */
#define FX_CONSTANT 0x517cc1b727220a95ull
INL u64 bitrotate5(u64 h) { return (h << 5) | (h >> 59); }
INL u64 fx_hash_limbs(const u64 *d, const int N, u64 h)
{
    for (int i = 0; i < N; ++i) h = (bitrotate5(h) ^ d[i]) * FX_CONSTANT;
    return h;
}

/* src/construction_utils.jl:56-69 (Copyable) and :41-54 (FourToTwo).
 * `from` is the 1-based index of the first symbol.  Returns 0, or the 1-based
 * position of the first uncertain symbol (FourToTwo only); *bad_enc gets its
 * 4-bit encoding. */
INL u64 unsafe_extract(u64 *d, int K, const int N, const u64 *words, u64 from,
                       int src_bits, u64 *bad_enc)
{
    for (int i = 0; i < N; ++i) d[i] = 0;
    for (u64 i = from; i < from + (u64)K; ++i) {
        u64 enc = extract_encoded_element(words, i, src_bits);
        if (src_bits == 4) {
            if (count_ones(enc) != 1) { *bad_enc = enc; return i; }
            enc = trailing_zeros(enc);
        }
        leftshift_carry(d, N, 2, enc);
    }
    return 0;
}

/* ======================================================================
 * Iterators.  One sequence: `words` is LongSequence.data, `first` is the
 * 0-based symbol offset of the sequence start inside `words` (0 for a
 * LongSequence; non-zero models a LongSubSeq view), `len` its length.
 * The output alphabet is always 2-bit (DNAAlphabet{2} / RNAAlphabet{2} are
 * bit-identical); src_bits = 2 is the Copyable scheme, 4 is FourToTwo.
 * ====================================================================== */

enum { KO_FW = 0, KO_FWRV = 1, KO_CANON = 2 };

/* FwKmers.jl:62-66,88-94,104-115 ; CanonicalKmers.jl:61-66,94-105,131-144,220-225.
 * mode FW: out_a = fw kmers.  FWRV: out_a = fw, out_b = rv.  CANON: out_a = canonical.
 * out_hash (optional) = fx_hash of what lands in out_a (h0 = 0).
 * Outputs are N limbs per k-mer, head first.  Returns status; *n_out = k-mers
 * yielded before any error; on KO_E_AMBIGUOUS *err_pos is the 1-based symbol
 * index (relative to the sequence) and *err_enc the offending nibble. */
INL int iterate_one(const u64 *words, u64 first, u64 len, const int src_bits, int K, const int N,
                    const int mode, u64 *out_a, u64 *out_b, u64 *out_hash,
                    u64 *n_out, u64 *err_pos, u64 *err_enc)
{
    u64 fw[KO_MAX_LIMBS], rv[KO_MAX_LIMBS];
    u64 n = 0;
    *n_out = 0;
    if (len < (u64)K) return KO_OK;
    u64 bad = 0;
    u64 p = unsafe_extract(fw, K, N, words, first + 1, src_bits, &bad);
    if (p) { *err_pos = p - first; *err_enc = bad; return KO_E_AMBIGUOUS; }
    if (mode != KO_FW) {
        for (int i = 0; i < N; ++i) rv[i] = fw[i];
        reverse_complement2(rv, K, N);
    }
    u64 i = (u64)K + 1; /* next symbol to read, 1-based */
    for (;;) {
        const u64 *a = fw;
        if (mode == KO_CANON) a = (cmp_limbs(fw, rv, N) == -1) ? fw : rv;
        for (int j = 0; j < N; ++j) out_a[n * (u64)N + j] = a[j];
        if (mode == KO_FWRV)
            for (int j = 0; j < N; ++j) out_b[n * (u64)N + j] = rv[j];
        if (out_hash) out_hash[n] = fx_hash_limbs(a, N, 0);
        ++n;
        if (i > len) break;
        u64 enc = extract_encoded_element(words, first + i, src_bits);
        if (src_bits == 4) {
            if (count_ones(enc) != 1) {
                *n_out = n; *err_pos = i; *err_enc = enc;
                return KO_E_AMBIGUOUS;
            }
            enc = trailing_zeros(enc);
        }
        shift_encoding(fw, K, N, enc);
        if (mode != KO_FW) shift_first_encoding(rv, K, N, enc ^ 3u);
        ++i;
    }
    *n_out = n;
    return KO_OK;
}

/* UnambiguousKmers.jl:64-77 (2-bit source) and :79-86,134-148 (4-bit source).
 * out_kmer: N limbs per element; out_pos: 1-based start index.  With
 * out_kmer == NULL the state machine runs but only counts. */
INL void unambiguous_one(const u64 *words, u64 first, u64 len, int src_bits, int K, const int N,
                         u64 *out_kmer, int64_t *out_pos, u64 *n_out)
{
    u64 kmer[KO_MAX_LIMBS];
    u64 n = 0;
    if (src_bits == 2) {
        /* Copyable: every window, index = i - K + 1 */
        if (len >= (u64)K) {
            u64 bad;
            unsafe_extract(kmer, K, N, words, first + 1, 2, &bad);
            if (out_kmer) {
                for (int j = 0; j < N; ++j) out_kmer[j] = kmer[j];
                out_pos[0] = 1;
            }
            n = 1;
            for (u64 i = (u64)K + 1; i <= len; ++i) {
                shift_encoding(kmer, K, N, extract_encoded_element(words, first + i, 2));
                if (out_kmer) {
                    for (int j = 0; j < N; ++j) out_kmer[n * (u64)N + j] = kmer[j];
                    out_pos[n] = (int64_t)(i - (u64)K + 1);
                }
                ++n;
            }
        }
        *n_out = n;
        return;
    }
    /* FourToTwo skip/restart loop */
    for (int j = 0; j < N; ++j) kmer[j] = 0;
    u64 remaining = (u64)K, index = 1;
    for (;;) {
        while (remaining != 0) {
            if (index > len) { *n_out = n; return; }
            u64 enc = extract_encoded_element(words, first + index, 4);
            /* shifted in even when ambiguous: tz(0) = 64 is OR-ed in as-is and
             * is flushed out by the K certain symbols that must follow. */
            shift_encoding(kmer, K, N, trailing_zeros(enc));
            index += 1;
            remaining = (count_ones(enc) == 1) ? remaining - 1 : (u64)K;
        }
        if (out_kmer) {
            for (int j = 0; j < N; ++j) out_kmer[n * (u64)N + j] = kmer[j];
            out_pos[n] = (int64_t)(index - (u64)K);
        }
        ++n;
        remaining = 1;
    }
}

/* Constant (mode, src_bits) instantiation: Julia specialises iterate() on the iterator type and
 * the RecodingScheme (construction.jl:75-100 is constant-folded), so the CPU baseline gets the
 * same treatment.  MM and SB are literal constants inside CALL. */
#define DISPATCH_MODE_BITS(mode, src_bits, CALL)                                   \
    switch ((mode) * 8 + (src_bits)) {                                             \
    case KO_FW * 8 + 2: { enum { MM = KO_FW, SB = 2 }; CALL; } break;              \
    case KO_FW * 8 + 4: { enum { MM = KO_FW, SB = 4 }; CALL; } break;              \
    case KO_FWRV * 8 + 2: { enum { MM = KO_FWRV, SB = 2 }; CALL; } break;          \
    case KO_FWRV * 8 + 4: { enum { MM = KO_FWRV, SB = 4 }; CALL; } break;          \
    case KO_CANON * 8 + 2: { enum { MM = KO_CANON, SB = 2 }; CALL; } break;        \
    default: { enum { MM = KO_CANON, SB = 4 }; CALL; } break;                      \
    }

/* Constant-N instantiation so the compiler unrolls the limb loops the way
 * Julia specialises on NTuple{N,UInt64}. */
#define DISPATCH_N(N, CALL)                                         \
    switch (N) {                                                    \
    case 1: { enum { NN = 1 }; CALL; } break;                       \
    case 2: { enum { NN = 2 }; CALL; } break;                       \
    case 3: { enum { NN = 3 }; CALL; } break;                       \
    case 4: { enum { NN = 4 }; CALL; } break;                       \
    case 5: { enum { NN = 5 }; CALL; } break;                       \
    case 6: { enum { NN = 6 }; CALL; } break;                       \
    case 7: { enum { NN = 7 }; CALL; } break;                       \
    default: { enum { NN = 8 }; CALL; } break;                      \
    }

static int check_k(int K)
{
    if (K < 1) return KO_E_BAD_K;
    if (n_limbs(K, 2) > KO_MAX_LIMBS) return KO_E_BAD_K;
    return KO_OK;
}

/* ---------------------------- exported API ----------------------------- */

int ko_n_limbs(int K) { return n_limbs(K, 2); }

/* FwKmers.jl:40-43 */
uint64_t ko_n_windows(uint64_t len, int K) { return len >= (u64)K ? len - (u64)K + 1 : 0; }

int ko_iterate(const uint64_t *words, uint64_t first, uint64_t len, int src_bits, int K, int mode,
               uint64_t *out_a, uint64_t *out_b, uint64_t *out_hash,
               uint64_t *n_out, uint64_t *err_pos, uint64_t *err_enc)
{
    int st = check_k(K);
    if (st) return st;
    int N = n_limbs(K, 2);
    if (mode < KO_FW || mode > KO_CANON || (src_bits != 2 && src_bits != 4)) return KO_E_BAD_K;
    DISPATCH_MODE_BITS(mode, src_bits,
                       DISPATCH_N(N, st = iterate_one(words, first, len, SB, K, NN, MM, out_a, out_b,
                                                      out_hash, n_out, err_pos, err_enc)));
    return st;
}

int ko_unambiguous(const uint64_t *words, uint64_t first, uint64_t len, int src_bits, int K,
                   uint64_t *out_kmer, int64_t *out_pos, uint64_t *n_out)
{
    int st = check_k(K);
    if (st) return st;
    int N = n_limbs(K, 2);
    DISPATCH_N(N, unambiguous_one(words, first, len, src_bits, K, NN, out_kmer, out_pos, n_out));
    return KO_OK;
}

/* fx_hash.(v) over an array of n k-mers of N limbs each (src/kmer.jl:255-261) */
void ko_fx_hash(const uint64_t *kmers, uint64_t n, int N, uint64_t h0, uint64_t *out)
{
    for (u64 i = 0; i < n; ++i) {
        u64 h = h0;
        for (int j = 0; j < N; ++j) h = (bitrotate5(h) ^ kmers[i * (u64)N + j]) * FX_CONSTANT;
        out[i] = h;
    }
}

/* Single-kmer transformations (src/transformations.jl) for the KAT tests. */
void ko_reverse_complement(uint64_t *d, int K)
{
    int N = n_limbs(K, 2);
    DISPATCH_N(N, reverse_complement2(d, K, NN));
}
void ko_reverse(uint64_t *d, int K)
{
    int N = n_limbs(K, 2);
    DISPATCH_N(N, reverse2(d, K, NN));
}
void ko_complement(uint64_t *d, int K)
{
    int N = n_limbs(K, 2);
    DISPATCH_N(N, complement2(d, K, NN));
}
int ko_cmp(const uint64_t *a, const uint64_t *b, int N) { return cmp_limbs(a, b, N); }
void ko_shift_encoding(uint64_t *d, int K, uint64_t enc)
{
    int N = n_limbs(K, 2);
    DISPATCH_N(N, shift_encoding(d, K, NN, enc));
}
void ko_shift_first_encoding(uint64_t *d, int K, uint64_t enc)
{
    int N = n_limbs(K, 2);
    DISPATCH_N(N, shift_first_encoding(d, K, NN, enc));
}
/* Kmer{A,K,N}(seq[from:from+K-1]) through unsafe_extract; returns 0 or the
 * position of the first uncertain symbol. */
uint64_t ko_unsafe_extract(uint64_t *d, int K, const uint64_t *words, uint64_t from, int src_bits,
                           uint64_t *bad_enc)
{
    int N = n_limbs(K, 2);
    u64 r = 0;
    DISPATCH_N(N, r = unsafe_extract(d, K, NN, words, from, src_bits, bad_enc));
    return r;
}

/* ----------------------------------------------------------------------
 * Batch drivers (the CPU baseline).  A read set is word-aligned CSR:
 * read r occupies words [word_off[r], word_off[r+1]) and has seq_len[r]
 * symbols; if word_off == NULL the set is uniform (uniform_len symbols,
 * uniform_stride words per read).  Output element offsets out_off[r]
 * (exclusive prefix sum of per-read output counts) must be supplied for the
 * fixed-count modes; reads are independent, so the loop is an OpenMP
 * `parallel for` over reads, each read walked with the reference's serial
 * per-symbol recurrence.  Returns status; on an ambiguity error reports the
 * first failing read in iteration order.
 * ---------------------------------------------------------------------- */
int ko_batch_iterate(const uint64_t *words, uint64_t n_seqs, const uint64_t *word_off,
                     const uint64_t *seq_len, uint64_t uniform_len, uint64_t uniform_stride,
                     int src_bits, int K, int mode, const uint64_t *out_off,
                     uint64_t *out_a, uint64_t *out_b, uint64_t *out_hash,
                     uint64_t *err_seq, uint64_t *err_pos, uint64_t *err_enc, int threads)
{
    int st0 = check_k(K);
    if (st0) return st0;
    if (mode < KO_FW || mode > KO_CANON || (src_bits != 2 && src_bits != 4)) return KO_E_BAD_K;
    const int N = n_limbs(K, 2);
    int failed = 0;
    u64 best_seq = ~(u64)0, best_pos = 0, best_enc = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)n_seqs; ++r) {
        u64 wo = word_off ? word_off[r] : (u64)r * uniform_stride;
        u64 len = seq_len ? seq_len[r] : uniform_len;
        u64 oo = out_off ? out_off[r] : (u64)r * ko_n_windows(uniform_len, K);
        u64 n = 0, ep = 0, ee = 0;
        int st = KO_OK;
        DISPATCH_MODE_BITS(mode, src_bits,
                           DISPATCH_N(N, st = iterate_one(words + wo, 0, len, SB, K, NN, MM,
                                                          out_a + oo * (u64)N, out_b ? out_b + oo * (u64)N : NULL,
                                                          out_hash ? out_hash + oo : NULL, &n, &ep, &ee)));
        if (st == KO_E_AMBIGUOUS) {
#pragma omp critical
            {
                failed = 1;
                if ((u64)r < best_seq) { best_seq = (u64)r; best_pos = ep; best_enc = ee; }
            }
        }
    }
    if (failed) {
        *err_seq = best_seq; *err_pos = best_pos; *err_enc = best_enc;
        return KO_E_AMBIGUOUS;
    }
    return KO_OK;
}

/* Unambiguous over a read set: count per read (parallel), exclusive scan,
 * write per read at its offset (parallel), so the output is the in-order
 * concatenation over reads.  out_seq_off (n_seqs+1 entries) receives the
 * per-read output offsets; out_seq_off[n_seqs] is the total.  With
 * out_kmer == NULL only the offsets are produced (the caller sizes buffers
 * from them and calls again). */
int ko_batch_unambiguous(const uint64_t *words, uint64_t n_seqs, const uint64_t *word_off,
                         const uint64_t *seq_len, uint64_t uniform_len, uint64_t uniform_stride,
                         int src_bits, int K, uint64_t *out_seq_off,
                         uint64_t *out_kmer, int64_t *out_pos, int threads)
{
    int st0 = check_k(K);
    if (st0) return st0;
    const int N = n_limbs(K, 2);
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)n_seqs; ++r) {
        u64 wo = word_off ? word_off[r] : (u64)r * uniform_stride;
        u64 len = seq_len ? seq_len[r] : uniform_len;
        u64 n = 0;
        DISPATCH_N(N, unambiguous_one(words + wo, 0, len, src_bits, K, NN, NULL, NULL, &n));
        out_seq_off[r + 1] = n;
    }
    out_seq_off[0] = 0;
    for (u64 r = 0; r < n_seqs; ++r) out_seq_off[r + 1] += out_seq_off[r];
    if (!out_kmer) return KO_OK;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)n_seqs; ++r) {
        u64 wo = word_off ? word_off[r] : (u64)r * uniform_stride;
        u64 len = seq_len ? seq_len[r] : uniform_len;
        u64 oo = out_seq_off[r], n = 0;
        DISPATCH_N(N, unambiguous_one(words + wo, 0, len, src_bits, K, NN,
                                      out_kmer + oo * (u64)N, out_pos + oo, &n));
    }
    return KO_OK;
}

int ko_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ======================================================================
 * k-mers over the 4-bit alphabets: Kmer{DNAAlphabet{4},K,N}, N = cld(4K, 64) (src/kmer.jl:117-137
 * with bps = 4).  Sources: 4-bit LongSequence (Copyable) or 2-bit LongSequence (TwoToFour).
 * BioSequences (not vendored) supplies complement / reversebits for 4-bit encodings:
 *   complement of a 4-bit nucleotide reverses its four bits (A=0001 <-> T=1000, C=0010 <-> G=0100,
 *   ambiguity sets likewise, N and gap are fixed points); complement_bitpar does so per nibble;
 *   reversebits(x, BitsPerSymbol{4}) reverses the order of the 16 nibbles.
 * Pinned by the reference's FwRvIterator{DNAAlphabet{4},3}("AGCGT") doctest
 * (src/iterators/CanonicalKmers.jl:13-18) and by the string-level definitions of the tests.
 * ====================================================================== */
INL u64 complement_nibble(u64 e) { return ((e & 1) << 3) | ((e & 2) << 1) | ((e & 4) >> 1) | ((e & 8) >> 3); }

INL u64 complement_bitpar4(u64 x)
{
    return ((x & 0x1111111111111111ull) << 3) | ((x & 0x8888888888888888ull) >> 3) |
           ((x & 0x2222222222222222ull) << 1) | ((x & 0x4444444444444444ull) >> 1);
}

INL u64 reversebits4(u64 x)
{
    x = ((x >> 4) & 0x0f0f0f0f0f0f0f0full) | ((x & 0x0f0f0f0f0f0f0f0full) << 4);
    return __builtin_bswap64(x);
}

/* src/construction_utils.jl:129-134 with bps = 4 */
INL void shift_encoding4(u64 *d, int K, const int N, u64 enc)
{
    leftshift_carry(d, N, 4, enc);
    d[0] &= get_mask(K, N, 4);
}

/* src/kmer.jl:511-518 with bps = 4 */
INL void shift_first_encoding4(u64 *d, int K, const int N, u64 enc)
{
    rightshift_carry(d, N, 4, 0);
    d[0] |= left_shift(enc, (unsigned)((elements_in_head(K, N, 4) - 1) * 4));
}

/* src/transformations.jl:14-18 (no head mask: the complement of gap is gap), :1-10, :32-34 */
INL void reverse_complement4(u64 *d, int K, const int N)
{
    u64 t[KO_MAX_LIMBS];
    for (int i = 0; i < N; ++i) d[i] = complement_bitpar4(d[i]);
    for (int i = 0; i < N; ++i) t[i] = reversebits4(d[N - 1 - i]);
    rightshift_carry(t, N, (unsigned)bits_unused(K, N, 4), 0);
    for (int i = 0; i < N; ++i) d[i] = t[i];
}

/* FwKmers.jl:62-66 (first window), :88-94 (Copyable), :96-102 (TwoToFour);
 * CanonicalKmers.jl:61-66, :107-120 (Copyable, 4-bit), :122-129 (TwoToFour), :220-225.
 * src_bits = 4: Copyable; src_bits = 2: TwoToFour (enc4 = 1 << enc2, construction_utils.jl:27-39). */
static void iterate4_one(const u64 *words, u64 first, u64 len, int src_bits, int K, int N, int mode,
                         u64 *out_a, u64 *out_b, u64 *out_hash, u64 *n_out)
{
    u64 fw[KO_MAX_LIMBS], rv[KO_MAX_LIMBS];
    u64 n = 0;
    *n_out = 0;
    if (len < (u64)K) return;
    for (int i = 0; i < N; ++i) fw[i] = 0;
    for (u64 i = first + 1; i < first + 1 + (u64)K; ++i) {
        u64 enc = extract_encoded_element(words, i, src_bits);
        if (src_bits == 2) enc = left_shift(1, (unsigned)enc);
        leftshift_carry(fw, N, 4, enc);
    }
    if (mode != KO_FW) {
        for (int i = 0; i < N; ++i) rv[i] = fw[i];
        reverse_complement4(rv, K, N);
    }
    u64 i = (u64)K + 1;
    for (;;) {
        const u64 *a = fw;
        if (mode == KO_CANON) a = (cmp_limbs(fw, rv, N) == -1) ? fw : rv;
        for (int j = 0; j < N; ++j) out_a[n * (u64)N + j] = a[j];
        if (mode == KO_FWRV)
            for (int j = 0; j < N; ++j) out_b[n * (u64)N + j] = rv[j];
        if (out_hash) out_hash[n] = fx_hash_limbs(a, N, 0);
        ++n;
        if (i > len) break;
        u64 enc = extract_encoded_element(words, first + i, src_bits);
        u64 rc;
        if (src_bits == 2) {
            rc = left_shift(1, (unsigned)(enc ^ 3u));
            enc = left_shift(1, (unsigned)enc);
        } else {
            rc = complement_nibble(enc);
        }
        shift_encoding4(fw, K, N, enc);
        if (mode != KO_FW) shift_first_encoding4(rv, K, N, rc);
        ++i;
    }
    *n_out = n;
}

int ko_n_limbs4(int K) { return n_limbs(K, 4); }

int ko_iterate4(const uint64_t *words, uint64_t first, uint64_t len, int src_bits, int K, int mode,
                uint64_t *out_a, uint64_t *out_b, uint64_t *out_hash, uint64_t *n_out)
{
    if (K < 1 || n_limbs(K, 4) > KO_MAX_LIMBS) return KO_E_BAD_K;
    if (mode < KO_FW || mode > KO_CANON || (src_bits != 2 && src_bits != 4)) return KO_E_BAD_K;
    iterate4_one(words, first, len, src_bits, K, n_limbs(K, 4), mode, out_a, out_b, out_hash, n_out);
    return KO_OK;
}

/* ======================================================================
 * ASCII sources (the AsciiEncode recoding scheme, src/construction.jl:95-96).
 *
 * BioSequences.ascii_encode(A, byte) is not in the reference tree; restated from
 * BioSequences v3 (src/alphabet.jl): a 256-entry table filled with 0x80, then for
 * every symbol of the alphabet its character and its lowercase character map to
 * the symbol's encoding.  For DNAAlphabet{2}: ACGT/acgt -> 0..3; for
 * RNAAlphabet{2}: ACGU/acgu -> 0..3; everything else >= 0x80 (EncodeError).
 * Pinned by the reference's tests only as far as: lowercase is accepted
 * (test/runtests.jl:713-719) and 'P' is rejected (:722-724, :844-846).
 * ====================================================================== */
INL unsigned ascii_encode2(int rna, unsigned b)
{
    switch (b) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return rna ? 0x80u : 3;
    case 'U': case 'u': return rna ? 3 : 0x80u;
    default: return 0x80u;
    }
}

/* src/iterators/common.jl:22-32  ASCII_SKIPPING_LUT */
INL unsigned ascii_skipping_lut(unsigned b)
{
    switch (b) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    case '-':
    case 'M': case 'R': case 'S': case 'V': case 'W': case 'Y': case 'H': case 'K': case 'D': case 'B': case 'N':
    case 'm': case 'r': case 's': case 'v': case 'w': case 'y': case 'h': case 'k': case 'd': case 'b': case 'n':
        return 0xf0;
    default: return 0xff;
    }
}

/* FwKmers.jl:69-78,117-129 ; CanonicalKmers.jl:69-79,146-174,220-225 ; construction_utils.jl:71-88 */
static int ascii_iterate_n(const uint8_t *src, u64 len, int rna, int K, const int N, int mode,
                           u64 *out_a, u64 *out_b, u64 *out_hash, u64 *n_out, u64 *err_pos, u64 *err_byte)
{
    u64 fw[KO_MAX_LIMBS], rv[KO_MAX_LIMBS];
    u64 n = 0;
    *n_out = 0;
    if (len < (u64)K) return KO_OK;
    for (int i = 0; i < N; ++i) fw[i] = 0;
    for (u64 i = 1; i <= (u64)K; ++i) { /* unsafe_extract(::AsciiEncode) */
        unsigned enc = ascii_encode2(rna, src[i - 1]);
        if (enc > 0x7f) { *err_pos = i; *err_byte = src[i - 1]; return KO_E_AMBIGUOUS; }
        leftshift_carry(fw, N, 2, enc);
    }
    if (mode != KO_FW) {
        for (int i = 0; i < N; ++i) rv[i] = fw[i];
        reverse_complement2(rv, K, N);
    }
    u64 i = (u64)K + 1;
    for (;;) {
        const u64 *a = fw;
        if (mode == KO_CANON) a = (cmp_limbs(fw, rv, N) == -1) ? fw : rv;
        for (int j = 0; j < N; ++j) out_a[n * (u64)N + j] = a[j];
        if (mode == KO_FWRV)
            for (int j = 0; j < N; ++j) out_b[n * (u64)N + j] = rv[j];
        if (out_hash) out_hash[n] = fx_hash_limbs(a, N, 0);
        ++n;
        if (i > len) break;
        unsigned enc = ascii_encode2(rna, src[i - 1]);
        if (enc > 0x7f) { *n_out = n; *err_pos = i; *err_byte = src[i - 1]; return KO_E_AMBIGUOUS; }
        shift_encoding(fw, K, N, enc);
        if (mode != KO_FW) shift_first_encoding(rv, K, N, enc ^ 3u);
        ++i;
    }
    *n_out = n;
    return KO_OK;
}

int ko_ascii_iterate(const uint8_t *src, uint64_t len, int rna, int K, int mode,
                     uint64_t *out_a, uint64_t *out_b, uint64_t *out_hash,
                     uint64_t *n_out, uint64_t *err_pos, uint64_t *err_byte)
{
    int st = check_k(K);
    if (st) return st;
    return ascii_iterate_n(src, len, rna, K, n_limbs(K, 2), mode, out_a, out_b, out_hash, n_out, err_pos, err_byte);
}

/* UnambiguousKmers.jl:79-86 (initial state) and :109-132 (the ASCII loop).  out_kmer may be NULL (count only). */
int ko_ascii_unambiguous(const uint8_t *src, uint64_t len, int K, uint64_t *out_kmer, int64_t *out_pos,
                         uint64_t *n_out, uint64_t *err_pos, uint64_t *err_byte)
{
    int st = check_k(K);
    if (st) return st;
    const int N = n_limbs(K, 2);
    u64 kmer[KO_MAX_LIMBS];
    for (int j = 0; j < N; ++j) kmer[j] = 0;
    u64 n = 0, remaining = (u64)K, index = 1;
    for (;;) {
        while (remaining != 0) {
            if (index > len) { *n_out = n; return KO_OK; }
            unsigned byte = src[index - 1];
            index += 1;
            unsigned enc = ascii_skipping_lut(byte);
            if (enc == 0xff) { *n_out = n; *err_pos = index - 1; *err_byte = byte; return KO_E_AMBIGUOUS; }
            else if (enc == 0xf0) remaining = (u64)K;
            else { remaining -= 1; shift_encoding(kmer, K, N, enc); }
        }
        if (out_kmer) {
            for (int j = 0; j < N; ++j) out_kmer[n * (u64)N + j] = kmer[j];
            out_pos[n] = (int64_t)(index - (u64)K);
        }
        ++n;
        remaining = 1;
    }
}


/* ======================================================================
 * SpacedKmers{A,K,J} (src/iterators/SpacedKmers.jl:22-139): k-mers at the 1-based starts 1, 1+J, 1+2J, ...
 * (length: L < K ? 0 : div(L - K, J) + 1, :36-40).  The state machine is restated literally:
 *   first element        unsafe_extract(R, T, src, 1); next_index = 1 + max(J, K)            (:92-105, :110-119)
 *   later elements       stop when i > lastindex - min(K, J) + 1                              (:131)
 *                        J >= K: unsafe_extract(R, T, src, i)            (i = the start of the k-mer)
 *                        J <  K: unsafe_shift_from(R, kmer, src, i, Val(J)) (i = the first NEW symbol)   (:134-138)
 * for every recoding scheme the nucleotide alphabets have (construction.jl:75-100): Copyable (2 -> 2, 4 -> 4),
 * TwoToFour, FourToTwo (uncertain symbol -> EncodeError) and AsciiEncode (invalid byte -> EncodeError) into
 * 2-bit or 4-bit k-mers.
 *
 * BioSequences.ascii_encode for the 4-bit alphabets is not in the reference tree; restated from BioSequences v3 /
 * BioSymbols: every symbol of the alphabet -- A C M G R S V T W Y H K D B N and the gap '-', U instead of T for
 * RNA -- in either case maps to its encoding (A=1 C=2 G=4 T/U=8, ambiguity codes = the OR of their bases, N=15,
 * gap=0); everything else is > 0x7f.  The reference pins '-', 'N', 'K', 'W' being accepted through
 * test/runtests.jl:855-863 (SpacedKmers{DNAAlphabet{4}} over codeunits("TA-NGAKATCGAWTAGA")).
 * ====================================================================== */
INL unsigned ascii_encode4(int rna, unsigned b)
{
    if (b >= 'a' && b <= 'z') b -= 32;
    switch (b) {
    case '-': return 0;
    case 'A': return 1;
    case 'C': return 2;
    case 'M': return 3;
    case 'G': return 4;
    case 'R': return 5;
    case 'S': return 6;
    case 'V': return 7;
    case 'T': return rna ? 0x80u : 8;
    case 'U': return rna ? 8 : 0x80u;
    case 'W': return 9;
    case 'Y': return 10;
    case 'H': return 11;
    case 'K': return 12;
    case 'D': return 13;
    case 'B': return 14;
    case 'N': return 15;
    default: return 0x80u;
    }
}

/* one symbol (1-based index i of the sequence) recoded into the k-mer alphabet; returns 0 on success, else the
 * offending source encoding / byte through *bad */
INL int spaced_symbol(const void *src, int src_bits, u64 first, u64 i, int rna, int kmer_bits, u64 *enc, u64 *bad)
{
    if (src_bits == 8) {
        const unsigned byte = ((const uint8_t *)src)[first + i - 1];
        const unsigned e = kmer_bits == 2 ? ascii_encode2(rna, byte) : ascii_encode4(rna, byte);
        if (e > 0x7f) { *bad = byte; return 1; }
        *enc = e;
        return 0;
    }
    u64 e = extract_encoded_element((const u64 *)src, first + i, src_bits);
    if (src_bits == 4 && kmer_bits == 2) { /* FourToTwo */
        if (count_ones(e) != 1) { *bad = e; return 1; }
        e = trailing_zeros(e);
    } else if (src_bits == 2 && kmer_bits == 4) { /* TwoToFour */
        e = left_shift(1, (unsigned)e);
    }
    *enc = e;
    return 0;
}

/* src: LongSequence words (src_bits 2 / 4) or ASCII bytes (src_bits 8); first = 0-based symbol offset of the
 * sequence in src; kmer_bits = 2 or 4.  out: N limbs per k-mer.  On KO_E_AMBIGUOUS *n_out = k-mers yielded before
 * the error, *err_pos the 1-based position of the symbol that raised it, *err_enc its encoding / byte. */
int ko_spaced(const void *src, int src_bits, uint64_t first, uint64_t len, int rna, int K, int J, int kmer_bits,
              uint64_t *out, uint64_t *n_out, uint64_t *err_pos, uint64_t *err_enc)
{
    *n_out = 0;
    if (K < 1 || J < 1 || (kmer_bits != 2 && kmer_bits != 4) || n_limbs(K, kmer_bits) > KO_MAX_LIMBS) return KO_E_BAD_K;
    if (src_bits != 2 && src_bits != 4 && src_bits != 8) return KO_E_BAD_K;
    const int N = n_limbs(K, kmer_bits);
    u64 kmer[KO_MAX_LIMBS];
    u64 n = 0, enc = 0, bad = 0;
    if (len < (u64)K) return KO_OK;
    /* unsafe_extract at 1 */
    for (int j = 0; j < N; ++j) kmer[j] = 0;
    for (u64 i = 1; i <= (u64)K; ++i) {
        if (spaced_symbol(src, src_bits, first, i, rna, kmer_bits, &enc, &bad)) { *err_pos = i; *err_enc = bad; return KO_E_AMBIGUOUS; }
        leftshift_carry(kmer, N, (unsigned)kmer_bits, enc);
    }
    u64 i = 1 + (u64)(J > K ? J : K);
    const u64 mkj = (u64)(K < J ? K : J);
    for (;;) {
        for (int j = 0; j < N; ++j) out[n * (u64)N + j] = kmer[j];
        ++n;
        if (i + mkj > len + 1) break; /* i > lastindex - min(K, J) + 1 */
        if (J >= K) {
            for (int j = 0; j < N; ++j) kmer[j] = 0;
            for (u64 q = i; q < i + (u64)K; ++q) {
                if (spaced_symbol(src, src_bits, first, q, rna, kmer_bits, &enc, &bad)) { *n_out = n; *err_pos = q; *err_enc = bad; return KO_E_AMBIGUOUS; }
                leftshift_carry(kmer, N, (unsigned)kmer_bits, enc);
            }
        } else {
            for (u64 q = i; q < i + (u64)J; ++q) {
                if (spaced_symbol(src, src_bits, first, q, rna, kmer_bits, &enc, &bad)) { *n_out = n; *err_pos = q; *err_enc = bad; return KO_E_AMBIGUOUS; }
                if (kmer_bits == 2) shift_encoding(kmer, K, N, enc); else shift_encoding4(kmer, K, N, enc);
            }
        }
        i += (u64)J;
    }
    *n_out = n;
    return KO_OK;
}

/* FwKmers over an ASCII source into 4-bit k-mers (FwKmers.jl:69-78,117-129 with ascii_encode of a 4-bit alphabet):
 * the spaced iterator with J = 1 is the same state machine (unsafe_shift_from of one symbol = shift_encoding). */

/* ======================================================================
 * Base.hash(x::Kmer, h::UInt) = hash(x.data, h ⊻ (ksize(typeof(x)) % UInt))   -- src/kmer.jl:206.
 * hash(::NTuple{N,UInt64}, ::UInt) is Julia Base, not the reference; restated for Julia 1.10 / 1.11:
 *   tuple.jl     hash(::Tuple{}, h) = h + tuplehash_seed (0x77cfa1eef01bca90 on 64-bit)
 *                hash(t::Tuple, h)  = hash(t[1], hash(tail(t), h))
 *   hashing.jl   hash(x::UInt64, h) = hash_uint64(x) - 3h ; hash_uint64 = hash_64_64 (below)
 * Pinned by the value the reference documents: hash(mer"UGCUGUAC"r) == 0xe5057d38c8907b22
 * (docs/src/hashing.md:18-20).  Julia >= 1.12 changed integer hashing (the reference warns,
 * hashing.md:10-13): those versions are NOT covered.
 * ====================================================================== */
INL u64 jl_hash_64_64(u64 a)
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

void ko_base_hash(const uint64_t *kmers, uint64_t n, int N, int K, uint64_t h0, uint64_t *out)
{
    for (u64 i = 0; i < n; ++i) {
        u64 acc = (h0 ^ (u64)K) + 0x77cfa1eef01bca90ull;
        for (int j = N - 1; j >= 0; --j) acc = jl_hash_64_64(kmers[i * (u64)N + j]) - 3 * acc;
        out[i] = acc;
    }
}
