/* collect_canonical.c -- the C ABI from plain C: collect(CanonicalDNAMers{5}(dna"TAGCTAGGACA")) and fx_hash of
 * every k-mer, through one kmc_extract_host call with host buffers (include/kmerscuda.h).
 *
 *   gcc -std=c99 -Iinclude examples/collect_canonical.c -Lkmers.jl_b200 -lkmerscuda -Wl,-rpath,$PWD/kmers.jl_b200 -o collect_canonical
 *
 * Expected output on a machine with a GPU (limbs derived from the reference's definitions, SURVEY.md 8c;
 * CanonicalKmers.jl:199-225, kmer.jl:255-261):
 *   7 canonical 5-mers
 *   0x9c 0x9c 0x1c9 0x172 0x328 0xa1 0x284
 * Without a CUDA device the library refuses to work (there is no CPU fallback) and the program says so. */
#include <inttypes.h>
#include <stdio.h>
#include <string.h>

#include "kmerscuda.h"

/* LongSequence{DNAAlphabet{2}} layout: symbol i in bits [2i mod 64, +2) of word i div 32; A C G T = 0 1 2 3 */
static uint64_t pack2(const char *s, uint64_t *words, uint64_t n_words)
{
    uint64_t n = strlen(s), i;
    memset(words, 0, n_words * sizeof *words);
    for (i = 0; i < n; ++i) {
        uint64_t code = s[i] == 'A' ? 0 : s[i] == 'C' ? 1 : s[i] == 'G' ? 2 : 3;
        words[i / 32] |= code << (2 * (i % 32));
    }
    return n;
}

int main(void)
{
    const char *dna = "TAGCTAGGACA";
    enum { K = 5 };
    uint64_t words[1], kmers[16], hashes[16];
    kmc_ctx *ctx = NULL;
    kmc_seqs seqs;
    kmc_out out;
    kmc_result res;
    uint64_t len, n, i;
    int32_t st;

    printf("libkmerscuda ABI version %d\n", (int)kmc_version());
    st = kmc_ctx_create(0, &ctx);
    if (st != KMC_OK) {
        printf("no CUDA device (%s): libkmerscuda has no CPU fallback\n", kmc_status_string(st));
        return 0;
    }
    len = pack2(dna, words, 1);
    n = len - K + 1;

    memset(&seqs, 0, sizeof seqs);
    seqs.words = words;
    seqs.n_words = 1;
    seqs.n_seqs = 1;
    seqs.uniform_len = len;
    seqs.uniform_stride_words = 1;
    seqs.src_bits = 2;

    memset(&out, 0, sizeof out);
    out.a = kmers;
    out.hash = hashes;
    out.capacity = n;

    st = kmc_extract_host(ctx, &seqs, K, KMC_CANON, KMC_HASH_FX, &out, &res);
    if (st != KMC_OK) {
        fprintf(stderr, "kmc_extract_host: %s (%s)\n", kmc_status_string(st), kmc_last_error(ctx));
        kmc_ctx_destroy(ctx);
        return 1;
    }
    printf("%" PRIu64 " canonical %d-mers\n", res.n_written, (int)K);
    for (i = 0; i < res.n_written; ++i) printf("%s0x%" PRIx64, i ? " " : "", kmers[i]);
    printf("\n");
    for (i = 0; i < res.n_written; ++i) printf("%sfx_hash 0x%016" PRIx64, i ? "\n" : "", hashes[i]);
    printf("\n");
    kmc_ctx_destroy(ctx);
    return 0;
}
