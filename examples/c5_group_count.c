/* c5_group_count.c -- config C5 of BASELINE.json from plain C, no Python and no torch in the loop: the canonical
 * 31-mer hash-bucket count table over synthetic 150 bp reads on every GPU of the box, the per-GPU tables merged
 * by NCCL inside libkmerscuda (kmc_group_bucket_count: the merge of a finished table range overlaps the count of
 * the later ranges).  This is what a Julia session does through ccall (julia/KmersCUDA: bucket_count(...; group)).
 *
 *   gcc -std=c99 -O2 -Iinclude examples/c5_group_count.c -Lkmers.jl_b200 -lkmerscuda -Wl,-rpath,$PWD/kmers.jl_b200 -o c5_group_count
 *   ./c5_group_count [reads_per_gpu = 25000000] [bucket_bits = 28] [gpus = all]
 *
 * Prints one line per step and a summary; checks that the merged table holds exactly as many counts as there are
 * k-mers on all GPUs and that every GPU holds the same merged table.  Without a CUDA device the library refuses to
 * work (there is no CPU fallback) and the program says so. */
#define _POSIX_C_SOURCE 200112L
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "kmerscuda.h"

enum { K = 31, READ_LEN = 150, STRIDE = 5, MAX_GPUS = 16 };

static uint64_t splitmix64(uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static double now_ms(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec * 1e3 + t.tv_nsec * 1e-6;
}

#define CHECK(ctx, call)                                                                             \
    do {                                                                                             \
        int32_t st__ = (call);                                                                       \
        if (st__ != KMC_OK) {                                                                        \
            fprintf(stderr, "%s: %s (%s)\n", #call, kmc_status_string(st__), kmc_last_error(ctx));   \
            return 1;                                                                                \
        }                                                                                            \
    } while (0)

int main(int argc, char **argv)
{
    uint64_t reads = argc > 1 ? strtoull(argv[1], NULL, 10) : 25000000ull;
    int bits = argc > 2 ? atoi(argv[2]) : 28;
    int32_t n_dev = 0, want = argc > 3 ? atoi(argv[3]) : 0;
    kmc_group *group = NULL;
    kmc_ctx *ctx[MAX_GPUS];
    kmc_seqs seqs[MAX_GPUS];
    kmc_result res[MAX_GPUS];
    uint32_t *tables[MAX_GPUS];
    uint64_t *d_words[MAX_GPUS];
    uint64_t *h_words;
    const uint64_t n_words = reads * STRIDE, n_kmers = reads * (READ_LEN - K + 1), table_bytes = 4ull << bits;
    int32_t st, g;
    int step;

    printf("libkmerscuda ABI version %d\n", (int)kmc_version());
    st = kmc_device_count(&n_dev);
    if (st != KMC_OK || n_dev == 0) {
        printf("no CUDA device (%s): libkmerscuda has no CPU fallback\n", kmc_status_string(st ? st : KMC_E_NO_DEVICE));
        return 0;
    }
    if (want > 0 && want < n_dev) n_dev = want;
    if (n_dev > MAX_GPUS) n_dev = MAX_GPUS;
    st = kmc_group_create(n_dev, NULL, &group);
    if (st != KMC_OK) {
        fprintf(stderr, "kmc_group_create(%d): %s\n", (int)n_dev, kmc_status_string(st));
        return 1;
    }
    /* synthetic reads: word j of GPU g = splitmix64(seed + g * 2^40 + j), the trailing bits of every read zeroed */
    h_words = (uint64_t *)malloc(n_words * sizeof *h_words);
    if (!h_words) return 1;
    for (g = 0; g < n_dev; ++g) {
        uint64_t j;
        CHECK(NULL, kmc_group_ctx(group, g, &ctx[g]));
        for (j = 0; j < n_words; ++j) h_words[j] = splitmix64(439824ull + ((uint64_t)g << 40) + j);
        for (j = STRIDE - 1; j < n_words; j += STRIDE) h_words[j] &= (1ull << (2 * (READ_LEN - 32 * (STRIDE - 1)))) - 1;
        CHECK(ctx[g], kmc_malloc(ctx[g], n_words * 8, (void **)&d_words[g]));
        CHECK(ctx[g], kmc_malloc(ctx[g], table_bytes, (void **)&tables[g]));
        CHECK(ctx[g], kmc_upload(ctx[g], d_words[g], h_words, n_words * 8));
        CHECK(ctx[g], kmc_sync(ctx[g]));
        memset(&seqs[g], 0, sizeof seqs[g]);
        seqs[g].words = d_words[g];
        seqs[g].n_words = n_words;
        seqs[g].n_seqs = reads;
        seqs[g].uniform_len = READ_LEN;
        seqs[g].uniform_stride_words = STRIDE;
        seqs[g].src_bits = 2;
    }
    free(h_words);

    for (step = 0; step < 4; ++step) {
        double t0, t1;
        float dev_ms = 0.f;
        for (g = 0; g < n_dev; ++g) CHECK(ctx[g], kmc_memset(ctx[g], tables[g], 0, table_bytes));
        CHECK(ctx[0], kmc_group_sync(group));
        t0 = now_ms();
        CHECK(ctx[0], kmc_group_bucket_count(group, seqs, K, bits, tables, res)); /* count on every GPU + NCCL merge */
        t1 = now_ms();
        for (g = 0; g < n_dev; ++g)
            if (res[g].kernel_ms > dev_ms) dev_ms = res[g].kernel_ms;
        printf("step %d: %d GPU(s), %" PRIu64 " reads each, B = %d: %.2f ms on the devices (max), %.2f ms wall -> %.1f G k-mers/s\n",
               step, (int)n_dev, reads, bits, dev_ms, t1 - t0, (double)n_kmers * n_dev / (dev_ms * 1e-3) / 1e9);
    }
    /* parity: the merged table sums to the k-mers of all GPUs, and every GPU holds the same table (fingerprints) */
    {
        uint64_t dg0[2] = {0, 0}, dg[2], sum = 0;
        uint32_t *h = (uint32_t *)malloc(table_bytes);
        uint64_t i;
        if (!h) return 1;
        CHECK(ctx[0], kmc_download(ctx[0], h, tables[0], table_bytes));
        for (i = 0; i < (1ull << bits); ++i) sum += h[i];
        free(h);
        for (g = 0; g < n_dev; ++g) {
            CHECK(ctx[g], kmc_digest(ctx[g], (const uint64_t *)tables[g], (1ull << bits) / 2, dg));
            if (g == 0) memcpy(dg0, dg, sizeof dg);
            if (memcmp(dg0, dg, sizeof dg) != 0) {
                fprintf(stderr, "GPU %d holds a different merged table\n", (int)g);
                return 1;
            }
        }
        printf("merged table total %" PRIu64 " (expected %" PRIu64 "): %s; identical on all %d GPU(s)\n", sum, n_kmers * n_dev,
               sum == n_kmers * n_dev ? "ok" : "MISMATCH", (int)n_dev);
        if (sum != n_kmers * n_dev) return 1;
    }
    for (g = 0; g < n_dev; ++g) {
        kmc_free(ctx[g], d_words[g]);
        kmc_free(ctx[g], tables[g]);
    }
    kmc_group_destroy(group);
    return 0;
}
