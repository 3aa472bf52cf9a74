"""Minimizers on the device (kmc_minimizers) against the prose definition of the reference
("the minimum of W consecutive kmers" under the fx_hash ordering,
/root/reference/docs/src/replacements.md:28-58, test/benchmark.jl:96-119), evaluated naively per
window from the string."""
import numpy as np
import pytest

import kmertools as kt

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kc():
    import kmerscuda
    return kmerscuda


def naive_minimizers(s, k, w, step, canonical):
    out = []
    span = k + w - 1
    for i in range(0, len(s) - span + 1, step):
        best = None
        for j in range(w):
            sub = s[i + j:i + j + k]
            km = kt.kmer_int(sub)
            if canonical:
                km = min(km, kt.kmer_int(kt.revcomp(sub)))
            h = kt.fx_hash((km,))
            if best is None or h < best[0]:
                best = (h, km, i + j + 1)
        out.append(best)
    return out


@pytest.mark.parametrize("k,w", [(8, 20), (5, 9), (31, 10), (32, 33), (15, 10), (1, 1), (1, 64), (21, 1)])
def test_single_sequence(kc, k, w):
    rng = np.random.default_rng(k * 100 + w)
    for length in sorted({0, k + w - 2, k + w - 1, k + w, 333, 2000}):
        s = kt.random_dna(rng, max(length, 0))
        rs = kc.ReadSet.single(kc.LongSequence(kc.DNAAlphabet2, kt.pack2(s), len(s)))
        for step in sorted({1, w, 7}):
            for canonical in (False, True):
                want = naive_minimizers(s, k, w, step, canonical)
                km, idx, h, off = kc.minimizers(rs, k, w, step, canonical=canonical, hash=True)
                assert km.tolist() == [x[1] for x in want], (k, w, length, step, canonical)
                assert idx.tolist() == [x[2] for x in want]
                assert h.tolist() == [x[0] for x in want]
                assert off.tolist() == [0, len(want)]


FX = np.uint64(0x517cc1b727220a95)


def fast_minimizers(codes, k, w, canonical):
    """Vectorised restatement of the same definition for dense windows (step 1) over 2-bit codes:
    (kmers, 1-based starts, hashes) per window start."""
    n = len(codes)
    nk = n - k + 1
    c = codes.astype(np.uint64)
    fw = np.zeros(nk, dtype=np.uint64)
    rc = np.zeros(nk, dtype=np.uint64)
    for j in range(k):
        fw |= c[j:j + nk] << np.uint64(2 * (k - 1 - j))
        rc |= (np.uint64(3) - c[j:j + nk]) << np.uint64(2 * j)
    km = np.minimum(fw, rc) if canonical else fw
    with np.errstate(over="ignore"):
        h = km * FX
    nwin = nk - w + 1
    best_h, best_j = h[:nwin].copy(), np.zeros(nwin, dtype=np.int64)
    for j in range(1, w):
        take = h[j:j + nwin] < best_h
        best_h = np.where(take, h[j:j + nwin], best_h)
        best_j = np.where(take, j, best_j)
    idx = np.arange(nwin, dtype=np.int64) + best_j
    return km[idx], idx + 1, best_h


@pytest.mark.parametrize("w", [1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 23, 31])
@pytest.mark.parametrize("k", [1, 2, 13, 16, 17, 31, 32])
def test_dense_windows_all_classes(kc, k, w):
    """Every W class of the register-blocked sliding-minimum kernel, K at the word / limb edges, a
    sequence long enough to span many thread blocks and a length that leaves a partial group."""
    from kmerscuda import _abi
    import ctypes as C
    rng = np.random.default_rng(1000 * k + w)
    n = 70_001 + 3 * w
    codes = rng.integers(0, 4, size=n).astype(np.uint8)
    # a long homopolymer run: equal hashes inside a window, the first must win
    codes[5000:5200] = 2
    words = kt.pack_codes(codes.astype(np.uint64), 2)
    rs = kc.ReadSet.single(kc.LongSequence(kc.DNAAlphabet2, words, n))
    ctx = kc.default_context()
    drs = kc.DeviceReadSet(ctx, rs)
    for canonical in (False, True):
        wk, wi, wh = fast_minimizers(codes, k, w, canonical)
        km, idx, h, off = kc.minimizers(drs, k, w, 1, canonical=canonical, hash=True)
        assert np.array_equal(km, wk) and np.array_equal(idx, wi) and np.array_equal(h, wh)
        # without the index stream (another kernel instantiation), unaligned output base
        cap = len(wk)
        da = ctx.alloc(8 * (cap + 1))
        out = _abi.kmc_out(da.ptr + 8, None, None, None, None, cap, 0)
        res = _abi.kmc_result()
        st = ctx.lib.kmc_minimizers(ctx.handle, C.byref(drs.desc), k, w, 1, 2 if canonical else 0, 0, C.byref(out), C.byref(res))
        assert st == 0 and res.n_written == cap
        assert np.array_equal(da.download(np.uint64, cap, 8), wk)


def test_dense_windows_uniform_reads(kc):
    rng = np.random.default_rng(99)
    n_reads, length, stride = 3000, 150, 5
    codes = rng.integers(0, 4, size=(n_reads, stride * 32)).astype(np.uint64)
    words = np.concatenate([kt.pack_codes(row, 2) for row in codes])
    rs = kc.ReadSet(2, words, n_reads, uniform_len=length, uniform_stride_words=stride)
    for k, w in ((31, 10), (21, 11), (15, 4)):
        want = [fast_minimizers(row[:length], k, w, True) for row in codes]
        km, idx, h, so = kc.minimizers(rs, k, w, 1, canonical=True, hash=True)
        assert np.array_equal(km, np.concatenate([x[0] for x in want]))
        assert np.array_equal(idx, np.concatenate([x[1] for x in want]))
        assert np.array_equal(h, np.concatenate([x[2] for x in want]))
        assert so.tolist() == (np.arange(n_reads + 1) * (length - k - w + 2)).tolist()


def test_read_sets(kc):
    rng = np.random.default_rng(8)
    k, w = 15, 10
    lens = [0, 5, k + w - 2, k + w - 1, k + w, 150, 151] + rng.integers(0, 300, size=400).tolist()
    seqs = [kt.random_dna(rng, int(n)) for n in lens]
    packed = [kt.pack2(s) for s in seqs]
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(p) for p in packed])
    words = np.concatenate([p for p in packed if len(p)] + [np.zeros(1, np.uint64)])
    rs = kc.ReadSet(2, words, len(seqs), seq_word_offset=off[:-1].copy(), seq_len=np.array(lens, dtype=np.uint64))
    for step in (1, 10):
        want = [naive_minimizers(s, k, w, step, True) for s in seqs]
        km, idx, h, so = kc.minimizers(rs, k, w, step, canonical=True, hash=True)
        flat = [x for ws in want for x in ws]
        assert km.tolist() == [x[1] for x in flat] and idx.tolist() == [x[2] for x in flat]
        assert so.tolist() == np.concatenate([[0], np.cumsum([len(ws) for ws in want])]).tolist()
    # uniform read set (the C2 shape), benchmark.jl's K = 8, W = 20, step 20
    n_reads, length, stride = 500, 150, 5
    seqs = [kt.random_dna(rng, length) for _ in range(n_reads)]
    words = np.concatenate([kt.pack2(s) for s in seqs])
    rs = kc.ReadSet(2, words, n_reads, uniform_len=length, uniform_stride_words=stride)
    km, idx, h, so = kc.minimizers(rs, 8, 20, 20)
    flat = [x for s in seqs for x in naive_minimizers(s, 8, 20, 20, False)]
    assert km.tolist() == [x[1] for x in flat] and idx.tolist() == [x[2] for x in flat] and h is None


def test_argument_checks(kc):
    rs = kc.ReadSet.single(kc.LongDNA2("ACGT" * 50))
    with pytest.raises(kc.KmersCUDAError):
        kc.minimizers(rs, 33, 5)
    with pytest.raises(kc.KmersCUDAError):
        kc.minimizers(rs, 31, 40)


def test_spaced_kmers(kc):
    """SpacedKmers{A,K,J} (/root/reference/src/iterators/SpacedKmers.jl:13-21, 57-81) = windows of one k-mer, step J."""
    rows = lambda a: [tuple(int(v) for v in r) for r in a]
    assert rows(kc.SpacedDNAMers(3, 2, kc.LongDNA2("AGCGTATA")).collect()) == [kt.kmer_limbs(x) for x in ("AGC", "CGT", "TAT")]
    assert rows(kc.each_codon(kc.LongDNA2("TGACGATCGAC")).collect()) == [kt.kmer_limbs(x) for x in ("TGA", "CGA", "TCG")]
    rng = np.random.default_rng(2)
    for k, j in ((7, 7), (31, 5), (32, 1), (3, 3), (1, 4)):
        for n in (0, k - 1, k, k + j - 1, k + j, 1000):
            s = kt.random_dna(rng, max(n, 0))
            it = kc.SpacedDNAMers(k, j, kc.LongDNA2(s))
            want = [kt.kmer_limbs(s[i:i + k]) for i in range(0, len(s) - k + 1, j)]
            assert rows(it.collect()) == want and len(it) == len(want)
