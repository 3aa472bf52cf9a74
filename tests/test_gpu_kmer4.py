"""GPU parity for k-mers over the 4-bit alphabets (KMC_KMER4): FwKmers / FwRvIterator / CanonicalKmers
yielding Kmer{DNAAlphabet{4},K,N} from 4-bit sources (Copyable, /root/reference/src/iterators/FwKmers.jl:88-94,
CanonicalKmers.jl:107-120) and from 2-bit sources (TwoToFour, FwKmers.jl:96-102, CanonicalKmers.jl:122-129),
through the C ABI, bit-exact against the oracle."""
import numpy as np
import pytest

import kmertools as kt
from oracle import oracle as ko

pytestmark = pytest.mark.gpu
FW, FWRV, CANON, UNAMBIG = 0, 1, 2, 3


@pytest.fixture(scope="module")
def kc():
    import kmerscuda
    return kmerscuda


def rows(a):
    return [tuple(int(v) for v in r) for r in a]


def test_fwrv_doctest(kc):
    # FwRvIterator{DNAAlphabet{4},3}("AGCGT")  CanonicalKmers.jl:13-18
    got = kc.FwRvIterator(kc.DNAAlphabet4, 3, kc.LongDNA4("AGCGT")).collect()
    assert rows(got[:, 0, :]) == [kt.kmer4_limbs(x) for x in ("AGC", "GCG", "CGT")]
    assert rows(got[:, 1, :]) == [kt.kmer4_limbs(x) for x in ("GCT", "CGC", "ACG")]
    # the same k-mers from a 2-bit source (TwoToFour)
    got = kc.FwRvIterator(kc.DNAAlphabet4, 3, kc.LongDNA2("AGCGT")).collect()
    assert rows(got[:, 0, :]) == [kt.kmer4_limbs(x) for x in ("AGC", "GCG", "CGT")]
    assert rows(got[:, 1, :]) == [kt.kmer4_limbs(x) for x in ("GCT", "CGC", "ACG")]
    with pytest.raises(TypeError):
        kc.UnambiguousKmers(kc.DNAAlphabet4, 3, kc.LongDNA4("AGCGT")).collect()
    with pytest.raises(ValueError):
        kc.FwKmers(kc.DNAAlphabet4, 65, kc.LongDNA4("A" * 100)).collect()


@pytest.mark.parametrize("src_bits", [4, 2])
@pytest.mark.parametrize("k", [1, 2, 8, 9, 15, 16, 17, 24, 25, 31, 32, 33, 47, 48, 49, 57, 63, 64])
def test_single_sequence(kc, k, src_bits):
    rng = np.random.default_rng(1000 * src_bits + k)
    A = kc.DNAAlphabet4 if src_bits == 4 else kc.DNAAlphabet2
    for n in sorted({0, k - 1, k, k + 1, 1000, 70_003}):
        n = max(n, 0)
        codes = rng.integers(0, 16 if src_bits == 4 else 4, size=n).astype(np.uint64)
        w = kt.pack_codes(codes, src_bits) if n else np.zeros(1, np.uint64)
        rs = kc.ReadSet.single(kc.LongSequence(A, w, n))
        for mode in (FW, FWRV, CANON):
            a, b, h = ko.iterate4(w, n, k, mode, src_bits=src_bits, want_hash=True)
            e = kc.extract(mode, rs, k, A=kc.DNAAlphabet4, hash=True)
            assert e.n == a.shape[0]
            assert np.array_equal(e.kmers, a) and np.array_equal(e.hash, h), (k, n, mode)
            if mode == FWRV:
                assert np.array_equal(e.rv, b)
                t = kc.extract(mode, rs, k, A=kc.DNAAlphabet4, aos=True)
                assert np.array_equal(t.kmers[:, 0, :], a) and np.array_equal(t.kmers[:, 1, :], b)


def test_subsequence_view(kc):
    rng = np.random.default_rng(3)
    n, k = 5000, 21
    codes = rng.integers(0, 16, size=n).astype(np.uint64)
    w = kt.pack_codes(codes, 4)
    seq = kc.LongSequence(kc.DNAAlphabet4, w, n)
    for first, length in ((1, 100), (15, 1000), (16, 999), (37, n - 37)):
        rs = kc.ReadSet.single(seq, first_symbol_offset=first, length=length)
        a, _, h = ko.iterate4(w, length, k, CANON, src_bits=4, first=first, want_hash=True)
        e = kc.extract(CANON, rs, k, A=kc.DNAAlphabet4, hash=True)
        assert np.array_equal(e.kmers, a) and np.array_equal(e.hash, h)


@pytest.mark.parametrize("src_bits", [4, 2])
def test_read_sets(kc, src_bits):
    rng = np.random.default_rng(40 + src_bits)
    per = 64 // src_bits
    for k in (5, 16, 31, 40):
        # ragged
        lens = np.concatenate([[0, k - 1, k, k + 1], rng.integers(0, 400, size=300)]).astype(np.uint64)
        off = np.zeros(len(lens) + 1, dtype=np.uint64)
        off[1:] = np.cumsum((lens + per - 1) // per)
        codes = rng.integers(0, 16 if src_bits == 4 else 4, size=int(off[-1]) * per).astype(np.uint64)
        words = kt.pack_codes(codes, src_bits)
        rs = kc.ReadSet(src_bits, np.concatenate([words, np.zeros(1, np.uint64)]), len(lens), seq_word_offset=off[:-1].copy(),
                        seq_len=lens)
        want = [ko.iterate4(words[int(o):], int(n), k, CANON, src_bits=src_bits, want_hash=True) for o, n in zip(off[:-1], lens)]
        for host_path in (False, True):
            e = kc.extract(CANON, rs, k, A=kc.DNAAlphabet4, hash=True, want_seq_offsets=True, host_path=host_path)
            assert np.array_equal(e.kmers, np.concatenate([x[0] for x in want]))
            assert np.array_equal(e.hash, np.concatenate([x[2] for x in want]))
            assert e.seq_out_offset.tolist() == np.concatenate([[0], np.cumsum([len(x[0]) for x in want])]).tolist()
        # uniform
        n_reads, length = 777, 150
        stride = (length + per - 1) // per
        codes = rng.integers(0, 16 if src_bits == 4 else 4, size=n_reads * stride * per).astype(np.uint64)
        words = kt.pack_codes(codes, src_bits)
        rs = kc.ReadSet(src_bits, words, n_reads, uniform_len=length, uniform_stride_words=stride)
        want = [ko.iterate4(words[r * stride:], length, k, FWRV, src_bits=src_bits) for r in range(n_reads)]
        e = kc.extract(FWRV, rs, k, A=kc.DNAAlphabet4)
        assert np.array_equal(e.kmers, np.concatenate([x[0] for x in want]))
        assert np.array_equal(e.rv, np.concatenate([x[1] for x in want]))


def test_host_path_long_sequence(kc):
    """Chunked host pipeline (several chunks) for a 4-bit source -> 4-bit k-mers."""
    rng = np.random.default_rng(9)
    n, k = 9_000_000, 31
    codes = rng.integers(0, 16, size=n).astype(np.uint64)
    w = kt.pack_codes(codes, 4)
    rs = kc.ReadSet.single(kc.LongSequence(kc.DNAAlphabet4, w, n))
    a, _, h = ko.iterate4(w, n, k, CANON, src_bits=4, want_hash=True)
    e = kc.extract(CANON, rs, k, A=kc.DNAAlphabet4, hash=True, host_path=True)
    assert np.array_equal(e.kmers, a) and np.array_equal(e.hash, h)
