"""GPU parity for SpacedKmers{A,K,J} / each_codon (/root/reference/src/iterators/SpacedKmers.jl:22-139) through
kmc_extract_spaced, and for ASCII sources into k-mers over the 4-bit alphabets (FwKmers.jl:117-129,
CanonicalKmers.jl:146-174): every recoding scheme of the nucleotide alphabets (construction.jl:75-100), bit-exact against
the oracle, the reference's docstring examples and the sequences of its own test (test/runtests.jl:849-889)."""
import numpy as np
import pytest

import kmertools as kt
from oracle import oracle as ko

pytestmark = pytest.mark.gpu
FW, FWRV, CANON = 0, 1, 2


@pytest.fixture(scope="module")
def kc():
    import kmerscuda
    return kmerscuda


def rows(a):
    return [tuple(int(v) for v in r) for r in a]


def seq4(kc, s):
    return kc.LongSequence(kc.DNAAlphabet4, kt.pack4(s) if s else np.zeros(1, np.uint64), len(s))


def seq2(kc, s):
    return kc.LongSequence(kc.DNAAlphabet2, kt.pack2(s) if s else np.zeros(1, np.uint64), len(s))


def test_reference_examples(kc):
    # SpacedKmers.jl:15-20, :70-76
    assert rows(kc.SpacedDNAMers(3, 2, "AGCGTATA").collect()) == [kt.kmer_limbs(x) for x in ("AGC", "CGT", "TAT")]
    assert rows(kc.each_codon("TGACGATCGAC").collect()) == [kt.kmer_limbs(x) for x in ("TGA", "CGA", "TCG")]
    assert len(kc.SpacedDNAMers(3, 2, "AGCGTATA")) == 3
    # test/runtests.jl:866-867: EncodeError on the W of the second window; a W BETWEEN windows is never read
    with pytest.raises(kc.EncodeError) as ei:
        kc.SpacedDNAMers(3, 4, "TAGAWWWW").collect()
    assert (ei.value.position, ei.value.symbol) == (5, "W")
    assert rows(kc.SpacedDNAMers(3, 4, "TAGWTAG").collect()) == [kt.kmer_limbs("TAG")] * 2
    with pytest.raises(ValueError):
        kc.SpacedDNAMers(3, 0, "ACGT")
    # test/runtests.jl:855-865
    s4 = "TA-NGAKATCGAWTAGA"
    for k, j in ((3, 2), (2, 4), (3, 3)):
        want = [kt.kmer4_limbs(s4[i:i + k]) for i in range(0, len(s4) - k + 1, j)]
        assert rows(kc.SpacedKmers(kc.DNAAlphabet4, k, j, s4).collect()) == want
        assert rows(kc.SpacedKmers(kc.DNAAlphabet4, k, j, s4.encode()).collect()) == want
        assert rows(kc.SpacedKmers(kc.DNAAlphabet4, k, j, seq4(kc, s4)).collect()) == want
        sr = "AUGCUGAUGAGUCGUAG"
        wr = [kt.kmer_limbs(sr.replace("U", "T")[i:i + k]) for i in range(0, len(sr) - k + 1, j)]
        assert rows(kc.SpacedKmers(kc.RNAAlphabet2, k, j, sr).collect()) == wr
        with pytest.raises(kc.EncodeError):
            kc.SpacedKmers(kc.DNAAlphabet2, k, j, sr).collect()
    assert rows(kc.SpacedDNAMers(4, 3, seq4(kc, "TAGTCGTAGTAG")).collect()) == rows(ko.spaced(kt.pack4("TAGTCGTAGTAG"), 12, 4, 3, src_bits=4))


@pytest.mark.parametrize("k", [1, 3, 16, 31, 32, 33, 63, 64, 65, 97, 128])
@pytest.mark.parametrize("j", [1, 2, 3, 31, 40, 200])
def test_every_scheme_vs_oracle(kc, k, j):
    rng = np.random.default_rng(77 * k + j)
    for n in (0, k - 1, k, k + j, 5000 + k):
        s = kt.random_dna(rng, max(n, 0))
        w2, w4 = (kt.pack2(s), kt.pack4(s)) if s else (np.zeros(1, np.uint64), np.zeros(1, np.uint64))
        want = ko.spaced(w2, len(s), k, j)
        for src in (seq2(kc, s), seq4(kc, s), s, s.lower().encode()):
            e = kc.extract_spaced(kc.ReadSet.ascii(src) if not isinstance(src, kc.LongSequence) else kc.ReadSet.single(src), k, j,
                                  hash=True)
            assert np.array_equal(e.kmers, want)
            assert np.array_equal(e.hash, ko.fx_hash(want) if want.size else np.zeros(0, np.uint64))
        if k <= 64:
            iu = kt.random_iupac(rng, len(s))
            wi = kt.pack4(iu) if iu else np.zeros(1, np.uint64)
            want4 = ko.spaced(wi, len(iu), k, j, src_bits=4, kmer_bits=4)
            assert np.array_equal(kc.SpacedKmers(kc.DNAAlphabet4, k, j, seq4(kc, iu)).collect(), want4)       # Copyable 4 -> 4
            assert np.array_equal(kc.SpacedKmers(kc.DNAAlphabet4, k, j, iu.lower()).collect(), want4)       # AsciiEncode -> 4-bit
            assert np.array_equal(kc.SpacedKmers(kc.DNAAlphabet4, k, j, seq2(kc, s)).collect(),               # TwoToFour
                                  ko.spaced(w2, len(s), k, j, kmer_bits=4))


def test_read_sets_views_and_errors(kc):
    rng = np.random.default_rng(9)
    k, j = 21, 5
    # ragged 4-bit reads with the occasional N: the first offending window decides
    lens = [0, 3, 20, 21, 22, 26, 150, 151] + rng.integers(0, 300, size=200).tolist()
    codes = [np.uint64(1) << rng.integers(0, 4, size=int(n)).astype(np.uint64) for n in lens]
    packed = [kt.pack_codes(c, 4) if len(c) else np.zeros(0, np.uint64) for c in codes]
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(p) for p in packed])
    words = np.concatenate([p for p in packed if len(p)] + [np.zeros(1, np.uint64)])
    rs = kc.ReadSet(4, words, len(lens), seq_word_offset=off[:-1].copy(), seq_len=np.array(lens, dtype=np.uint64))
    per = [ko.spaced(words[int(off[i]):int(off[i + 1]) + 1], int(lens[i]), k, j, src_bits=4) for i in range(len(lens))]
    e = kc.extract_spaced(rs, k, j, hash=True, want_seq_offsets=True)
    assert np.array_equal(e.kmers, np.concatenate(per))
    assert np.array_equal(e.seq_out_offset, np.concatenate([[0], np.cumsum([p.shape[0] for p in per])]).astype(np.uint64))
    # plant an N inside a sampled window of read 150 and one BETWEEN windows of an earlier read (J > K there)
    bad = words.copy()
    r = next(i for i in range(100, len(lens)) if lens[i] >= 60)
    pos = 7
    bad[int(off[r]) + pos // 16] |= np.uint64(15) << np.uint64(4 * (pos % 16))
    with pytest.raises(kc.EncodeError) as ei:
        kc.extract_spaced(kc.ReadSet(4, bad, len(lens), seq_word_offset=off[:-1].copy(), seq_len=np.array(lens, dtype=np.uint64)), k, j)
    assert (ei.value.seq_index, ei.value.position, ei.value.symbol) == (r, pos + 1, "N")
    # uniform 2-bit reads and a view with first_symbol_offset
    n_reads, length, stride = 500, 150, 5
    w = rng.integers(0, 2**64, size=n_reads * stride, dtype=np.uint64)
    for first in (0, 3):
        rsu = kc.ReadSet(2, w, n_reads, uniform_len=length - first, uniform_stride_words=stride, first_symbol_offset=first)
        want = np.concatenate([ko.spaced(w[i * stride:(i + 1) * stride], length - first, 31, 7, first=first) for i in range(n_reads)])
        assert np.array_equal(kc.extract_spaced(rsu, 31, 7).kmers, want)
    # a step larger than K over a 4-bit source with N between the windows: no error
    s = ("ACGTACGTAC" + "NNNNN") * 40
    assert np.array_equal(kc.SpacedDNAMers(10, 15, seq4(kc, s)).collect(), ko.spaced(kt.pack4(s), len(s), 10, 15, src_bits=4))


@pytest.mark.parametrize("k", [1, 5, 16, 17, 32, 33, 64])
def test_ascii_sources_into_4bit_kmers(kc, k):
    """FwKmers / FwRvIterator / CanonicalKmers{DNAAlphabet{4}} over String / codeunits (FwKmers.jl:117-129,
    CanonicalKmers.jl:146-174): the same k-mers as over the LongDNA{4} of the same letters; any byte that is no IUPAC
    letter or gap is an EncodeError at that byte."""
    rng = np.random.default_rng(k)
    for n in (0, k - 1, k, 3 * k + 11, 4000):
        iu = kt.random_iupac(rng, n)
        mixed = "".join(c.lower() if rng.random() < 0.4 else c for c in iu)
        wi = kt.pack4(iu) if iu else np.zeros(1, np.uint64)
        for mode, omode in ((FW, ko.FW), (FWRV, ko.FWRV), (CANON, ko.CANON)):
            a, b, h = ko.iterate4(wi, len(iu), k, omode, want_hash=True)
            e = kc.extract(mode, kc.ReadSet.ascii(mixed), k, A=kc.DNAAlphabet4, hash=True)
            assert np.array_equal(e.kmers, a) and np.array_equal(e.hash, h)
            if mode == FWRV:
                assert np.array_equal(e.rv, b)
            eh = kc.extract(mode, kc.ReadSet.ascii(mixed.encode()), k, A=kc.DNAAlphabet4, hash=True, host_path=True)
            assert np.array_equal(eh.kmers, a) and np.array_equal(eh.hash, h)
    # errors: position and byte; U is a letter of RNA only; sequences shorter than K are never read
    s = "ACGTNNKM-ACGT" * 3
    bad = s[:20] + "X" + s[21:]
    if len(bad) >= k:
        for host_path in (False, True):
            with pytest.raises(kc.EncodeError) as ei:
                kc.extract(FW, kc.ReadSet.ascii(bad), k, A=kc.DNAAlphabet4, host_path=host_path)
            assert (ei.value.position, ei.value.symbol) == (21, "X")
    with pytest.raises(kc.EncodeError):
        kc.extract(FW, kc.ReadSet.ascii("ACGU" * 20), k, A=kc.DNAAlphabet4)
    want, _, _ = ko.iterate4(kt.pack4("ACGT" * 20), 80, k, ko.FW)
    assert np.array_equal(kc.extract(FW, kc.ReadSet.ascii("ACGU" * 20), k, A=kc.RNAAlphabet4).kmers, want)
    short = kc.ReadSet.from_strings(["XX", "ACGTNACGT" * 10])
    if k > 2:
        e = kc.extract(FW, short, k, A=kc.DNAAlphabet4)
        assert e.n == 90 - k + 1
