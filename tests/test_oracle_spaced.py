"""CPU: the oracle's restatement of SpacedKmers{A,K,J} (/root/reference/src/iterators/SpacedKmers.jl:22-139) and of
ASCII sources into k-mers over the 4-bit alphabets, against the reference's docstring examples, the sequences of its own
test (test/runtests.jl:849-868: `collect(SpacedKmers{A,k,space}(seq)) == [T(seq[i:i+k-1]) for i in 1:space:length(seq)-k+1]`)
and an independent string-level definition, for every recoding scheme of the nucleotide alphabets."""
import numpy as np
import pytest

import kmertools as kt
from oracle import oracle as ko


def rows(a):
    return [tuple(int(v) for v in r) for r in a]


def naive(s, k, j, bits):
    """The reference test's definition: the k-mer of s[i:i+k] for i in 1:j:(L-k+1)."""
    f = kt.kmer_limbs if bits == 2 else kt.kmer4_limbs
    return [f(s[i:i + k]) for i in range(0, len(s) - k + 1, j)]


def test_docstring_examples():
    # collect(SpacedDNAMers{3, 2}("AGCGTATA")) -> AGC CGT TAT                    SpacedKmers.jl:15-20
    assert rows(ko.spaced(b"AGCGTATA", 8, 3, 2, src_bits=8)) == [kt.kmer_limbs(x) for x in ("AGC", "CGT", "TAT")]
    # collect(each_codon(DNA, "TGACGATCGAC")) -> TGA CGA TCG                     SpacedKmers.jl:70-76
    assert rows(ko.spaced(b"TGACGATCGAC", 11, 3, 3, src_bits=8)) == [kt.kmer_limbs(x) for x in ("TGA", "CGA", "TCG")]
    # SpacedDNAMers{3, 4}("TAGAWWWW") throws EncodeError                          test/runtests.jl:866-867
    with pytest.raises(ko.AmbiguousError) as ei:
        ko.spaced(b"TAGAWWWW", 8, 3, 4, src_bits=8)
    assert (ei.value.pos, chr(ei.value.enc), ei.value.n_before) == (5, "W", 1)
    # ... and the symbols BETWEEN windows are never read: J = 4 > K = 3 skips position 4
    assert rows(ko.spaced(b"TAGWTAG", 7, 3, 4, src_bits=8)) == [kt.kmer_limbs("TAG")] * 2


@pytest.mark.parametrize("k,j", [(3, 2), (2, 4), (3, 3)])
def test_reference_test_sequences(k, j):
    # test/runtests.jl:855-865
    s4 = "TA-NGAKATCGAWTAGA"  # DNAAlphabet{4}, as a String and as codeunits
    assert rows(ko.spaced(s4.encode(), len(s4), k, j, src_bits=8, kmer_bits=4)) == naive(s4, k, j, 4)
    assert rows(ko.spaced(kt.pack4(s4), len(s4), k, j, src_bits=4, kmer_bits=4)) == naive(s4, k, j, 4)
    sr = "AUGCUGAUGAGUCGUAG"  # RNAAlphabet{2}
    assert rows(ko.spaced(sr.encode(), len(sr), k, j, src_bits=8, rna=True)) == naive(sr.replace("U", "T"), k, j, 2)
    with pytest.raises(ko.AmbiguousError):  # U is no DNA letter
        ko.spaced(sr.encode(), len(sr), k, j, src_bits=8, rna=False)
    # test_naive_spaced(DNAAlphabet{2}, rna"UAGUCGUAGUAG", 4, 3): a 4-bit source into 2-bit k-mers (FourToTwo)
    s = "TAGTCGTAGTAG"
    assert rows(ko.spaced(kt.pack4(s), len(s), 4, 3, src_bits=4)) == naive(s, 4, 3, 2)
    # test_naive_spaced(RNAAlphabet{4}, dna"TAGCCWKMMNAGCTV", 2, 3): Copyable 4 -> 4
    s = "TAGCCWKMMNAGCTV"
    assert rows(ko.spaced(kt.pack4(s), len(s), 2, 3, src_bits=4, kmer_bits=4)) == naive(s, 2, 3, 4)


@pytest.mark.parametrize("k", [1, 3, 16, 31, 32, 33, 64, 65, 128])
@pytest.mark.parametrize("j", [1, 2, 3, 7, 31, 32, 40, 200])
def test_every_scheme_matches_the_string_definition(k, j):
    rng = np.random.default_rng(1000 * k + j)
    for n in (0, k - 1, k, k + 1, k + j, 3 * k + 2 * j + 5, 777):
        s = kt.random_dna(rng, max(n, 0))
        want2 = naive(s, k, j, 2)
        assert rows(ko.spaced(kt.pack2(s) if s else np.zeros(1, np.uint64), len(s), k, j)) == want2                       # Copyable
        assert rows(ko.spaced(kt.pack4(s) if s else np.zeros(1, np.uint64), len(s), k, j, src_bits=4)) == want2           # FourToTwo
        low = "".join(c.lower() if rng.random() < 0.3 else c for c in s)
        assert rows(ko.spaced(low.encode(), len(s), k, j, src_bits=8)) == want2                                          # AsciiEncode
        if k <= 64:
            want4 = naive(s, k, j, 4)
            assert rows(ko.spaced(kt.pack2(s) if s else np.zeros(1, np.uint64), len(s), k, j, kmer_bits=4)) == want4      # TwoToFour
            iu = kt.random_iupac(rng, len(s))
            assert rows(ko.spaced(kt.pack4(iu) if iu else np.zeros(1, np.uint64), len(iu), k, j, src_bits=4, kmer_bits=4)) == naive(iu, k, j, 4)
            assert rows(ko.spaced(iu.encode(), len(iu), k, j, src_bits=8, kmer_bits=4)) == naive(iu, k, j, 4)


def test_j1_is_fwkmers_and_views():
    rng = np.random.default_rng(3)
    s = kt.random_dna(rng, 300)
    w = kt.pack2(s)
    for k in (5, 31, 33):
        a, _, _ = ko.iterate(w, len(s), k, ko.FW)
        assert np.array_equal(ko.spaced(w, len(s), k, 1), a)
        assert rows(ko.spaced(w, len(s) - 7, k, 3, first=7)) == naive(s[7:], k, 3, 2)


def test_errors_are_raised_where_the_walk_meets_them():
    # J < K: every symbol up to the last window is read, in order
    s = "ACGTACGTNACGTACGT"
    with pytest.raises(ko.AmbiguousError) as ei:
        ko.spaced(kt.pack4(s), len(s), 5, 2, src_bits=4)
    assert (ei.value.pos, ei.value.enc, ei.value.n_before) == (9, 15, 2)  # windows 1-5 and 3-7 were yielded
    # J >= K: only the windows are read; an N between them is never seen
    s = "ACGNNACGNNACG"
    assert rows(ko.spaced(kt.pack4(s), len(s), 3, 5, src_bits=4)) == [kt.kmer_limbs("ACG")] * 3
    with pytest.raises(ko.AmbiguousError) as ei:
        ko.spaced(kt.pack4(s), len(s), 4, 5, src_bits=4)
    assert (ei.value.pos, ei.value.n_before) == (4, 0)
    # invalid byte for a 4-bit alphabet
    with pytest.raises(ko.AmbiguousError) as ei:
        ko.spaced(b"ACGTNN-KXA", 10, 2, 2, src_bits=8, kmer_bits=4)
    assert (ei.value.pos, chr(ei.value.enc)) == (9, "X")
