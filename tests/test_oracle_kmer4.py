"""CPU: the oracle's restatement of k-mers over the 4-bit alphabets (Copyable 4 -> 4 and TwoToFour,
/root/reference/src/iterators/FwKmers.jl:88-102, CanonicalKmers.jl:107-129, transformations.jl:14-18)
against the reference's own doctest and an independent string-level definition."""
import numpy as np
import pytest

import kmertools as kt
from oracle import oracle as ko


def rows(a):
    return [tuple(int(v) for v in r) for r in a]


def pack4(s):
    return kt.pack_codes(np.array([kt.CODE4[c] for c in s], dtype=np.uint64), 4)


def test_fwrv_doctest_is_a_4bit_kat():
    # FwRvIterator{DNAAlphabet{4},3}("AGCGT") -> (AGC, GCT), (GCG, CGC), (CGT, ACG)   CanonicalKmers.jl:13-18
    a, b, _ = ko.iterate4(pack4("AGCGT"), 5, 3, ko.FWRV)
    assert rows(a) == [kt.kmer4_limbs(x) for x in ("AGC", "GCG", "CGT")]
    assert rows(b) == [kt.kmer4_limbs(x) for x in ("GCT", "CGC", "ACG")]


@pytest.mark.parametrize("k", [1, 2, 15, 16, 17, 31, 32, 33, 48, 49, 64])
def test_copyable_4bit_matches_string_definition(k):
    rng = np.random.default_rng(k)
    for n in (0, k - 1, k, k + 1, 3 * k + 7):
        s = kt.random_iupac(rng, max(n, 0))
        w = pack4(s) if s else np.zeros(1, np.uint64)
        want = kt.naive_fwrv4(s, k)
        a, b, h = ko.iterate4(w, len(s), k, ko.FWRV, want_hash=True)
        assert rows(a) == [x[0] for x in want] and rows(b) == [x[1] for x in want]
        assert h.tolist() == [kt.fx_hash(x[0]) for x in want]
        c, _, hc = ko.iterate4(w, len(s), k, ko.CANON, want_hash=True)
        assert rows(c) == [min(x) for x in want]  # tuples compare head first, like cmp(x.data, y.data)
        assert hc.tolist() == [kt.fx_hash(min(x)) for x in want]
        f, _, _ = ko.iterate4(w, len(s), k, ko.FW)
        assert rows(f) == [x[0] for x in want]


@pytest.mark.parametrize("k", [1, 7, 16, 17, 32, 33, 64])
def test_two_to_four_matches_string_definition(k):
    rng = np.random.default_rng(100 + k)
    s = kt.random_dna(rng, 2 * k + 41)
    w = kt.pack2(s)
    want = kt.naive_fwrv4(s, k)
    a, b, _ = ko.iterate4(w, len(s), k, ko.FWRV, src_bits=2)
    assert rows(a) == [x[0] for x in want] and rows(b) == [x[1] for x in want]
    c, _, _ = ko.iterate4(w, len(s), k, ko.CANON, src_bits=2)
    assert rows(c) == [min(x) for x in want]
    # a view at a non-zero offset (LongSubSeq)
    a2, _, _ = ko.iterate4(w, len(s) - 5, k, ko.FW, src_bits=2, first=5)
    assert rows(a2) == [x[0] for x in want[5:]]
