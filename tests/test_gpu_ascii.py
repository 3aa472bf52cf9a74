"""GPU parity for ASCII sources (the AsciiEncode recoding scheme): String / codeunits / byte-vector
sources recoded on the device, strict FwKmers / FwRvIterator / CanonicalKmers
(/root/reference/src/iterators/FwKmers.jl:117-129, CanonicalKmers.jl:146-174) and UnambiguousKmers with
the ASCII skipping table (UnambiguousKmers.jl:109-132, iterators/common.jl:22-32), through the C ABI,
bit-exact against the oracle and the reference's own examples."""
import json
import os

import numpy as np
import pytest

import kmertools as kt
from oracle import oracle as ko

pytestmark = pytest.mark.gpu

KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))
DIFF = KATS["differential_sequences"]
FW, FWRV, CANON, UNAMBIG = 0, 1, 2, 3


@pytest.fixture(scope="module")
def kc():
    import kmerscuda
    return kmerscuda


def rows(a):
    return [tuple(int(v) for v in r) for r in a]


def rand_ascii(rng, n, amb=0.0, lower=0.3, rna=False):
    base = rng.choice(list("ACGU" if rna else "ACGT"), size=n)
    if amb > 0:
        base = np.where(rng.random(n) < amb, rng.choice(list("MRSVWYHKDBN-"), size=n), base)
    s = "".join(base.tolist())
    low = rng.random(n) < lower
    return "".join(c.lower() if l else c for c, l in zip(s, low))


def test_reference_examples(kc):
    # CanonicalKmers.jl:190-197: CanonicalRNAMers{3}("AGCGA") -> AGC, CGC, CGA
    assert rows(kc.CanonicalRNAMers(3, "AGCGA").collect()) == [kt.kmer_limbs(x) for x in ("AGC", "CGC", "CGA")]
    # docs/src/iteration.md:51 and the canonical differential sequences given as strings / codeunits
    for s in DIFF["canonical"]["seqs"]:
        k = DIFF["canonical"]["k"]
        d = s.upper().replace("U", "T")
        A = kc.CanonicalRNAMers if "U" in s else kc.CanonicalDNAMers
        assert rows(A(k, s).collect()) == kt.naive_canonical(d, k)
        assert rows(A(k, s.encode()).collect()) == kt.naive_canonical(d, k)
        assert rows(A(k, np.frombuffer(s.lower().encode(), np.uint8)).collect()) == kt.naive_canonical(d, k)
    # StringViews test (test/runtests.jl:892-899): FwRvIterator{DNAAlphabet{2},9} over bytes
    s = DIFF["fwrv_k9"]["seqs"][0]
    fr = kc.FwRvDNAIterator(9, np.frombuffer(s.encode(), np.uint8)).collect()
    want = kt.naive_fwrv(s, 9)
    assert rows(fr[:, 0, :]) == [w[0] for w in want] and rows(fr[:, 1, :]) == [w[1] for w in want]
    # UnambiguousKmers over String sources, K = 4 (test/runtests.jl:812-821)
    for s in DIFF["unambiguous_4bit"]["seqs"] + DIFF["unambiguous_4bit_k4"]["seqs"]:
        A = kc.UnambiguousRNAMers if "U" in s else kc.UnambiguousDNAMers
        km, pos = A(4, s).collect()
        want = kt.naive_unambiguous(s.upper().replace("U", "T"), 4)
        assert rows(km) == [w[0] for w in want] and pos.tolist() == [w[1] for w in want]
    # bad byte in ASCII (test/runtests.jl:722-724, 844-846)
    for it in (kc.FwDNAMers, kc.CanonicalDNAMers, kc.UnambiguousDNAMers):
        with pytest.raises(kc.EncodeError) as ei:
            it(3, "TAGTCGTAGPATGC").collect()
        assert ei.value.symbol == "P" and ei.value.position == 10
    # strict: an ambiguity letter is an error for a 2-bit alphabet; U is not DNA, T is not RNA
    for it, s, sym, pos in ((kc.FwDNAMers, "ACGTNACGT", "N", 5), (kc.FwDNAMers, "ACGUACG", "U", 4),
                            (kc.FwRNAMers, "ACGUTACG", "T", 5)):
        with pytest.raises(kc.EncodeError) as ei:
            it(3, s).collect()
        assert (ei.value.symbol, ei.value.position) == (sym, pos)
    # shorter than K: strict touches nothing; UnambiguousKmers still reads (and rejects) every byte
    assert kc.FwDNAMers(5, "AP").collect().shape[0] == 0
    with pytest.raises(kc.EncodeError):
        kc.UnambiguousDNAMers(5, "AP").collect()
    assert kc.UnambiguousDNAMers(5, "ANNA").collect()[0].shape[0] == 0


@pytest.mark.parametrize("k", [1, 2, 5, 16, 31, 32, 33, 63, 64, 65, 97, 128])
def test_single_source_vs_oracle(kc, k):
    rng = np.random.default_rng(0xA5C11 + k)
    for rna in (False, True):
        for length in sorted({0, 1, k - 1, k, k + 1, k + 15, k + 31, k + 32, k + 33, 3 * k + 257, 5000}):
            s = rand_ascii(rng, length, rna=rna)
            A = kc.RNAAlphabet2 if rna else kc.DNAAlphabet2
            rs = kc.ReadSet.ascii(s)
            a, b, h = ko.ascii_iterate(s, k, ko.FWRV, rna=rna, want_hash=True)
            e = kc.extract(FWRV, rs, k, A=A, hash=True)
            assert np.array_equal(e.kmers, a) and np.array_equal(e.rv, b) and np.array_equal(e.hash, h)
            c, _, hc = ko.ascii_iterate(s, k, ko.CANON, rna=rna, want_hash=True)
            e = kc.extract(CANON, rs, k, A=A, hash=True)
            assert np.array_equal(e.kmers, c) and np.array_equal(e.hash, hc)
            for amb in (0.02, 0.3):
                sa = rand_ascii(rng, length, amb=amb, rna=rna)
                km, pos = ko.ascii_unambiguous(sa, k)
                e = kc.extract(UNAMBIG, kc.ReadSet.ascii(sa), k, A=A, hash=True)
                assert e.n == km.shape[0] and np.array_equal(e.kmers, km) and np.array_equal(e.index, pos)
                assert np.array_equal(e.hash, ko.fx_hash(km) if km.size else np.zeros(0, np.uint64))
                if length >= k and any(ch in "MRSVWYHKDBN-mrsvwyhkdbn" for ch in sa):
                    with pytest.raises(ko.AmbiguousError) as oi:
                        ko.ascii_iterate(sa, k, ko.FW, rna=rna)
                    with pytest.raises(kc.EncodeError) as ei:
                        kc.extract(FW, kc.ReadSet.ascii(sa), k, A=A)
                    assert ei.value.position == oi.value.pos and ei.value.symbol == chr(oi.value.enc)


def test_every_byte_value(kc):
    """All 256 byte values through both tables."""
    k = 2
    for b in range(256):
        s = bytes([65, 67, b, 71, 84, 65])
        for rna in (False, True):
            A = kc.RNAAlphabet2 if rna else kc.DNAAlphabet2
            src = s.replace(b"T", b"U") if rna else s
            try:
                want = ko.ascii_iterate(src, k, ko.FW, rna=rna)[0]
            except ko.AmbiguousError as err:
                with pytest.raises(kc.EncodeError) as ei:
                    kc.extract(FW, kc.ReadSet.ascii(src), k, A=A)
                assert ei.value.position == err.pos and ord(ei.value.symbol) == err.enc
            else:
                assert np.array_equal(kc.extract(FW, kc.ReadSet.ascii(src), k, A=A).kmers, want)
        try:
            km, pos = ko.ascii_unambiguous(s, k)
        except ko.AmbiguousError as err:
            with pytest.raises(kc.EncodeError) as ei:
                kc.extract(UNAMBIG, kc.ReadSet.ascii(s), k)
            assert ei.value.position == err.pos and ord(ei.value.symbol) == err.enc
        else:
            e = kc.extract(UNAMBIG, kc.ReadSet.ascii(s), k)
            assert np.array_equal(e.kmers, km) and np.array_equal(e.index, pos)


def test_string_read_set_and_views(kc):
    rng = np.random.default_rng(12)
    k = 31
    strs = [rand_ascii(rng, int(n), amb=0.01) for n in rng.integers(0, 400, size=3000)]
    rs = kc.ReadSet.from_strings(strs)
    kms, poss, offs = [], [], [0]
    for s in strs:
        km, pos = ko.ascii_unambiguous(s, k)
        kms.append(km)
        poss.append(pos)
        offs.append(offs[-1] + km.shape[0])
    e = kc.extract(UNAMBIG, rs, k, hash=True, want_seq_offsets=True)
    assert np.array_equal(e.kmers, np.concatenate(kms)) and np.array_equal(e.index, np.concatenate(poss))
    assert e.seq_out_offset.tolist() == offs
    e = kc.extract(UNAMBIG, rs, k, host_path=True, want_seq_offsets=True)
    assert np.array_equal(e.kmers, np.concatenate(kms)) and e.seq_out_offset.tolist() == offs
    clean = [rand_ascii(rng, int(n)) for n in rng.integers(0, 400, size=3000)]
    rs = kc.ReadSet.from_strings(clean)
    parts = [ko.ascii_iterate(s, k, ko.CANON, want_hash=True) for s in clean]
    for host_path in (False, True):
        e = kc.extract(CANON, rs, k, hash=True, host_path=host_path)
        assert np.array_equal(e.kmers, np.concatenate([p[0] for p in parts]))
        assert np.array_equal(e.hash, np.concatenate([p[2] for p in parts]))
    # an error in read 1234 is reported with its read index
    bad = list(clean)
    while len(bad[1234]) < k:
        bad[1234] += "ACGT" * 10
    bad[1234] = bad[1234][:7] + "!" + bad[1234][8:]
    for mode in (FW, UNAMBIG):
        for host_path in (False, True):
            with pytest.raises(kc.EncodeError) as ei:
                kc.extract(mode, kc.ReadSet.from_strings(bad), k, host_path=host_path)
            assert (ei.value.seq_index, ei.value.position, ei.value.symbol) == (1234, 8, "!")
    # SubString-style views: unaligned start inside a larger buffer
    big = rand_ascii(rng, 3000, amb=0.01)
    for first in (1, 3, 17, 31, 33):
        sub = big[first:first + 1500]
        km, pos = ko.ascii_unambiguous(sub, k)
        e = kc.extract(UNAMBIG, kc.ReadSet.ascii(big, first_symbol_offset=first, length=1500), k)
        assert np.array_equal(e.kmers, km) and np.array_equal(e.index, pos)


def test_host_path_long_ascii_sequence(kc):
    rng = np.random.default_rng(3)
    n, k = 9_000_017, 31
    codes = rng.integers(0, 4, size=n)
    arr = np.frombuffer(b"ACGT", np.uint8)[codes].copy()
    arr[rng.random(n) < 0.01] = ord("N")
    arr[rng.random(n) < 0.2] |= 0x20  # lowercase
    km, pos = ko.ascii_unambiguous(arr, k)
    e = kc.extract(UNAMBIG, kc.ReadSet.ascii(arr), k, hash=True, host_path=True)
    assert e.n == km.shape[0] and np.array_equal(e.kmers, km) and np.array_equal(e.index, pos)
    assert np.array_equal(e.hash, ko.fx_hash(km))
    clean = np.frombuffer(b"ACGT", np.uint8)[codes].copy()
    a, _, h = ko.ascii_iterate(clean, k, ko.CANON, want_hash=True)
    for host_path in (False, True):
        e = kc.extract(CANON, kc.ReadSet.ascii(clean), k, hash=True, host_path=host_path)
        assert np.array_equal(e.kmers, a) and np.array_equal(e.hash, h)
    clean[8_765_432] = ord("x")
    with pytest.raises(kc.EncodeError) as ei:
        kc.extract(CANON, kc.ReadSet.ascii(clean), k, host_path=True)
    assert ei.value.position == 8_765_433 and ei.value.symbol == "x"
    with pytest.raises(kc.EncodeError) as ei:
        kc.extract(UNAMBIG, kc.ReadSet.ascii(clean), k, host_path=True)
    assert ei.value.position == 8_765_433 and ei.value.symbol == "x"
