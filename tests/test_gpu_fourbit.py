"""GPU parity for 4-bit sources (the FourToTwo recoding scheme): strict FwKmers / FwRvIterator /
CanonicalKmers (an uncertain symbol is an EncodeError, /root/reference/src/iterators/FwKmers.jl:104-115,
CanonicalKmers.jl:131-144) and UnambiguousKmers (skip / restart, UnambiguousKmers.jl:134-148),
through the C ABI, bit-exact against the oracle and against the reference's own examples.
"""
import ctypes as C
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


@pytest.fixture(scope="module")
def ctx(kc):
    return kc.default_context()


def dna(s):
    return s.upper().replace("U", "T")


def rows(a):
    return [tuple(int(v) for v in r) for r in a]


def random_codes4(rng, n, ambiguous):
    """4-bit encodings: one-hot bases, with probability `ambiguous` any non-one-hot nibble
    (IUPAC ambiguity codes, N = 15, gap = 0)."""
    codes = np.uint64(1) << rng.integers(0, 4, size=n).astype(np.uint64)
    if ambiguous > 0:
        amb = rng.choice(np.array([0, 3, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15], dtype=np.uint64), size=n)
        codes = np.where(rng.random(n) < ambiguous, amb, codes)
    return codes.astype(np.uint64)


# ----------------------------------------------------------------------------- golden vectors
def test_kat_unambiguous_docstring(kc):
    for e in KATS["unambiguous"] + KATS["unambiguous_starts"]:
        s = dna(e["seq"])
        km, pos = kc.UnambiguousDNAMers(e["k"], kc.LongDNA4(s)).collect()
        if "expect" in e:
            assert rows(km) == [kt.kmer_limbs(dna(x[0])) for x in e["expect"]]
            assert pos.tolist() == [x[1] for x in e["expect"]]
        else:
            assert pos.tolist() == e["starts"]
            assert rows(km) == [kt.kmer_limbs(s[p - 1:p - 1 + e["k"]]) for p in e["starts"]]


@pytest.mark.parametrize("key", ["unambiguous_4bit", "unambiguous_4bit_k4"])
def test_reference_differential_unambiguous(kc, key):
    e = DIFF[key]
    for s in map(dna, e["seqs"]):
        for k in (e["k"], 1, 2, 5):
            want = kt.naive_unambiguous(s, k)
            km, pos = kc.UnambiguousDNAMers(k, kc.LongDNA4(s)).collect()
            assert rows(km) == [w[0] for w in want] and pos.tolist() == [w[1] for w in want]


@pytest.mark.parametrize("key", ["fw_4bit", "fw_4to2", "fwrv", "canonical"])
def test_reference_differential_strict(kc, key):
    e = DIFF[key]
    k = e["k"]
    for s in map(dna, e["seqs"]):
        seq = kc.LongDNA4(s)
        assert rows(kc.FwDNAMers(k, seq).collect()) == kt.naive_fw(s, k)
        fr = kc.FwRvDNAIterator(k, seq).collect()
        want = kt.naive_fwrv(s, k)
        assert rows(fr[:, 0, :]) == [w[0] for w in want] and rows(fr[:, 1, :]) == [w[1] for w in want]
        assert rows(kc.CanonicalDNAMers(k, seq).collect()) == kt.naive_canonical(s, k)


def test_kat_strict_errors(kc):
    """The reference throws EncodeError at the first uncertain symbol the iterator reaches."""
    for e in KATS["strict_4to2_errors"]:
        s = dna(e["seq"])
        for it in (kc.FwDNAMers, kc.CanonicalDNAMers, kc.FwRvDNAIterator):
            with pytest.raises(kc.EncodeError) as ei:
                it(e["k"], kc.LongDNA4(s)).collect()
            assert ei.value.position == e["pos"], e
            assert ei.value.symbol == dna(e["bad_symbol"]), e
    # shorter than K: nothing is touched, nothing is thrown (FwKmers.jl:62-66)
    assert kc.FwDNAMers(5, kc.LongDNA4("ANNA")).collect().shape[0] == 0


# ------------------------------------------------------------------- randomised, vs the oracle
KS = [1, 2, 5, 15, 16, 17, 31, 32, 33, 48, 63, 64, 65, 96, 97, 127, 128]


@pytest.mark.parametrize("k", KS)
def test_single_sequence_unambiguous(kc, k):
    rng = np.random.default_rng(0x4B17 + k)
    for amb in (0.0, 0.01, 0.2):
        for length in sorted({0, 1, k - 1, k, k + 1, k + 15, k + 16, k + 17, 3 * k + 257, 4099}):
            codes = random_codes4(rng, length, amb)
            w = kt.pack_codes(codes, 4) if length else np.zeros(0, np.uint64)
            seq = kc.LongSequence(kc.DNAAlphabet4, w, length)
            rs = kc.ReadSet.single(seq)
            km, pos = ko.unambiguous(w, length, k, src_bits=4)
            e = kc.extract(UNAMBIG, rs, k, hash=True)
            assert e.n == km.shape[0], (k, amb, length)
            assert np.array_equal(e.kmers, km) and np.array_equal(e.index, pos)
            assert np.array_equal(e.hash, ko.fx_hash(km) if km.size else np.zeros(0, np.uint64))
            e = kc.extract(UNAMBIG, rs, k, aos=True)
            N = kc.n_limbs(k)
            assert np.array_equal(e.kmers[:, :N], km) and np.array_equal(e.kmers[:, N].astype(np.int64), pos)
            n = C.c_uint64()
            ctx = kc.default_context()
            drs = kc.DeviceReadSet(ctx, rs)
            assert ctx.lib.kmc_count(ctx.handle, C.byref(drs.desc), k, UNAMBIG, C.byref(n)) == 0
            assert n.value == km.shape[0]


@pytest.mark.parametrize("k", [1, 5, 31, 32, 33, 63, 64, 97, 128])
def test_single_sequence_strict(kc, k):
    rng = np.random.default_rng(0x57 + k)
    for length in sorted({0, k - 1, k, k + 1, k + 16, 3 * k + 257, 2500}):
        codes = random_codes4(rng, length, 0.0)
        w = kt.pack_codes(codes, 4) if length else np.zeros(0, np.uint64)
        rs = kc.ReadSet.single(kc.LongSequence(kc.DNAAlphabet4, w, length))
        a, b, h = ko.iterate(w, length, k, ko.FWRV, src_bits=4, want_hash=True)
        e = kc.extract(FWRV, rs, k, hash=True)
        assert np.array_equal(e.kmers, a) and np.array_equal(e.rv, b) and np.array_equal(e.hash, h)
        c, _, hc = ko.iterate(w, length, k, ko.CANON, src_bits=4, want_hash=True)
        e = kc.extract(CANON, rs, k, hash=True)
        assert np.array_equal(e.kmers, c) and np.array_equal(e.hash, hc)
        if length >= k:
            # one uncertain symbol anywhere: the error names exactly that symbol
            for p in sorted({0, k - 1, min(k, length - 1), length // 2, length - 1}):
                bad = codes.copy()
                bad[p] = [15, 0, 9, 6][p % 4]
                if p + 7 < length:
                    bad[p + 7] = 15  # a later one must not be reported
                wb = kt.pack_codes(bad, 4)
                rsb = kc.ReadSet.single(kc.LongSequence(kc.DNAAlphabet4, wb, length))
                with pytest.raises(ko.AmbiguousError) as oi:
                    ko.iterate(wb, length, k, ko.FW, src_bits=4)
                for mode in (FW, CANON):
                    with pytest.raises(kc.EncodeError) as ei:
                        kc.extract(mode, rsb, k)
                    assert ei.value.position == p + 1 == oi.value.pos
                    assert ei.value.symbol == "-ACMGRSVTWYHKDBN"[int(bad[p])]


@pytest.mark.parametrize("k,amb", [(1, 0.5), (3, 0.2), (31, 0.03), (64, 0.01)])
def test_dense_ambiguity_many_runs(kc, k, amb):
    """Lots of short runs (more runs per tile than the kernel stages in shared memory)."""
    rng = np.random.default_rng(31 * k)
    n = 300_000
    codes = random_codes4(rng, n, amb)
    w = kt.pack_codes(codes, 4)
    rs = kc.ReadSet.single(kc.LongSequence(kc.DNAAlphabet4, w, n))
    km, pos = ko.unambiguous(w, n, k, src_bits=4)
    for aos in (False, True):
        e = kc.extract(UNAMBIG, rs, k, hash=True, aos=aos)
        N = kc.n_limbs(k)
        assert e.n == km.shape[0]
        if aos:
            assert np.array_equal(e.kmers[:, :N], km) and np.array_equal(e.kmers[:, N].astype(np.int64), pos)
        else:
            assert np.array_equal(e.kmers, km) and np.array_equal(e.index, pos)
        assert np.array_equal(e.hash, ko.fx_hash(km))
    e = kc.extract(UNAMBIG, rs, k, host_path=True)
    assert np.array_equal(e.kmers, km) and np.array_equal(e.index, pos)


def test_unambiguous_capacity_is_respected(kc, ctx):
    """The compaction finds its output positions on the device; a buffer that is too small must fail
    with KMC_E_OUT_TOO_SMALL and nothing may be written past its end (unaligned base on purpose)."""
    from kmerscuda import _abi
    rng = np.random.default_rng(77)
    n, k = 200_000, 31
    codes = random_codes4(rng, n, 0.01)
    w = kt.pack_codes(codes, 4)
    rs = kc.ReadSet.single(kc.LongSequence(kc.DNAAlphabet4, w, n))
    km, pos = ko.unambiguous(w, n, k, src_bits=4)
    total = km.shape[0]
    drs = kc.DeviceReadSet(ctx, rs)
    cap = total - 1000
    guard = 4096
    sentinel = np.full(cap + guard + 1, 0xDEADBEEFDEADBEEF, dtype=np.uint64)
    da, di = ctx.to_device(sentinel), ctx.to_device(sentinel)
    res = _abi.kmc_result()
    out = _abi.kmc_out(da.ptr + 8, None, None, di.ptr + 8, None, cap, 0)
    st = ctx.lib.kmc_extract(ctx.handle, C.byref(drs.desc), k, UNAMBIG, 0, C.byref(out), C.byref(res))
    assert st == _abi.KMC_E_OUT_TOO_SMALL and res.n_written == 0
    assert np.all(da.download(np.uint64, guard, 8 * (cap + 1)) == 0xDEADBEEFDEADBEEF)
    assert np.all(di.download(np.uint64, guard, 8 * (cap + 1)) == 0xDEADBEEFDEADBEEF)
    assert da.download(np.uint64, 1)[0] == 0xDEADBEEFDEADBEEF
    # exactly large enough: everything arrives, in order
    out = _abi.kmc_out(da.ptr + 8, None, None, di.ptr + 8, None, total, 0)
    da2, di2 = ctx.alloc(8 * (total + 1)), ctx.alloc(8 * (total + 1))
    out = _abi.kmc_out(da2.ptr + 8, None, None, di2.ptr + 8, None, total, 0)
    st = ctx.lib.kmc_extract(ctx.handle, C.byref(drs.desc), k, UNAMBIG, 0, C.byref(out), C.byref(res))
    assert st == 0 and res.n_written == total
    assert np.array_equal(da2.download(np.uint64, total, 8), km[:, 0])
    assert np.array_equal(di2.download(np.int64, total, 8), pos)


def make_ragged4(rng, lens, amb):
    codes = [random_codes4(rng, n, amb) for n in lens]
    packed = [kt.pack_codes(c, 4) if len(c) else np.zeros(0, np.uint64) for c in codes]
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(p) for p in packed])
    words = np.concatenate([p for p in packed if len(p)] + [np.zeros(1, np.uint64)])
    return words, off[:-1].copy(), np.array(lens, dtype=np.uint64)


@pytest.mark.parametrize("k", [1, 7, 31, 32, 33, 63, 64, 127])
def test_ragged_read_set_4bit(kc, k):
    rng = np.random.default_rng(400 + k)
    lens = [0, 1, k - 1, k, k + 1, k + 2, 150, 151, 0, 0, k, 40, 999, k + 5] + rng.integers(0, 400, size=300).tolist()
    lens = [max(0, int(x)) for x in lens]
    for amb in (0.0, 0.02):
        words, off, ln = make_ragged4(rng, lens, amb)
        rs = kc.ReadSet(4, words, len(lens), seq_word_offset=off, seq_len=ln)
        km, pos, out_off = ko.batch_unambiguous(words, len(lens), k, word_off=off, seq_len=ln, src_bits=4)
        e = kc.extract(UNAMBIG, rs, k, hash=True, want_seq_offsets=True)
        assert e.n == km.shape[0]
        assert np.array_equal(e.kmers, km) and np.array_equal(e.index, pos)
        assert np.array_equal(e.seq_out_offset, out_off)
        assert np.array_equal(e.hash, ko.fx_hash(km))
        if amb == 0.0:
            a, _, h, woff = ko.batch_iterate(words, len(lens), k, ko.CANON, word_off=off, seq_len=ln, src_bits=4,
                                             want_hash=True)
            e = kc.extract(CANON, rs, k, hash=True, want_seq_offsets=True)
            assert np.array_equal(e.kmers, a) and np.array_equal(e.hash, h) and np.array_equal(e.seq_out_offset, woff)
        else:
            with pytest.raises(ko.AmbiguousError) as oi:
                ko.batch_iterate(words, len(lens), k, ko.FW, word_off=off, seq_len=ln, src_bits=4, threads=1)
            with pytest.raises(kc.EncodeError) as ei:
                kc.extract(FW, rs, k)
            assert (ei.value.seq_index, ei.value.position) == (oi.value.seq, oi.value.pos)


@pytest.mark.parametrize("k,length", [(31, 150), (31, 151), (21, 100), (63, 150), (5, 36), (31, 31), (31, 30), (97, 300)])
def test_uniform_read_set_4bit(kc, k, length):
    rng = np.random.default_rng(k * 1000 + length)
    n_reads = 3000
    stride = (length + 15) // 16 + (1 if length % 7 == 0 else 0)
    codes = np.ones((n_reads, stride * 16), dtype=np.uint64)
    codes[:, :length] = random_codes4(rng, n_reads * length, 0.01).reshape(n_reads, length)
    codes[:, length:] = 0
    words = kt.pack_codes(codes.reshape(-1), 4)
    rs = kc.ReadSet(4, words, n_reads, uniform_len=length, uniform_stride_words=stride)
    km, pos, out_off = ko.batch_unambiguous(words, n_reads, k, uniform_len=length, uniform_stride=stride, src_bits=4)
    for aos in (False, True):
        e = kc.extract(UNAMBIG, rs, k, hash=True, aos=aos, want_seq_offsets=True)
        N = kc.n_limbs(k)
        assert e.n == km.shape[0]
        if aos:
            assert np.array_equal(e.kmers[:, :N], km) and np.array_equal(e.kmers[:, N].astype(np.int64), pos)
        else:
            assert np.array_equal(e.kmers, km) and np.array_equal(e.index, pos)
        assert np.array_equal(e.seq_out_offset, out_off)
        assert np.array_equal(e.hash, ko.fx_hash(km) if km.size else np.zeros(0, np.uint64))


def test_subsequence_views_4bit(kc):
    rng = np.random.default_rng(78)
    codes = random_codes4(rng, 900, 0.02)
    w = kt.pack_codes(codes, 4)
    for first in (1, 7, 15, 16, 17, 33):
        for k in (5, 31, 33, 64):
            n = 500
            rs = kc.ReadSet(4, w, 1, uniform_len=n, uniform_stride_words=w.size, first_symbol_offset=first)
            km, pos = ko.unambiguous(w, n, k, src_bits=4, first=first)
            e = kc.extract(UNAMBIG, rs, k)
            assert np.array_equal(e.kmers, km) and np.array_equal(e.index, pos)


# --------------------------------------------------------------------- pipelined host path
def synth_4bit(n, seed=439824):
    """per base u = splitmix64(seed + i); N with probability 1 %, else uniform ACGT."""
    u = kt.splitmix64(np.arange(n, dtype=np.uint64) + np.uint64(seed))
    return np.where(u % np.uint64(100) == 0, np.uint64(15), np.uint64(1) << ((u >> np.uint64(32)) & np.uint64(3)))


@pytest.mark.parametrize("k", [31, 63])
def test_host_path_single_sequence_4bit(kc, k):
    n = 9_000_011  # > 4 Mi windows: several chunks, each with its own count pass
    w = kt.pack_codes(synth_4bit(n), 4)
    rs = kc.ReadSet.single(kc.LongSequence(kc.DNAAlphabet4, w, n))
    km, pos = ko.unambiguous(w, n, k, src_bits=4)
    e = kc.extract(UNAMBIG, rs, k, hash=True, host_path=True, want_seq_offsets=True)
    assert e.n == km.shape[0]
    assert np.array_equal(e.kmers, km) and np.array_equal(e.index, pos) and np.array_equal(e.hash, ko.fx_hash(km))
    assert e.seq_out_offset.tolist() == [0, km.shape[0]]
    e = kc.extract(UNAMBIG, rs, k, aos=True, host_path=True)
    N = kc.n_limbs(k)
    assert np.array_equal(e.kmers[:, :N], km) and np.array_equal(e.kmers[:, N].astype(np.int64), pos)
    e2 = kc.extract(UNAMBIG, rs, k, hash=True)  # device path, one launch
    assert np.array_equal(e2.kmers, km) and np.array_equal(e2.index, pos)
    # strict over the same sequence: the first N is reported, wherever the chunk boundary lies
    with pytest.raises(ko.AmbiguousError) as oi:
        ko.iterate(w, n, k, ko.FW, src_bits=4)
    for host_path in (False, True):
        with pytest.raises(kc.EncodeError) as ei:
            kc.extract(CANON, rs, k, host_path=host_path)
        assert ei.value.position == oi.value.pos and ei.value.symbol == "N"


def test_host_path_strict_error_in_late_chunk(kc):
    k, n = 31, 9_000_000
    codes = np.uint64(1) << (kt.splitmix64(np.arange(n, dtype=np.uint64)) & np.uint64(3))
    p = 8_500_123
    codes[p] = 10  # Y
    w = kt.pack_codes(codes, 4)
    rs = kc.ReadSet.single(kc.LongSequence(kc.DNAAlphabet4, w, n))
    with pytest.raises(kc.EncodeError) as ei:
        kc.extract(FW, rs, k, host_path=True)
    assert ei.value.position == p + 1 and ei.value.symbol == "Y"
    codes[p] = 4
    w = kt.pack_codes(codes, 4)
    rs = kc.ReadSet.single(kc.LongSequence(kc.DNAAlphabet4, w, n))
    a, _, h = ko.iterate(w, n, k, ko.CANON, src_bits=4, want_hash=True)
    e = kc.extract(CANON, rs, k, hash=True, host_path=True)
    assert np.array_equal(e.kmers, a) and np.array_equal(e.hash, h)


def test_host_path_uniform_reads_4bit(kc):
    k, length, stride, n_reads = 31, 150, 10, 90_000  # 10.8 M windows -> 3 chunks
    codes = np.zeros((n_reads, stride * 16), dtype=np.uint64)
    codes[:, :length] = synth_4bit(n_reads * length).reshape(n_reads, length)
    words = kt.pack_codes(codes.reshape(-1), 4)
    rs = kc.ReadSet(4, words, n_reads, uniform_len=length, uniform_stride_words=stride)
    km, pos, out_off = ko.batch_unambiguous(words, n_reads, k, uniform_len=length, uniform_stride=stride, src_bits=4)
    e = kc.extract(UNAMBIG, rs, k, hash=True, host_path=True, want_seq_offsets=True)
    assert e.n == km.shape[0]
    assert np.array_equal(e.kmers, km) and np.array_equal(e.index, pos) and np.array_equal(e.hash, ko.fx_hash(km))
    assert np.array_equal(e.seq_out_offset, out_off)


def test_host_path_ragged_reads_4bit(kc):
    rng = np.random.default_rng(22)
    k = 31
    lens = rng.integers(0, 400, size=60_000).tolist()
    words, off, ln = make_ragged4(rng, lens, 0.01)
    rs = kc.ReadSet(4, words, len(lens), seq_word_offset=off, seq_len=ln)
    km, pos, out_off = ko.batch_unambiguous(words, len(lens), k, word_off=off, seq_len=ln, src_bits=4)
    e = kc.extract(UNAMBIG, rs, k, hash=True, host_path=True, want_seq_offsets=True)
    assert e.n == km.shape[0] and np.array_equal(e.kmers, km) and np.array_equal(e.index, pos)
    assert np.array_equal(e.seq_out_offset, out_off)


# ----------------------------------------------- full-size properties (BASELINE config C3 shape)
def test_full_size_properties_c3(kc, ctx):
    """UnambiguousDNAMers{31} over 10 M x 150 bp 4-bit reads with 1 % N, outputs resident on the
    device.  Checks: kmc_count == n_written == last per-read offset; per-read offsets are
    monotone; the fused hash equals kmc_fx_hash of the emitted k-mers; every emitted k-mer equals
    the strict FwKmers k-mer at its index (device-side gather); sampled reads equal the oracle."""
    import torch
    from kmerscuda import _abi
    free, _ = torch.cuda.mem_get_info()
    n_reads = 10_000_000 if free > 80e9 else 500_000
    k, length, stride = 31, 150, 10
    wpr = length - k + 1
    g = torch.Generator(device="cuda").manual_seed(439824)
    base = torch.randint(0, 4, (n_reads, stride * 16), dtype=torch.int64, device="cuda", generator=g)
    nib = torch.ones_like(base) << base
    nib[torch.rand(nib.shape, device="cuda", generator=g) < 0.01] = 15
    nib[:, length:] = 0
    shifts = (torch.arange(16, device="cuda", dtype=torch.int64) * 4)
    words = (nib.view(n_reads, stride, 16) << shifts).sum(dim=2).reshape(-1).contiguous()  # disjoint bits: sum == or
    del base, nib
    cap = n_reads * wpr
    km = torch.empty(cap, dtype=torch.int64, device="cuda")
    idx = torch.empty(cap, dtype=torch.int64, device="cuda")
    hsh = torch.empty(cap, dtype=torch.int64, device="cuda")
    soff = torch.empty(n_reads + 1, dtype=torch.int64, device="cuda")
    desc = _abi.kmc_seqs(words.data_ptr(), words.numel(), n_reads, None, None, length, stride, 4, 0)
    out = _abi.kmc_out(km.data_ptr(), None, hsh.data_ptr(), idx.data_ptr(), soff.data_ptr(), cap, 0)
    res = _abi.kmc_result()
    torch.cuda.synchronize()
    n = C.c_uint64()
    ctx._check(ctx.lib.kmc_count(ctx.handle, C.byref(desc), k, UNAMBIG, C.byref(n)))
    ctx._check(ctx.lib.kmc_extract(ctx.handle, C.byref(desc), k, UNAMBIG, _abi.KMC_HASH_FX, C.byref(out), C.byref(res)))
    total = int(res.n_written)
    assert total == n.value and 0.6 * cap < total < 0.85 * cap  # ~0.99^31 = 73 % of the windows survive
    assert int(soff[-1]) == total and int(soff[0]) == 0
    assert bool((soff[1:] >= soff[:-1]).all())
    h2 = torch.empty(total, dtype=torch.int64, device="cuda")
    ctx._check(ctx.lib.kmc_fx_hash(ctx.handle, km.data_ptr(), total, 1, 0, h2.data_ptr()))
    ctx.sync()
    assert torch.equal(hsh[:total], h2)
    del h2
    # indices are 1-based window starts, strictly increasing inside a read
    per_read = (soff[1:] - soff[:-1])
    assert int(per_read.max()) <= wpr
    assert int(idx[:total].min()) >= 1 and int(idx[:total].max()) <= wpr
    # sampled reads against the oracle
    rng = np.random.default_rng(3)
    soff_h = soff.cpu().numpy()
    for r in np.sort(rng.choice(n_reads, size=1500, replace=False)):
        w = words[r * stride:(r + 1) * stride].cpu().numpy().view(np.uint64)
        okm, opos = ko.unambiguous(w, length, k, src_bits=4)
        lo, hi = int(soff_h[r]), int(soff_h[r + 1])
        assert hi - lo == okm.shape[0]
        assert np.array_equal(km[lo:hi].cpu().numpy().view(np.uint64), okm[:, 0])
        assert np.array_equal(idx[lo:hi].cpu().numpy(), opos)
