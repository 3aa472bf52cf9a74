"""CPU check of the device primitives themselves: kmers.jl_b200/csrc/kmer_core.cuh (block load and alignment,
block_kmers, limbs_less, fx_hash) compiled for the host with portable definitions of the CUDA intrinsics
(tests/host_core/core_host.cpp) and compared, work item by work item, with the oracle -- the same closed form the
kernels run, for every (limbs, block width) class of the launcher tables and both k-mer alphabets.  Test
infrastructure only: the library has no CPU path and nothing in it links this code."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import kmertools as kt
from oracle import oracle as ko

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def core(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    so = str(tmp_path_factory.mktemp("core_host") / "libcore_host.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas", "-o", so,
                    os.path.join(ROOT, "tests", "host_core", "core_host.cpp")], check=True)
    lib = C.CDLL(so)
    lib.core_item_windows.restype = C.c_int
    lib.core_item_windows.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.POINTER(C.c_int)]
    lib.core_recode_word.restype = None
    lib.core_recode_word.argtypes = [C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.core_valid_start_word.restype = C.c_uint32
    lib.core_valid_start_word.argtypes = [C.c_void_p, C.c_int]
    lib.core_onehot8.restype = C.c_uint32
    lib.core_onehot8.argtypes = [C.c_uint32]
    lib.core_ascii_luts.restype = None
    lib.core_ascii_luts.argtypes = [C.c_void_p]
    for f in (lib.core_base_hash, lib.core_fx_hash):
        f.restype = C.c_uint64
        f.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
    return lib


def item(core, words, bit, k, bps):
    """(G, fw[G][N], rv[G][N], canon[G][N], hash[G]) of the work item starting at stream bit `bit`."""
    w32 = np.ascontiguousarray(words).view(np.uint32)
    fw, rv, ca = (np.zeros(8 * 4, dtype=np.uint64) for _ in range(3))
    h = np.zeros(8, dtype=np.uint64)
    n = C.c_int(0)
    g = core.core_item_windows(w32.ctypes.data, w32.size, bit, k, bps, fw.ctypes.data, rv.ctypes.data, ca.ctypes.data,
                               h.ctypes.data, C.byref(n))
    assert g > 0, f"no instantiation for K={k}, bps={bps}"
    N = n.value
    return g, fw[:g * N].reshape(g, N), rv[:g * N].reshape(g, N), ca[:g * N].reshape(g, N), h[:g]


def starts(rng, n_windows):
    fixed = [0, 1, 2, 7, 15, 16, 17, 31, 32, 33, 63, 64, 65]
    return sorted({p for p in fixed if p < n_windows} | {int(p) for p in rng.integers(0, max(n_windows, 1), size=24)})


@pytest.mark.parametrize("k", [1, 2, 5, 15, 16, 17, 29, 31, 32, 33, 47, 48, 49, 63, 64, 65, 80, 95, 96, 97, 112, 127, 128])
def test_two_bit_items_match_the_oracle(core, k):
    rng = np.random.default_rng(1000 + k)
    n = 700
    words = rng.integers(0, 2**64, size=(n + 31) // 32 + 1, dtype=np.uint64)
    fw, rv, _ = ko.iterate(words, n, k, ko.FWRV)
    ca, _, h = ko.iterate(words, n, k, ko.CANON, want_hash=True)
    nwin = n - k + 1
    for p in starts(rng, nwin):
        g, f, r, c, hh = item(core, words, 2 * p, k, 2)
        m = min(g, nwin - p)
        assert np.array_equal(f[:m], fw[p:p + m]) and np.array_equal(r[:m], rv[p:p + m]), (k, p)
        assert np.array_equal(c[:m], ca[p:p + m]) and np.array_equal(hh[:m], h[p:p + m]), (k, p)


def test_aligned_locator_divides_exactly(core):
    """extract_aligned_kernel finds the read of a work item with a multiply-high by floor(2^32 / gprm) and one correction step
    (kmer_core.cuh: AlignedLocator).  Against divmod for group counts from 1 to 2^31 - 1 -- powers of two and their
    neighbours, primes, C2's 15 -- and items up to 2^32 - 1, incl. the multiples of gprm and their neighbours."""
    core.core_aligned_bit.restype = C.c_uint64
    core.core_aligned_bit.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
    rng = np.random.default_rng(77)
    gprms = {1, 2, 3, 5, 7, 15, 16, 17, 120, 255, 256, 257, 1000003, 2**31 - 1, 2**30, 2**30 + 1, 3 * 2**29, 12345678, 250_000_000}
    gprms |= {int(x) for x in rng.integers(1, 2**31, size=40)} | {2**b + d for b in range(1, 31) for d in (-1, 0, 1) if 2**b + d >= 1}
    gi = C.c_uint32(0)
    for gprm in sorted(gprms):
        items = {0, 1, gprm - 1, gprm, gprm + 1, 2**32 - 1, 2**32 - 2, 2**31, 2**31 - 1}
        top = (2**32 - 1) // gprm
        for q in {1, 2, top, top - 1, max(top // 2, 1), int(rng.integers(0, top + 1))}:
            items |= {q * gprm - 1, q * gprm, q * gprm + 1}
        items |= {int(x) for x in rng.integers(0, 2**32, size=50)}
        read_bits, first = int(rng.integers(1, 2**32)), int(rng.integers(0, 2**20))
        for item in items:
            if not 0 <= item < 2**32:
                continue
            r, g = divmod(item, gprm)
            if g * 16 >= 2**32:
                continue  # (the launcher only takes sets whose windows per read fit the 32-bit arithmetic)
            bit = core.core_aligned_bit(item, gprm, read_bits, first, C.byref(gi))
            assert gi.value == g, (gprm, item)
            assert bit == (r * read_bits + 2 * first + g * 16) % 2**64, (gprm, item)


@pytest.mark.parametrize("k", list(range(1, 33)))
def test_two_window_items_of_the_tuple_kernels_match_the_oracle(core, k):
    """The Julia tuple layouts run the lean kernel with groups of two windows (geometry(k, 2, 2)): a block shape of its own
    (NX = 1, 2, 3; the half-width head mask for K >= 31), every K of a one-limb k-mer, every alignment of the block."""
    core.core_item_windows_g2.restype = C.c_int
    core.core_item_windows_g2.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(3000 + k)
    n = 300
    words = rng.integers(0, 2**64, size=(n + 31) // 32 + 1, dtype=np.uint64)
    w32 = np.ascontiguousarray(words).view(np.uint32)
    fw, rv, _ = ko.iterate(words, n, k, ko.FWRV)
    nwin = n - k + 1
    f, r = np.zeros(2, dtype=np.uint64), np.zeros(2, dtype=np.uint64)
    for p in range(0, min(nwin - 1, 130)):
        assert core.core_item_windows_g2(w32.ctypes.data, w32.size, 2 * p, k, f.ctypes.data, r.ctypes.data) == 2
        assert np.array_equal(f, fw[p:p + 2, 0]) and np.array_equal(r, rv[p:p + 2, 0]), (k, p)


@pytest.mark.parametrize("k", [1, 3, 8, 15, 16, 17, 24, 31, 32, 33, 40, 47, 48, 49, 63, 64])
def test_four_bit_alphabet_items_match_the_oracle(core, k):
    """Kmer{DNAAlphabet{4}}: rev4 / comp4 (any IUPAC symbol, N and gap are legal symbols of the k-mer)."""
    rng = np.random.default_rng(2000 + k)
    n = 400
    codes = rng.integers(0, 16, size=n).astype(np.uint64)
    words = np.concatenate([kt.pack_codes(codes, 4), np.zeros(1, dtype=np.uint64)])
    fw, rv, _ = ko.iterate4(words, n, k, ko.FWRV)
    ca, _, h = ko.iterate4(words, n, k, ko.CANON, want_hash=True)
    nwin = n - k + 1
    for p in starts(rng, nwin):
        g, f, r, c, hh = item(core, words, 4 * p, k, 4)
        m = min(g, nwin - p)
        assert np.array_equal(f[:m], fw[p:p + m]) and np.array_equal(r[:m], rv[p:p + m]), (k, p)
        assert np.array_equal(c[:m], ca[p:p + m]) and np.array_equal(hh[:m], h[p:p + m]), (k, p)


def test_loads_near_the_ends_of_the_buffer_are_clamped(core):
    """Items whose block reaches past the last word (or starts before the first) read clamped words; the windows
    that exist are still exact."""
    rng = np.random.default_rng(5)
    k, n = 31, 64
    words = rng.integers(0, 2**64, size=2, dtype=np.uint64)  # exactly the 64 symbols, no slack word
    fw, rv, _ = ko.iterate(words, n, k, ko.FWRV)
    nwin = n - k + 1
    for p in range(nwin - 9, nwin):
        g, f, r, _, _ = item(core, words, 2 * p, k, 2)
        m = min(g, nwin - p)
        assert np.array_equal(f[:m], fw[p:p + m]) and np.array_equal(r[:m], rv[p:p + m]), p


def test_recode_word_is_trailing_zeros_and_the_uncertainty_flag(core):
    """FourToTwo (construction_utils.jl:41-54): code = trailing_zeros(enc) for a one-hot nibble; flag <=>
    count_ones(enc) != 1 (IUPAC sets, N = 15, gap = 0)."""
    rng = np.random.default_rng(9)
    words = [0, 2**64 - 1, 0x8421842184218421, 0x1248124812481248] + [int(x) for x in rng.integers(0, 2**64, size=2000, dtype=np.uint64)]
    # half of the random words: only certain symbols, as real reads mostly are
    words += [int(sum((1 << int(c)) << (4 * i) for i, c in enumerate(rng.integers(0, 4, size=16)))) for _ in range(500)]
    for w in words:
        codes, flags = C.c_uint32(0), C.c_uint32(0)
        core.core_recode_word(w, C.byref(codes), C.byref(flags))
        for i in range(16):
            enc = (w >> (4 * i)) & 15
            certain = bin(enc).count("1") == 1
            assert ((flags.value >> i) & 1) == (0 if certain else 1), (hex(w), i)
            if certain:
                assert ((codes.value >> (2 * i)) & 3) == enc.bit_length() - 1, (hex(w), i)
        assert flags.value >> 16 == 0


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 34, 47, 63, 64, 65, 66, 95, 96, 97, 98, 127, 128])
def test_valid_start_word_is_the_sliding_window_of_the_flags(core, k):
    """bit t set <=> none of the symbols [t, t + K) is flagged (UnambiguousKmers.jl:134-148: every symbol of an
    emitted k-mer is certain), for every class of the word-count specialisation."""
    rng = np.random.default_rng(300 + k)
    for density in (0.0, 0.01, 0.05, 0.3, 1.0):
        for _ in range(40):
            bits = (rng.random(160) < density).astype(np.uint8)
            a = np.zeros(6, dtype=np.uint32)
            for i in range(160):
                if bits[i]:
                    a[i // 32] |= np.uint32(1) << np.uint32(i % 32)
            got = core.core_valid_start_word(a.ctypes.data, k)
            want = 0
            for t in range(32):
                if not bits[t:t + k].any():
                    want |= 1 << t
            assert got == want, (k, density, [hex(int(x)) for x in a])


def test_hash_primitives_reproduce_the_reference_values(core):
    """fx_hash (src/kmer.jl:255-261) and Base.hash (src/kmer.jl:206, Julia 1.10 / 1.11) as the device code computes
    them: the reference's known answers (test/runtests.jl:903-910, docs/src/hashing.md:18-20) and the oracle."""
    import json
    kats = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_kats.json")))
    for e in kats["fx_hash"]:
        limbs = [int(x, 16) for x in e["limbs"]] if "limbs" in e else (kt.kmer_limbs(e["kmer"].replace("U", "T")) if e["kmer"] else [])
        a = np.array(limbs, dtype=np.uint64)
        assert core.core_fx_hash(a.ctypes.data if a.size else None, a.size, 0) == int(e["hash"], 16), e
    for e in kats["base_hash"]:
        s = e["kmer"].replace("U", "T")
        a = np.array(kt.kmer_limbs(s), dtype=np.uint64)
        assert core.core_base_hash(a.ctypes.data, a.size, 0 ^ len(s)) == int(e["hash"], 16), e
    rng = np.random.default_rng(77)
    for n, k in ((1, 31), (2, 63), (3, 90), (4, 128)):
        km = rng.integers(0, 2**64, size=(50, n), dtype=np.uint64)
        for h0 in (0, 0x1234_5678_9ABC_DEF0):
            want_fx, want_b = ko.fx_hash(km, h0), ko.base_hash(km, k, h0)
            for i in range(km.shape[0]):
                row = np.ascontiguousarray(km[i])
                assert core.core_fx_hash(row.ctypes.data, n, h0) == int(want_fx[i])
                assert core.core_base_hash(row.ctypes.data, n, h0 ^ k) == int(want_b[i])


def test_two_to_four_expansion(core):
    """TwoToFour (src/construction_utils.jl:35): enc4 = 1 << enc2, eight symbols at a time, every input."""
    for s in range(1 << 16):
        want = 0
        for i in range(8):
            want |= (1 << ((s >> (2 * i)) & 3)) << (4 * i)
        assert core.core_onehot8(s) == want, hex(s)


def test_ascii_tables_follow_the_reference(core):
    """Strict tables: BioSequences' ascii_encode for the 2-bit alphabets (ACGT / ACGU, either case, anything else is an
    error).  Skipping table: ASCII_SKIPPING_LUT (src/iterators/common.jl:22-32) -- Aa, Cc, Gg, TtUu are 0..3, the
    ambiguity letters and the gap (either case) are skipped, every other byte is an error."""
    out = np.zeros(768, dtype=np.uint8)
    core.core_ascii_luts(out.ctypes.data)
    dna, rna, skip = out[:256], out[256:512], out[512:]
    SKIP, ERR = 0x40, 0x80
    want_skip = {}
    for code, letters in enumerate(("Aa", "cC", "gG", "TtUu")):
        for ch in letters:
            want_skip[ord(ch)] = code
    for ch in "-MRSVWYHKDBN":
        want_skip[ord(ch)] = SKIP
        want_skip[ord(ch.lower())] = SKIP
    for b in range(256):
        assert skip[b] == want_skip.get(b, ERR), b
        c = chr(b)
        assert dna[b] == ("ACGT".index(c.upper()) if c.upper() in "ACGT" and c.isalpha() else ERR), b
        assert rna[b] == ("ACGU".index(c.upper()) if c.upper() in "ACGU" and c.isalpha() else ERR), b


def test_positioned_ascii_tables_recode_like_the_byte_tables(core):
    """ascii_recode_kernel ORs eight pre-placed 32-bit entries per byte pair (ascii_luts.h: make_positioned) and assembles a
    group of 32 bytes with byte permutes: the result must be what the plain byte tables say, symbol by symbol -- every byte
    value in every position of a group, for the three tables."""
    core.core_ascii_group_positioned.restype = None
    core.core_ascii_group_positioned.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    luts = np.zeros(768, dtype=np.uint8)
    core.core_ascii_luts(luts.ctypes.data)
    rng = np.random.default_rng(11)
    out = np.zeros(4, dtype=np.uint32)
    for which in range(3):
        lut = luts[256 * which:256 * which + 256]
        groups = [np.roll(np.arange(256, dtype=np.uint8), s)[i:i + 32].copy() for s in range(32) for i in range(0, 256, 32)]
        groups += [rng.integers(0, 256, size=32, dtype=np.uint8) for _ in range(200)]
        groups += [np.frombuffer(b"ACGTNacgtnUu-RYKM" * 2, dtype=np.uint8)[:32].copy()]
        for g in groups:
            core.core_ascii_group_positioned(which, g.ctypes.data, out.ctypes.data)
            codes = int(out[0]) | (int(out[1]) << 32)
            for t in range(32):
                e = int(lut[g[t]])
                assert (codes >> (2 * t)) & 3 == e & 3, (which, t, g[t])
                assert (int(out[2]) >> t) & 1 == (1 if e >> 6 else 0), (which, t, g[t])
                assert (int(out[3]) >> t) & 1 == e >> 7, (which, t, g[t])


def test_ascii_tables_of_the_4bit_alphabets(core):
    """The tables behind k-mers over DNAAlphabet{4} / RNAAlphabet{4} from ASCII bytes: every symbol of the alphabet -- the gap
    and A C M G R S V T/U W Y H K D B N, encodings 0..15 (BioSymbols) -- in either case; any other byte is an error.  Checked
    against the encodings the reference's own sequences use (kmertools.CODE4) and against the oracle's restatement, byte by
    byte (SpacedKmers with K = J = 1 yields the encoding of each byte or raises)."""
    core.core_ascii_luts4.restype = None
    core.core_ascii_luts4.argtypes = [C.c_void_p]
    out = np.zeros(512, dtype=np.uint8)
    core.core_ascii_luts4(out.ctypes.data)
    for rna, lut in ((False, out[:256]), (True, out[256:])):
        for b in range(256):
            c = chr(b).upper()
            ok = c in kt.CODE4 and c != ("T" if rna else "U") and (chr(b).isalpha() or chr(b) == "-")
            assert lut[b] == (kt.CODE4[c] if ok else 0x80), (rna, b)
            try:
                got = int(ko.spaced(bytes([b]), 1, 1, 1, src_bits=8, kmer_bits=4, rna=rna)[0, 0])
            except ko.AmbiguousError:
                got = 0x80
            assert lut[b] == got, (rna, b)
