"""UnambiguousKmers over recoded (4-bit / ASCII) sources: the two compaction paths behind kmc_extract.

  * lin_compact_kernel (lincompact.cuh): sets whose sequences lie ascending and disjoint in the buffer are
    compacted in SOURCE order -- the cases here are the ones its index arithmetic has to get right: work items
    that straddle two (or many) sequences, uniform strides that are not multiples of the item width, views with
    a first_symbol_offset, every limb count, both output layouts;
  * compact_kernel (compact_kernels.cuh): every other layout -- sequences out of order or overlapping in the
    buffer -- and the same suites again with KMC_LINEAR=0, so that the general path stays covered.

Everything is compared with the oracle (/root/reference/src/iterators/UnambiguousKmers.jl:109-148 restated)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import kmertools as kt
from oracle import oracle as ko

pytestmark = pytest.mark.gpu
UNAMBIG = 3
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def kc():
    import kmerscuda
    return kmerscuda


def codes4(rng, n, amb):
    c = np.uint64(1) << rng.integers(0, 4, size=n).astype(np.uint64)
    return np.where(rng.random(n) < amb, np.uint64(15), c).astype(np.uint64)


def ascii_reads(rng, lens, amb):
    out = []
    for n in lens:
        s = rng.choice(np.frombuffer(b"ACGTacgtU", dtype=np.uint8), size=int(n))
        s = np.where(rng.random(int(n)) < amb, np.uint8(ord("N")), s)
        out.append(bytes(s.astype(np.uint8)))
    return out


def check(kc, rs, per_read, k, **kw):
    """per_read: list of (kmers, 1-based positions) per sequence, from the oracle."""
    km = np.concatenate([p[0] for p in per_read]) if per_read else np.zeros((0, kc.n_limbs(k)), np.uint64)
    pos = np.concatenate([p[1] for p in per_read]) if per_read else np.zeros(0, np.int64)
    off = np.concatenate([[0], np.cumsum([p[0].shape[0] for p in per_read])]).astype(np.uint64)
    N = kc.n_limbs(k)
    for aos in (False, True):
        for hash_ in (False, True):
            e = kc.extract(UNAMBIG, rs, k, aos=aos, hash=hash_, want_seq_offsets=True, **kw)
            assert e.n == km.shape[0]
            if aos:
                assert np.array_equal(e.kmers[:, :N], km) and np.array_equal(e.kmers[:, N].astype(np.int64), pos)
            else:
                assert np.array_equal(e.kmers, km) and np.array_equal(e.index, pos)
            assert np.array_equal(e.seq_out_offset, off)
            if hash_:
                assert np.array_equal(e.hash, ko.fx_hash(km) if km.size else np.zeros(0, np.uint64))
    import ctypes as C
    ctx = kc.default_context()
    drs = kc.DeviceReadSet(ctx, rs)
    cnt = C.c_uint64(0)
    ctx._check(ctx.lib.kmc_count(ctx.handle, C.byref(drs.desc), k, UNAMBIG, C.byref(cnt)))
    assert cnt.value == km.shape[0]


@pytest.mark.parametrize("k,length,stride", [(31, 150, 150), (31, 150, 157), (5, 37, 37), (2, 9, 9), (1, 8, 8), (31, 31, 33),
                                             (63, 150, 151), (97, 300, 301), (3, 10, 4097), (31, 150, 5000),
                                             # K = 32 N (every bit of the run's last word is used); odd and tiny numbers of windows per read
                                             (32, 150, 150), (64, 200, 203), (96, 250, 250), (128, 300, 307), (32, 33, 40), (32, 34, 34),
                                             (30, 150, 150), (33, 151, 160)])
def test_uniform_ascii_reads_straddle_items(kc, k, length, stride):
    """Uniform ASCII reads: one symbol per offset unit, so a stride that is not a multiple of the item width makes
    items straddle two reads, and the per-slot position has to wrap (strides below and above 4096 take different
    arithmetic in the kernel)."""
    rng = np.random.default_rng(k * 7919 + stride)
    n_reads = 700 if stride < 1000 else 40
    reads = ascii_reads(rng, [length] * n_reads, 0.02)
    buf = np.full(n_reads * stride + 1, ord("A"), dtype=np.uint8)
    for i, r in enumerate(reads):
        buf[i * stride:i * stride + length] = np.frombuffer(r, dtype=np.uint8)
    rs = kc.ReadSet(8, buf, n_reads, uniform_len=length, uniform_stride_words=stride)
    check(kc, rs, [ko.ascii_unambiguous(r, k) for r in reads], k)


@pytest.mark.parametrize("k", [1, 4, 31, 33, 64, 65, 97, 128])
def test_ragged_ascii_reads_straddle_items(kc, k):
    """Concatenated strings of every length (no padding at all): items straddle two or many sequences."""
    rng = np.random.default_rng(31 * k)
    lens = [0, 1, 2, 1, 1, 3, k - 1, k, k + 1, 2 * k, 7, 8, 9, 0, 0, 5, 150, 151, 1000] + rng.integers(0, 260, size=400).tolist()
    reads = ascii_reads(rng, [max(0, int(x)) for x in lens], 0.03)
    rs = kc.ReadSet.from_strings(reads)
    check(kc, rs, [ko.ascii_unambiguous(r, k) for r in reads], k)


@pytest.mark.parametrize("first", [1, 3, 15, 16, 21])
def test_uniform_4bit_reads_with_first_symbol_offset(kc, first):
    rng = np.random.default_rng(first)
    k, length, n_reads = 31, 150, 500
    stride = (first + length + 15) // 16
    codes = np.ones((n_reads, stride * 16), dtype=np.uint64)
    codes[:, :] = codes4(rng, n_reads * stride * 16, 0.01).reshape(n_reads, -1)
    words = kt.pack_codes(codes.reshape(-1), 4)
    rs = kc.ReadSet(4, words, n_reads, uniform_len=length, uniform_stride_words=stride, first_symbol_offset=first)
    per = []
    for r in range(n_reads):
        per.append(ko.unambiguous(words[r * stride:(r + 1) * stride], length, k, src_bits=4, first=first))
    check(kc, rs, per, k)


def test_sequences_out_of_order_or_overlapping_take_the_general_path(kc):
    """Offsets that are not ascending, and views that overlap in the buffer: the output order is the order of the
    sequences, not of the buffer, so the source-order compaction must not be used (lin_prepare detects it)."""
    rng = np.random.default_rng(2024)
    k = 21
    lens = rng.integers(0, 300, size=200)
    nw = (lens + 15) // 16
    codes = [codes4(rng, int(n), 0.02) for n in lens]
    packed = [kt.pack_codes(c, 4) if len(c) else np.zeros(0, np.uint64) for c in codes]
    order = rng.permutation(len(lens))  # where each sequence lies in the buffer
    off = np.zeros(len(lens), dtype=np.uint64)
    at = 0
    for i in order:
        off[i] = at
        at += int(nw[i])
    words = np.zeros(at + 1, dtype=np.uint64)
    for i in range(len(lens)):
        words[int(off[i]):int(off[i]) + int(nw[i])] = packed[i]
    rs = kc.ReadSet(4, words, len(lens), seq_word_offset=off, seq_len=lens.astype(np.uint64))
    per = [ko.unambiguous(words[int(off[i]):int(off[i]) + int(nw[i]) + 1], int(lens[i]), k, src_bits=4) for i in range(len(lens))]
    check(kc, rs, per, k)
    # overlapping views of one buffer (every view starts one word after the previous one and is 10 words long)
    base = kt.pack_codes(codes4(rng, 16 * 400, 0.01), 4)
    n = 380
    off2 = np.arange(n, dtype=np.uint64)
    ln2 = np.full(n, 150, dtype=np.uint64)
    rs2 = kc.ReadSet(4, base, n, seq_word_offset=off2, seq_len=ln2)
    per2 = [ko.unambiguous(base[i:i + 11], 150, k, src_bits=4) for i in range(n)]
    check(kc, rs2, per2, k)
    check(kc, rs2, per2, k, host_path=True)  # the host pipeline checks the layout on the host
    # the same overlap as a uniform set (stride 1 word < 150 symbols)
    rs3 = kc.ReadSet(4, base, n, uniform_len=150, uniform_stride_words=1)
    check(kc, rs3, per2, k)


def test_tiny_uniform_strides(kc):
    """Strides shorter than a work item (fewer symbols than G): handled by the general path."""
    rng = np.random.default_rng(5)
    for k, length, stride in ((1, 3, 3), (2, 4, 5), (2, 2, 2)):
        reads = ascii_reads(rng, [length] * 900, 0.05)
        buf = np.full(900 * stride + 1, ord("A"), dtype=np.uint8)
        for i, r in enumerate(reads):
            buf[i * stride:i * stride + length] = np.frombuffer(r, dtype=np.uint8)
        rs = kc.ReadSet(8, buf, 900, uniform_len=length, uniform_stride_words=stride)
        check(kc, rs, [ko.ascii_unambiguous(r, k) for r in reads], k)


def test_one_long_sequence_and_a_tail_chunk(kc):
    """One sequence whose windows do not fill the last chunk of 2048 positions; K = 31 and K = 32."""
    rng = np.random.default_rng(77)
    for k, n in ((31, 70_001), (32, 5000), (64, 2048 + 63)):
        words = kt.pack_codes(codes4(rng, n, 0.01), 4)
        rs = kc.ReadSet(4, words, 1, uniform_len=n, uniform_stride_words=len(words))
        check(kc, rs, [ko.unambiguous(words, n, k, src_bits=4)], k)


def test_general_path_still_passes_its_suites():
    """KMC_LINEAR=0 keeps every set on compact_kernel: the 4-bit / ASCII UnambiguousKmers suites once more on it."""
    env = dict(os.environ, KMC_LINEAR="0")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider",
                        os.path.join(ROOT, "tests", "test_gpu_fourbit.py"), os.path.join(ROOT, "tests", "test_gpu_ascii.py"),
                        "-k", "unambiguous or ragged or uniform or views or capacity or dense or string_read_set or host_path"],
                       env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
