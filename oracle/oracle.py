"""ctypes binding of the CPU oracle (oracle/kmers_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(kmers.jl_b200/kmerscuda) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkmers_oracle.so")

KO_OK, KO_E_BAD_K, KO_E_AMBIGUOUS = 0, 1, 3
FW, FWRV, CANON = 0, 1, 2


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "kmers_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        cmd = ["make", "-C", _HERE] + (["-B"] if force else [])
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None
_u64p = C.POINTER(C.c_uint64)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.ko_n_limbs.restype = C.c_int
        L.ko_n_windows.restype = C.c_uint64
        L.ko_n_windows.argtypes = [C.c_uint64, C.c_int]
        L.ko_iterate4.restype = C.c_int
        L.ko_iterate4.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        L.ko_iterate.restype = C.c_int
        L.ko_iterate.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, _u64p, _u64p, _u64p]
        L.ko_unambiguous.restype = C.c_int
        L.ko_unambiguous.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, _u64p]
        L.ko_fx_hash.restype = None
        L.ko_fx_hash.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_void_p]
        for name in ("ko_reverse_complement", "ko_reverse", "ko_complement"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [C.c_void_p, C.c_int]
        L.ko_cmp.restype = C.c_int
        L.ko_cmp.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ko_shift_encoding.restype = None
        L.ko_shift_encoding.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        L.ko_shift_first_encoding.restype = None
        L.ko_shift_first_encoding.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        L.ko_unsafe_extract.restype = C.c_uint64
        L.ko_unsafe_extract.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_int, _u64p]
        L.ko_batch_iterate.restype = C.c_int
        L.ko_batch_iterate.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64,
                                       C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, _u64p, _u64p, _u64p, C.c_int]
        L.ko_batch_unambiguous.restype = C.c_int
        L.ko_batch_unambiguous.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64,
                                           C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int]
        L.ko_ascii_iterate.restype = C.c_int
        L.ko_ascii_iterate.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, _u64p, _u64p, _u64p]
        L.ko_ascii_unambiguous.restype = C.c_int
        L.ko_ascii_unambiguous.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, _u64p, _u64p, _u64p]
        L.ko_base_hash.restype = None
        L.ko_base_hash.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.c_void_p]
        L.ko_max_threads.restype = C.c_int
        _lib = L
    return _lib


class AmbiguousError(Exception):
    """Mirror of BioSequences.EncodeError raised through construction.jl:108-110."""

    def __init__(self, seq, pos, enc, n_before=0):
        super().__init__(f"cannot encode 4-bit symbol 0x{enc:x} at position {pos} (sequence {seq})")
        self.seq, self.pos, self.enc, self.n_before = seq, pos, enc, n_before


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def n_limbs(k: int) -> int:
    return lib().ko_n_limbs(k)


def _words(words):
    w = np.ascontiguousarray(words, dtype=np.uint64)
    if w.size == 0:
        w = np.zeros(1, dtype=np.uint64)
    return w


def iterate(words, length, k, mode, src_bits=2, first=0, want_hash=False):
    """One sequence through the literal iterator state machine.

    Returns (a, b, hash): a = fw (FW/FWRV) or canonical (CANON) limbs [n, N];
    b = rv limbs for FWRV else None; hash = fx_hash(a) or None."""
    L = lib()
    if k < 1:
        raise ValueError("K must be at least 1")
    N = n_limbs(k)
    n = max(0, length - k + 1)
    w = _words(words)
    a = np.zeros((n, N), dtype=np.uint64)
    b = np.zeros((n, N), dtype=np.uint64) if mode == FWRV else None
    h = np.zeros(n, dtype=np.uint64) if want_hash else None
    n_out, ep, ee = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    st = L.ko_iterate(_ptr(w), first, length, src_bits, k, mode, _ptr(a), _ptr(b), _ptr(h),
                      C.byref(n_out), C.byref(ep), C.byref(ee))
    if st == KO_E_AMBIGUOUS:
        raise AmbiguousError(0, ep.value, ee.value, n_out.value)
    if st != KO_OK:
        raise ValueError(f"oracle status {st}")
    assert n_out.value == n
    return a, b, h


def n_limbs4(k):
    """N of Kmer{<:NucleicAcidAlphabet{4},K,N} (src/kmer.jl:117-137, bps = 4)."""
    return (4 * k + 63) // 64


def iterate4(words, length, k, mode, src_bits=4, first=0, want_hash=False):
    """One sequence -> k-mers over a 4-bit alphabet (Copyable from a 4-bit source, TwoToFour from a
    2-bit source).  Returns (a, b, hash) like iterate()."""
    L = lib()
    if k < 1:
        raise ValueError("K must be at least 1")
    N = n_limbs4(k)
    n = max(0, length - k + 1)
    w = _words(words)
    a = np.zeros((n, N), dtype=np.uint64)
    b = np.zeros((n, N), dtype=np.uint64) if mode == FWRV else None
    h = np.zeros(n, dtype=np.uint64) if want_hash else None
    n_out = C.c_uint64(0)
    st = L.ko_iterate4(_ptr(w), first, length, src_bits, k, mode, _ptr(a), _ptr(b), _ptr(h), C.byref(n_out))
    if st != KO_OK:
        raise ValueError(f"oracle status {st}")
    assert n_out.value == n
    return a, b, h


def unambiguous(words, length, k, src_bits=4, first=0):
    L = lib()
    N = n_limbs(k)
    n = max(0, length - k + 1)
    w = _words(words)
    km = np.zeros((max(n, 1), N), dtype=np.uint64)
    pos = np.zeros(max(n, 1), dtype=np.int64)
    n_out = C.c_uint64(0)
    st = L.ko_unambiguous(_ptr(w), first, length, src_bits, k, _ptr(km), _ptr(pos), C.byref(n_out))
    if st != KO_OK:
        raise ValueError(f"oracle status {st}")
    return km[: n_out.value].copy(), pos[: n_out.value].copy()


def _bytes(src):
    if isinstance(src, str):
        src = src.encode("latin-1")
    b = np.frombuffer(bytes(src), dtype=np.uint8) if not isinstance(src, np.ndarray) else np.ascontiguousarray(src, np.uint8)
    return b if b.size else np.zeros(1, np.uint8), (len(src) if not isinstance(src, np.ndarray) else int(src.size))


def ascii_iterate(src, k, mode, rna=False, want_hash=False):
    """FwKmers / FwRvIterator / CanonicalKmers over an ASCII source (String, codeunits, ...).
    Raises AmbiguousError(pos = 1-based byte index, enc = the byte) on an invalid byte."""
    L = lib()
    if k < 1:
        raise ValueError("K must be at least 1")
    b, length = _bytes(src)
    N = n_limbs(k)
    n = max(0, length - k + 1)
    a = np.zeros((n, N), dtype=np.uint64)
    bb = np.zeros((n, N), dtype=np.uint64) if mode == FWRV else None
    h = np.zeros(n, dtype=np.uint64) if want_hash else None
    n_out, ep, ee = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    st = L.ko_ascii_iterate(_ptr(b), length, int(rna), k, mode, _ptr(a), _ptr(bb), _ptr(h), C.byref(n_out), C.byref(ep), C.byref(ee))
    if st == KO_E_AMBIGUOUS:
        raise AmbiguousError(0, ep.value, ee.value, n_out.value)
    if st != KO_OK:
        raise ValueError(f"oracle status {st}")
    return a, bb, h


def ascii_unambiguous(src, k):
    """UnambiguousKmers over an ASCII source: (kmers, 1-based starts)."""
    L = lib()
    b, length = _bytes(src)
    N = n_limbs(k)
    n = max(0, length - k + 1)
    km = np.zeros((max(n, 1), N), dtype=np.uint64)
    pos = np.zeros(max(n, 1), dtype=np.int64)
    n_out, ep, ee = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    st = L.ko_ascii_unambiguous(_ptr(b), length, k, _ptr(km), _ptr(pos), C.byref(n_out), C.byref(ep), C.byref(ee))
    if st == KO_E_AMBIGUOUS:
        raise AmbiguousError(0, ep.value, ee.value, n_out.value)
    if st != KO_OK:
        raise ValueError(f"oracle status {st}")
    return km[: n_out.value].copy(), pos[: n_out.value].copy()


def fx_hash(kmers, h0=0):
    km = np.ascontiguousarray(kmers, dtype=np.uint64)
    if km.ndim == 1:
        km = km.reshape(-1, 1)
    out = np.zeros(km.shape[0], dtype=np.uint64)
    lib().ko_fx_hash(_ptr(km), km.shape[0], km.shape[1], h0, _ptr(out))
    return out


def base_hash(kmers, k, h0=0):
    """Base.hash.(kmers, h0) (src/kmer.jl:206), Julia 1.10 / 1.11 tuple + UInt64 hashing."""
    km = np.ascontiguousarray(kmers, dtype=np.uint64)
    if km.ndim == 1:
        km = km.reshape(-1, 1)
    out = np.zeros(km.shape[0], dtype=np.uint64)
    lib().ko_base_hash(_ptr(km) if km.size else None, km.shape[0], km.shape[1], k, h0, _ptr(out))
    return out


def _limbs(x, k):
    a = np.array(x, dtype=np.uint64).reshape(-1)
    assert a.size == n_limbs(k)
    return a


def reverse_complement(limbs, k):
    a = _limbs(limbs, k).copy()
    lib().ko_reverse_complement(_ptr(a), k)
    return tuple(int(v) for v in a)


def reverse(limbs, k):
    a = _limbs(limbs, k).copy()
    lib().ko_reverse(_ptr(a), k)
    return tuple(int(v) for v in a)


def complement(limbs, k):
    a = _limbs(limbs, k).copy()
    lib().ko_complement(_ptr(a), k)
    return tuple(int(v) for v in a)


def cmp(a, b):
    x = np.array(a, dtype=np.uint64)
    y = np.array(b, dtype=np.uint64)
    return lib().ko_cmp(_ptr(x), _ptr(y), len(x))


def shift_encoding(limbs, k, enc):
    a = _limbs(limbs, k).copy()
    lib().ko_shift_encoding(_ptr(a), k, enc)
    return tuple(int(v) for v in a)


def shift_first_encoding(limbs, k, enc):
    a = _limbs(limbs, k).copy()
    lib().ko_shift_first_encoding(_ptr(a), k, enc)
    return tuple(int(v) for v in a)


def unsafe_extract(words, k, from_index, src_bits):
    a = np.zeros(n_limbs(k), dtype=np.uint64)
    bad = C.c_uint64(0)
    w = _words(words)
    p = lib().ko_unsafe_extract(_ptr(a), k, _ptr(w), from_index, src_bits, C.byref(bad))
    if p:
        raise AmbiguousError(0, p, bad.value)
    return tuple(int(v) for v in a)


def window_offsets(seq_len, k):
    """Exclusive prefix sum of per-read window counts (FwKmers.jl:40-43)."""
    seq_len = np.asarray(seq_len, dtype=np.uint64)
    cnt = np.where(seq_len >= k, seq_len - np.uint64(k) + np.uint64(1), np.uint64(0)).astype(np.uint64)
    off = np.zeros(len(seq_len) + 1, dtype=np.uint64)
    np.cumsum(cnt, out=off[1:])
    return off


def batch_iterate(words, n_seqs, k, mode, *, word_off=None, seq_len=None, uniform_len=0,
                  uniform_stride=0, src_bits=2, want_hash=False, threads=0, out=None):
    """Read set through the per-read serial recurrence, OpenMP over reads.
    Returns (a, b, hash, out_off)."""
    L = lib()
    N = n_limbs(k)
    w = _words(words)
    if word_off is not None:
        word_off = np.ascontiguousarray(word_off, dtype=np.uint64)
        seq_len = np.ascontiguousarray(seq_len, dtype=np.uint64)
        out_off = window_offsets(seq_len, k)
        total = int(out_off[-1])
    else:
        out_off = None
        total = n_seqs * max(0, uniform_len - k + 1)
    if out is not None:
        a, b, h = out
    else:
        a = np.empty((total, N), dtype=np.uint64)
        b = np.empty((total, N), dtype=np.uint64) if mode == FWRV else None
        h = np.empty(total, dtype=np.uint64) if want_hash else None
    es, ep, ee = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    st = L.ko_batch_iterate(_ptr(w), n_seqs, _ptr(word_off), _ptr(seq_len), uniform_len, uniform_stride,
                            src_bits, k, mode, _ptr(out_off), _ptr(a), _ptr(b), _ptr(h),
                            C.byref(es), C.byref(ep), C.byref(ee), threads)
    if st == KO_E_AMBIGUOUS:
        raise AmbiguousError(es.value, ep.value, ee.value)
    if st != KO_OK:
        raise ValueError(f"oracle status {st}")
    return a, b, h, out_off


def batch_unambiguous(words, n_seqs, k, *, word_off=None, seq_len=None, uniform_len=0,
                      uniform_stride=0, src_bits=4, threads=0):
    L = lib()
    N = n_limbs(k)
    w = _words(words)
    if word_off is not None:
        word_off = np.ascontiguousarray(word_off, dtype=np.uint64)
        seq_len = np.ascontiguousarray(seq_len, dtype=np.uint64)
    off = np.zeros(n_seqs + 1, dtype=np.uint64)
    args = (_ptr(w), n_seqs, _ptr(word_off), _ptr(seq_len), uniform_len, uniform_stride, src_bits, k)
    st = L.ko_batch_unambiguous(*args, _ptr(off), None, None, threads)
    if st != KO_OK:
        raise ValueError(f"oracle status {st}")
    total = int(off[-1])
    km = np.zeros((max(total, 1), N), dtype=np.uint64)
    pos = np.zeros(max(total, 1), dtype=np.int64)
    st = L.ko_batch_unambiguous(*args, _ptr(off), _ptr(km), _ptr(pos), threads)
    if st != KO_OK:
        raise ValueError(f"oracle status {st}")
    return km[:total], pos[:total], off


def spaced(src, length, k, j, *, src_bits=2, kmer_bits=2, first=0, rna=False):
    """SpacedKmers{A,K,J} over one sequence: LongSequence words (src_bits 2 / 4) or ASCII bytes (src_bits 8, `src`
    bytes / str / uint8 array); kmer_bits = 2 or 4.  Raises AmbiguousError like the strict iterators."""
    L = lib()
    if src_bits == 8:
        b, _ = _bytes(src)
        buf = b
    else:
        buf = _words(src)
    N = (k * kmer_bits + 63) // 64
    n = 0 if length < k else (length - k) // j + 1
    out = np.zeros((max(n, 1), N), dtype=np.uint64)
    n_out, ep, ee = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    L.ko_spaced.restype = C.c_int
    L.ko_spaced.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                            _u64p, _u64p, _u64p]
    st = L.ko_spaced(_ptr(buf), src_bits, first, length, int(rna), k, j, kmer_bits, _ptr(out), C.byref(n_out), C.byref(ep), C.byref(ee))
    if st == KO_E_AMBIGUOUS:
        raise AmbiguousError(0, ep.value, ee.value, n_out.value)
    if st != KO_OK:
        raise ValueError(f"oracle status {st}")
    assert n_out.value == n
    return out[:n]


def max_threads() -> int:
    return lib().ko_max_threads()


# ---- consumers of the stream (docs examples) -------------------------------------------------
def minhash_sketch(words, n, k, s, canonical=True):
    """Bottom-s MinHash sketch under fx_hash: `sketch(fx_hash, CanonicalDNAMers{K}(seq), s)`
    (/root/reference/docs/src/minhash.md:31-36).  MinHash.jl is an external package that is not
    vendored; its published definition (Mash, Ondov et al. 2016) is restated: the s smallest
    distinct hash values over all k-mers, ascending."""
    if n < k:
        return np.zeros(0, dtype=np.uint64)
    _, _, h = iterate(words, n, k, CANON if canonical else FW, want_hash=True)
    return np.unique(h)[:s]


def composition(words, n, k, canonical=False):
    """`counts[as_integer(kmer) + 1] += 1` for every k-mer (/root/reference/docs/src/composition.md:28-39;
    as_integer: src/kmer.jl:305-326 -- the k-mer's bits as one integer), 0-based here."""
    if n < k:
        return np.zeros(4**k, dtype=np.uint32)
    a, _, _ = iterate(words, n, k, CANON if canonical else FW)
    return np.bincount(a[:, -1].astype(np.int64), minlength=4**k).astype(np.uint32)
