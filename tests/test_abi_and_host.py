"""CPU-side checks (no GPU): the C-ABI library loads and exports exactly what include/kmerscuda.h
declares, the ctypes structs match the header's layout, and the host mirror keeps the reference's
argument checks and error behaviour."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import kmertools as kt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "kmerscuda.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kmc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import kmerscuda
    from kmerscuda import _abi
    lib = _abi.load()
    declared = header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in kmerscuda.h but not exported"
    assert sorted(_abi.SIGNATURES) == declared, "ctypes binding and header disagree"
    assert lib.kmc_version() == 100


def test_struct_layout_matches_header():
    from kmerscuda import _abi
    # natural alignment of the C structs in the header
    assert C.sizeof(_abi.kmc_seqs) == 64
    assert _abi.kmc_seqs.src_bits.offset == 56 and _abi.kmc_seqs.first_symbol_offset.offset == 60
    assert C.sizeof(_abi.kmc_out) == 56
    assert C.sizeof(_abi.kmc_result) == 64 and _abi.kmc_result.digest.offset == 32
    assert _abi.kmc_result.err_sym.offset == 24 and _abi.kmc_result.kernel_ms.offset == 28


def test_status_strings_and_constants():
    from kmerscuda import _abi
    lib = _abi.load()
    assert lib.kmc_status_string(0) == b"ok"
    assert b"at least 1" in lib.kmc_status_string(_abi.KMC_E_BAD_K)
    hdr = open(HEADER).read()
    for name in ("KMC_FW", "KMC_FWRV", "KMC_CANON", "KMC_UNAMBIG", "KMC_E_BAD_K", "KMC_E_AMBIGUOUS",
                 "KMC_E_OUT_TOO_SMALL", "KMC_MAX_K"):
        m = re.search(rf"#define {name} (\w+)", hdr)
        assert int(m.group(1), 0) == getattr(_abi, name)


def test_null_context_is_rejected_without_a_gpu():
    from kmerscuda import _abi
    lib = _abi.load()
    assert lib.kmc_sync(None) == _abi.KMC_E_BAD_ARG
    assert lib.kmc_extract(None, None, 31, 0, 0, None, None) == _abi.KMC_E_BAD_ARG


def test_longsequence_packing_matches_biosequences_layout():
    import kmerscuda as kc
    rng = np.random.default_rng(1)
    for n in (0, 1, 31, 32, 33, 64, 100):
        s = kt.random_dna(rng, n, ambiguous=0.1)
        s2 = "".join(c if c in "ACGT" else "A" for c in s)
        assert np.array_equal(kc.LongDNA2(s2).data, kt.pack2(s2))
        assert np.array_equal(kc.LongDNA4(s).data, kt.pack4(s))
    # derived from the reference's definitions: LongDNA{2}("TAGCTAGGACA").data == [0x4a363]
    assert kc.LongDNA2("TAGCTAGGACA").data.tolist() == [0x4A363]
    with pytest.raises(kc.EncodeError):
        kc.LongDNA2("ACGN")


def test_k_checks_mirror_the_reference():
    import kmerscuda as kc
    seq = kc.LongDNA2("ACGT")
    for ctor in (kc.FwDNAMers, kc.FwRvDNAIterator, kc.CanonicalDNAMers, kc.UnambiguousDNAMers):
        with pytest.raises(ValueError, match="K must be at least 1"):  # FwKmers.jl:32-33
            ctor(0, seq)
        with pytest.raises(TypeError, match="K must be an Int"):
            ctor(2.5, seq)
    assert len(kc.FwDNAMers(3, seq)) == 2 and len(kc.FwDNAMers(5, seq)) == 0  # FwKmers.jl:40-43
    assert len(kc.UnambiguousDNAMers(2, seq)) == 3
    with pytest.raises(TypeError):  # SizeUnknown for 4-bit sources, UnambiguousKmers.jl:33-37
        len(kc.UnambiguousDNAMers(2, kc.LongDNA4("ACGT")))
    assert [kc.n_limbs(k) for k in (1, 32, 33, 64, 65, 128)] == [1, 1, 2, 2, 3, 4]


def test_readset_layout():
    import kmerscuda as kc
    seqs = [kc.LongDNA2("ACGT" * 10), kc.LongDNA2(""), kc.LongDNA2("A" * 33)]
    rs = kc.ReadSet.from_sequences(seqs)
    assert rs.seq_word_offset.tolist() == [0, 2, 2] and rs.seq_len.tolist() == [40, 0, 33]
    assert rs.words.size == 4
    assert rs.window_counts(31).tolist() == [10, 0, 3]


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="checks the no-device failure mode")
def test_no_cpu_fallback_without_a_device():
    import kmerscuda as kc
    with pytest.raises(kc.KmersCUDAError, match="no CPU fallback"):
        kc.CanonicalDNAMers(3, kc.LongDNA2("ACGTACGT")).collect()


def test_header_is_plain_c_and_the_c_examples_link(tmp_path):
    """include/kmerscuda.h is C99 (what a cgo / ccall / ctypes binding needs), and the examples -- collect_canonical.c
    (one kmc_extract_host call) and c5_group_count.c (the multi-GPU count + NCCL merge, no Python in the loop) -- build
    against the shared library with gcc alone; without a device they report that there is no CPU fallback."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    lib_dir = os.path.join(ROOT, "kmers.jl_b200")
    for name in ("collect_canonical", "c5_group_count"):
        exe = str(tmp_path / name)
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", name + ".c"), "-L", lib_dir, "-lkmerscuda",
                        "-Wl,-rpath," + lib_dir, "-o", exe], check=True)
        if __import__("torch").cuda.is_available():
            continue  # the device path of the examples is exercised by tests/test_gpu_multi.py
        r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
        assert r.returncode == 0
        assert "ABI version 100" in r.stdout and "no CPU fallback" in r.stdout
