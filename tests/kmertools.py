"""Independent, naive (string-level) definitions used by the tests.

Nothing here shares code with the oracle or the CUDA library: k-mers are built
per window from the *string* slice, the way the reference's own tests build their
expected vectors (`[Kmer{A,K}(seq[i:i+K-1]) for i in ...]`,
/root/reference/test/runtests.jl:674-690, 739-761, 774-787, 804-837).
"""
from __future__ import annotations

import numpy as np

MASK64 = (1 << 64) - 1

# 2-bit encoding A,C,G,T/U = 0,1,2,3 (pinned by as_integer(mer"AACT"d) == 0x07,
# /root/reference/src/kmer.jl:288-289, and the fx_hash KATs).
CODE2 = {"A": 0, "C": 1, "G": 2, "T": 3, "U": 3}
# 4-bit one-hot encoding, IUPAC = bitwise OR, gap = 0, N = 15 (BioSymbols).
CODE4 = {
    "-": 0, "A": 1, "C": 2, "M": 3, "G": 4, "R": 5, "S": 6, "V": 7,
    "T": 8, "U": 8, "W": 9, "Y": 10, "H": 11, "K": 12, "D": 13, "B": 14, "N": 15,
}
COMPLEMENT = {"A": "T", "C": "G", "G": "C", "T": "A", "U": "A"}


def n_limbs(k: int) -> int:
    return (2 * k + 63) // 64


def pack2(seq: str) -> np.ndarray:
    """LongSequence{DNAAlphabet{2}}.data: symbol i (0-based) at bits [2i mod 64, +2) of word i//32."""
    words = [0] * ((len(seq) + 31) // 32)
    for i, c in enumerate(seq.upper()):
        words[i // 32] |= CODE2[c] << (2 * (i % 32))
    return np.array(words, dtype=np.uint64)


def pack4(seq: str) -> np.ndarray:
    """LongSequence{DNAAlphabet{4}}.data: symbol i (0-based) at bits [4i mod 64, +4) of word i//16."""
    words = [0] * ((len(seq) + 15) // 16)
    for i, c in enumerate(seq.upper()):
        words[i // 16] |= CODE4[c] << (4 * (i % 16))
    return np.array(words, dtype=np.uint64)


def pack_codes(codes: np.ndarray, bits: int) -> np.ndarray:
    """Vectorised packer for integer code arrays (bits = 2 or 4)."""
    per = 64 // bits
    n = len(codes)
    nw = (n + per - 1) // per
    padded = np.zeros(nw * per, dtype=np.uint64)
    padded[:n] = codes.astype(np.uint64)
    shifts = (np.arange(per, dtype=np.uint64) * np.uint64(bits))
    return np.bitwise_or.reduce(padded.reshape(nw, per) << shifts, axis=1).astype(np.uint64)


def kmer_int(s: str) -> int:
    """Kmer{DNAAlphabet{2},K} as one big integer: first symbol in the highest bits."""
    v = 0
    for c in s.upper():
        v = (v << 2) | CODE2[c]
    return v


def kmer_limbs(s: str) -> tuple[int, ...]:
    """NTuple{N,UInt64} limbs, head (most significant) first; unused bits are the top of limb 1."""
    n = n_limbs(len(s))
    v = kmer_int(s)
    return tuple((v >> (64 * (n - 1 - i))) & MASK64 for i in range(n))


def revcomp(s: str) -> str:
    return "".join(COMPLEMENT[c] for c in reversed(s.upper()))


def is_certain(c: str) -> bool:
    return bin(CODE4[c.upper()]).count("1") == 1


def naive_fw(seq: str, k: int) -> list[tuple[int, ...]]:
    return [kmer_limbs(seq[i:i + k]) for i in range(len(seq) - k + 1)]


def naive_fwrv(seq: str, k: int):
    return [(kmer_limbs(seq[i:i + k]), kmer_limbs(revcomp(seq[i:i + k]))) for i in range(len(seq) - k + 1)]


def naive_canonical(seq: str, k: int) -> list[tuple[int, ...]]:
    out = []
    for i in range(len(seq) - k + 1):
        w = seq[i:i + k]
        out.append(min(kmer_limbs(w), kmer_limbs(revcomp(w))))
    return out


def naive_unambiguous(seq: str, k: int):
    """(kmer, 1-based start) for every window whose symbols are all certain
    (/root/reference/test/runtests.jl:804-811)."""
    out = []
    for i in range(len(seq) - k + 1):
        w = seq[i:i + k]
        if all(is_certain(c) for c in w):
            out.append((kmer_limbs(w), i + 1))
    return out


FX_CONSTANT = 0x517CC1B727220A95


def fx_hash(limbs, h: int = 0) -> int:
    """/root/reference/src/kmer.jl:255-261."""
    for x in limbs:
        h = ((((h << 5) | (h >> 59)) & MASK64) ^ x) * FX_CONSTANT & MASK64
    return h


def splitmix64(x: np.ndarray) -> np.ndarray:
    """Counter-based generator used for all synthetic inputs."""
    with np.errstate(over="ignore"):
        z = x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def random_dna(rng: np.random.Generator, n: int, ambiguous: float = 0.0) -> str:
    """Random DNA string; with probability `ambiguous` a symbol is drawn from the
    IUPAC ambiguity codes and gap (cf. /root/reference/test/utils.jl:22-24)."""
    base = rng.choice(list("ACGT"), size=n)
    if ambiguous > 0:
        amb = rng.choice(list("MRSVWYHKDBN-"), size=n)
        mask = rng.random(n) < ambiguous
        base = np.where(mask, amb, base)
    return "".join(base.tolist())


# ---- k-mers over the 4-bit alphabets (independent string-level definitions) ---------------------
# IUPAC complements (BioSymbols: complement of an ambiguity set = the set of complements)
COMPLEMENT4 = {"A": "T", "C": "G", "G": "C", "T": "A", "M": "K", "K": "M", "R": "Y", "Y": "R", "W": "W", "S": "S",
               "V": "B", "B": "V", "H": "D", "D": "H", "N": "N", "-": "-"}
SYM4 = "-ACMGRSVTWYHKDBN"


def n_limbs4(k: int) -> int:
    return (4 * k + 63) // 64


def kmer4_limbs(s: str) -> tuple:
    """Kmer{DNAAlphabet{4},K,N}.data of a string: first symbol in the highest used nibble."""
    v = 0
    for c in s.upper().replace("U", "T"):
        v = (v << 4) | CODE4[c]
    N = n_limbs4(len(s))
    return tuple((v >> (64 * (N - 1 - i))) & (2**64 - 1) for i in range(N))


def revcomp4(s: str) -> str:
    return "".join(COMPLEMENT4[c] for c in reversed(s.upper().replace("U", "T")))


def naive_fwrv4(seq: str, k: int):
    return [(kmer4_limbs(seq[i:i + k]), kmer4_limbs(revcomp4(seq[i:i + k]))) for i in range(len(seq) - k + 1)]


def random_iupac(rng: np.random.Generator, n: int) -> str:
    return "".join(rng.choice(list(SYM4), size=n).tolist())
