// ascii_luts.h -- the byte -> 2-bit code tables of the ASCII sources (AsciiEncode, src/construction.jl:95-96): plain
// host code, uploaded to constant memory by ascii.cu; a header so that the CPU tests can read the same tables.
#pragma once
#include <cstdint>

namespace kmc {

// LUT entry: bits 0-1 = 2-bit code, bit 6 = skip (ambiguity code / gap), bit 7 = error
constexpr uint8_t kSkip = 0x40, kErr = 0x80;

// the 4-bit tables (k-mers over DNAAlphabet{4} / RNAAlphabet{4} from ASCII sources): bits 0-3 = the 4-bit encoding,
// bit 7 = error
struct AsciiLuts {
    uint8_t strict_dna[256], strict_rna[256], skipping[256], dna4[256], rna4[256];
};

inline AsciiLuts make_luts()
{
    AsciiLuts l;
    for (int i = 0; i < 256; ++i) l.strict_dna[i] = l.strict_rna[i] = l.skipping[i] = l.dna4[i] = l.rna4[i] = kErr;
    // BioSequences.ascii_encode of a 4-bit alphabet (restated, BioSequences v3 / BioSymbols): every symbol of the
    // alphabet -- the gap, A C M G R S V T/U W Y H K D B N in encoding order 0..15 -- in either case
    {
        const char *sym4 = "-ACMGRSVTWYHKDBN";
        for (int e = 0; e < 16; ++e) {
            const char c = sym4[e];
            const char lower = (c >= 'A' && c <= 'Z') ? static_cast<char>(c | 0x20) : c;
            const char cr = c == 'T' ? 'U' : c, lr = c == 'T' ? 'u' : lower;
            l.dna4[static_cast<uint8_t>(c)] = l.dna4[static_cast<uint8_t>(lower)] = static_cast<uint8_t>(e);
            l.rna4[static_cast<uint8_t>(cr)] = l.rna4[static_cast<uint8_t>(lr)] = static_cast<uint8_t>(e);
        }
    }
    const char *dna = "ACGT", *rna = "ACGU";
    for (int c = 0; c < 4; ++c) {
        l.strict_dna[static_cast<uint8_t>(dna[c])] = l.strict_dna[static_cast<uint8_t>(dna[c] | 0x20)] = static_cast<uint8_t>(c);
        l.strict_rna[static_cast<uint8_t>(rna[c])] = l.strict_rna[static_cast<uint8_t>(rna[c] | 0x20)] = static_cast<uint8_t>(c);
    }
    // iterators/common.jl:22-32
    const char *codes[4] = {"Aa", "cC", "gG", "TtUu"};
    for (int c = 0; c < 4; ++c)
        for (const char *q = codes[c]; *q; ++q) l.skipping[static_cast<uint8_t>(*q)] = static_cast<uint8_t>(c);
    for (const char *q = "-MRSVWYHKDBN"; *q; ++q) {
        l.skipping[static_cast<uint8_t>(*q)] = kSkip;
        const char lower = (*q >= 'A' && *q <= 'Z') ? static_cast<char>(*q | 0x20) : *q; // lowercase('-') == '-'
        l.skipping[static_cast<uint8_t>(lower)] = kSkip;
    }
    return l;
}

// The 2-bit tables once more, "positioned": eight 256-entry tables of 32-bit words, one per (byte position in a 32-bit
// word, parity of the word within a pair).  The entry of byte c in table (pos, par) has the byte's three results already
// in the place they take in the result F of an eight-byte pair -- OR-ing the eight entries of a pair IS the recoding:
//   bits  0- 7  the four 2-bit codes of the even word   (code << 2 pos, if par == 0)
//   bits 24-31  the four 2-bit codes of the odd word    (code << 2 pos, if par == 1)
//   bits  8-15  "not a base" flags of the eight bytes   (bit 8 + 4 par + pos)
//   bits 16-23  error flags of the eight bytes          (bit 16 + 4 par + pos)
// so ascii_recode_kernel spends a shift, a mask and a shared-memory load per byte and a handful of byte permutes per 32
// bytes instead of a dozen instructions per byte.
inline void make_positioned(const uint8_t (&lut)[256], uint32_t (&out)[8][256])
{
    for (int par = 0; par < 2; ++par)
        for (int pos = 0; pos < 4; ++pos)
            for (int c = 0; c < 256; ++c) {
                const uint32_t e = lut[c], code = e & 3u, fb = (e >> 6) ? 1u : 0u, fe = e >> 7;
                out[4 * par + pos][c] = ((code << (2 * pos)) << (par ? 24 : 0)) | (fb << (8 + 4 * par + pos)) | (fe << (16 + 4 * par + pos));
            }
}

} // namespace kmc
