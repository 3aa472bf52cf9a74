# tests/golden/make_julia_fixtures.jl -- dumps, from a REAL Julia + BioSequences + Kmers installation, the facts this
# repository can only pin indirectly today (VERDICT r1: rows a14 / a15):
#   * LongSequence.data words for given strings (the bit layout the device kernels assume),
#   * collect(...) of the four iterators, SpacedKmers and the 4-bit alphabets on seeded inputs,
#   * fx_hash and Base.hash values (one- and two-limb k-mers), with the Julia version that produced them.
#
#   julia --project=<env with Kmers, BioSequences, StableRNGs> tests/golden/make_julia_fixtures.jl > tests/golden/julia_fixtures.json
#
# tests/test_oracle_golden.py picks tests/golden/julia_fixtures.json up when it exists and checks the oracle (and,
# with -m gpu, the device) against every entry.  NOT RUN HERE: there is no Julia in the build image or on the GPU box.
using Kmers, BioSequences, StableRNGs

const RNG = StableRNG(0xccfb2d5055d8c990)      # the reference's test seed, test/runtests.jl:11
hex(x) = "0x" * string(x, base = 16)
limbs(m) = "[" * join((hex(l) for l in m.data), ", ") * "]"
jstr(s) = "\"" * String(s) * "\""

entries = String[]
for len in (0, 1, 11, 31, 32, 33, 64, 150, 1000)
    s = randdnaseq(RNG, len)
    s2, s4 = LongDNA{2}(s), LongDNA{4}(s)
    len > 3 && (s4[len ÷ 2] = DNA_N)
    push!(entries, "{\"kind\": \"layout\", \"seq\": $(jstr(s2)), \"data2\": [" * join(hex.(s2.data), ", ") * "], \"seq4\": $(jstr(s4)), \"data4\": [" *
                   join(hex.(s4.data), ", ") * "]}")
    for K in (1, 5, 31, 32, 33, 63, 64)
        len >= K || continue
        fw = collect(FwDNAMers{K}(s2))
        ca = collect(CanonicalDNAMers{K}(s2))
        un = collect(UnambiguousDNAMers{K}(s4))
        push!(entries, "{\"kind\": \"iterators\", \"seq\": $(jstr(s2)), \"seq4\": $(jstr(s4)), \"k\": $K, \"fw\": [" * join(limbs.(fw), ", ") *
                       "], \"canonical\": [" * join(limbs.(ca), ", ") * "], \"unambiguous\": [" *
                       join(("[" * limbs(m) * ", $i]" for (m, i) in un), ", ") * "], \"fx_hash\": [" * join(hex.(fx_hash.(ca)), ", ") *
                       "], \"base_hash\": [" * join(hex.(hash.(ca)), ", ") * "], \"base_hash_h7\": [" * join(hex.(hash.(ca, UInt(7))), ", ") * "]}")
    end
    for (K, J) in ((3, 2), (2, 4), (3, 3), (31, 7))
        len >= K || continue
        sp = collect(SpacedDNAMers{K, J}(s2))
        sp4 = collect(SpacedKmers{DNAAlphabet{4}, K, J}(s4))
        push!(entries, "{\"kind\": \"spaced\", \"seq\": $(jstr(s2)), \"seq4\": $(jstr(s4)), \"k\": $K, \"j\": $J, \"spaced\": [" * join(limbs.(sp), ", ") *
                       "], \"spaced4\": [" * join(limbs.(sp4), ", ") * "]}")
    end
end
println("{\"julia\": \"$(VERSION)\", \"kmers\": \"$(pkgversion(Kmers))\", \"biosequences\": \"$(pkgversion(BioSequences))\", \"entries\": [")
println(join(entries, ",\n"))
println("]}")
