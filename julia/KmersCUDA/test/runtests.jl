# Parity of KmersCUDA.collect against the reference's own iterators (run where Julia, Kmers.jl,
# BioSequences.jl, libkmerscuda.so and a B200 are all present; not runnable in the build image).
using Test, BioSequences, Kmers, KmersCUDA, StableRNGs

const RNG = StableRNG(0xccfb2d5055d8c990)   # the reference's test seed, test/runtests.jl:11

@testset "collect == reference" begin
    for len in (0, 30, 31, 32, 150, 10_007), K in (1, 5, 31, 32, 33, 63, 64)
        s2 = randdnaseq(RNG, len); s2 = LongDNA{2}(s2)
        for I in (FwDNAMers{K}, CanonicalDNAMers{K}, FwRvIterator{DNAAlphabet{2}, K}, UnambiguousDNAMers{K})
            @test KmersCUDA.collect(I(s2)) == collect(I(s2))
        end
        s4 = LongDNA{4}(s2)
        len > 3 && (s4[len ÷ 2] = DNA_N)
        @test KmersCUDA.collect(UnambiguousDNAMers{K}(s4)) == collect(UnambiguousDNAMers{K}(s4))
        if len >= K && len > 3
            @test_throws BioSequences.EncodeError KmersCUDA.collect(FwDNAMers{K}(s4))
        end
    end
end

@testset "fx_hash" begin
    v = collect(CanonicalDNAMers{31}(LongDNA{2}(randdnaseq(RNG, 1000))))
    @test KmersCUDA.fx_hash_device(v) == fx_hash.(v)
    @test KmersCUDA.fx_hash_device([mer"TAGCTAG"d])[1] == 0xa76409341339d05a   # test/runtests.jl:907
end

@testset "count_kmers == Dict loop" begin
    s = LongDNA{2}(randdnaseq(RNG, 5_000))
    want = Dict{eltype(CanonicalDNAMers{11}(s)), Int}()
    for m in CanonicalDNAMers{11}(s)
        want[m] = get(want, m, 0) + 1
    end
    @test KmersCUDA.count_kmers(CanonicalDNAMers{11}(s)) == want
    t = KmersCUDA.bucket_count(CanonicalDNAMers{11}(s), 12)
    @test sum(t) == length(s) - 10
end
