# baseline/julia_ref.jl -- the TRUE reference arm: Kmers.jl's own iterators over the bench.py workload, all host threads.
#
#   JULIA_NUM_THREADS=$(nproc) julia --project=<env with Kmers, BioSequences> baseline/julia_ref.jl [reads=10000000] [steps=5]
#
# bench.py --impl reference runs this instead of the C restatement (oracle/) whenever `julia` is on PATH and Kmers.jl
# loads.  Neither the build image nor the GPU box of this project has Julia, so this file HAS NOT BEEN EXECUTED; it is
# kept so that the reference arm becomes the real reference the moment a toolchain exists.
#
# Workload (SURVEY.md 8d, bench.py): reads of 150 bp whose packed words are word[j] = splitmix64(439824 + j), 5 words
# per read, the trailing bits of the last word zeroed; per read collect CanonicalDNAMers{31} and fx_hash of every
# k-mer (src/iterators/CanonicalKmers.jl:199-225, src/kmer.jl:255-261); output written to preallocated vectors.
using Kmers, BioSequences

const K = 31
const READ_LEN = 150
const STRIDE = 5
const WPR = READ_LEN - K + 1
const SEED = UInt64(439824)

@inline function splitmix64(x::UInt64)
    z = x + 0x9e3779b97f4a7c15
    z = (z ⊻ (z >> 30)) * 0xbf58476d1ce4e5b9
    z = (z ⊻ (z >> 27)) * 0x94d049bb133111eb
    return z ⊻ (z >> 31)
end

function synth_reads(n::Int)
    tail_mask = (UInt64(1) << (2 * (READ_LEN - 32 * (STRIDE - 1)))) - 1
    reads = Vector{LongDNA{2}}(undef, n)
    Threads.@threads for r in 1:n
        data = Vector{UInt64}(undef, STRIDE)
        for w in 1:STRIDE
            data[w] = splitmix64(SEED + UInt64((r - 1) * STRIDE + (w - 1)))
        end
        data[STRIDE] &= tail_mask
        reads[r] = LongDNA{2}(data, UInt(READ_LEN))     # the constructor the reference itself uses, src/construction.jl:299
    end
    return reads
end

function step!(canon::Vector{DNAKmer{K, 1}}, hashes::Vector{UInt64}, reads)
    Threads.@threads for r in eachindex(reads)
        base = (r - 1) * WPR
        i = 0
        for m in CanonicalDNAMers{K}(reads[r])
            i += 1
            @inbounds canon[base + i] = m
            @inbounds hashes[base + i] = fx_hash(m)
        end
    end
    return nothing
end

function main()
    n = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 10_000_000
    steps = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 5
    reads = synth_reads(n)
    canon = Vector{DNAKmer{K, 1}}(undef, n * WPR)
    hashes = Vector{UInt64}(undef, n * WPR)
    step!(canon, hashes, reads)                          # warm-up (compilation, page faults)
    t0 = time_ns()
    for _ in 1:steps
        step!(canon, hashes, reads)
    end
    dt = (time_ns() - t0) / 1e9
    value = n * WPR * steps / dt
    # fingerprints of both streams (xor and wrapping sum): bench.py compares them with the GPU's
    xa = reduce(⊻, (x.data[1] for x in canon); init = UInt64(0))
    sa = reduce(+, (x.data[1] for x in canon); init = UInt64(0))
    xh = reduce(⊻, hashes; init = UInt64(0))
    sh = reduce(+, hashes; init = UInt64(0))
    println("{\"impl\": \"reference\", \"kind\": \"reference\", \"value\": $value, \"unit\": \"kmers/s\", \"cores\": $(Threads.nthreads()), ",
        "\"reads\": $n, \"steps\": $steps, \"ms_per_step\": $(dt / steps * 1e3), ",
        "\"digest\": [$(xa), $(sa), $(xh), $(sh)], \"julia\": \"$(VERSION)\"}")
end

main()
