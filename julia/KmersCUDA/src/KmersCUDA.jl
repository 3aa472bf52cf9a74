"""
    KmersCUDA

Drop-in GPU (NVIDIA B200, sm_100a) replacement for the k-mer *extraction* path of Kmers.jl:
`collect` of `FwKmers`, `FwRvIterator`, `CanonicalKmers` and `UnambiguousKmers` over
`LongSequence{<:NucleicAcidAlphabet{2|4}}`, plus `fx_hash`.  Every number is computed by
`libkmerscuda.so` (hand-written CUDA) through `ccall`; there is no CUDA.jl codegen and no CPU
fallback.  Results are bit-identical to the reference's own iterators.

NOTE: this module is written against include/kmerscuda.h but could not be executed in the build
environment (no Julia toolchain there); the Python mirror `kmers.jl_b200/kmerscuda` binds the
same symbols and is what the test-suite exercises.
"""
module KmersCUDA

using BioSequences
using Kmers
using Kmers: FwKmers, FwRvIterator, CanonicalKmers, UnambiguousKmers, Kmer, derive_type
using Libdl

export fx_hash_device, hash_device, extract, set_library!, minhash_sketch, composition, minimizers, count_kmers, bucket_count

# ---------------------------------------------------------------------------------------------
# library handle
# ---------------------------------------------------------------------------------------------
const LIB = Ref{String}(get(ENV, "KMERSCUDA_LIB", "libkmerscuda.so"))
set_library!(path::AbstractString) = (LIB[] = String(path))

const KMC_OK = Int32(0)
const KMC_E_BAD_K = Int32(1)
const KMC_E_AMBIGUOUS = Int32(3)
const KMC_FW, KMC_FWRV, KMC_CANON, KMC_UNAMBIG = Int32(0), Int32(1), Int32(2), Int32(3)
const KMC_HASH_FX, KMC_AOS = UInt32(1), UInt32(2)
const KMC_KMER4 = UInt32(0x40)   # k-mers over a 4-bit alphabet (Copyable 4 -> 4, TwoToFour)

# struct kmc_seqs / kmc_out / kmc_result of include/kmerscuda.h (same field order and sizes).  KmcResult is
# mutable and passed as Ref{KmcResult}: ccall hands the library the address of the object itself.
struct KmcSeqs
    words::Ptr{UInt64}
    n_words::UInt64
    n_seqs::UInt64
    seq_word_offset::Ptr{UInt64}
    seq_len::Ptr{UInt64}
    uniform_len::UInt64
    uniform_stride_words::UInt64
    src_bits::UInt32
    first_symbol_offset::UInt32
end

struct KmcOut
    a::Ptr{UInt64}
    b::Ptr{UInt64}
    hash::Ptr{UInt64}
    index::Ptr{Int64}
    seq_out_offset::Ptr{UInt64}
    capacity::UInt64
    index_base::Int64
end

mutable struct KmcResult
    n_written::UInt64
    err_seq::UInt64
    err_pos::UInt64
    err_sym::UInt32
    kernel_ms::Float32
    digest::NTuple{4, UInt64}
    KmcResult() = new(0, 0, 0, 0, 0.0f0, (0, 0, 0, 0))
end

mutable struct Context
    handle::Ptr{Cvoid}
    function Context(device::Integer = 0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        st = ccall((:kmc_ctx_create, LIB[]), Int32, (Int32, Ptr{Ptr{Cvoid}}), device, h)
        st == KMC_OK || error("KmersCUDA: kmc_ctx_create failed ($(status_string(st))); a CUDA device is required, there is no CPU fallback")
        ctx = new(h[])
        finalizer(c -> ccall((:kmc_ctx_destroy, LIB[]), Int32, (Ptr{Cvoid},), c.handle), ctx)
        return ctx
    end
end

status_string(st::Int32) = unsafe_string(ccall((:kmc_status_string, LIB[]), Cstring, (Int32,), st))
last_error(ctx::Context) = unsafe_string(ccall((:kmc_last_error, LIB[]), Cstring, (Ptr{Cvoid},), ctx.handle))

const DEFAULT_CTX = Ref{Union{Nothing, Context}}(nothing)
default_context() = something(DEFAULT_CTX[], (DEFAULT_CTX[] = Context(0)))

# ---------------------------------------------------------------------------------------------
# type parameters -> runtime integers
# ---------------------------------------------------------------------------------------------
src_bits(::Type{<:LongSequence{<:NucleicAcidAlphabet{2}}}) = UInt32(2)
src_bits(::Type{<:LongSequence{<:NucleicAcidAlphabet{4}}}) = UInt32(4)

mode_of(::FwKmers) = KMC_FW
mode_of(::FwRvIterator) = KMC_FWRV
mode_of(::CanonicalKmers) = KMC_CANON
mode_of(::UnambiguousKmers) = KMC_UNAMBIG

source(it::Union{FwKmers, FwRvIterator}) = it.seq
source(it::CanonicalKmers) = it.it.seq
source(it::UnambiguousKmers) = it.it.seq

# FwRvIterator and UnambiguousKmers are not AbstractKmerIterators in the reference
# (CanonicalKmers.jl:25, UnambiguousKmers.jl:29), so the parameters are read per type.
const AnyIter{A, K} = Union{FwKmers{A, K}, FwRvIterator{A, K}, CanonicalKmers{A, K}, UnambiguousKmers{A, K}}
ksize_of(::AnyIter{A, K}) where {A, K} = K
ksize_of(::Type{<:AnyIter{A, K}}) where {A, K} = K
alphabet_of(::Type{<:AnyIter{A, K}}) where {A, K} = A
mode_of(::Type{<:FwKmers}) = KMC_FW
mode_of(::Type{<:FwRvIterator}) = KMC_FWRV
mode_of(::Type{<:CanonicalKmers}) = KMC_CANON
mode_of(::Type{<:UnambiguousKmers}) = KMC_UNAMBIG

function throw_status(ctx::Context, st::Int32, res::KmcResult, ::Type{A}) where {A}
    if st == KMC_E_AMBIGUOUS
        # what src/construction.jl:108-110 throws: EncodeError(Alphabet, symbol)
        sym = A <: RNAAlphabet ? reinterpret(RNA, UInt8(res.err_sym)) : reinterpret(DNA, UInt8(res.err_sym))
        throw(BioSequences.EncodeError(A(), sym))
    elseif st == KMC_E_BAD_K
        error("K must be at least 1")            # src/iterators/FwKmers.jl:32-33
    else
        error("libkmerscuda status $st: $(last_error(ctx))")
    end
end

# ---------------------------------------------------------------------------------------------
# collect(it): one kmc_extract_host call straight into the memory of the result Vector
# ---------------------------------------------------------------------------------------------
"""
    KmersCUDA.collect(it) -> Vector{eltype}

Same result as `Base.collect(it)` for `FwKmers`, `FwRvIterator`, `CanonicalKmers` and
`UnambiguousKmers` whose source is a `LongSequence` over a 2- or 4-bit nucleotide alphabet.  The
k-mer alphabet may be 2-bit (Copyable / FourToTwo) or, for the first three iterators, 4-bit
(Copyable 4 -> 4 and TwoToFour: `KMC_KMER4`, K <= 64).  `Kmer{A,K,N}` and tuples of `Kmer`/`Int`
are isbits, so the device writes the Julia element layout directly into the vector (KMC_AOS).
"""
function collect(it::Union{FwKmers{A, K}, FwRvIterator{A, K}, CanonicalKmers{A, K}};
        ctx::Context = default_context()) where {A <: NucleicAcidAlphabet{4}, K}
    seq = source(it)
    seq isa LongSequence || throw(ArgumentError("4-bit k-mers are accelerated for LongSequence sources"))
    K <= 64 || throw(ArgumentError("KmersCUDA handles K <= 64 for k-mers over a 4-bit alphabet"))
    T = derive_type(Kmer{A, K})
    mode = mode_of(it)
    len = length(seq)
    nwin = max(0, len - K + 1)
    words = seq.data
    ET = mode == KMC_FWRV ? Tuple{T, T} : T
    out = Vector{ET}(undef, nwin)
    res = KmcResult()
    GC.@preserve words out begin
        s = Ref(KmcSeqs(pointer(words), length(words), 1, C_NULL, C_NULL, len, length(words), src_bits(typeof(seq)), 0))
        o = Ref(KmcOut(Ptr{UInt64}(pointer(out)), C_NULL, C_NULL, C_NULL, C_NULL, nwin, 0))
        st = ccall((:kmc_extract_host, LIB[]), Int32,
            (Ptr{Cvoid}, Ptr{KmcSeqs}, Int32, Int32, UInt32, Ptr{KmcOut}, Ref{KmcResult}),
            ctx.handle, s, K, mode, KMC_AOS | KMC_KMER4, o, res)
        st == KMC_OK || throw_status(ctx, st, res, A)
    end
    return out
end

function collect(it::Union{FwKmers{A, K}, FwRvIterator{A, K}, CanonicalKmers{A, K}, UnambiguousKmers{A, K}};
        ctx::Context = default_context()) where {A <: NucleicAcidAlphabet{2}, K}
    seq = source(it)
    seq isa AsciiSource && return collect_ascii(it, seq; ctx)
    seq isa LongSequence || throw(ArgumentError("KmersCUDA accelerates LongSequence and ASCII sources"))
    T = derive_type(Kmer{A, K})
    mode = mode_of(it)
    len = length(seq)
    nwin = max(0, len - K + 1)
    words = seq.data
    bits = src_bits(typeof(seq))
    ET = mode == KMC_FWRV ? Tuple{T, T} : mode == KMC_UNAMBIG ? Tuple{T, Int} : T
    cap = nwin
    if mode == KMC_UNAMBIG && bits == UInt32(4)
        cap = count(it; ctx)   # Base.IteratorSize is SizeUnknown (UnambiguousKmers.jl:33-37)
    end
    out = Vector{ET}(undef, cap)
    res = KmcResult()
    GC.@preserve words out begin
        s = Ref(KmcSeqs(pointer(words), length(words), 1, C_NULL, C_NULL, len, length(words), bits, 0))
        o = Ref(KmcOut(Ptr{UInt64}(pointer(out)), C_NULL, C_NULL, C_NULL, C_NULL, cap, 0))
        st = ccall((:kmc_extract_host, LIB[]), Int32,
            (Ptr{Cvoid}, Ptr{KmcSeqs}, Int32, Int32, UInt32, Ptr{KmcOut}, Ref{KmcResult}),
            ctx.handle, s, K, mode, KMC_AOS, o, res)
        st == KMC_OK || throw_status(ctx, st, res, A)
    end
    resize!(out, res.n_written)
    return out
end

# ASCII sources (the AsciiEncode scheme, src/construction.jl:95-96): String, SubString{String},
# codeunits and byte vectors go to the device as they are (src_bits = 8); nothing is packed on the host.
const AsciiSource = Union{String, SubString{String}, Base.CodeUnits{UInt8, String}, Vector{UInt8}}
ascii_bytes(s::Union{String, SubString{String}}) = codeunits(s)
ascii_bytes(s) = s
# first byte of the source (the caller holds the source with GC.@preserve)
ascii_pointer(s::Union{String, SubString{String}, Vector{UInt8}}) = Ptr{UInt8}(pointer(s))
ascii_pointer(s::Base.CodeUnits{UInt8, String}) = Ptr{UInt8}(pointer(s.s))

function collect_ascii(it::AnyIter{A, K}, src::AsciiSource; ctx::Context = default_context()) where {A <: NucleicAcidAlphabet{2}, K}
    T = derive_type(Kmer{A, K})
    mode = mode_of(it)
    bytes = ascii_bytes(src)
    len = length(bytes)
    ET = mode == KMC_FWRV ? Tuple{T, T} : mode == KMC_UNAMBIG ? Tuple{T, Int} : T
    out = Vector{ET}(undef, max(0, len - K + 1))     # upper bound; UnambiguousKmers may return fewer
    res = KmcResult()
    flags = KMC_AOS | (A <: RNAAlphabet ? UInt32(0x20) : UInt32(0))   # KMC_RNA: U, not T, is the fourth letter
    GC.@preserve src out begin
        s = Ref(KmcSeqs(Ptr{UInt64}(ascii_pointer(src)), len, 1, C_NULL, C_NULL, len, max(len, 1), UInt32(8), 0))
        o = Ref(KmcOut(Ptr{UInt64}(pointer(out)), C_NULL, C_NULL, C_NULL, C_NULL, length(out), 0))
        st = ccall((:kmc_extract_host, LIB[]), Int32,
            (Ptr{Cvoid}, Ptr{KmcSeqs}, Int32, Int32, UInt32, Ptr{KmcOut}, Ref{KmcResult}),
            ctx.handle, s, K, mode, flags, o, res)
        if st == KMC_E_AMBIGUOUS   # FwKmers.jl:124-126 / UnambiguousKmers.jl:123-124
            throw(BioSequences.EncodeError(A(), repr(UInt8(res.err_sym))))
        end
        st == KMC_OK || throw_status(ctx, st, res, A)
    end
    resize!(out, res.n_written)
    return out
end

"Number of elements `collect(it)` returns (runs the device count pass for 4-bit UnambiguousKmers)."
function count(it; ctx::Context = default_context())
    seq = source(it)
    words = seq.data
    n = Ref{UInt64}(0)
    GC.@preserve words begin
        d_words = device_upload(ctx, words)
        s = Ref(KmcSeqs(d_words, length(words), 1, C_NULL, C_NULL, length(seq), length(words), src_bits(typeof(seq)), 0))
        st = ccall((:kmc_count, LIB[]), Int32, (Ptr{Cvoid}, Ptr{KmcSeqs}, Int32, Int32, Ptr{UInt64}),
            ctx.handle, s, ksize_of(it), mode_of(it), n)
        device_free(ctx, d_words)
        st == KMC_OK || error("libkmerscuda status $st: $(last_error(ctx))")
    end
    return Int(n[])
end

function device_upload(ctx::Context, v::Vector{UInt64})
    p = Ref{Ptr{Cvoid}}(C_NULL)
    ccall((:kmc_malloc, LIB[]), Int32, (Ptr{Cvoid}, UInt64, Ptr{Ptr{Cvoid}}), ctx.handle, sizeof(v), p)
    ccall((:kmc_upload, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, UInt64), ctx.handle, p[], v, sizeof(v))
    ccall((:kmc_sync, LIB[]), Int32, (Ptr{Cvoid},), ctx.handle)
    return Ptr{UInt64}(p[])
end
device_free(ctx::Context, p) = ccall((:kmc_free, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), ctx.handle, p)

# ---------------------------------------------------------------------------------------------
# batched extraction over a read set: what `[collect(Iter(r)) for r in reads]` concatenates to
# ---------------------------------------------------------------------------------------------
"""
    extract(Iter, reads::Vector{<:LongSequence}; hash=false) -> (kmers, hashes, offsets)

`Iter` is e.g. `CanonicalDNAMers{31}`.  Packs the reads into one word-aligned CSR buffer
(multithreaded gather on the host) and makes ONE library call.  `offsets[i]+1 : offsets[i+1]`
are the elements of read `i`.
"""
function extract(::Type{I}, reads::Vector{S}; hash::Bool = false,
        ctx::Context = default_context()) where {I <: AnyIter, S <: LongSequence}
    A, K, mode = alphabet_of(I), ksize_of(I), mode_of(I)
    T = derive_type(Kmer{A, K})
    n = length(reads)
    lens = UInt64[length(r) for r in reads]
    nw = UInt64[length(r.data) for r in reads]
    woff = cumsum(vcat(UInt64(0), nw))
    words = Vector{UInt64}(undef, woff[end] + 1)
    Threads.@threads for i in 1:n
        copyto!(words, woff[i] + 1, reads[i].data, 1, nw[i])
    end
    cap = sum(l -> l >= K ? Int(l) - K + 1 : 0, lens; init = 0)
    ET = mode == KMC_FWRV ? Tuple{T, T} : mode == KMC_UNAMBIG ? Tuple{T, Int} : T
    out = Vector{ET}(undef, cap)
    hashes = hash ? Vector{UInt64}(undef, cap) : UInt64[]
    offsets = Vector{UInt64}(undef, n + 1)
    res = KmcResult()
    GC.@preserve words lens woff out hashes offsets begin
        s = Ref(KmcSeqs(pointer(words), length(words), n, pointer(woff), pointer(lens), 0, 0, src_bits(S), 0))
        o = Ref(KmcOut(Ptr{UInt64}(pointer(out)), C_NULL, hash ? pointer(hashes) : C_NULL, C_NULL,
            pointer(offsets), cap, 0))
        st = ccall((:kmc_extract_host, LIB[]), Int32,
            (Ptr{Cvoid}, Ptr{KmcSeqs}, Int32, Int32, UInt32, Ptr{KmcOut}, Ref{KmcResult}),
            ctx.handle, s, K, mode, KMC_AOS | (hash ? KMC_HASH_FX : UInt32(0)), o, res)
        st == KMC_OK || throw_status(ctx, st, res, A)
    end
    resize!(out, res.n_written)
    hash && resize!(hashes, res.n_written)
    return out, hashes, offsets
end

# ---------------------------------------------------------------------------------------------
# fx_hash.(v) on the device (src/kmer.jl:255-261)
# ---------------------------------------------------------------------------------------------
function fx_hash_device(v::Vector{Kmer{A, K, N}}, h::UInt64 = UInt64(0); ctx::Context = default_context()) where {A, K, N}
    out = Vector{UInt64}(undef, length(v))
    isempty(v) && return out
    GC.@preserve v out begin
        dk, dout = Ref{Ptr{Cvoid}}(C_NULL), Ref{Ptr{Cvoid}}(C_NULL)
        ccall((:kmc_malloc, LIB[]), Int32, (Ptr{Cvoid}, UInt64, Ptr{Ptr{Cvoid}}), ctx.handle, sizeof(v), dk)
        ccall((:kmc_malloc, LIB[]), Int32, (Ptr{Cvoid}, UInt64, Ptr{Ptr{Cvoid}}), ctx.handle, sizeof(out), dout)
        ccall((:kmc_upload, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, UInt64), ctx.handle, dk[], v, sizeof(v))
        st = ccall((:kmc_fx_hash, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, UInt64, Int32, UInt64, Ptr{Cvoid}),
            ctx.handle, dk[], length(v), N, h, dout[])
        ccall((:kmc_download, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, UInt64), ctx.handle, out, dout[], sizeof(out))
        ccall((:kmc_free, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), ctx.handle, dk[])
        ccall((:kmc_free, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), ctx.handle, dout[])
        st == KMC_OK || error("libkmerscuda status $st: $(last_error(ctx))")
    end
    return out
end

# Base.hash(kmer, h) = hash(kmer.data, h ⊻ K) (src/kmer.jl:206) depends on the Julia version's tuple
# and integer hashing.  libkmerscuda implements the Julia 1.10 / 1.11 definition (kmc_base_hash); it
# reproduces the reference's documented hash(mer"UGCUGUAC"r) == 0xe5057d38c8907b22
# (docs/src/hashing.md:18-20).  Whether THIS Julia agrees is checked once, on the host, before the
# device path is trusted; otherwise Base.hash is evaluated on the returned k-mers by Julia itself.
const BASE_HASH_ON_DEVICE = Ref{Union{Nothing, Bool}}(nothing)
function base_hash_matches()
    if BASE_HASH_ON_DEVICE[] === nothing
        probe = [mer"UGCUGUAC"r, mer"TAGCTAGGACATTTTAAACCCGGGTAGCTAGGACATTTTAAACC"d]
        BASE_HASH_ON_DEVICE[] = _hash_device(probe, UInt(0)) == hash.(probe)
    end
    return BASE_HASH_ON_DEVICE[]::Bool
end

function _hash_device(v::Vector{Kmer{A, K, N}}, h::UInt; ctx::Context = default_context()) where {A, K, N}
    out = Vector{UInt64}(undef, length(v))
    isempty(v) && return out
    GC.@preserve v out begin
        dk, dout = Ref{Ptr{Cvoid}}(C_NULL), Ref{Ptr{Cvoid}}(C_NULL)
        ccall((:kmc_malloc, LIB[]), Int32, (Ptr{Cvoid}, UInt64, Ptr{Ptr{Cvoid}}), ctx.handle, sizeof(v), dk)
        ccall((:kmc_malloc, LIB[]), Int32, (Ptr{Cvoid}, UInt64, Ptr{Ptr{Cvoid}}), ctx.handle, sizeof(out), dout)
        ccall((:kmc_upload, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, UInt64), ctx.handle, dk[], v, sizeof(v))
        st = ccall((:kmc_base_hash, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, UInt64, Int32, Int32, UInt64, Ptr{Cvoid}),
            ctx.handle, dk[], length(v), N, K, h, dout[])
        ccall((:kmc_download, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, UInt64), ctx.handle, out, dout[], sizeof(out))
        ccall((:kmc_free, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), ctx.handle, dk[])
        ccall((:kmc_free, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), ctx.handle, dout[])
        st == KMC_OK || error("libkmerscuda status $st: $(last_error(ctx))")
    end
    return out
end

"`hash.(v, h)` (Base.hash of k-mers): on the device when this Julia's hashing matches, else on the host."
hash_device(v::Vector{<:Kmer}, h::UInt = UInt(0); ctx::Context = default_context()) =
    base_hash_matches() ? _hash_device(v, h; ctx) : hash.(v, h)

# ---------------------------------------------------------------------------------------------
# consumers of the stream that never materialise it (device-resident sequence, small results)
# ---------------------------------------------------------------------------------------------
function with_device_sequence(f, ctx::Context, seq::LongSequence)
    words = seq.data
    dw = device_upload(ctx, words)
    try
        s = Ref(KmcSeqs(dw, length(words), 1, C_NULL, C_NULL, length(seq), length(words), src_bits(typeof(seq)), 0))
        return f(s)
    finally
        device_free(ctx, dw)
    end
end

"""
    minhash_sketch(CanonicalKmers{A,K}(seq) | FwKmers{A,K}(seq), s) -> Vector{UInt64}

Bottom-`s` MinHash sketch under `fx_hash` -- `sketch(fx_hash, CanonicalDNAMers{K}(seq), s)` of
docs/src/minhash.md:31-36 -- the `s` smallest distinct hash values, ascending (`kmc_minhash_sketch`).
"""
function minhash_sketch(it::Union{FwKmers{A, K}, CanonicalKmers{A, K}}, s::Integer;
        ctx::Context = default_context()) where {A <: NucleicAcidAlphabet{2}, K}
    out = Vector{UInt64}(undef, s)
    res = KmcResult()
    with_device_sequence(ctx, source(it)) do d
        dout = Ref{Ptr{Cvoid}}(C_NULL)
        ccall((:kmc_malloc, LIB[]), Int32, (Ptr{Cvoid}, UInt64, Ptr{Ptr{Cvoid}}), ctx.handle, 8 * s, dout)
        st = ccall((:kmc_minhash_sketch, LIB[]), Int32,
            (Ptr{Cvoid}, Ptr{KmcSeqs}, Int32, Int32, UInt64, Ptr{Cvoid}, Ref{KmcResult}),
            ctx.handle, d, K, mode_of(typeof(it)), s, dout[], res)
        st == KMC_OK && ccall((:kmc_download, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, UInt64),
            ctx.handle, out, dout[], 8 * res.n_written)
        device_free(ctx, dout[])
        st == KMC_OK || throw_status(ctx, st, res, A)
    end
    resize!(out, res.n_written)
    return out
end

"""
    composition(FwKmers{A,K}(seq) | CanonicalKmers{A,K}(seq)) -> Vector{UInt32}  (length 4^K, K <= 14)

`counts[as_integer(kmer) + 1] += 1` for every k-mer (docs/src/composition.md:28-39), `kmc_composition`.
"""
function composition(it::Union{FwKmers{A, K}, CanonicalKmers{A, K}};
        ctx::Context = default_context()) where {A <: NucleicAcidAlphabet{2}, K}
    counts = zeros(UInt32, 4^K)
    res = KmcResult()
    with_device_sequence(ctx, source(it)) do d
        dt = Ref{Ptr{Cvoid}}(C_NULL)
        ccall((:kmc_malloc, LIB[]), Int32, (Ptr{Cvoid}, UInt64, Ptr{Ptr{Cvoid}}), ctx.handle, sizeof(counts), dt)
        ccall((:kmc_memset, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, UInt64), ctx.handle, dt[], 0, sizeof(counts))
        st = ccall((:kmc_composition, LIB[]), Int32,
            (Ptr{Cvoid}, Ptr{KmcSeqs}, Int32, Int32, Ptr{Cvoid}, Ref{KmcResult}),
            ctx.handle, d, K, mode_of(typeof(it)), dt[], res)
        st == KMC_OK && ccall((:kmc_download, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, UInt64),
            ctx.handle, counts, dt[], sizeof(counts))
        device_free(ctx, dt[])
        st == KMC_OK || throw_status(ctx, st, res, A)
    end
    return counts
end

"""
    minimizers(FwKmers{A,K}(seq) | CanonicalKmers{A,K}(seq), W; step = 1) -> (Vector{Kmer}, Vector{Int})

For every window start 1, 1+step, ...: the k-mer with the smallest `fx_hash` among `W` consecutive
k-mers and its 1-based start (docs/src/replacements.md:28-58), `kmc_minimizers`; K <= 32, K + W - 1 <= 64.
"""
function minimizers(it::Union{FwKmers{A, K}, CanonicalKmers{A, K}}, W::Integer; step::Integer = 1,
        ctx::Context = default_context()) where {A <: NucleicAcidAlphabet{2}, K}
    seq = source(it)
    T = derive_type(Kmer{A, K})
    span = K + W - 1
    n = length(seq) >= span ? (length(seq) - span) ÷ step + 1 : 0
    kmers = Vector{T}(undef, n)
    starts = Vector{Int}(undef, n)
    res = KmcResult()
    with_device_sequence(ctx, seq) do d
        dk, di = Ref{Ptr{Cvoid}}(C_NULL), Ref{Ptr{Cvoid}}(C_NULL)
        ccall((:kmc_malloc, LIB[]), Int32, (Ptr{Cvoid}, UInt64, Ptr{Ptr{Cvoid}}), ctx.handle, 8 * max(n, 1), dk)
        ccall((:kmc_malloc, LIB[]), Int32, (Ptr{Cvoid}, UInt64, Ptr{Ptr{Cvoid}}), ctx.handle, 8 * max(n, 1), di)
        o = Ref(KmcOut(Ptr{UInt64}(dk[]), C_NULL, C_NULL, Ptr{Int64}(di[]), C_NULL, n, 0))
        st = ccall((:kmc_minimizers, LIB[]), Int32,
            (Ptr{Cvoid}, Ptr{KmcSeqs}, Int32, Int32, Int32, Int32, UInt32, Ptr{KmcOut}, Ref{KmcResult}),
            ctx.handle, d, K, W, step, mode_of(typeof(it)), UInt32(0), o, res)
        if st == KMC_OK
            ccall((:kmc_download, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, UInt64), ctx.handle, kmers, dk[], 8 * n)
            ccall((:kmc_download, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, UInt64), ctx.handle, starts, di[], 8 * n)
        end
        device_free(ctx, dk[])
        device_free(ctx, di[])
        st == KMC_OK || throw_status(ctx, st, res, A)
    end
    return kmers, starts
end

"""
    count_kmers(FwKmers{A,K}(seq) | CanonicalKmers{A,K}(seq); log2_capacity) -> Dict{Kmer{A,K,1}, Int}

Exact k-mer counts -- the `Dict` a loop over the iterator would build -- through `kmc_kmer_count` (open addressing
in device memory keyed by the k-mer, K <= 31, or K = 32 for canonical k-mers) and `kmc_kmer_table_export`.
`log2_capacity` defaults to twice the number of windows, rounded up to a power of two.
"""
function count_kmers(it::Union{FwKmers{A, K}, CanonicalKmers{A, K}};
        log2_capacity::Integer = max(10, ceil(Int, log2(2 * max(1, length(source(it)))))),
        ctx::Context = default_context()) where {A <: NucleicAcidAlphabet{2}, K}
    K <= 32 || throw(ArgumentError("count_kmers handles one-limb k-mers (K <= 32)"))
    T = derive_type(Kmer{A, K})
    slots = UInt64(1) << log2_capacity
    res = KmcResult()
    keys, vals = UInt64[], UInt32[]
    with_device_sequence(ctx, source(it)) do d
        dk, dv = Ref{Ptr{Cvoid}}(C_NULL), Ref{Ptr{Cvoid}}(C_NULL)
        ccall((:kmc_malloc, LIB[]), Int32, (Ptr{Cvoid}, UInt64, Ptr{Ptr{Cvoid}}), ctx.handle, 8 * slots, dk)
        ccall((:kmc_malloc, LIB[]), Int32, (Ptr{Cvoid}, UInt64, Ptr{Ptr{Cvoid}}), ctx.handle, 4 * slots, dv)
        ccall((:kmc_memset, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, UInt64), ctx.handle, dk[], 0xff, 8 * slots)  # free slot = ~0
        ccall((:kmc_memset, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, UInt64), ctx.handle, dv[], 0, 4 * slots)
        st = ccall((:kmc_kmer_count, LIB[]), Int32,
            (Ptr{Cvoid}, Ptr{KmcSeqs}, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, UInt32, Ref{KmcResult}),
            ctx.handle, d, K, mode_of(typeof(it)), dk[], dv[], UInt32(log2_capacity), res)
        if st == KMC_OK
            n = res.digest[1]                      # keys this call added = every key of a fresh table
            resize!(keys, n); resize!(vals, n)
            ok, ov, nout = Ref{Ptr{Cvoid}}(C_NULL), Ref{Ptr{Cvoid}}(C_NULL), Ref{UInt64}(0)
            ccall((:kmc_malloc, LIB[]), Int32, (Ptr{Cvoid}, UInt64, Ptr{Ptr{Cvoid}}), ctx.handle, 8 * max(n, 1), ok)
            ccall((:kmc_malloc, LIB[]), Int32, (Ptr{Cvoid}, UInt64, Ptr{Ptr{Cvoid}}), ctx.handle, 4 * max(n, 1), ov)
            st = ccall((:kmc_kmer_table_export, LIB[]), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, UInt32, Ptr{Cvoid}, Ptr{Cvoid}, UInt64, Ptr{UInt64}),
                ctx.handle, dk[], dv[], UInt32(log2_capacity), ok[], ov[], n, nout)
            if st == KMC_OK && n > 0
                ccall((:kmc_download, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, UInt64), ctx.handle, keys, ok[], 8 * n)
                ccall((:kmc_download, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, UInt64), ctx.handle, vals, ov[], 4 * n)
            end
            device_free(ctx, ok[])
            device_free(ctx, ov[])
        end
        device_free(ctx, dk[])
        device_free(ctx, dv[])
        st == KMC_OK || throw_status(ctx, st, res, A)
    end
    kmers = reinterpret(T, keys)               # Kmer{A,K,1} is one UInt64 limb
    return Dict{T, Int}(kmers[i] => Int(vals[i]) for i in eachindex(vals))
end

"""
    bucket_count(CanonicalKmers{A,K}(seq), bits) -> Vector{UInt32}   (length 2^bits)

`table[fx_hash(canonical k-mer) >> (64 - bits) + 1] += 1` for every window (`kmc_bucket_count`; the count table of
the multi-GPU configuration, whose per-GPU tables a caller sums with its own collective).
"""
function bucket_count(it::CanonicalKmers{A, K}, bits::Integer;
        ctx::Context = default_context()) where {A <: NucleicAcidAlphabet{2}, K}
    table = zeros(UInt32, 1 << bits)
    res = KmcResult()
    with_device_sequence(ctx, source(it)) do d
        dt = Ref{Ptr{Cvoid}}(C_NULL)
        ccall((:kmc_malloc, LIB[]), Int32, (Ptr{Cvoid}, UInt64, Ptr{Ptr{Cvoid}}), ctx.handle, sizeof(table), dt)
        ccall((:kmc_memset, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, UInt64), ctx.handle, dt[], 0, sizeof(table))
        st = ccall((:kmc_bucket_count, LIB[]), Int32,
            (Ptr{Cvoid}, Ptr{KmcSeqs}, Int32, Int32, Ptr{Cvoid}, Ref{KmcResult}),
            ctx.handle, d, K, bits, dt[], res)
        st == KMC_OK && ccall((:kmc_download, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, UInt64),
            ctx.handle, table, dt[], sizeof(table))
        device_free(ctx, dt[])
        st == KMC_OK || throw_status(ctx, st, res, A)
    end
    return table
end

end # module
