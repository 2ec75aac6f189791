# B200NFFT.jl -- Julia glue that puts libnfftb200.so behind the AbstractNFFTs plan API.
#
# STATUS: UNTESTED.  No `julia` binary exists in the build image or on the GPU box, so this file has never been parsed
# or run; the executable twin of exactly the same calls is nfft.jl_b200/plan.py (ctypes), which the GPU test-suite
# drives.  Every ccall below binds an entry point of include/nfftb200.h with the argument order of that header.
#
# Reference interfaces mirrored (all under /root/reference):
#   backend struct / activate! / backend()      src/NFFT.jl:40-42
#   plan_nfft(::Backend, ::Type, k, N; kw...)   src/NFFT.jl:49-58, ext/NFFTGPUArraysExt/implementation.jl:21-30
#   NFFTPlan fields read by tools               src/implementation.jl:16-43 (N, NOut, J, k, Ñ, dims, params, tmpVec)
#   Base.copy                                   src/implementation.jl:45-66
#   size_in / size_out / nodes! / mul! (x2)     AbstractNFFTs/src/interface.jl:170-211, src/implementation.jl:108-193
#   directional plans (dims=...)                src/directional.jl:58-156, test/accuracy.jl:83-163
#   convolve! family                            AbstractNFFTs/src/interface.jl:217-243
#   foreign-handle finalizer precedent          Wrappers/FINUFFT.jl:52-56
#
# Memory safety: Julia arrays are passed to ccall as arrays (the `Ptr{T}` argument conversion roots them for the
# duration of the call); where a raw address has to be formed first (device arrays, temporaries) the owner is kept
# alive with GC.@preserve.
module B200NFFT

using AbstractNFFTs
using AbstractNFFTs: AbstractNFFTBackend, AbstractNFFTPlan, PrecomputeFlags, TimingStats, accuracyParams,
                     FULL, TENSOR, LINEAR, POLYNOMIAL
using LinearAlgebra
import LinearAlgebra: mul!, Adjoint

export B200Backend, B200NFFTPlan

const libnfftb200 = get(ENV, "NFFTB200_LIB", "libnfftb200.so")

struct B200Backend <: AbstractNFFTBackend end
activate!() = AbstractNFFTs.set_active_backend!(B200NFFT)
backend() = B200Backend()

# `where` argument of the C ABI (include/nfftb200.h)
const HOST = Cint(0)
const DEVICE = Cint(1)
const HOST_ASYNC = Cint(2)

function check(p::Ptr{Cvoid}, st::Integer)
    st == 0 && return nothing
    msg = unsafe_string(ccall((:nfftb200_last_error, libnfftb200), Cstring, (Ptr{Cvoid},), p))
    if st == 1 || st == 2 || st == 8
        throw(ArgumentError(msg))                  # src/utils.jl:50, src/precomputation.jl:19-21, src/convolution.jl:52
    elseif st == 3
        throw(DimensionMismatch(msg))              # src/utils.jl:101-103
    else
        error("nfftb200 status $st: $msg")
    end
end

# the subset of NFFTParams (src/implementation.jl:3-14) that tools read through `p.params`
struct B200Params{T}
    m::Int
    σ::T
    reltol::Float64
    window::Symbol
    LUTSize::Int64
    precompute::PrecomputeFlags
    blockSize::Vector{Int64}
    blocking::Bool
    sortNodes::Bool
    storeDeconvolutionIdx::Bool
end

# D = number of transformed dimensions, DIM = ndims of the user's array (= D unless dims selects a subset)
mutable struct B200NFFTPlan{T,D,DIM,AT<:AbstractArray} <: AbstractNFFTPlan{T,DIM,1}
    handle::Ptr{Cvoid}
    N::NTuple{DIM,Int64}              # user-facing array size
    NOut::Tuple                       # J in place of the first transformed dim, the others removed (src/precomputation.jl:40-50)
    J::Int64
    k::Matrix{T}
    Ñ::NTuple{D,Int64}                # oversampled grid of the transformed dims
    dims::UnitRange{Int64}
    params::B200Params{T}
    ntransforms::Int                  # product of the untransformed dims (x user-requested batch)
    device::Int
    user_block::Union{Nothing,Vector{Int64}}
    tmpVec::AT                        # array-type witness for NFFTTools.sdc (samplingDensity.jl:63-64); the grid itself lives in the library
end

# order of the NFFTB200_* window enum in include/nfftb200.h
const WINDOWS = (:kaiser_bessel, :gauss, :spline, :kaiser_bessel_rev, :cosh_type, :exp_sqrt)   # :exp_sqrt is not in the reference

dtype_code(::Type{Float32}) = Cint(0)
dtype_code(::Type{Float64}) = Cint(1)

function B200NFFTPlan(k::Matrix{T}, N::NTuple{DIM,Int}; dims::Union{Integer,UnitRange{Int64}}=1:DIM,
                      window::Symbol=:kaiser_bessel, precompute::PrecomputeFlags=POLYNOMIAL,
                      ntransforms::Int=1, blockSize=nothing, device::Int=0, arraytype::Type=Array,
                      sortNodes=false, storeDeconvolutionIdx=true, blocking=true, fftflags=nothing,
                      kwargs...) where {T<:Union{Float32,Float64},DIM}
    dimsr = dims isa Integer ? (Int64(dims):Int64(dims)) : dims
    D = length(dimsr)
    (first(dimsr) >= 1 && last(dimsr) <= DIM) || throw(ArgumentError("dims $(dimsr) out of range for an array with $DIM dimensions"))
    size(k, 1) == D || throw(ArgumentError("Nodes x have dimension $(size(k,1)) != $D"))   # src/precomputation.jl:19-21
    wcode = findfirst(==(window), WINDOWS)                                                  # src/windowFunctions.jl:4-19
    wcode === nothing && error("Window $(window) not yet implemented!")
    m, σ, reltol = accuracyParams(; kwargs...)                                              # AbstractNFFTs/src/misc.jl:66-81
    others = [d for d in 1:DIM if !(d in dimsr)]
    B = ntransforms * prod(Int64[N[d] for d in others]; init=Int64(1))
    Nt = Int64[N[d] for d in dimsr]
    ub = blockSize === nothing ? nothing : collect(Int64, blockSize)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    st = if ub === nothing
        ccall((:nfftb200_plan_create, libnfftb200), Cint,
              (Ref{Ptr{Cvoid}}, Cint, Ptr{Int64}, Cint, Cint, Cdouble, Cint, Cint, Cint, Ptr{Int64}, Cint),
              h, D, Nt, dtype_code(T), m, σ, wcode - 1, Int(precompute), B, C_NULL, device)
    else
        ccall((:nfftb200_plan_create, libnfftb200), Cint,
              (Ref{Ptr{Cvoid}}, Cint, Ptr{Int64}, Cint, Cint, Cdouble, Cint, Cint, Cint, Ptr{Int64}, Cint),
              h, D, Nt, dtype_code(T), m, σ, wcode - 1, Int(precompute), B, ub, device)   # `ub` is rooted by the ccall
    end
    check(C_NULL, st)
    Ñv = zeros(Int64, D); bsv = zeros(Int64, D)
    nt = Ref{Int64}(0); lut = Ref{Int64}(0); sg = Ref{Cdouble}(0); M = Ref{Int64}(0)
    ccall((:nfftb200_get_info, libnfftb200), Cint,
          (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ref{Int64}, Ref{Int64}, Ref{Cdouble}, Ref{Int64}),
          h[], Ñv, bsv, nt, lut, sg, M)
    J = size(k, 2)
    params = B200Params{T}(m, T(sg[]), reltol, window, lut[], precompute, bsv, Bool(blocking), Bool(sortNodes),
                           Bool(storeDeconvolutionIdx))
    witness = arraytype{Complex{T},D}(undef, ntuple(_ -> 0, D))
    p = B200NFFTPlan{T,D,DIM,typeof(witness)}(h[], N, out_size(N, dimsr, J), J, k, Tuple(Ñv), dimsr, params, B, device, ub, witness)
    finalizer(p) do q
        q.handle == C_NULL || ccall((:nfftb200_destroy, libnfftb200), Cint, (Ptr{Cvoid},), q.handle)
        q.handle = C_NULL
    end
    AbstractNFFTs.nodes!(p, k)
    return p
end

# NOut of src/precomputation.jl:40-50
function out_size(N::NTuple{DIM,Int}, dims::UnitRange{Int64}, J::Integer) where {DIM}
    out = Int64[]
    taken = false
    for d in 1:DIM
        if !(d in dims)
            push!(out, N[d])
        elseif !taken
            push!(out, J); taken = true
        end
    end
    return Tuple(out)
end

function AbstractNFFTs.plan_nfft(::B200Backend, arr::Type{<:AbstractArray}, k::Matrix{T}, N::NTuple{D,Int}, rest...;
                                 timing::Union{Nothing,TimingStats}=nothing, kargs...) where {T,D}
    t = @elapsed p = B200NFFTPlan(k, N, rest...; kargs...)
    timing !== nothing && (timing.pre = t)
    return p
end

AbstractNFFTs.size_in(p::B200NFFTPlan) = p.N
AbstractNFFTs.size_out(p::B200NFFTPlan) = p.NOut

function AbstractNFFTs.nodes!(p::B200NFFTPlan{T}, k::Matrix{T}) where {T}
    st = ccall((:nfftb200_set_nodes, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{T}, Int64, Cint),
               p.handle, k, size(k, 2), HOST)
    check(p.handle, st)
    p.k = k; p.J = size(k, 2); p.NOut = out_size(p.N, p.dims, p.J)
    return p
end

# ---- where does a buffer live, and its raw address ------------------------------------------------------------
# Any dense array type other than Array (CuArray, ...) is taken to live on the plan's device.  CUDA.jl's pointer() is
# a CuPtr, which converts to an integer but not to Ptr{Cvoid}.
loc(::Array) = HOST
loc(::AbstractArray) = DEVICE
rawptr(x::Array) = Ptr{Cvoid}(pointer(x))
rawptr(x::AbstractArray) = Ptr{Cvoid}(UInt(pointer(x)))

function same_side(a, b)
    loc(a) == loc(b) || throw(ArgumentError("input and output must both be host arrays or both be device arrays"))
    return loc(a)
end

function fill_timing!(p, timing::TimingStats)
    t = zeros(Cdouble, 7)
    ccall((:nfftb200_get_timing, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), p.handle, t)
    timing.conv, timing.fft, timing.deconv = t[2], t[3], t[4]
    timing.conv_adjoint, timing.fft_adjoint, timing.deconv_adjoint = t[5], t[6], t[7]
end

# ---- directional plans: bring the transformed dims to the front (batch slowest), src/directional.jl ------------
is_plain(p::B200NFFTPlan{T,D,DIM}) where {T,D,DIM} = p.dims == 1:D      # leading dims: the library's layout already
others(p::B200NFFTPlan{T,D,DIM}) where {T,D,DIM} = [d for d in 1:DIM if !(d in p.dims)]

function to_internal_image(p::B200NFFTPlan, f)
    is_plain(p) && return f
    return permutedims(f, (collect(p.dims)..., others(p)...))            # (N[dims]..., N[others]...)
end
function from_internal_image!(f, p::B200NFFTPlan, fi)
    is_plain(p) && return f
    permutedims!(f, reshape(fi, (ntuple(i -> p.N[p.dims[i]], length(p.dims))..., ntuple(i -> p.N[others(p)[i]], length(others(p)))...)),
                 invperm([collect(p.dims)..., others(p)...]))
    return f
end
# user layout of fHat: the untransformed dims before dims, J, the untransformed dims after; library layout: (J, others...)
function out_perm(p::B200NFFTPlan)
    npre = first(p.dims) - 1
    no = length(others(p))
    return (collect(2:npre+1)..., 1, collect(npre+2:no+1)...)           # internal (J, o1, o2, ...) -> user order
end
function to_internal_out(p::B200NFFTPlan, fHat)
    is_plain(p) && return fHat
    return permutedims(fHat, invperm(collect(out_perm(p))))
end
function from_internal_out!(fHat, p::B200NFFTPlan, fi)
    is_plain(p) && return fHat
    o = others(p)
    permutedims!(fHat, reshape(fi, (p.J, ntuple(i -> p.N[o[i]], length(o))...)), out_perm(p))
    return fHat
end

function LinearAlgebra.mul!(fHat::AbstractArray{Complex{T}}, p::B200NFFTPlan{T}, f::AbstractArray{Complex{T}};
                            verbose=false, timing::Union{Nothing,TimingStats}=nothing, async::Bool=false) where {T}
    (size_in(p) == size(f) && size_out(p) == size(fHat)) ||
        throw(DimensionMismatch("Data is not consistent with NFFTPlan"))            # src/utils.jl:98-105
    side = same_side(f, fHat)
    plain = is_plain(p)
    (async && (!plain || side != HOST)) && throw(ArgumentError("async needs host arrays and a plan over the leading dims"))
    fi = to_internal_image(p, f)
    fo = plain ? fHat : similar(fHat, (p.J, p.ntransforms))
    ccall((:nfftb200_set_timing, libnfftb200), Cint, (Ptr{Cvoid}, Cint), p.handle, timing !== nothing)
    GC.@preserve fi fo begin
        st = ccall((:nfftb200_exec_forward, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
                   p.handle, rawptr(fi), rawptr(fo), async ? HOST_ASYNC : side)
    end
    check(p.handle, st)
    timing !== nothing && fill_timing!(p, timing)
    return from_internal_out!(fHat, p, fo)
end

function LinearAlgebra.mul!(f::AbstractArray{Complex{T}}, pl::Adjoint{Complex{T},<:B200NFFTPlan{T}},
                            fHat::AbstractArray{Complex{T}}; verbose=false,
                            timing::Union{Nothing,TimingStats}=nothing, async::Bool=false) where {T}
    p = pl.parent
    (size_in(p) == size(f) && size_out(p) == size(fHat)) ||
        throw(DimensionMismatch("Data is not consistent with NFFTPlan"))
    side = same_side(f, fHat)
    plain = is_plain(p)
    (async && (!plain || side != HOST)) && throw(ArgumentError("async needs host arrays and a plan over the leading dims"))
    hi = to_internal_out(p, fHat)
    fo = plain ? f : similar(f, (ntuple(i -> p.N[p.dims[i]], length(p.dims))..., p.ntransforms))
    ccall((:nfftb200_set_timing, libnfftb200), Cint, (Ptr{Cvoid}, Cint), p.handle, timing !== nothing)
    GC.@preserve hi fo begin
        st = ccall((:nfftb200_exec_adjoint, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
                   p.handle, rawptr(hi), rawptr(fo), async ? HOST_ASYNC : side)
    end
    check(p.handle, st)
    timing !== nothing && fill_timing!(p, timing)
    return from_internal_image!(f, p, fo)
end

"wait for every queued transform of the plan (needed after `async=true` calls before the outputs are read)"
sync(p::B200NFFTPlan) = check(p.handle, ccall((:nfftb200_sync, libnfftb200), Cint, (Ptr{Cvoid},), p.handle))

const RealOrComplex{T} = Union{T,Complex{T}}

function AbstractNFFTs.convolve!(p::B200NFFTPlan{T,D}, g::AbstractArray{<:RealOrComplex{T},D},
                                 fHat::AbstractVector{<:RealOrComplex{T}}) where {T,D}
    size(g) == p.Ñ || throw(DimensionMismatch("size(g)=$(size(g)) ≠ Ñ = $(p.Ñ)"))
    size(fHat) == (p.J,) || throw(DimensionMismatch("size(fHat)=$(size(fHat)) ≠ J = $(p.J)"))
    (eltype(g) <: Complex && eltype(fHat) <: Real) &&
        throw(ArgumentError("Complex input g requires Complex output fHat"))        # src/convolution.jl:47-53
    cplx = eltype(fHat) <: Complex
    gg = (cplx && eltype(g) <: Real) ? Complex{T}.(g) : g
    side = same_side(gg, fHat)
    GC.@preserve gg fHat begin
        st = ccall((:nfftb200_convolve, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint),
                   p.handle, rawptr(gg), rawptr(fHat), cplx, side)
    end
    check(p.handle, st)
    return fHat
end

function AbstractNFFTs.convolve_transpose!(p::B200NFFTPlan{T,D}, fHat::AbstractVector{<:RealOrComplex{T}},
                                           g::AbstractArray{<:RealOrComplex{T},D}) where {T,D}
    size(g) == p.Ñ || throw(DimensionMismatch("size(g)=$(size(g)) ≠ Ñ = $(p.Ñ)"))
    size(fHat) == (p.J,) || throw(DimensionMismatch("size(fHat)=$(size(fHat)) ≠ J = $(p.J)"))
    (eltype(fHat) <: Complex && eltype(g) <: Real) &&
        throw(ArgumentError("Complex input fHat requires Complex output g"))        # src/convolution.jl:143-149
    cplx = eltype(g) <: Complex
    ff = (cplx && eltype(fHat) <: Real) ? Complex{T}.(fHat) : fHat
    side = same_side(ff, g)
    GC.@preserve ff g begin
        st = ccall((:nfftb200_convolve_transpose, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint),
                   p.handle, rawptr(ff), rawptr(g), cplx, side)
    end
    check(p.handle, st)
    return g
end

function AbstractNFFTs.deconvolve!(p::B200NFFTPlan{T,D}, f::AbstractArray{Complex{T},D},
                                   g::AbstractArray{Complex{T},D}) where {T,D}
    side = same_side(f, g)
    GC.@preserve f g begin
        st = ccall((:nfftb200_deconvolve, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
                   p.handle, rawptr(f), rawptr(g), side)
    end
    check(p.handle, st)
    return
end

function AbstractNFFTs.deconvolve_transpose!(p::B200NFFTPlan{T,D}, g::AbstractArray{Complex{T},D},
                                             f::AbstractArray{Complex{T},D}) where {T,D}
    side = same_side(g, f)
    GC.@preserve f g begin
        st = ccall((:nfftb200_deconvolve_transpose, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
                   p.handle, rawptr(g), rawptr(f), side)
    end
    check(p.handle, st)
    return
end

"0-based tile-major node permutation (concat of nodesInBlock, src/precomputation.jl:501-504) and tile offsets"
function permutation(p::B200NFFTPlan)
    nt = Ref{Int64}(0)
    ccall((:nfftb200_get_info, libnfftb200), Cint,
          (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ref{Int64}, Ptr{Int64}, Ptr{Cdouble}, Ptr{Int64}),
          p.handle, C_NULL, C_NULL, nt, C_NULL, C_NULL, C_NULL)
    perm = zeros(Int64, p.J); ts = zeros(Int64, nt[] + 1)
    check(p.handle, ccall((:nfftb200_get_permutation, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}),
                          p.handle, perm, ts))
    return perm, ts
end

"device address of the oversampled grid (the reference's p.tmpVec), e.g. for unsafe_wrap(CuArray, CuPtr{Complex{T}}(UInt(ptr)), p.Ñ)"
function grid_pointer(p::B200NFFTPlan)
    ptr = Ref{Ptr{Cvoid}}(C_NULL)
    check(p.handle, ccall((:nfftb200_get_grid, libnfftb200), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), p.handle, ptr))
    return ptr[]
end

function Base.show(io::IO, p::B200NFFTPlan)
    print(io, "B200NFFTPlan with ", p.J, " sampling points for an input array of size", p.N,
          " and an output array of size", p.NOut, " with dims ", p.dims)
end

# a plan is single-stream state (src/implementation.jl:26,36); copy re-plans with every parameter of the original, like
# Base.copy(::NFFTPlan) (:45-66): window, precompute flag, blockSize, dims, device and array type are all carried over
function Base.copy(p::B200NFFTPlan{T,D,DIM,AT}) where {T,D,DIM,AT}
    batch = p.ntransforms ÷ prod(Int64[p.N[d] for d in others(p)]; init=Int64(1))
    return B200NFFTPlan(p.k, p.N; dims=p.dims, m=p.params.m, σ=Float64(p.params.σ), window=p.params.window,
                        precompute=p.params.precompute, ntransforms=batch, blockSize=p.user_block, device=p.device,
                        arraytype=Base.typename(AT).wrapper, sortNodes=p.params.sortNodes, blocking=p.params.blocking)
end

# ---- multi-GPU: one process per GPU, the communicator is bootstrapped from a 128-byte id (include/nfftb200.h) ----
const SHARD_BATCH = Cint(1)
const SHARD_NODES = Cint(2)

"128-byte NCCL unique id; create it on rank 0 and broadcast it with the host's own transport (MPI.jl, Distributed, ...)"
function comm_unique_id()
    id = zeros(UInt8, 128)
    check(C_NULL, ccall((:nfftb200_comm_unique_id, libnfftb200), Cint, (Ptr{UInt8},), id))
    return id
end

"collective: attach the plan to a communicator of `nranks` processes; mode = SHARD_BATCH or SHARD_NODES.  Call nodes! afterwards."
function comm_init!(p::B200NFFTPlan, id::Vector{UInt8}, rank::Integer, nranks::Integer, mode::Integer)
    length(id) == 128 || throw(ArgumentError("the communicator id has 128 bytes"))
    check(p.handle, ccall((:nfftb200_comm_init, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint, Cint),
                          p.handle, id, rank, nranks, mode))
    return p
end

"true if the node-sharded plan exchanges grid data over CUDA-IPC peer memory inside the kernels (no NCCL collective on the path)"
comm_is_fused(p::B200NFFTPlan) = ccall((:nfftb200_comm_is_fused, libnfftb200), Cint, (Ptr{Cvoid},), p.handle) != 0

# ---- sampling density compensation, NFFTTools/src/samplingDensity.jl:59-155 ---------------------------------
"sdc(p; iters=20): Pipe-Menon weights, all iterations on the device (NFFTTools.sdc works too, through convolve!)"
function sdc(p::B200NFFTPlan{T}; iters::Int=20) where {T}
    w = Vector{T}(undef, p.J)
    check(p.handle, ccall((:nfftb200_sdc, libnfftb200), Cint, (Ptr{Cvoid}, Cint, Ptr{T}, Cint),
                          p.handle, iters, w, HOST))
    return w
end

# ---- Toeplitz (Gram) operator, NFFTTools/src/Toeplitz.jl ------------------------------------------------------
"calculateToeplitzKernel!(f, p, tr, fftplan) (NFFTTools/src/Toeplitz.jl:131-137): the FFT plan lives in the library"
function calculateToeplitzKernel!(f::AbstractArray{Complex{T},D}, p::B200NFFTPlan{T,D}, tr::Matrix{T}, fftplan=nothing) where {T,D}
    AbstractNFFTs.nodes!(p, tr)
    size(f) == p.N || throw(DimensionMismatch("Toeplitz kernel has size $(size(f)) != $(p.N)"))
    GC.@preserve f begin
        st = ccall((:nfftb200_toeplitz_kernel, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint), p.handle, rawptr(f), loc(f))
    end
    check(p.handle, st)
    return f
end

"calculateToeplitzKernel(shape, tr; m=4, σ=2.0, window=:kaiser_bessel) (NFFTTools/src/Toeplitz.jl:86-93)"
function calculateToeplitzKernel(shape::NTuple{D,Int}, tr::Matrix{T}; m=4, σ=2.0, window=:kaiser_bessel, kwargs...) where {T,D}
    p = B200NFFTPlan(tr, 2 .* shape; m, σ, window, kwargs...)
    return calculateToeplitzKernel!(Array{Complex{T}}(undef, 2 .* shape), p, tr)
end

"owner of fftplan, ifftplan, xOS1, xOS2 and a device copy of λ for convolveToeplitzKernel! (Toeplitz.jl:230-244)"
mutable struct ToeplitzOperator{T,D}
    handle::Ptr{Cvoid}
    shape::NTuple{D,Int}
end

function ToeplitzOperator(λ::AbstractArray{Complex{T},D}; ntransforms::Int=1, device::Int=0) where {T,D}
    shape = size(λ) .÷ 2
    h = Ref{Ptr{Cvoid}}(C_NULL)
    shp = collect(Int64, shape)
    check(C_NULL, ccall((:nfftb200_toeplitz_create, libnfftb200), Cint,
                        (Ref{Ptr{Cvoid}}, Cint, Ptr{Int64}, Cint, Cint, Cint),
                        h, D, shp, dtype_code(T), ntransforms, device))
    op = ToeplitzOperator{T,D}(h[], shape)
    finalizer(op) do q
        q.handle == C_NULL || ccall((:nfftb200_toeplitz_destroy, libnfftb200), Cint, (Ptr{Cvoid},), q.handle)
        q.handle = C_NULL
    end
    GC.@preserve λ begin
        st = ccall((:nfftb200_toeplitz_set_kernel, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint), op.handle, rawptr(λ), loc(λ))
    end
    check(C_NULL, st)
    return op
end

"convolveToeplitzKernel!(y, λ) with the plans and work arrays held by `op`"
function convolveToeplitzKernel!(y::AbstractArray{Complex{T}}, op::ToeplitzOperator{T}) where {T}
    GC.@preserve y begin
        st = ccall((:nfftb200_toeplitz_apply, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint), op.handle, rawptr(y), loc(y))
    end
    check(C_NULL, st)
    return y
end
convolveToeplitzKernel!(y::AbstractArray{Complex{T},D}, λ::AbstractArray{Complex{T},D}) where {T,D} =
    convolveToeplitzKernel!(y, ToeplitzOperator(λ))

end # module
