# B200NFFT.jl -- Julia glue that puts libnfftb200.so behind the AbstractNFFTs plan API.
#
# NOTE: no `julia` binary exists in the build image or on the GPU box, so this file is syntax-reviewed only;
# the executable twin of exactly the same calls is nfft.jl_b200/plan.py (ctypes).  Every ccall below binds an
# entry point of include/nfftb200.h.
#
# Reference interfaces mirrored (all under /root/reference):
#   backend struct / activate! / backend()      src/NFFT.jl:40-42
#   plan_nfft(::Backend, ::Type, k, N; kw...)   src/NFFT.jl:49-58, ext/NFFTGPUArraysExt/implementation.jl:21-30
#   size_in / size_out / nodes! / mul! (x2)     AbstractNFFTs/src/interface.jl:170-211, src/implementation.jl:108-193
#   convolve! family                            AbstractNFFTs/src/interface.jl:217-243
#   foreign-handle finalizer precedent          Wrappers/FINUFFT.jl:52-56
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

const HOST = Cint(0)
const DEVICE = Cint(1)

function check(p::Ptr{Cvoid}, st::Cint)
    st == 0 && return
    msg = unsafe_string(ccall((:nfftb200_last_error, libnfftb200), Cstring, (Ptr{Cvoid},), p))
    if st == 1 || st == 2 || st == 8
        throw(ArgumentError(msg))                  # src/utils.jl:50, src/precomputation.jl:19-21, src/convolution.jl:52
    elseif st == 3
        throw(DimensionMismatch(msg))              # src/utils.jl:101-103
    else
        error("nfftb200 status $st: $msg")
    end
end

mutable struct B200NFFTPlan{T,D} <: AbstractNFFTPlan{T,D,1}
    handle::Ptr{Cvoid}
    N::NTuple{D,Int64}
    NOut::NTuple{1,Int64}
    J::Int64
    k::Matrix{T}
    Ñ::NTuple{D,Int64}
    dims::UnitRange{Int64}
    m::Int
    σ::Float64
    reltol::Float64
    precompute::PrecomputeFlags
    ntransforms::Int
end

# order of the NFFTB200_* window enum in include/nfftb200.h
const WINDOWS = (:kaiser_bessel, :gauss, :spline, :kaiser_bessel_rev, :cosh_type)

dtype_code(::Type{Float32}) = Cint(0)
dtype_code(::Type{Float64}) = Cint(1)

function B200NFFTPlan(k::Matrix{T}, N::NTuple{D,Int}; dims::Union{Integer,UnitRange{Int64}}=1:D,
                      window::Symbol=:kaiser_bessel, precompute::PrecomputeFlags=POLYNOMIAL,
                      ntransforms::Int=1, blockSize=nothing, device::Int=0,
                      sortNodes=false, storeDeconvolutionIdx=false, blocking=true, fftflags=nothing,
                      kwargs...) where {T<:Union{Float32,Float64},D}
    dims == 1:D || error("GPU NFFT does not work along directions right now!")   # ext/...:35-37
    wcode = findfirst(==(window), WINDOWS)                                        # src/windowFunctions.jl:4-19
    wcode === nothing && error("Window $(window) not yet implemented!")
    size(k, 1) == D || throw(ArgumentError("Nodes x have dimension $(size(k,1)) != $D"))
    m, σ, reltol = accuracyParams(; kwargs...)                                    # AbstractNFFTs/src/misc.jl:66-81
    h = Ref{Ptr{Cvoid}}(C_NULL)
    Nv = collect(Int64, N)
    bs = blockSize === nothing ? C_NULL : pointer(collect(Int64, blockSize))
    st = ccall((:nfftb200_plan_create, libnfftb200), Cint,
               (Ref{Ptr{Cvoid}}, Cint, Ptr{Int64}, Cint, Cint, Cdouble, Cint, Cint, Cint, Ptr{Int64}, Cint),
               h, D, Nv, dtype_code(T), m, σ, wcode - 1, Int(precompute), ntransforms, bs, device)
    check(C_NULL, st)
    Ñv = zeros(Int64, D); bsv = zeros(Int64, D)
    nt = Ref{Int64}(0); lut = Ref{Int64}(0); sg = Ref{Cdouble}(0); M = Ref{Int64}(0)
    ccall((:nfftb200_get_info, libnfftb200), Cint,
          (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ref{Int64}, Ref{Int64}, Ref{Cdouble}, Ref{Int64}),
          h[], Ñv, bsv, nt, lut, sg, M)
    p = B200NFFTPlan{T,D}(h[], N, (size(k, 2),), size(k, 2), k, Tuple(Ñv), 1:D, m, sg[], reltol, precompute, ntransforms)
    finalizer(p) do q
        q.handle == C_NULL || ccall((:nfftb200_destroy, libnfftb200), Cint, (Ptr{Cvoid},), q.handle)
        q.handle = C_NULL
    end
    AbstractNFFTs.nodes!(p, k)
    return p
end

function AbstractNFFTs.plan_nfft(::B200Backend, ::Type{<:AbstractArray}, k::Matrix{T}, N::NTuple{D,Int}, rest...;
                                 timing::Union{Nothing,TimingStats}=nothing, kargs...) where {T,D}
    t = @elapsed p = B200NFFTPlan(k, N, rest...; kargs...)
    timing !== nothing && (timing.pre = t)
    return p
end

AbstractNFFTs.size_in(p::B200NFFTPlan) = p.ntransforms == 1 ? p.N : (p.N..., p.ntransforms)
AbstractNFFTs.size_out(p::B200NFFTPlan) = p.ntransforms == 1 ? p.NOut : (p.J, p.ntransforms)

function AbstractNFFTs.nodes!(p::B200NFFTPlan{T}, k::Matrix{T}) where {T}
    st = ccall((:nfftb200_set_nodes, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{T}, Int64, Cint),
               p.handle, k, size(k, 2), HOST)
    check(p.handle, st)
    p.k = k; p.J = size(k, 2); p.NOut = (p.J,)
    return p
end

# device pointers: any array type with `pointer` living on the plan's device (CuArray) passes DEVICE
where(::Array) = HOST
where(::AbstractArray) = DEVICE

function fill_timing!(p, timing::TimingStats)
    t = zeros(Cdouble, 7)
    ccall((:nfftb200_get_timing, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), p.handle, t)
    timing.conv, timing.fft, timing.deconv = t[2], t[3], t[4]
    timing.conv_adjoint, timing.fft_adjoint, timing.deconv_adjoint = t[5], t[6], t[7]
end

function LinearAlgebra.mul!(fHat::AbstractArray{Complex{T}}, p::B200NFFTPlan{T}, f::AbstractArray{Complex{T}};
                            verbose=false, timing::Union{Nothing,TimingStats}=nothing) where {T}
    (size_in(p) == size(f) && size_out(p) == size(fHat)) ||
        throw(DimensionMismatch("Data is not consistent with NFFTPlan"))            # src/utils.jl:98-105
    ccall((:nfftb200_set_timing, libnfftb200), Cint, (Ptr{Cvoid}, Cint), p.handle, timing !== nothing)
    st = ccall((:nfftb200_exec_forward, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
               p.handle, pointer(f), pointer(fHat), where(fHat))
    check(p.handle, st)
    timing !== nothing && fill_timing!(p, timing)
    return fHat
end

function LinearAlgebra.mul!(f::AbstractArray{Complex{T}}, pl::Adjoint{Complex{T},<:B200NFFTPlan{T}},
                            fHat::AbstractArray{Complex{T}}; verbose=false,
                            timing::Union{Nothing,TimingStats}=nothing) where {T}
    p = pl.parent
    (size_in(p) == size(f) && size_out(p) == size(fHat)) ||
        throw(DimensionMismatch("Data is not consistent with NFFTPlan"))
    ccall((:nfftb200_set_timing, libnfftb200), Cint, (Ptr{Cvoid}, Cint), p.handle, timing !== nothing)
    st = ccall((:nfftb200_exec_adjoint, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
               p.handle, pointer(fHat), pointer(f), where(f))
    check(p.handle, st)
    timing !== nothing && fill_timing!(p, timing)
    return f
end

const RealOrComplex{T} = Union{T,Complex{T}}

function AbstractNFFTs.convolve!(p::B200NFFTPlan{T,D}, g::AbstractArray{<:RealOrComplex{T},D},
                                 fHat::AbstractVector{<:RealOrComplex{T}}) where {T,D}
    size(g) == p.Ñ || throw(DimensionMismatch("size(g)=$(size(g)) ≠ Ñ = $(p.Ñ)"))
    size(fHat) == (p.J,) || throw(DimensionMismatch("size(fHat)=$(size(fHat)) ≠ J = $(p.J)"))
    (eltype(g) <: Complex && eltype(fHat) <: Real) &&
        throw(ArgumentError("Complex input g requires Complex output fHat"))        # src/convolution.jl:47-53
    cplx = eltype(fHat) <: Complex
    gg = (cplx && eltype(g) <: Real) ? Complex{T}.(g) : g
    st = ccall((:nfftb200_convolve, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint),
               p.handle, pointer(gg), pointer(fHat), cplx, where(fHat))
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
    st = ccall((:nfftb200_convolve_transpose, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint),
               p.handle, pointer(ff), pointer(g), cplx, where(g))
    check(p.handle, st)
    return g
end

function AbstractNFFTs.deconvolve!(p::B200NFFTPlan{T,D}, f::AbstractArray{Complex{T},D},
                                   g::AbstractArray{Complex{T},D}) where {T,D}
    st = ccall((:nfftb200_deconvolve, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
               p.handle, pointer(f), pointer(g), where(g))
    check(p.handle, st)
    return
end

function AbstractNFFTs.deconvolve_transpose!(p::B200NFFTPlan{T,D}, g::AbstractArray{Complex{T},D},
                                             f::AbstractArray{Complex{T},D}) where {T,D}
    st = ccall((:nfftb200_deconvolve_transpose, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
               p.handle, pointer(g), pointer(f), where(f))
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

function Base.show(io::IO, p::B200NFFTPlan{T,D}) where {T,D}
    print(io, "B200NFFTPlan with ", p.J, " sampling points for an input array of size", p.N,
          " and an output array of size", p.NOut, " with dims ", p.dims)
end

# a plan is single-stream state (src/implementation.jl:26,36); copy re-plans like Base.copy(::NFFTPlan) (:45-66)
Base.copy(p::B200NFFTPlan{T,D}) where {T,D} =
    B200NFFTPlan(p.k, p.N; m=p.m, σ=p.σ, precompute=p.precompute, ntransforms=p.ntransforms)

# ---- sampling density compensation, NFFTTools/src/samplingDensity.jl:59-155 ---------------------------------
"sdc(p; iters=20): Pipe-Menon weights, all iterations on the device (NFFTTools.sdc works too, through convolve!)"
function sdc(p::B200NFFTPlan{T,D}; iters::Int=20) where {T,D}
    w = Vector{T}(undef, p.J)
    check(p.handle, ccall((:nfftb200_sdc, libnfftb200), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint),
                          p.handle, iters, w, HOST))
    return w
end

# ---- Toeplitz (Gram) operator, NFFTTools/src/Toeplitz.jl ------------------------------------------------------
"calculateToeplitzKernel!(f, p, tr, fftplan) (NFFTTools/src/Toeplitz.jl:131-137): the FFT plan lives in the library"
function calculateToeplitzKernel!(f::AbstractArray{Complex{T},D}, p::B200NFFTPlan{T,D}, tr::Matrix{T}, fftplan=nothing) where {T,D}
    AbstractNFFTs.nodes!(p, tr)
    size(f) == p.N || throw(DimensionMismatch("Toeplitz kernel has size $(size(f)) != $(p.N)"))
    check(p.handle, ccall((:nfftb200_toeplitz_kernel, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint),
                          p.handle, pointer(f), where(f)))
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
    check(C_NULL, ccall((:nfftb200_toeplitz_create, libnfftb200), Cint,
                        (Ref{Ptr{Cvoid}}, Cint, Ptr{Int64}, Cint, Cint, Cint),
                        h, D, collect(Int64, shape), dtype_code(T), ntransforms, device))
    op = ToeplitzOperator{T,D}(h[], shape)
    finalizer(op) do q
        q.handle == C_NULL || ccall((:nfftb200_toeplitz_destroy, libnfftb200), Cint, (Ptr{Cvoid},), q.handle)
        q.handle = C_NULL
    end
    check(C_NULL, ccall((:nfftb200_toeplitz_set_kernel, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint),
                        op.handle, pointer(λ), where(λ)))
    return op
end

"convolveToeplitzKernel!(y, λ) with the plans and work arrays held by `op`"
function convolveToeplitzKernel!(y::AbstractArray{Complex{T}}, op::ToeplitzOperator{T}) where {T}
    check(C_NULL, ccall((:nfftb200_toeplitz_apply, libnfftb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint),
                        op.handle, pointer(y), where(y)))
    return y
end
convolveToeplitzKernel!(y::AbstractArray{Complex{T},D}, λ::AbstractArray{Complex{T},D}) where {T,D} =
    convolveToeplitzKernel!(y, ToeplitzOperator(λ))

end # module
