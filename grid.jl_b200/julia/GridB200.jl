# GridB200.jl -- Julia (>= 1.6) host side of the B200 engine: Grid.jl's names over `ccall`s into
# libgridb200.so (include/gridb200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no `julia` binary (see DESIGN.md);
# the Python mirror grid.jl_b200/api.py binds the very same entry points with ctypes and is what
# the tests drive.  This file is deliberately thin -- every decision lives behind the C ABI.
#
# Reference surface being preserved (timholy/Grid.jl): src/Grid.jl:28-70 (exports),
# src/interp.jl:42-72,146-306 (InterpGrid, getindex, valgrad, valgradhess), src/coord.jl:3-69
# (CoordInterpGrid), src/interp.jl:909-953 (interp_invert!), src/restrict_prolong.jl:10-185,322-355.
module GridB200

export BoundaryCondition, BCnil, BCnan, BCna, BCreflect, BCperiodic, BCnearest, BCfill,
       InterpType, InterpNearest, InterpLinear, InterpQuadratic, InterpCubic,
       AbstractInterpGrid, InterpGrid, CoordInterpGrid,
       valgrad, valgradhess, getindex!, valgrad!, valgradhess!, valgrad_packed!,
       interp_invert!, pad1, restrict, prolong, restrict_size, prolong_size,
       restrictb, restrict_extrap, prolongb, interp_channels!, InterpIrregular, InterpForward, InterpBackward,
       filledges!, pin!, unpin!, GridComm, replicate, getindex_sharded!, valgrad_sharded!, restrict3_sharded, prolong3_sharded, restrict3_sharded_levels, prolong3_sharded_levels, peer_halo!

const LIB = get(ENV, "GRIDB200_LIB", joinpath(@__DIR__, "..", "libgridb200.so"))

abstract type BoundaryCondition end
struct BCnil <: BoundaryCondition end
struct BCnan <: BoundaryCondition end
struct BCna <: BoundaryCondition end
struct BCreflect <: BoundaryCondition end
struct BCperiodic <: BoundaryCondition end
struct BCnearest <: BoundaryCondition end
struct BCfill <: BoundaryCondition end
abstract type InterpType end
struct InterpNearest <: InterpType end
struct InterpLinear <: InterpType end
struct InterpQuadratic <: InterpType end
struct InterpCubic <: InterpType end

bccode(::Type{BCnil}) = 0; bccode(::Type{BCnan}) = 1; bccode(::Type{BCna}) = 2
bccode(::Type{BCreflect}) = 3; bccode(::Type{BCperiodic}) = 4; bccode(::Type{BCnearest}) = 5; bccode(::Type{BCfill}) = 6
itcode(::Type{InterpNearest}) = 0; itcode(::Type{InterpLinear}) = 1; itcode(::Type{InterpQuadratic}) = 2; itcode(::Type{InterpCubic}) = 3
dtcode(::Type{Float32}) = 0; dtcode(::Type{Float64}) = 1
const GB_HOST, GB_DEVICE = 0, 1
const OUT_VAL, OUT_GRAD, OUT_HESS, OUT_PACKED = 1, 2, 4, 8
const GB_FAST, GB_TILED, GB_PLAIN, GB_SORTED = 1, 2, 4, 8     # gb_eval flags (include/gridb200.h)

lasterror() = unsafe_string(ccall((:gb_last_error, LIB), Cstring, ()))
function check(rc::Cint, bad::Int64 = -1)
    rc == 0 && return
    msg = lasterror()
    rc == 2 && error("Invalid location" * (bad >= 0 ? " (query $(bad + 1))" : ""))   # src/interp.jl:402-404
    rc == 3 && throw(MethodError(getindex, ()))                                      # no method in the reference
    rc == 5 && throw(DimensionMismatch(msg))                                         # src/coord.jl:7
    error(msg)                                                                       # ErrorException
end

abstract type AbstractInterpGrid{T,N,BC<:BoundaryCondition,IT<:InterpType} <: AbstractArray{T,N} end

mutable struct InterpGrid{T<:AbstractFloat,N,BC<:BoundaryCondition,IT<:InterpType} <: AbstractInterpGrid{T,N,BC,IT}
    handle::Ptr{Cvoid}
    dims::NTuple{N,Int}
    fillval::T
end

function _create(A::Array{T,N}, ::Type{BC}, ::Type{IT}, fill::Float64) where {T,N,BC,IT}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    dims = Int64[size(A)...]
    GC.@preserve A dims begin
        check(ccall((:gb_grid_create, LIB), Cint,
                    (Ref{Ptr{Cvoid}}, Cint, Cint, Ptr{Int64}, Ptr{Cvoid}, Cint, Cint, Cint, Cdouble, Ptr{Cvoid}),
                    h, dtcode(T), N, dims, A, GB_HOST, bccode(BC), itcode(IT), fill, C_NULL))
    end
    G = InterpGrid{T,N,BC,IT}(h[], size(A), T(fill))
    finalizer(g -> ccall((:gb_grid_destroy, LIB), Cint, (Ptr{Cvoid},), g.handle), G)
    G
end
# src/interp.jl:48-57 (generic), :58-65 (Cubic pads inside the library), :66-72 (fill value)
function InterpGrid(A::Array{T,N}, ::Type{BC}, ::Type{IT}) where {T<:AbstractFloat,N,BC<:BoundaryCondition,IT<:InterpType}
    BC == BCfill && error("Construct BCfill InterpGrids by supplying the fill value")
    _create(A, BC, IT, NaN)
end
InterpGrid(A::Array{T,N}, f::Number, ::Type{IT}) where {T<:AbstractFloat,N,IT<:InterpType} = _create(A, BCfill, IT, Float64(f))

Base.size(G::InterpGrid) = G.dims

# ---- batched entry points (new; semantics = a loop of the scalar call) -------------------------
# X is N x nq (one query per column), host memory.
function _eval!(G::InterpGrid{T,N,BC}, mask::Int, X::Matrix{XT}, v, g, h; affine = C_NULL, flags::Int = 0) where {T,N,BC,XT<:Union{Float32,Float64}}
    size(X, 1) == N || error("Incorrect number of dimensions supplied")           # src/interp.jl:88-90
    nq = size(X, 2)
    bad = Ref{Int64}(-1)
    GC.@preserve X v g h begin
        rc = ccall((:gb_eval, LIB), Cint,
                   (Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cdouble}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ref{Int64}, Ptr{Cvoid}),
                   G.handle, mask, nq, X, dtcode(XT), N, 1, affine,
                   v === nothing ? C_NULL : pointer(v), g === nothing ? C_NULL : pointer(g), h === nothing ? C_NULL : pointer(h),
                   GB_HOST, flags, bad, C_NULL)
    end
    check(rc, bad[])
end
getindex!(out::Vector{T}, G::InterpGrid{T,N}, X::Matrix) where {T,N} = (length(out) == size(X, 2) || error("output length"); _eval!(G, OUT_VAL, X, out, nothing, nothing); out)
function valgrad!(v::Vector{T}, g::Matrix{T}, G::InterpGrid{T,N}, X::Matrix) where {T,N}
    size(g) == (N, size(X, 2)) || error("Wrong number of components for the gradient")   # src/interp.jl:159-161
    _eval!(G, OUT_VAL | OUT_GRAD, X, v, g, nothing); v
end
# valgrad_packed!(VG, G, X): VG is (N+1) x nq -- VG[1,q] the value, VG[2:end,q] the gradient of query q (GB_OUT_PACKED: one
# full-sector write per query on the device; the library's tiled path skips its unpack pass)
function valgrad_packed!(VG::Matrix{T}, G::InterpGrid{T,N}, X::Matrix) where {T,N}
    size(VG) == (N + 1, size(X, 2)) || error("Wrong number of components for the gradient")
    _eval!(G, OUT_VAL | OUT_GRAD | OUT_PACKED, X, VG, nothing, nothing); VG
end
function valgradhess!(v::Vector{T}, g::Matrix{T}, h::Array{T,3}, G::InterpGrid{T,N}, X::Matrix) where {T,N}
    size(g) == (N, size(X, 2)) || error("Wrong number of components for the gradient")   # :206-208
    size(h) == (N, N, size(X, 2)) || error("Wrong number of components for the hessian") # :209-211
    _eval!(G, OUT_VAL | OUT_GRAD | OUT_HESS, X, v, g, h); v
end

# ---- scalar API (src/interp.jl:146-270) ---------------------------------------------------------
function Base.getindex(G::InterpGrid{T,N}, x::Real...) where {T,N}
    if N == 1 && length(x) == 2
        x[2] == 1 || throw(BoundsError())                                               # :147-153
        x = x[1:1]
    end
    length(x) == N || error("Incorrect number of dimensions supplied")
    out = Vector{T}(undef, 1)
    getindex!(out, G, reshape(Float64[x...], N, 1))[1]
end
function valgrad(G::InterpGrid{T,N}, x::Real...) where {T,N}
    v = Vector{T}(undef, 1); g = Matrix{T}(undef, N, 1)
    valgrad!(v, g, G, reshape(Float64[x...], N, 1))
    N == 1 ? (v[1], g[1]) : (v[1], vec(g))
end
function valgrad(g::AbstractVector{T}, G::InterpGrid{T,N}, x::Real...) where {T,N}
    length(g) == N || error("Wrong number of components for the gradient")
    v, gg = valgrad(G, x...); g .= gg; v
end
function valgradhess(G::InterpGrid{T,N}, x::Real...) where {T,N}
    v = Vector{T}(undef, 1); g = Matrix{T}(undef, N, 1); h = Array{T,3}(undef, N, N, 1)
    valgradhess!(v, g, h, G, reshape(Float64[x...], N, 1))
    N == 1 ? (v[1], g[1], h[1]) : (v[1], vec(g), h[:, :, 1])
end
# tensor-product getindex (src/interp.jl:273-306)
function Base.getindex(G::InterpGrid{T,N}, xs::AbstractVector{<:Real}...) where {T,N}
    length(xs) == 1 && N > 1 && error("Linear indexing not supported")
    length(xs) == N || error("Dimensionality mismatch")
    vecs = [Vector{Float64}(x) for x in xs]
    lens = Int64[length(v) for v in vecs]
    out = Array{T,N}(undef, lens...)
    ptrs = Ptr{Cvoid}[pointer(v) for v in vecs]
    GC.@preserve vecs out ptrs lens check(ccall((:gb_eval_outer, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Int64}, Ptr{Ptr{Cvoid}}, Cint, Ptr{Cdouble}, Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}),
        G.handle, lens, ptrs, 1, C_NULL, out, GB_HOST, 0, C_NULL))
    out
end

# ---- CoordInterpGrid (src/coord.jl) ------------------------------------------------------------------
struct CoordInterpGrid{T,N,BC,IT,R} <: AbstractInterpGrid{T,N,BC,IT}
    coord::R
    grid::InterpGrid{T,N,BC,IT}
    affine::Matrix{Float64}      # 4 x N: kind, start, step, divisor  (coordlookup :30-32)
end
# Julia >= 1.0 float ranges are StepRangeLen; the reference's are Julia 0.4 FloatRange(start, step, len, divisor) with
# INTEGER-valued start / step after the "lifting" of colon(start, step, stop) (base/range.jl of 0.4), and coordlookup
# computes (divisor*x - start)/step + 1 (src/coord.jl:30): -1.0:0.1:1.0 is FloatRange(-10, 1, 21, 10), i.e.
# (10x + 10)/1 + 1 -- NOT (x + 1)/0.1 + 1.  The lifting is reproduced here so that index coordinates carry the same bits.
function _rat(x::Float64)                              # Base.rat of Julia 0.4: continued-fraction rational of a float
    y = x; a = d = 1; b = c = 0
    m = 16777216.0                                     # maxintfloat(Float32)
    while abs(y) <= m
        f = trunc(Int, y); y -= f
        a, c = f * a + c, a
        b, d = f * b + d, b
        max(abs(a), abs(b)) > 16777216 && return c, d
        (b != 0 && Float64(a) / Float64(b) == x) && break
        y == 0 && break
        y = inv(y)
    end
    return a, b
end
function _lift(start::Float64, step::Float64, stop::Float64)   # -> (start, step, len, divisor) of the FloatRange
    step == 0 && error("range step cannot be zero")
    start == stop && return (start, step, 1, 1.0)
    (0 < step) != (start < stop) && return (start, step, 0, 1.0)
    r = (stop - start) / step
    n = round(r)
    lo = prevfloat((prevfloat(stop) - nextfloat(start)) / n)
    hi = nextfloat((nextfloat(stop) - prevfloat(start)) / n)
    if lo <= step <= hi
        a0, b = _rat(start); a = Float64(a0)
        if b != 0 && a / Float64(b) == start
            c0, d = _rat(step); c = Float64(c0)
            if d != 0 && c / Float64(d) == step
                e = lcm(b, d)
                a *= div(e, b); c *= div(e, d)
                (a + n * c) / Float64(e) == stop && return (a, c, Int(n) + 1, Float64(e))
            end
        end
    end
    return (start, step, floor(Int, r) + 1, 1.0)
end
function _row(r::StepRangeLen)
    a, c, len, e = _lift(Float64(first(r)), Float64(step(r)), Float64(last(r)))
    len == length(r) || return Float64[0, Float64(first(r)), Float64(step(r)), 1]   # not a colon-built range: plain affine map
    Float64[0, a, c, e]
end
_row(r::StepRange{<:Integer}) = Float64[1, first(r), step(r), 1]
_row(r::UnitRange{<:Integer}) = Float64[2, first(r), 1, 1]
function CoordInterpGrid(coord::NTuple{N,AbstractRange}, grid::InterpGrid{T,N,BC,IT}) where {T,N,BC,IT}
    map(length, coord) == size(grid) || throw(DimensionMismatch("Coordinate lengths do not match grid size."))
    CoordInterpGrid{T,N,BC,IT,typeof(coord)}(coord, grid, hcat(map(_row, coord)...))
end
CoordInterpGrid(coord::NTuple{N,AbstractRange}, A::Array{T,N}, args...) where {T<:AbstractFloat,N} = CoordInterpGrid(coord, InterpGrid(A, args...))
CoordInterpGrid(coord::AbstractRange, A::Vector{T}, args...) where {T<:AbstractFloat} = CoordInterpGrid((coord,), InterpGrid(A, args...))
Base.size(C::CoordInterpGrid) = size(C.grid)
getindex!(out::Vector{T}, C::CoordInterpGrid{T,N}, X::Matrix) where {T,N} = (_eval!(C.grid, OUT_VAL, X, out, nothing, nothing; affine = pointer(C.affine)); out)
valgrad!(v::Vector{T}, g::Matrix{T}, C::CoordInterpGrid{T,N}, X::Matrix) where {T,N} = (_eval!(C.grid, OUT_VAL | OUT_GRAD, X, v, g, nothing; affine = pointer(C.affine)); v)
Base.getindex(C::CoordInterpGrid{T,N}, x::Real...) where {T,N} = getindex!(Vector{T}(undef, 1), C, reshape(Float64[x...], N, 1))[1]
function valgrad(C::CoordInterpGrid{T,N}, x::Real...) where {T,N}
    v = Vector{T}(undef, 1); g = Matrix{T}(undef, N, 1)
    valgrad!(v, g, C, reshape(Float64[x...], N, 1))
    N == 1 ? (v[1], g[1]) : (v[1], vec(g))
end

# ---- prefilter / padding ----------------------------------------------------------------------------
# fast = true: the line-parallel GB_FAST solver where it applies (few long lines); a few ulp of max|A| instead of bit-identical
function interp_invert!(A::Array{T,N}, ::Type{BC}, ::Type{IT}, dimlist = 1:N; fast::Bool = false) where {T<:Union{Float32,Float64},N,BC<:BoundaryCondition,IT<:InterpType}
    dims = Int64[size(A)...]; dl = Cint[d - 1 for d in dimlist]
    GC.@preserve A dims dl check(ccall((:gb_invert_opt, LIB), Cint,
        (Cint, Cint, Ptr{Int64}, Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cint}, Cint, Cint, Ptr{Cvoid}),
        dtcode(T), N, dims, A, GB_HOST, bccode(BC), itcode(IT), dl, length(dl), fast ? GB_FAST : 0, C_NULL))
    A
end
# filledges!(A, val), src/utilities.jl:22-33
function filledges!(A::Array{T,N}, val) where {T<:Union{Float32,Float64},N}
    dims = Int64[size(A)...]
    GC.@preserve A dims check(ccall((:gb_filledges, LIB), Cint, (Cint, Cint, Ptr{Int64}, Ptr{Cvoid}, Cdouble, Cint, Ptr{Cvoid}),
        dtcode(T), N, dims, A, Float64(val), GB_HOST, C_NULL))
    A
end
# A Julia Array is pageable memory: large host-buffer calls are staged by the library through its own pinned slots with a few
# host threads (C2 batch: 19 ms against 7.5 ms for pinned buffers; INTEGRATION.md).
# Arrays that are passed again and again (query batches, output buffers of an optimisation loop) can be page-locked once:
pin!(A::Array) = (check(ccall((:gb_host_register, LIB), Cint, (Ptr{Cvoid}, Csize_t), A, sizeof(A))); A)
unpin!(A::Array) = (check(ccall((:gb_host_unregister, LIB), Cint, (Ptr{Cvoid},), A)); A)
function pad1(A::Array{T,N}, f::Number) where {T<:Union{Float32,Float64},N}
    out = Array{T,N}(undef, (size(A) .+ 2)...); dims = Int64[size(A)...]
    GC.@preserve A out dims check(ccall((:gb_pad1, LIB), Cint, (Cint, Cint, Ptr{Int64}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
        dtcode(T), N, dims, A, Float64(f), out, GB_HOST, C_NULL))
    out
end

# ---- restrict / prolong (src/restrict_prolong.jl) --------------------------------------------------
restrict_size(len::Int) = isodd(len) ? div(len + 1, 2) : div(len, 2) + 1
function prolong_size(sz::Int, len::Int)
    newsz = isodd(len) ? 2sz - 1 : 2(sz - 1)
    newsz == len || error("Cannot prolong a dimension of length ", sz, " to ", len)
    newsz
end
function restrict(A::Array{T,N}, dim::Integer, scale::Real = one(T)) where {T<:Union{Float32,Float64},N}
    sz = [size(A)...]; sz[dim] = restrict_size(sz[dim]); R = Array{T,N}(undef, sz...); dims = Int64[size(A)...]
    GC.@preserve A R dims check(ccall((:gb_restrict, LIB), Cint, (Cint, Cint, Ptr{Int64}, Ptr{Cvoid}, Cint, Cdouble, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
        dtcode(T), N, dims, A, dim - 1, Float64(scale), R, GB_HOST, C_NULL))
    R
end
# One column of flags that restricts every dimension of a 3-D array is ONE pass over the fine grid (gb_restrict3: the
# intermediates are rounded as the three calls would round them, so the bits are the same).
function restrict3(A::Array{T,3}, scale::Real = one(T)) where {T<:Union{Float32,Float64}}
    R = Array{T,3}(undef, map(restrict_size, size(A))...); dims = Int64[size(A)...]
    GC.@preserve A R dims check(ccall((:gb_restrict3, LIB), Cint, (Cint, Ptr{Int64}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
        dtcode(T), dims, A, Float64(scale), R, GB_HOST, C_NULL))
    R
end
function restrict12(A::Array{T,2}, scale::Real = one(T)) where {T<:Union{Float32,Float64}}
    R = Array{T,2}(undef, map(restrict_size, size(A))...); dims = Int64[size(A)...]
    GC.@preserve A R dims check(ccall((:gb_restrict12, LIB), Cint, (Cint, Cint, Ptr{Int64}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
        dtcode(T), 2, dims, A, Float64(scale), R, GB_HOST, C_NULL))
    R
end
# every column of flags all-true on a 3-D array: all levels in ONE library call, the levels in between stay on the device
function restrict3_levels(A::Array{T,3}, nlevels::Integer, scale::Real = one(T)) where {T<:Union{Float32,Float64}}
    sz = size(A)
    for _ = 1:nlevels
        sz = map(restrict_size, sz)
    end
    R = Array{T,3}(undef, sz...); dims = Int64[size(A)...]
    GC.@preserve A R dims check(ccall((:gb_restrict3_levels, LIB), Cint, (Cint, Ptr{Int64}, Cint, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
        dtcode(T), dims, nlevels, A, Float64(scale), R, GB_HOST, C_NULL))
    R
end
function restrict(A::Array, flag::Union{Array{Bool},BitArray}, scale::Real = one(eltype(A)))     # mapdim, src/utilities.jl:7-20
    ndims(A) == size(flag, 1) || error("Boolean flag must have as many rows as dimensions of A")
    if ndims(A) == 3 && eltype(A) <: Union{Float32,Float64} && size(flag, 2) > 1 && all(flag)
        return restrict3_levels(A, size(flag, 2), scale)
    end
    for j = 1:size(flag, 2)
        if ndims(A) == 3 && eltype(A) <: Union{Float32,Float64} && all(flag[:, j])
            A = restrict3(A, scale)
        elseif ndims(A) == 2 && eltype(A) <: Union{Float32,Float64} && all(flag[:, j]) && minimum(size(A)) >= 2
            A = restrict12(A, scale)         # image pyramids: both dimensions in one pass (gb_restrict12)
        else
            for i = 1:size(flag, 1)
                flag[i, j] && (A = restrict(A, i, scale))
            end
        end
    end
    A
end
function prolong(A::Array{T,N}, dim::Integer, len::Integer) where {T<:Union{Float32,Float64},N}
    sz = [size(A)...]; sz[dim] = prolong_size(sz[dim], len); P = Array{T,N}(undef, sz...); dims = Int64[size(A)...]
    GC.@preserve A P dims check(ccall((:gb_prolong, LIB), Cint, (Cint, Cint, Ptr{Int64}, Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
        dtcode(T), N, dims, A, dim - 1, len, P, GB_HOST, C_NULL))
    P
end
# prolong(A, [s1,s2,s3]) of a 3-D array as one pass (gb_prolong3; dimensions whose target is not larger are kept)
function prolong3(A::Array{T,3}, target::AbstractVector{<:Integer}) where {T<:Union{Float32,Float64}}
    osz = Int64[target[i] > size(A, i) ? prolong_size(size(A, i), Int(target[i])) : size(A, i) for i = 1:3]
    P = Array{T,3}(undef, osz...); dims = Int64[size(A)...]
    GC.@preserve A P dims osz check(ccall((:gb_prolong3, LIB), Cint, (Cint, Ptr{Int64}, Ptr{Cvoid}, Ptr{Int64}, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
        dtcode(T), dims, A, osz, P, GB_HOST, C_NULL))
    P
end
function prolong12(A::Array{T,2}, m1::Integer, m2::Integer) where {T<:Union{Float32,Float64}}
    P = Array{T,2}(undef, prolong_size(size(A, 1), Int(m1)), prolong_size(size(A, 2), Int(m2))); dims = Int64[size(A)...]
    GC.@preserve A P dims check(ccall((:gb_prolong12, LIB), Cint, (Cint, Cint, Ptr{Int64}, Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
        dtcode(T), 2, dims, A, Int64(m1), Int64(m2), P, GB_HOST, C_NULL))
    P
end
# every column of sizes enlarges every dimension of a 3-D array: all levels in ONE library call
function prolong3_levels(A::Array{T,3}, sz::Matrix{Int}) where {T<:Union{Float32,Float64}}
    cur = [size(A)...]
    for j = 1:size(sz, 2), i = 1:3
        cur[i] = prolong_size(cur[i], sz[i, j])
    end
    P = Array{T,3}(undef, cur...); dims = Int64[size(A)...]; od = Int64[sz...]
    GC.@preserve A P dims od check(ccall((:gb_prolong3_levels, LIB), Cint, (Cint, Ptr{Int64}, Cint, Ptr{Int64}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
        dtcode(T), dims, size(sz, 2), od, A, P, GB_HOST, C_NULL))
    P
end
function _all_grow(A::Array, sz::Array{Int})
    cur = [size(A)...]
    for j = 1:size(sz, 2)
        all(sz[:, j] .> cur) || return false
        cur = sz[:, j]
    end
    true
end
function prolong(A::Array, sz::Array{Int})
    ndims(A) == size(sz, 1) || error("Array of sizes must have as many rows as dimensions of A")
    if ndims(A) == 3 && eltype(A) <: Union{Float32,Float64} && ndims(sz) == 2 && size(sz, 2) > 1 && _all_grow(A, sz)
        return prolong3_levels(A, sz)
    end
    for j = 1:size(sz, 2)
        if ndims(A) == 3 && eltype(A) <: Union{Float32,Float64}
            A = prolong3(A, sz[:, j])
        elseif ndims(A) == 2 && eltype(A) <: Union{Float32,Float64} && sz[1, j] > size(A, 1) && sz[2, j] > size(A, 2)
            A = prolong12(A, sz[1, j], sz[2, j])
        else
            for i = 1:size(sz, 1)
                sz[i, j] > size(A, i) && (A = prolong(A, i, sz[i, j]))
            end
        end
    end
    A
end

# ---- restrictb / restrict_extrap / prolongb (src/restrict_prolong.jl:187-320) ------------------------
for (fname, sym) in ((:restrictb, :gb_restrictb), (:restrict_extrap, :gb_restrict_extrap))
    @eval begin
        function $fname(A::Array{T,N}, dim::Integer, scale::Real = one(T)) where {T<:Union{Float32,Float64},N}
            sz = [size(A)...]; sz[dim] = restrict_size(sz[dim]); R = Array{T,N}(undef, sz...); dims = Int64[size(A)...]
            GC.@preserve A R dims check(ccall(($(QuoteNode(sym)), LIB), Cint, (Cint, Cint, Ptr{Int64}, Ptr{Cvoid}, Cint, Cdouble, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
                dtcode(T), N, dims, A, dim - 1, Float64(scale), R, GB_HOST, C_NULL))
            R
        end
        function $fname(A::Array, flag::Union{Array{Bool},BitArray}, scale::Real = one(eltype(A)))
            ndims(A) == size(flag, 1) || error("Boolean flag must have as many rows as dimensions of A")
            for j = 1:size(flag, 2), i = 1:size(flag, 1)
                flag[i, j] && (A = $fname(A, i, scale))
            end
            A
        end
    end
end
function prolongb(A::Array{T,N}, dim::Integer, len::Integer) where {T<:Union{Float32,Float64},N}
    sz = [size(A)...]; sz[dim] = prolong_size(sz[dim], len); P = Array{T,N}(undef, sz...); dims = Int64[size(A)...]
    GC.@preserve A P dims check(ccall((:gb_prolongb, LIB), Cint, (Cint, Cint, Ptr{Int64}, Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
        dtcode(T), N, dims, A, dim - 1, len, P, GB_HOST, C_NULL))
    P
end
function prolongb(A::Array, sz::Array{Int})
    ndims(A) == size(sz, 1) || error("Array of sizes must have as many rows as dimensions of A")
    for j = 1:size(sz, 2), i = 1:size(sz, 1)
        sz[i, j] > size(A, i) && (A = prolongb(A, i, sz[i, j]))
    end
    A
end

# ---- multi-valued grids: the low-level interface, batched (README.md:151-209) -------------------------
# interp_channels!(out, coefs, BC, IT, X): coefs has size (dims..., nchan) and was prepared with
# interp_invert!(coefs, BC, IT, 1:N); X is N x nq in ARRAY index coordinates (what set_position takes);
# out[c, q] = interp(ic, coefs, 1 + (c-1)*stride) after set_position(ic, BC, false, false, X[:, q]).
function interp_channels!(out::Matrix{T}, coefs::Array{T}, ::Type{BC}, ::Type{IT}, X::Matrix{Float64}) where {T<:Union{Float32,Float64},BC,IT}
    N = ndims(coefs) - 1; nchan = size(coefs, N + 1); dims = Int64[size(coefs)[1:N]...]
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve coefs dims check(ccall((:gb_grid_create_from_coefs_channels, LIB), Cint,
        (Ref{Ptr{Cvoid}}, Cint, Cint, Ptr{Int64}, Int64, Ptr{Cvoid}, Cint, Cint, Cint, Cdouble, Cint, Ptr{Cvoid}),
        h, dtcode(T), N, dims, nchan, coefs, GB_HOST, bccode(BC), itcode(IT), NaN, 1, C_NULL))
    bad = Ref{Int64}(-1)
    rc = GC.@preserve X out ccall((:gb_eval, LIB), Cint,
        (Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cdouble}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ref{Int64}, Ptr{Cvoid}),
        h[], OUT_VAL, size(X, 2), X, 1, N, 1, C_NULL, out, C_NULL, C_NULL, GB_HOST, 0, bad, C_NULL)
    ccall((:gb_grid_destroy, LIB), Cint, (Ptr{Cvoid},), h[])
    check(rc, bad[])
    out
end

# ---- InterpIrregular (src/interp.jl:318-377) -----------------------------------------------------------
struct InterpForward <: InterpType end
struct InterpBackward <: InterpType end
itcode(::Type{InterpForward}) = 4; itcode(::Type{InterpBackward}) = 5
struct InterpIrregular{T<:Union{Float32,Float64},BC<:BoundaryCondition,IT<:InterpType}
    grid::Vector{T}
    coefs::Vector{T}
    fillval::T
end
function InterpIrregular(grid::Vector{T}, A::Vector{T}, ::Type{BC}, ::Type{IT}) where {T,BC<:BoundaryCondition,IT<:InterpType}
    (BC == BCreflect || BC == BCperiodic) && error("Sorry, reflecting or periodic boundary conditions not yet supported")
    issorted(grid) || error("Coordinates must be supplied in increasing order")
    InterpIrregular{T,BC,IT}(copy(grid), copy(A), T(NaN))
end
InterpIrregular(grid::Vector{T}, A::Vector{T}, f::Number, ::Type{IT}) where {T,IT<:InterpType} = InterpIrregular{T,BCfill,IT}(copy(grid), copy(A), T(f))
function getindex!(out::Vector{T}, G::InterpIrregular{T,BC,IT}, x::Vector{T}) where {T,BC,IT}
    bad = Ref{Int64}(-1)
    rc = GC.@preserve G x out ccall((:gb_interp_irregular, LIB), Cint,
        (Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cdouble, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ref{Int64}, Ptr{Cvoid}),
        dtcode(T), length(G.grid), G.grid, G.coefs, bccode(BC), itcode(IT), Float64(G.fillval), length(x), x, out, GB_HOST, bad, C_NULL)
    rc == 2 && throw(BoundsError(G.grid, bad[] + 1))   # src/interp.jl:356-360
    check(rc)
    out
end
Base.getindex(G::InterpIrregular{T}, x::Real) where {T} = getindex!(Vector{T}(undef, 1), G, T[x])[1]

# ---- multi-GPU: one Julia process drives every GPU of the box (gb_comm_init_all; SURVEY 8b/8e) -------------------------
# Point evaluation: the coefficients are replicated once (one ncclBroadcast), a host batch is split into contiguous ranges
# over the devices.  restrict / prolong of a 3-D array: slabs along dim 3, one halo exchange per level
# (src/restrict_prolong.jl:181-185, 87-100 spread over the devices; bit-identical to the single-device result).
mutable struct GridComm
    handle::Ptr{Cvoid}
    ndev::Int
end
function GridComm(ndev::Integer)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gb_comm_init_all, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint, Ptr{Cint}), h, ndev, C_NULL))
    c = GridComm(h[], ndev)
    finalizer(x -> ccall((:gb_comm_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle), c)
    c
end
# peer_halo!(comm, on): sharded restrict / prolong read their halo planes in place from the neighbouring devices' slabs through
# peer (NVLink) pointers -- no exchange; on by default where the devices can map each other's memory.  Returns the setting in force.
function peer_halo!(c::GridComm, on::Bool)
    e = Ref{Cint}(0)
    check(ccall((:gb_comm_peer_halo, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{Cint}), c.handle, on ? 1 : 0, e))
    e[] != 0
end
# replicate(comm, G): G lives on device 0; returns one handle per device (the first is G's own)
function replicate(c::GridComm, G::InterpGrid{T,N,BC,IT}) where {T,N,BC,IT}
    hs = fill(C_NULL, c.ndev); hs[1] = G.handle
    stored = Vector{Int64}(undef, 8); dt = Ref{Cint}(0); nd = Ref{Cint}(0)
    check(ccall((:gb_grid_info, LIB), Cint, (Ptr{Cvoid}, Ref{Cint}, Ref{Cint}, Ptr{Int64}, Ptr{Int64}, Ptr{Cint}, Ptr{Cint}, Ptr{Cdouble}),
                G.handle, dt, nd, C_NULL, stored, C_NULL, C_NULL, C_NULL))
    check(ccall((:gb_grid_replicate, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}, Cint, Cint, Ptr{Int64}, Cint, Cint, Cdouble),
                c.handle, 0, hs, dtcode(T), N, stored, bccode(BC), itcode(IT), Float64(G.fillval)))
    hs      # hs[2:end] are owned by the caller: gb_grid_destroy them when done
end
function _eval_sharded!(c::GridComm, hs::Vector{Ptr{Cvoid}}, mask::Int, X::Matrix{Float64}, v, g, h)
    bad = Ref{Int64}(-1)
    rc = GC.@preserve X v g h hs ccall((:gb_eval_sharded, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Cint, Int64, Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ref{Int64}),
        c.handle, hs, mask, size(X, 2), X, 1, C_NULL, v === nothing ? C_NULL : pointer(v), g === nothing ? C_NULL : pointer(g),
        h === nothing ? C_NULL : pointer(h), 0, bad)
    check(rc, bad[])
end
getindex_sharded!(out::Vector, c::GridComm, hs::Vector{Ptr{Cvoid}}, X::Matrix{Float64}) = (_eval_sharded!(c, hs, OUT_VAL, X, out, nothing, nothing); out)
valgrad_sharded!(v::Vector, g::Matrix, c::GridComm, hs::Vector{Ptr{Cvoid}}, X::Matrix{Float64}) = (_eval_sharded!(c, hs, OUT_VAL | OUT_GRAD, X, v, g, nothing); v)
# slabs[i]: device pointer (gb_device_alloc) to planes gb_shard_range(size3, i-1, ndev) of the global array on device i-1;
# the outputs are device pointers the caller allocated for planes gb_shard_range(restrict_size(size3), i-1, ndev)
function restrict3_sharded(c::GridComm, ::Type{T}, gdims::NTuple{3,Int}, slabs::Vector{Ptr{Cvoid}}, outs::Vector{Ptr{Cvoid}}, scale::Real = 1.0) where {T<:Union{Float32,Float64}}
    dims = Int64[gdims...]
    check(ccall((:gb_restrict3_sharded, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Ptr{Cvoid}}, Cdouble, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}),
                c.handle, dtcode(T), dims, slabs, Float64(scale), outs, C_NULL))
    outs
end
function prolong3_sharded(c::GridComm, ::Type{T}, gdims::NTuple{3,Int}, target::NTuple{3,Int}, slabs::Vector{Ptr{Cvoid}}, outs::Vector{Ptr{Cvoid}}) where {T<:Union{Float32,Float64}}
    dims = Int64[gdims...]; odims = Int64[target...]
    check(ccall((:gb_prolong3_sharded, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}),
                c.handle, dtcode(T), dims, slabs, odims, outs, C_NULL))
    outs
end

# Several levels in one call: restrict(A, flags) with `nlevels` all-true columns / prolong(A, sizes) with one column of sizes
# per level (mapdim).  slabs: the devices' planes of the array that enters the first level; outs: device pointers for their
# planes of the LAST level's result.  The levels in between never leave the library.
function restrict3_sharded_levels(c::GridComm, ::Type{T}, gdims::NTuple{3,Int}, nlevels::Integer, slabs::Vector{Ptr{Cvoid}}, outs::Vector{Ptr{Cvoid}},
                                  scale::Real = 1.0) where {T<:Union{Float32,Float64}}
    dims = Int64[gdims...]
    check(ccall((:gb_restrict3_sharded_levels, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}, Cint, Ptr{Ptr{Cvoid}}, Cdouble, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}),
                c.handle, dtcode(T), dims, nlevels, slabs, Float64(scale), outs, C_NULL))
    outs
end
function prolong3_sharded_levels(c::GridComm, ::Type{T}, gdims::NTuple{3,Int}, sizes::Matrix{Int}, slabs::Vector{Ptr{Cvoid}}, outs::Vector{Ptr{Cvoid}}) where {T<:Union{Float32,Float64}}
    size(sizes, 1) == 3 || error("Array of sizes must have as many rows as dimensions of A")
    dims = Int64[gdims...]; sz = Int64[sizes...]
    check(ccall((:gb_prolong3_sharded_levels, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}, Cint, Ptr{Int64}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}),
                c.handle, dtcode(T), dims, size(sizes, 2), sz, slabs, outs, C_NULL))
    outs
end

end # module
