# Drop-in Julia front end for the B200-native DCOL library (libdcol_b200.so).
#
# Keeps the call surface of kevin-tracy/DifferentiableCollisions.jl v0.1.5 —
#   primitives        src/primitives.jl:4-96, src/primitives_mrp.jl:2-96
#   proximity         src/proximity.jl:3-26
#   proximity_jacobian   src/proximity.jl:168-195
#   proximity_gradient   src/proximity_gradient.jl:80-111
#   helpers           src/problem_matrices.jl:1-27, src/misc_primitive_constructors.jl:1-33
# — and forwards every query through `ccall` to the C ABI of include/dcol_b200.h.  There is no Julia
# solver here and no CPU fallback: without the shared library and a CUDA device `__init__` throws (the only
# arithmetic on the Julia side packs shape tables and, for `verbose = true`, the dense (c, G, h) of one pair that
# `dcol_solve_socp_trace` then solves on the GPU).
#
# NOTE: Julia is not installed in the build image, so this file is written against the C header and
# has NOT been executed; the same ABI is exercised from Python (ctypes) by tests/ and bench.py.
# Set ENV["DCOL_B200_LIB"] to the path of libdcol_b200.so (default: next to this package).
module DifferentiableCollisions

using LinearAlgebra
using StaticArrays

# (The reference exports only its NT-scaling operators `*` and `\`, src/DifferentiableCollisions.jl:36; users call `dc.name`.
#  Those operators belong to the CPU solver and have no counterpart here, so nothing is exported.)

const _lib = Ref{String}("")
const _ctx = Ref{Ptr{Cvoid}}(C_NULL)
const _grp = Ref{Ptr{Cvoid}}(C_NULL)      # device group (dcol_group_*), set when DCOL_B200_DEVICES lists more than one GPU

const ATT_QUAT = Cint(0)
const ATT_MRP = Cint(1)
const MEM_HOST = Cint(0)
const MEM_DEVICE = Cint(1)

abstract type AbstractPrimitive end
abstract type AbstractPrimitiveMRP end
const AnyPrimitive = Union{AbstractPrimitive,AbstractPrimitiveMRP}

# ---------------------------------------------------------------------------------------------
# primitives: same names, fields and defaults as the reference (pose at the origin, identity attitude)
# ---------------------------------------------------------------------------------------------
_q0() = SA[1.0, 0.0, 0.0, 0.0]
_p0() = SA[0.0, 0.0, 0.0]
_r0() = SA[0.0, 0.0, 0.0]
_I3() = SMatrix{3,3,Float64,9}(I)

mutable struct Polytope{n,n3} <: AbstractPrimitive
    r::SVector{3,Float64}; q::SVector{4,Float64}
    A::SMatrix{n,3,Float64,n3}; b::SVector{n,Float64}
    r_offset::SVector{3,Float64}; Q_offset::SMatrix{3,3,Float64,9}
    Polytope(A::SMatrix{n,3,Float64,n3}, b::SVector{n,Float64}) where {n,n3} = new{n,n3}(_r0(), _q0(), A, b, _r0(), _I3())
end
mutable struct PolytopeMRP{n,n3} <: AbstractPrimitiveMRP
    r::SVector{3,Float64}; p::SVector{3,Float64}
    A::SMatrix{n,3,Float64,n3}; b::SVector{n,Float64}
    r_offset::SVector{3,Float64}; Q_offset::SMatrix{3,3,Float64,9}
    PolytopeMRP(A::SMatrix{n,3,Float64,n3}, b::SVector{n,Float64}) where {n,n3} = new{n,n3}(_r0(), _p0(), A, b, _r0(), _I3())
end
mutable struct Polygon{nh,nh2} <: AbstractPrimitive
    r::SVector{3,Float64}; q::SVector{4,Float64}
    A::SMatrix{nh,2,Float64,nh2}; b::SVector{nh,Float64}; R::Float64
    r_offset::SVector{3,Float64}; Q_offset::SMatrix{3,3,Float64,9}
    Polygon(A::SMatrix{nh,2,Float64,nh2}, b::SVector{nh,Float64}, R::Float64) where {nh,nh2} = new{nh,nh2}(_r0(), _q0(), A, b, R, _r0(), _I3())
end
mutable struct PolygonMRP{nh,nh2} <: AbstractPrimitiveMRP
    r::SVector{3,Float64}; p::SVector{3,Float64}
    A::SMatrix{nh,2,Float64,nh2}; b::SVector{nh,Float64}; R::Float64
    r_offset::SVector{3,Float64}; Q_offset::SMatrix{3,3,Float64,9}
    PolygonMRP(A::SMatrix{nh,2,Float64,nh2}, b::SVector{nh,Float64}, R::Float64) where {nh,nh2} = new{nh,nh2}(_r0(), _p0(), A, b, R, _r0(), _I3())
end

for (T, TM, geom) in ((:Capsule, :CapsuleMRP, (:R, :L)), (:Cylinder, :CylinderMRP, (:R, :L)),
                      (:Cone, :ConeMRP, (:H, :β)), (:Sphere, :SphereMRP, (:R,)))
    gf = [:($g::Float64) for g in geom]
    @eval begin
        mutable struct $T <: AbstractPrimitive
            r::SVector{3,Float64}; q::SVector{4,Float64}
            $(gf...)
            r_offset::SVector{3,Float64}; Q_offset::SMatrix{3,3,Float64,9}
            $T($(gf...)) = new(_r0(), _q0(), $(geom...), _r0(), _I3())
        end
        mutable struct $TM <: AbstractPrimitiveMRP
            r::SVector{3,Float64}; p::SVector{3,Float64}
            $(gf...)
            r_offset::SVector{3,Float64}; Q_offset::SMatrix{3,3,Float64,9}
            $TM($(gf...)) = new(_r0(), _p0(), $(geom...), _r0(), _I3())
        end
    end
end

# Ellipsoid: x'Px <= 1; keeps P, U = cholesky(P).U and F = eigen(P) like the reference.  Both
# `Ellipsoid(P)` and `EllipsoidMRP(P)` are constructible (the reference's MRP constructor is
# mis-named, src/primitives_mrp.jl:85).
mutable struct Ellipsoid <: AbstractPrimitive
    r::SVector{3,Float64}; q::SVector{4,Float64}
    P::SMatrix{3,3,Float64,9}; U::SMatrix{3,3,Float64,9}; F::Eigen
    r_offset::SVector{3,Float64}; Q_offset::SMatrix{3,3,Float64,9}
    Ellipsoid(P::SMatrix{3,3,Float64,9}) = new(_r0(), _q0(), P, SMatrix{3,3}(cholesky(Hermitian(P)).U), eigen(Matrix(P)), _r0(), _I3())
end
mutable struct EllipsoidMRP <: AbstractPrimitiveMRP
    r::SVector{3,Float64}; p::SVector{3,Float64}
    P::SMatrix{3,3,Float64,9}; U::SMatrix{3,3,Float64,9}; F::Eigen
    r_offset::SVector{3,Float64}; Q_offset::SMatrix{3,3,Float64,9}
    EllipsoidMRP(P::SMatrix{3,3,Float64,9}) = new(_r0(), _p0(), P, SMatrix{3,3}(cholesky(Hermitian(P)).U), eigen(Matrix(P)), _r0(), _I3())
end

# ---------------------------------------------------------------------------------------------
# attitude helpers (host conveniences used by callers; not on the device path)
# ---------------------------------------------------------------------------------------------
mrp_from_q(q::SVector{4,T}) where {T} = q[SA[2, 3, 4]] / (1 + q[1])
function q_from_mrp(p::SVector{3,T}) where {T}
    pp = dot(p, p)
    return (1 / (1 + pp)) * vcat(1 - pp, 2 * p)
end
function dcm_from_q(q::SVector{4,T}) where {T}
    s, v1, v2, v3 = normalize(q)
    return SA[2v1^2+2s^2-1 2(v1*v2-v3*s) 2(v1*v3+v2*s)
              2(v1*v2+v3*s) 2v2^2+2s^2-1 2(v2*v3-v1*s)
              2(v1*v3-v2*s) 2(v2*v3+v1*s) 2v3^2+2s^2-1]
end
dcm_from_mrp(p::SVector{3,T}) where {T} = dcm_from_q(q_from_mrp(p))

function create_rect_prism(len = 20.0, wid = 20.0, hei = 2.0; attitude = :MRP)
    A = SA[1.0 0 0; 0 1 0; 0 0 1; -1 0 0; 0 -1 0; 0 0 -1]
    b = SA[len / 2, wid / 2, hei / 2, len / 2, wid / 2, hei / 2]
    mass = len * wid * hei
    inertia = (mass / 12) * Diagonal(SA[wid^2 + hei^2, len^2 + hei^2, len^2 + wid^2])
    attitude == :MRP && return PolytopeMRP(A, b), mass, inertia
    attitude == :quat && return Polytope(A, b), mass, inertia
    error("attitude must be :MRP or :quat")
end
function create_n_sided(N, d)
    ns = [[cos(θ), sin(θ)] for θ in 0:(2π / N):(2π * (N - 1) / N)]
    A = vcat(transpose.(ns)...)
    return SMatrix{N,2}(A), SVector{N}(d * ones(N))
end

# mass and diagonal body-frame inertia (x = symmetry axis) of uniform-density cones, capsules and cylinders — the three
# shapes the reference covers (src/mass_properties.jl:2-48); the cylinder keeps the reference's 2R^2 coefficient
function mass_properties(c::Union{Cone,ConeMRP}; ρ = 1.0)
    rad = tan(c.β) * c.H
    m = ρ * π * rad^2 * c.H / 3
    it = m * (3rad^2 / 20 + 3c.H^2 / 80)
    return m, Diagonal(SA[0.3m * rad^2, it, it])
end
function mass_properties(c::Union{Capsule,CapsuleMRP}; ρ = 1.0)
    R, L = c.R, c.L
    mc = ρ * π * R^2 * L
    mh = ρ * 2π * R^3 / 3
    it = mc * (L^2 / 12 + R^2 / 4) + 2mh * (0.4R^2 + L^2 / 2 + 3L * R / 8)
    return mc + 2mh, Diagonal(SA[0.5mc * R^2 + 2mh * 0.4R^2, it, it])
end
function mass_properties(c::Union{Cylinder,CylinderMRP}; ρ = 1.0)
    m = ρ * π * c.R^2 * c.L
    it = m / 12 * (2c.R^2 + c.L^2)
    return m, Diagonal(SA[0.5m * c.R^2, it, it])
end

# ---------------------------------------------------------------------------------------------
# shape table packing (include/dcol_b200.h: dcol_set_shapes)
# ---------------------------------------------------------------------------------------------
_rowmajor(M) = vec(collect(transpose(M)))
_tail(P) = vcat(collect(P.r_offset), _rowmajor(P.Q_offset))
shape_type(::Union{Polytope,PolytopeMRP}) = Cint(0)
shape_type(::Union{Capsule,CapsuleMRP}) = Cint(1)
shape_type(::Union{Cylinder,CylinderMRP}) = Cint(2)
shape_type(::Union{Cone,ConeMRP}) = Cint(3)
shape_type(::Union{Sphere,SphereMRP}) = Cint(4)
shape_type(::Union{Ellipsoid,EllipsoidMRP}) = Cint(5)
shape_type(::Union{Polygon,PolygonMRP}) = Cint(6)
shape_params(P::Union{Polytope,PolytopeMRP}) = vcat(Float64(length(P.b)), _rowmajor(P.A), collect(P.b), _tail(P))
shape_params(P::Union{Capsule,CapsuleMRP,Cylinder,CylinderMRP}) = vcat(P.R, P.L, _tail(P))
shape_params(P::Union{Cone,ConeMRP}) = vcat(P.H, P.β, _tail(P))
shape_params(P::Union{Sphere,SphereMRP}) = vcat(P.R, _tail(P))
shape_params(P::Union{Ellipsoid,EllipsoidMRP}) = vcat(_rowmajor(P.U), _tail(P))
shape_params(P::Union{Polygon,PolygonMRP}) = vcat(Float64(length(P.b)), _rowmajor(P.A), collect(P.b), P.R, _tail(P))
att(P::AbstractPrimitive) = P.q
att(P::AbstractPrimitiveMRP) = P.p
att_kind(::AbstractPrimitive) = ATT_QUAT
att_kind(::AbstractPrimitiveMRP) = ATT_MRP

"Immutable table of shapes; `ShapeTable(prims)` packs geometry once, poses are passed per query."
struct ShapeTable
    types::Vector{Cint}
    offs::Vector{Int64}
    params::Vector{Float64}
end
function ShapeTable(prims::AbstractVector)
    types = Cint[]; offs = Int64[]; params = Float64[]
    for P in prims
        push!(types, shape_type(P)); push!(offs, length(params)); append!(params, shape_params(P))
    end
    return ShapeTable(types, offs, params)
end

_check(rc) = rc == 0 || error("dcol_b200: " * unsafe_string(ccall((:dcol_last_error, _lib[]), Cstring, (Ptr{Cvoid},), _ctx[])))

# Registry behind the single-pair calls: primitives with identical geometry share one shape id, and the device table
# is uploaded again only when a shape was added or a caller's own table replaced it (the reference's examples call
# `proximity` in their inner loops with the same handful of shapes: no upload per call).
const _reg = ShapeTable(Cint[], Int64[], Float64[])
const _reg_ids = Dict{Tuple{Cint,Vector{Float64}},Int32}()
const _reg_on_device = Ref(false)
function _registered!(P)::Int32
    key = (shape_type(P), shape_params(P))
    id = get(_reg_ids, key, Int32(-1))
    if id < 0
        id = Int32(length(_reg.types))
        push!(_reg.types, key[1]); push!(_reg.offs, length(_reg.params)); append!(_reg.params, key[2])
        _reg_ids[key] = id
        _reg_on_device[] = false
    end
    return id
end

function set_shapes!(t::ShapeTable)
    _reg_on_device[] = (t === _reg)
    _check(ccall((:dcol_set_shapes, _lib[]), Cint, (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{Int64}, Ptr{Float64}, Int64),
                 _ctx[], length(t.types), t.types, t.offs, t.params, length(t.params)))
    if _grp[] != C_NULL
        rc = ccall((:dcol_group_set_shapes, _lib[]), Cint, (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{Int64}, Ptr{Float64}, Int64),
                   _grp[], length(t.types), t.types, t.offs, t.params, length(t.params))
        rc == 0 || error("dcol_b200: " * unsafe_string(ccall((:dcol_group_last_error, _lib[]), Cstring, (Ptr{Cvoid},), _grp[])))
    end
    return nothing
end

function __init__()
    _lib[] = get(ENV, "DCOL_B200_LIB", joinpath(@__DIR__, "..", "..", "..", "differentiablecollisions.jl_b200", "libdcol_b200.so"))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:dcol_create, _lib[]), Cint, (Ref{Ptr{Cvoid}}, Cint), h, parse(Cint, get(ENV, "DCOL_B200_DEVICE", "0")))
    rc == 0 || error("dcol_b200: " * unsafe_string(ccall((:dcol_last_error, _lib[]), Cstring, (Ptr{Cvoid},), C_NULL)))
    _ctx[] = h[]
    atexit(() -> ccall((:dcol_destroy, _lib[]), Cvoid, (Ptr{Cvoid},), _ctx[]))
    # DCOL_B200_DEVICES=0,1,2,...: the batched queries (proximity_batch, ...) are cut into one index slice per listed GPU
    # and run concurrently from this process (dcol_group_query); single-pair calls and dcol_se2_constraints stay on _ctx
    devs = Cint[parse(Cint, d) for d in split(get(ENV, "DCOL_B200_DEVICES", ""), ','; keepempty = false)]
    if length(devs) > 1
        g = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:dcol_group_create, _lib[]), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cint}, Cint), g, devs, length(devs))
        rc == 0 || error("dcol_b200: " * unsafe_string(ccall((:dcol_group_last_error, _lib[]), Cstring, (Ptr{Cvoid},), C_NULL)))
        _grp[] = g[]
        atexit(() -> ccall((:dcol_group_destroy, _lib[]), Cvoid, (Ptr{Cvoid},), _grp[]))
    end
end

# ---------------------------------------------------------------------------------------------
# batched API (new): arrays of poses, Julia column-major => r is 3 x n, att is (4|3) x n,
# J comes back as 4 x nt x n — exactly the C layout, no transposes.
# ---------------------------------------------------------------------------------------------
struct BatchResult
    α::Vector{Float64}
    x::Matrix{Float64}              # 3 x n  (empty for gradient queries)
    deriv::Array{Float64}           # gradient: nt x n ; jacobian: 4 x nt x n ; proximity: empty
    status::Vector{Int32}           # 0 ok, 1 "pdip failed", 2 numerical breakdown
    iters::Vector{Int32}
end

function _query(sym::Symbol, shape1::Vector{Int32}, shape2::Vector{Int32}, r1, a1, r2, a2, kind::Cint, pdip_tol::Float64)
    n = length(shape1)
    nt = 6 + 2 * size(a1, 1)
    α = Vector{Float64}(undef, n); status = Vector{Int32}(undef, n); iters = Vector{Int32}(undef, n)
    x = Matrix{Float64}(undef, 3, sym === :dcol_proximity_gradient ? 0 : n)
    if _grp[] != C_NULL            # several GPUs: contiguous index slices, one per device, same results
        q = sym === :dcol_proximity ? 0 : (sym === :dcol_proximity_gradient ? 1 : 2)
        deriv = q == 0 ? Float64[] : (q == 1 ? Matrix{Float64}(undef, nt, n) : Array{Float64}(undef, 4, nt, n))
        rc = ccall((:dcol_group_query, _lib[]), Cint,
                   (Ptr{Cvoid}, Cint, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Float64,
                    Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                   _grp[], q, n, shape1, shape2, r1, a1, r2, a2, kind, pdip_tol, α, q == 1 ? C_NULL : x, q == 0 ? C_NULL : deriv, status, iters)
        rc == 0 || error("dcol_b200: " * unsafe_string(ccall((:dcol_group_last_error, _lib[]), Cstring, (Ptr{Cvoid},), _grp[])))
        return BatchResult(α, x, deriv, status, iters)
    end
    if sym === :dcol_proximity
        deriv = Float64[]
        _check(ccall((:dcol_proximity, _lib[]), Cint,
                     (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Float64, Cint,
                      Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                     _ctx[], n, shape1, shape2, r1, a1, r2, a2, kind, pdip_tol, MEM_HOST, α, x, status, iters))
    elseif sym === :dcol_proximity_gradient
        deriv = Matrix{Float64}(undef, nt, n)
        _check(ccall((:dcol_proximity_gradient, _lib[]), Cint,
                     (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Float64, Cint,
                      Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                     _ctx[], n, shape1, shape2, r1, a1, r2, a2, kind, pdip_tol, MEM_HOST, α, deriv, status, iters))
    else
        deriv = Array{Float64}(undef, 4, nt, n)
        _check(ccall((:dcol_proximity_jacobian, _lib[]), Cint,
                     (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Float64, Cint,
                      Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                     _ctx[], n, shape1, shape2, r1, a1, r2, a2, kind, pdip_tol, MEM_HOST, α, x, deriv, status, iters))
    end
    return BatchResult(α, x, deriv, status, iters)
end

"`proximity_batch(shape1, shape2, R1, Q1, R2, Q2; pdip_tol)`: shape ids are 0-based rows of the ShapeTable passed to `set_shapes!`."
proximity_batch(s1, s2, r1::Matrix{Float64}, a1::Matrix{Float64}, r2::Matrix{Float64}, a2::Matrix{Float64}; pdip_tol = 1e-6) =
    _query(:dcol_proximity, Int32.(s1), Int32.(s2), r1, a1, r2, a2, size(a1, 1) == 4 ? ATT_QUAT : ATT_MRP, Float64(pdip_tol))
proximity_gradient_batch(s1, s2, r1::Matrix{Float64}, a1::Matrix{Float64}, r2::Matrix{Float64}, a2::Matrix{Float64}; pdip_tol = 1e-6) =
    _query(:dcol_proximity_gradient, Int32.(s1), Int32.(s2), r1, a1, r2, a2, size(a1, 1) == 4 ? ATT_QUAT : ATT_MRP, Float64(pdip_tol))
proximity_jacobian_batch(s1, s2, r1::Matrix{Float64}, a1::Matrix{Float64}, r2::Matrix{Float64}, a2::Matrix{Float64}; pdip_tol = 1e-6) =
    _query(:dcol_proximity_jacobian, Int32.(s1), Int32.(s2), r1, a1, r2, a2, size(a1, 1) == 4 ? ATT_QUAT : ATT_MRP, Float64(pdip_tol))

"""
    scene_allpairs(kind, shape, r, att; pdip_tol) -> BatchResult

All `i<j` pairs of a scene in `combinations(1:N,2)` order (the pair list of the contact example); `kind` 0/1/2 =
proximity / gradient / jacobian, `shape` 0-based shape ids, `r` 3×N, `att` (4|3)×N.
"""
function scene_allpairs(kind::Integer, shape::Vector{Int32}, r::Matrix{Float64}, att::Matrix{Float64}; pdip_tol = 1e-6)
    nb = length(shape); na = size(att, 1); nt = 6 + 2na
    n = Int(ccall((:dcol_scene_pair_count, _lib[]), Int64, (Cint,), nb))
    α = Vector{Float64}(undef, n); status = Vector{Int32}(undef, n); iters = Vector{Int32}(undef, n)
    x = Matrix{Float64}(undef, 3, kind == 1 ? 0 : n)
    deriv = kind == 0 ? Float64[] : (kind == 1 ? Matrix{Float64}(undef, nt, n) : Array{Float64}(undef, 4, nt, n))
    if _grp[] != C_NULL      # several GPUs (DCOL_B200_DEVICES): every member solves one slice of the pair enumeration
        rc = ccall((:dcol_group_scene_allpairs, _lib[]), Cint,
                   (Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Cint, Float64,
                    Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                   _grp[], kind, nb, shape, r, att, na == 4 ? ATT_QUAT : ATT_MRP, Float64(pdip_tol),
                   α, kind == 1 ? C_NULL : x, kind == 0 ? C_NULL : deriv, status, iters)
        rc == 0 || error("dcol_b200: " * unsafe_string(ccall((:dcol_group_last_error, _lib[]), Cstring, (Ptr{Cvoid},), _grp[])))
        return BatchResult(α, x, deriv, status, iters)
    end
    _check(ccall((:dcol_scene_allpairs, _lib[]), Cint,
                 (Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Cint, Float64, Cint,
                  Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                 _ctx[], kind, nb, shape, r, att, na == 4 ? ATT_QUAT : ATT_MRP, Float64(pdip_tol), MEM_HOST,
                 α, kind == 1 ? C_NULL : x, kind == 0 ? C_NULL : deriv, status, iters))
    return BatchResult(α, x, deriv, status, iters)
end

"""
    scene_wrench(shape, r, att, λ; pdip_tol) -> (α, wrench, status, iters)

The accumulation loop of the contact example (`DELs[i] += E[1:6]; DELs[j] += E[7:12]` over all pairs, without its
constant factors `h * τ_mod`) as one call: `wrench[:, b] = Σ_k λ[k] * [∂α_k/∂r_b ; G(q_b)' ∂α_k/∂q_b]` over the pairs `k`
containing body `b`, `G(q) = L(q) * H`; for MRP attitudes the attitude part is the plain sum of `∂α_k/∂p_b`.
`λ` has one entry per pair in `combinations(1:N,2)` order.  The per-pair gradient rows stay on the GPU.
"""
function scene_wrench(shape::Vector{Int32}, r::Matrix{Float64}, att::Matrix{Float64}, λ::Vector{Float64}; pdip_tol = 1e-6)
    nb = length(shape); na = size(att, 1)
    n = Int(ccall((:dcol_scene_pair_count, _lib[]), Int64, (Cint,), nb))
    length(λ) == n || error("scene_wrench: λ must have one entry per pair ($n)")
    α = Vector{Float64}(undef, n); status = Vector{Int32}(undef, n); iters = Vector{Int32}(undef, n)
    wrench = Matrix{Float64}(undef, 6, nb)
    if _grp[] != C_NULL      # per-body partial sums of the members' pair slices, added in member order
        rc = ccall((:dcol_group_scene_wrench, _lib[]), Cint,
                   (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Float64,
                    Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                   _grp[], nb, shape, r, att, λ, na == 4 ? ATT_QUAT : ATT_MRP, Float64(pdip_tol), α, wrench, status, iters)
        rc == 0 || error("dcol_b200: " * unsafe_string(ccall((:dcol_group_last_error, _lib[]), Cstring, (Ptr{Cvoid},), _grp[])))
        return α, wrench, status, iters
    end
    _check(ccall((:dcol_scene_wrench, _lib[]), Cint,
                 (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Float64, Cint,
                  Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                 _ctx[], nb, shape, r, att, λ, na == 4 ? ATT_QUAT : ATT_MRP, Float64(pdip_tol), MEM_HOST,
                 α, wrench, status, iters))
    return α, wrench, status, iters
end

"""
    se2_constraints(xyθ, vehicle_shape, obs_shape, r_obs, p_obs; pdip_tol) -> (c, Jc, status)

`ineq_con_x` / `ineq_con_x_jac` of the piano-mover example for a batch of planar states: column `s` of `xyθ` (3×n) puts
the vehicle at `r = (x, y, 0)`, `p = (0, 0, tan(θ/4))`; `c[o, s] = 1 - α` against obstacle `o` and `Jc[:, o, s] = ∂c/∂(x, y, θ)`
(columns 1, 2 and 5 of the example's Jacobian rows).  Shape ids are 0-based.
"""
function se2_constraints(xyθ::Matrix{Float64}, vehicle_shape::Integer, obs_shape::Vector{Int32}, r_obs::Matrix{Float64},
                         p_obs::Matrix{Float64}; pdip_tol = 1e-6)
    n = size(xyθ, 2); K = length(obs_shape)
    c = Matrix{Float64}(undef, K, n); Jc = Array{Float64}(undef, 3, K, n); status = Matrix{Int32}(undef, K, n)
    _check(ccall((:dcol_se2_constraints, _lib[]), Cint,
                 (Ptr{Cvoid}, Int64, Ptr{Float64}, Int32, Cint, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Float64, Cint,
                  Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                 _ctx[], n, xyθ, vehicle_shape, K, obs_shape, r_obs, p_obs, Float64(pdip_tol), MEM_HOST, c, Jc, status))
    return c, Jc, status
end

"""
    se3_constraints(X, vehicle_shape, obs_shape, r_obs, p_obs; r_rows = 1:3, p_rows = 7:9, pdip_tol) -> (c, Jc, status)

`ineq_con_x` / `ineq_con_x_jac` of the 3-D trajectory examples (cone through a wall, quadrotor in a cluttered hallway) for a
batch of states: column `s` of `X` (nx×n; the examples' state `[r; v; p; ω]`) holds the vehicle's position in rows `r_rows`
and its MRP in rows `p_rows`; `c[o, s] = 1 - α` against obstacle `o`, `Jc[1:3, o, s] = -∂α/∂r`, `Jc[4:6, o, s] = -∂α/∂p` (the
non-zero blocks of the examples' rows `[-J[1:3]' 0 -J[4:6]' 0]`).  Shape ids are 0-based.
"""
function se3_constraints(X::Matrix{Float64}, vehicle_shape::Integer, obs_shape::Vector{Int32}, r_obs::Matrix{Float64},
                         p_obs::Matrix{Float64}; r_rows = 1:3, p_rows = 7:9, pdip_tol = 1e-6)
    nx, n = size(X); K = length(obs_shape)
    (length(r_rows) == 3 && length(p_rows) == 3) || error("se3_constraints: r_rows and p_rows must be three consecutive rows each")
    c = Matrix{Float64}(undef, K, n); Jc = Array{Float64}(undef, 6, K, n); status = Matrix{Int32}(undef, K, n)
    _check(ccall((:dcol_se3_constraints, _lib[]), Cint,
                 (Ptr{Cvoid}, Int64, Ptr{Float64}, Cint, Cint, Cint, Int32, Cint, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Float64, Cint,
                  Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                 _ctx[], n, X, nx, first(r_rows) - 1, first(p_rows) - 1, vehicle_shape, K, obs_shape, r_obs, p_obs, Float64(pdip_tol), MEM_HOST,
                 c, Jc, status))
    return c, Jc, status
end

"""
    query_indexed(kind, shape1, shape2, pose1, pose2, r1, att1, r2, att2; pdip_tol) -> BatchResult

Pose-table query: pair k uses column `pose1[k]+1` of `(r1, att1)` and `pose2[k]+1` of `(r2, att2)` (0-based indices, as in C).
"""
function query_indexed(kind::Integer, s1::Vector{Int32}, s2::Vector{Int32}, p1::Vector{Int32}, p2::Vector{Int32},
                       r1::Matrix{Float64}, a1::Matrix{Float64}, r2::Matrix{Float64}, a2::Matrix{Float64}; pdip_tol = 1e-6)
    n = length(s1); na = size(a1, 1); nt = 6 + 2na
    α = Vector{Float64}(undef, n); status = Vector{Int32}(undef, n); iters = Vector{Int32}(undef, n)
    x = Matrix{Float64}(undef, 3, kind == 1 ? 0 : n)
    deriv = kind == 0 ? Float64[] : (kind == 1 ? Matrix{Float64}(undef, nt, n) : Array{Float64}(undef, 4, nt, n))
    _check(ccall((:dcol_query_indexed, _lib[]), Cint,
                 (Ptr{Cvoid}, Cint, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Int64, Ptr{Float64}, Ptr{Float64},
                  Int64, Ptr{Float64}, Ptr{Float64}, Cint, Float64, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                 _ctx[], kind, n, s1, s2, p1, p2, size(r1, 2), r1, a1, size(r2, 2), r2, a2, na == 4 ? ATT_QUAT : ATT_MRP,
                 Float64(pdip_tol), MEM_HOST, α, kind == 1 ? C_NULL : x, kind == 0 ? C_NULL : deriv, status, iters))
    return BatchResult(α, x, deriv, status, iters)
end

"""
    submit!(res::BatchResult, kind, s1, s2, r1, a1, r2, a2; pdip_tol) -> ticket ;  wait!(ticket)

Asynchronous form (`dcol_submit` / `dcol_wait`): up to three submissions in flight, so that the PCIe copies of one
batch overlap the kernels of the previous one.  `res` and the input arrays must stay alive and untouched until
`wait!(ticket)` returns (use pinned memory, `pin!`, for the copies to be truly asynchronous).
"""
function submit!(res::BatchResult, kind::Integer, s1::Vector{Int32}, s2::Vector{Int32}, r1::Matrix{Float64}, a1::Matrix{Float64},
                 r2::Matrix{Float64}, a2::Matrix{Float64}; pdip_tol = 1e-6)
    t = ccall((:dcol_submit, _lib[]), Cint,
              (Ptr{Cvoid}, Cint, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Float64,
               Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
              _ctx[], kind, length(s1), s1, s2, r1, a1, r2, a2, size(a1, 1) == 4 ? ATT_QUAT : ATT_MRP, Float64(pdip_tol),
              res.α, kind == 1 ? C_NULL : res.x, kind == 0 ? C_NULL : res.deriv, res.status, res.iters)
    t >= 0 || _check(t)
    return t
end
wait!(ticket::Integer) = (_check(ccall((:dcol_wait, _lib[]), Cint, (Ptr{Cvoid}, Cint), _ctx[], ticket)); nothing)

"""
    pin!(a::Array) -> a ;  unpin!(a::Array)

Page-lock a Julia array in place (`dcol_host_register`): host-pointer calls then move it by DMA at the full PCIe rate
instead of through the driver's staging copies.  Pin input and output arrays once, reuse them across calls, and
`unpin!` them before they can be garbage-collected.
"""
function pin!(a::Array)
    _check(ccall((:dcol_host_register, _lib[]), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), _ctx[], a, sizeof(a)))
    return a
end
unpin!(a::Array) = (_check(ccall((:dcol_host_unregister, _lib[]), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), _ctx[], a)); nothing)

# ---------------------------------------------------------------------------------------------
# round-2 additions to the C ABI (include/dcol_b200.h): the solver on its own, opt-in cull / warm start, diagnostics
# ---------------------------------------------------------------------------------------------
"""
    solve_socp(c, G, h, idx_ort, idx_soc1, idx_soc2; pdip_tol = 1e-4) -> (x, s, z)

Same call as the reference's `solve_socp` (`src/solvers/coneqp/static_solver2.jl:106-191`); rows of `G`, `h` must be
ordered `[idx_ort; idx_soc1; idx_soc2]` as `combine_problem_matrices` produces them.  Runs `dcol_solve_socp` on one problem;
`verbose = true` runs `dcol_solve_socp_trace` and prints the reference's per-iteration table; the single-pair calls below
do that on the pair's dense problem (`pair_problem`) when called with `verbose = true`.
"""
# "%W.De" without a dependency on Printf: mantissa with D decimals, two-digit exponent, right-aligned in W characters
function _e(v::Float64, w::Int, d::Int)
    (isfinite(v) && v != 0) || return lpad(string(v), w)
    ex = floor(Int, log10(abs(v))); m = round(v / 10.0^ex; digits = d)
    abs(m) >= 10 && (m /= 10; ex += 1)
    ms = string(m); nd = d - (length(ms) - findfirst('.', ms))
    return lpad(ms * "0"^nd * "e" * (ex < 0 ? "-" : "+") * lpad(abs(ex), 2, "0"), w)
end

function solve_socp(c::AbstractVector, G::AbstractMatrix, h::AbstractVector, idx_ort, idx_soc1, idx_soc2; pdip_tol = 1e-4, verbose = false)
    ns, nx = size(G)
    Gr = Matrix{Float64}(transpose(G))                               # row-major ns x nx for the C side
    x = Vector{Float64}(undef, nx); s = Vector{Float64}(undef, ns); z = Vector{Float64}(undef, ns)
    status = Int32[0]; iters = Int32[0]
    if verbose                                                       # static_solver2.jl:117-120, 179-183
        trace = fill(NaN, 5, 50)                                     # DCOL_TRACE_COLS x DCOL_TRACE_ROWS, one column per iteration
        _check(ccall((:dcol_solve_socp_trace, _lib[]), Cint,
                     (Ptr{Cvoid}, Int64, Cint, Cint, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Cint,
                      Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}),
                     _ctx[], 1, nx, length(idx_ort), length(idx_soc1), length(idx_soc2), collect(Float64, c), Gr, collect(Float64, h),
                     Float64(pdip_tol), MEM_HOST, x, s, z, status, iters, trace))
        println("iter     objv        μ       |Gx+s-h|      |s∘z|      α")
        println("---------------------------------------------------------")
        for k in 1:min(iters[1], 50)
            println(lpad(k, 3), "   ", _e(trace[1, k], 10, 3), "  ", _e(trace[2, k], 9, 2), "  ", _e(trace[3, k], 9, 2), "  ",
                    _e(trace[4, k], 9, 2), "  ", trace[5, k] < 0 ? "" : " ", string(round(trace[5, k]; digits = 4)))
        end
    else
    _check(ccall((:dcol_solve_socp, _lib[]), Cint,
                 (Ptr{Cvoid}, Int64, Cint, Cint, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Cint,
                  Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                 _ctx[], 1, nx, length(idx_ort), length(idx_soc1), length(idx_soc2), collect(Float64, c), Gr, collect(Float64, h),
                 Float64(pdip_tol), MEM_HOST, x, s, z, status, iters))
    end
    status[1] == 1 && error("pdip failed")                           # static_solver2.jl:188
    status[1] == 0 || throw(DomainError(status[1], "numerical breakdown in the conic solver"))
    return x, s, z
end

"`scene_set_cull!(α_cut)`: opt-in broad phase of `scene_allpairs` / `scene_wrench` (`dcol_scene_set_cull`); pairs proven to have α > α_cut come back with status 3 and a lower bound of α.  0 switches it off."
function scene_set_cull!(α_cut::Real)
    _check(ccall((:dcol_scene_set_cull, _lib[]), Cint, (Ptr{Cvoid}, Float64), _ctx[], Float64(α_cut)))
    if _grp[] != C_NULL
        for i in 0:(ccall((:dcol_group_size, _lib[]), Cint, (Ptr{Cvoid},), _grp[]) - 1)
            ccall((:dcol_scene_set_cull, _lib[]), Cint, (Ptr{Cvoid}, Float64), ccall((:dcol_group_ctx, _lib[]), Ptr{Cvoid}, (Ptr{Cvoid}, Cint), _grp[], i), Float64(α_cut))
        end
    end
    nothing
end

"`set_mu_output!(μ)`: the following host-pointer queries also return the final duality measure per pair in `μ` (a `Vector{Float64}` the caller keeps alive); `nothing` switches it off."
set_mu_output!(μ::Union{Vector{Float64},Nothing}) =
    (_check(ccall((:dcol_set_mu_output, _lib[]), Cint, (Ptr{Cvoid}, Ptr{Float64}), _ctx[], μ === nothing ? C_NULL : μ)); nothing)

"`warm_stride()`: doubles per pair of a warm-start record for the current shape table (`dcol_warm_stride`).  The records live in DEVICE memory; allocate them with your CUDA binding of choice and pass the device pointer to `set_warm_start!`."
warm_stride() = Int(ccall((:dcol_warm_stride, _lib[]), Int64, (Ptr{Cvoid},), _ctx[]))
set_warm_start!(state::Ptr{Float64}, capacity_pairs::Integer; mu_save = 1e-2) =
    (_check(ccall((:dcol_set_warm_start, _lib[]), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Float64), _ctx[], state, capacity_pairs, Float64(mu_save))); nothing)


# ---------------------------------------------------------------------------------------------
# dense conic form of one pair (host side; only `verbose = true` and users of `solve_socp` need it: the batched
# kernels never form G).  Same blocks as src/problem_matrices.jl:30-191 and src/combine_problem_matrices.jl:15-86,
# with plain `Matrix` / `Vector` instead of static arrays: rows `h - G*y` in the orthant, then in a second-order
# cone; `y = [x; α; extra variables of primitive 1; extra variables of primitive 2]`.
# ---------------------------------------------------------------------------------------------
_dcm(P::AbstractPrimitive) = Matrix(dcm_from_q(P.q))
_dcm(P::AbstractPrimitiveMRP) = Matrix(dcm_from_mrp(P.p))
function _ball_block(R, r, E)                    # SOC rows [R α; x - r - E e]
    G = zeros(4, 4 + size(E, 2)); G[1, 4] = -R
    for i in 1:3; G[1 + i, i] = -1.0; end
    G[2:4, 5:end] .= E
    return G, vcat(0.0, -r)
end
function problem_matrices(P::AnyPrimitive)
    Qn = _dcm(P); r = Vector(P.r) + Qn * Vector(P.r_offset); Q = Qn * Matrix(P.Q_offset)
    t = shape_type(P); bx = Q[:, 1]
    if t == 0                                                        # polytope: A Q'(x - r) <= α b
        AQt = Matrix(P.A) * Q'
        return hcat(AQt, -Vector(P.b)), AQt * r, zeros(0, 4), zeros(0)
    elseif t == 4                                                    # sphere
        G, h = _ball_block(P.R, r, zeros(3, 0)); return zeros(0, 4), zeros(0), G, h
    elseif t == 1 || t == 2                                          # capsule / cylinder: axis point r + γ bx, |γ| <= α L/2
        G, h = _ball_block(P.R, r, reshape(bx, 3, 1))
        Go = [0 0 0 -P.L/2 1.0; 0 0 0 -P.L/2 -1.0]; ho = zeros(2)
        if t == 2                                                    # the two end planes
            Go = vcat(Go, hcat(-bx', -P.L / 2, 0.0), hcat(bx', -P.L / 2, 0.0)); ho = [0.0, 0.0, -dot(bx, r), dot(bx, r)]
        end
        return Go, ho, G, h
    elseif t == 3                                                    # cone
        tb = tan(P.β); EQt = Diagonal([tb, 1.0, 1.0]) * Q'
        return hcat(bx', -P.H / 4), [dot(bx, r)], hcat(-EQt, [-tb * 3 * P.H / 4, 0.0, 0.0]), -EQt * r
    elseif t == 6                                                    # padded polygon: in-plane point r + Q[:,1:2] e, A e <= α b
        nh = size(P.A, 1); G, h = _ball_block(P.R, r, Q[:, 1:2])
        return hcat(zeros(nh, 3), -Vector(P.b), Matrix(P.A)), zeros(nh), G, h
    else                                                             # ellipsoid: |U Q'(x - r)| <= α
        UQt = Matrix(P.U) * Q'; G = zeros(4, 4); G[1, 4] = -1.0; G[2:4, 1:3] .= -UQt
        return zeros(0, 4), zeros(0), G, vcat(0.0, -UQt * r)
    end
end
function combine_problem_matrices(Go1, ho1, Gs1, hs1, Go2, ho2, Gs2, hs2)
    e1 = max(size(Go1, 2), size(Gs1, 2)) - 4; e2 = max(size(Go2, 2), size(Gs2, 2)) - 4; nx = 4 + e1 + e2
    function widen(B, first, ne)
        out = zeros(size(B, 1), nx); out[:, 1:4] .= B[:, 1:4]; out[:, first:first+ne-1] .= B[:, 5:4+ne]; out
    end
    G = vcat(widen(Go1, 5, e1), widen(Go2, 5 + e1, e2), widen(Gs1, 5, e1), widen(Gs2, 5 + e1, e2))
    h = vcat(ho1, ho2, hs1, hs2); c = zeros(nx); c[4] = 1.0
    no = size(Go1, 1) + size(Go2, 1); d1 = size(Gs1, 1); d2 = size(Gs2, 1)
    return c, G, h, collect(1:no), collect(no+1:no+d1), collect(no+d1+1:no+d1+d2)
end
pair_problem(P1::AnyPrimitive, P2::AnyPrimitive) = combine_problem_matrices(problem_matrices(P1)..., problem_matrices(P2)...)

# ---------------------------------------------------------------------------------------------
# single-pair drop-in calls: a batch of one through the same entry points
# ---------------------------------------------------------------------------------------------
function _single(sym, P1, P2, pdip_tol, verbose = false)
    # verbose (static_solver2.jl:117-120, 179-183): the iteration table comes from the pair's dense problem solved by
    # dcol_solve_socp_trace; the values returned below are the fast path's
    verbose && solve_socp(pair_problem(P1, P2)...; pdip_tol = pdip_tol, verbose = true)
    s1 = _registered!(P1); s2 = _registered!(P2)
    _reg_on_device[] || set_shapes!(_reg)
    col(v) = reshape(collect(Float64, v), :, 1)
    res = _query(sym, Int32[s1], Int32[s2], col(P1.r), col(att(P1)), col(P2.r), col(att(P2)), att_kind(P1), Float64(pdip_tol))
    res.status[1] == 1 && error("pdip failed")                        # static_solver2.jl:188
    res.status[1] == 0 || throw(DomainError(res.status[1], "numerical breakdown in the conic solver"))
    return res
end

for (PT, nt) in ((:AbstractPrimitive, 14), (:AbstractPrimitiveMRP, 12))
    @eval begin
        function proximity(prim1::$PT, prim2::$PT; pdip_tol::Float64 = 1e-6, verbose::Bool = false)
            res = _single(:dcol_proximity, prim1, prim2, pdip_tol, verbose)
            return res.α[1], SVector{3}(res.x[:, 1])
        end
        function proximity_gradient(prim1::$PT, prim2::$PT; pdip_tol::Float64 = 1e-6, verbose::Bool = false)
            res = _single(:dcol_proximity_gradient, prim1, prim2, pdip_tol, verbose)
            return res.α[1], SVector{$nt}(res.deriv[:, 1])
        end
        function proximity_jacobian(prim1::$PT, prim2::$PT; pdip_tol::Float64 = 1e-6, verbose::Bool = false)
            res = _single(:dcol_proximity_jacobian, prim1, prim2, pdip_tol, verbose)
            return res.α[1], SVector{3}(res.x[:, 1]), SMatrix{4,$nt}(res.deriv[:, :, 1])
        end
    end
end

end # module
