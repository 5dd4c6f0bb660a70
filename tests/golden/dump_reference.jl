# Dump (alpha, x, grad, J) of the REAL DifferentiableCollisions.jl package for seeded poses, so that the C++ oracle and
# the CUDA path can be pinned to Julia output (SURVEY.md section 8c, VERDICT round 1 item 2e).  Julia is not installed in
# the build image, so this script has never been run there; anyone with Julia 1.8+ can close the gap:
#
#   julia --project=/path/to/DifferentiableCollisions.jl tests/golden/dump_reference.jl [N_POSES=10] > tests/golden/julia_reference.json
#
# tests/test_julia_reference.py consumes the file when it is present (and skips otherwise).  The output carries the inputs
# it used, so parity does not depend on the two random generators agreeing; the poses follow the pose law of
# test/polytope_derivs_test.jl:27-30 (r1 = 2 N(0,I), r2 = N(0,I), q = normalize(N(0,I4))) driven by SplitMix64 + Box-Muller
# so that the Python side (tests/golden/reference_inputs.py) can regenerate exactly the same numbers as a cross-check.
import DifferentiableCollisions as dc
using StaticArrays, LinearAlgebra, JLD2, Printf

const N_POSES = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 10
const PDIP_TOL = 1e-6

mutable struct SplitMix64; s::UInt64; end
function next_u64!(g::SplitMix64)
    g.s += 0x9e3779b97f4a7c15
    z = g.s
    z = (z ⊻ (z >> 30)) * 0xbf58476d1ce4e5b9
    z = (z ⊻ (z >> 27)) * 0x94d049bb133111eb
    return z ⊻ (z >> 31)
end
uniform!(g) = (Float64(next_u64!(g) >> 11) + 0.5) / 9007199254740992.0          # (0,1)
function normal!(g)                                                             # Box-Muller, cosine branch only
    u1 = uniform!(g); u2 = uniform!(g)
    return sqrt(-2.0 * log(u1)) * cos(2.0 * pi * u2)
end

mrp_from_q(q) = q[1] < 0 ? -q[SA[2, 3, 4]] / (1 - q[1]) : q[SA[2, 3, 4]] / (1 + q[1])   # flip q -> -q first: |p| <= 1

function shapes(variant::Symbol, att::Symbol)
    f = jldopen(joinpath(dirname(pathof(dc)), "..", "test", "example_socps", "polytopes.jld2"))
    A2 = SMatrix{8,3}(f["A2"]); b2 = SVector{8}(f["b2"])
    A5, b5 = dc.create_n_sided(5, 0.6)
    P = SA[1.1 0.1 0.1; 0.1 2.1 0.1; 0.1 0.1 0.6]
    q = att == :quat
    poly = variant == :cube ? dc.create_rect_prism(1.0, 1.0, 1.0; attitude = q ? :quat : :MRP)[1] :
           (q ? dc.Polytope(A2, b2) : dc.PolytopeMRP(A2, b2))
    caps = variant == :cube ? (q ? dc.Capsule(0.3, 1.2) : dc.CapsuleMRP(0.3, 1.2)) : (q ? dc.Capsule(0.2, 5.0) : dc.CapsuleMRP(0.2, 5.0))
    ell = q ? dc.Ellipsoid(P) : begin   # EllipsoidMRP's inner constructor is misnamed in v0.1.5 (src/primitives_mrp.jl:85); build it if that is fixed
        try dc.EllipsoidMRP(P) catch; nothing end
    end
    return ["polytope" => poly, "capsule" => caps,
            "cylinder" => (q ? dc.Cylinder(0.6, 3.0) : dc.CylinderMRP(0.6, 3.0)),
            "cone" => (q ? dc.Cone(2.0, deg2rad(22)) : dc.ConeMRP(2.0, deg2rad(22))),
            "sphere" => (q ? dc.Sphere(0.8) : dc.SphereMRP(0.8)),
            "ellipsoid" => ell,
            "polygon" => (q ? dc.Polygon(A5, b5, 0.2) : dc.PolygonMRP(A5, b5, 0.2))]
end

fmt(v) = "[" * join((@sprintf("%.17g", x) for x in v), ",") * "]"

function main()
    g = SplitMix64(UInt64(1234))
    first = true
    println("{\"package\":\"DifferentiableCollisions.jl\",\"pdip_tol\":$(PDIP_TOL),\"n_poses\":$(N_POSES),\"records\":[")
    for variant in (:cube, :scene), att in (:quat, :mrp)
        S = shapes(variant, att)
        for (n1, P1c) in S, (n2, P2c) in S
            (P1c === nothing || P2c === nothing) && continue
            P1 = deepcopy(P1c); P2 = deepcopy(P2c)
            for k in 1:N_POSES
                r1 = 2 * SA[normal!(g), normal!(g), normal!(g)]
                q1 = normalize(SA[normal!(g), normal!(g), normal!(g), normal!(g)])
                r2 = SA[normal!(g), normal!(g), normal!(g)]
                q2 = normalize(SA[normal!(g), normal!(g), normal!(g), normal!(g)])
                P1.r = r1; P2.r = r2
                if att == :quat
                    P1.q = q1; P2.q = q2; a1 = q1; a2 = q2
                else
                    a1 = mrp_from_q(q1); a2 = mrp_from_q(q2); P1.p = a1; P2.p = a2
                end
                rec = try
                    α, x = dc.proximity(P1, P2; pdip_tol = PDIP_TOL)
                    αg, grad = dc.proximity_gradient(P1, P2; pdip_tol = PDIP_TOL)
                    αj, xj, J = dc.proximity_jacobian(P1, P2; pdip_tol = PDIP_TOL)
                    "\"status\":0,\"alpha\":$(@sprintf("%.17g", α)),\"x\":$(fmt(x)),\"grad\":$(fmt(grad)),\"J_colmajor\":$(fmt(vec(J)))"
                catch err
                    "\"status\":1,\"error\":\"$(replace(sprint(showerror, err), '"' => '\''))\""
                end
                print(first ? "" : ",\n"); first = false
                print("{\"shapes\":\"$(variant)\",\"att\":\"$(att)\",\"p1\":\"$(n1)\",\"p2\":\"$(n2)\",\"r1\":$(fmt(r1)),\"a1\":$(fmt(a1)),\"r2\":$(fmt(r2)),\"a2\":$(fmt(a2)),$(rec)}")
            end
        end
    end
    println("\n]}")
end
main()
