"""Host-side mirror of the reference's primitive structs.

Mirrors `src/primitives.jl:4-96` (quaternion structs, `<: AbstractPrimitive`) and
`src/primitives_mrp.jl:2-96` (MRP twins, `<: AbstractPrimitiveMRP`): same names,
same mutable fields (`r`, `q`|`p`, geometry, `r_offset`, `Q_offset`), same
defaults (origin / identity attitude, `src/primitives.jl:12`).  The structs hold
no device state; `ShapeTable` (shapes.py) packs their geometry for the library.

Deviation from the reference, on purpose: `EllipsoidMRP(P)` is constructible
under its own name (the reference declares its inner constructor as
`Ellipsoid`, `src/primitives_mrp.jl:85`, which shadows the quaternion one).
"""
from __future__ import annotations

import math
import numpy as np

POLYTOPE, CAPSULE, CYLINDER, CONE, SPHERE, ELLIPSOID, POLYGON = range(7)
TYPE_NAMES = ["polytope", "capsule", "cylinder", "cone", "sphere", "ellipsoid", "polygon"]
ATT_QUAT, ATT_MRP = 0, 1


def _bytes(x):
    return x.tobytes() if type(x) is np.ndarray and x.dtype == np.float64 else np.asarray(x, dtype=np.float64).tobytes()


def _vec(x, n):
    a = np.asarray(x, dtype=np.float64).reshape(-1)
    if a.shape[0] != n:
        raise ValueError(f"expected length-{n} vector, got shape {np.shape(x)}")
    return a.copy()


class AbstractPrimitive:
    """Quaternion-attitude primitive (`src/primitives.jl:2`)."""
    att_kind = ATT_QUAT
    shape_type = -1

    def _init_pose(self):
        self.r = np.zeros(3)
        if self.att_kind == ATT_QUAT:
            self.q = np.array([1.0, 0.0, 0.0, 0.0])
        else:
            self.p = np.zeros(3)
        self.r_offset = np.zeros(3)
        self.Q_offset = np.eye(3)

    @property
    def att(self):
        return self.q if self.att_kind == ATT_QUAT else self.p

    def shape_params(self) -> np.ndarray:
        """Geometry packed as include/dcol_b200.h documents (without the offsets tail)."""
        raise NotImplementedError

    def packed(self) -> np.ndarray:
        return np.concatenate([self.shape_params(), _vec(self.r_offset, 3),
                               np.asarray(self.Q_offset, dtype=np.float64).reshape(9)])

    def shape_key(self):
        """Identity of the geometry (shape-table de-duplication); cheaper than packing: this sits in every single-pair call."""
        return (self.shape_type,) + self._geom_key() + (_bytes(self.r_offset), _bytes(self.Q_offset))

    def _geom_key(self):
        return (self.shape_params().tobytes(),)


class AbstractPrimitiveMRP(AbstractPrimitive):
    """MRP-attitude primitive (`src/primitives_mrp.jl:2`)."""
    att_kind = ATT_MRP


class _PolytopeBase:
    shape_type = POLYTOPE

    def __init__(self, A, b):  # src/primitives.jl:11-13
        A = np.asarray(A, dtype=np.float64)
        b = np.asarray(b, dtype=np.float64).reshape(-1)
        if A.ndim != 2 or A.shape[1] != 3 or A.shape[0] != b.shape[0]:
            raise ValueError("Polytope needs A (n x 3) and b (n)")
        self.A, self.b = A.copy(), b.copy()
        self._init_pose()

    def shape_params(self):
        return np.concatenate([[float(self.A.shape[0])], self.A.reshape(-1), self.b])

    def _geom_key(self):
        return (_bytes(self.A), _bytes(self.b))


class _CapsuleBase:
    shape_type = CAPSULE

    def __init__(self, R, L):  # src/primitives.jl:23-25
        self.R, self.L = float(R), float(L)
        self._init_pose()

    def shape_params(self):
        return np.array([self.R, self.L])

    def _geom_key(self):
        return (float(self.R), float(self.L))


class _CylinderBase(_CapsuleBase):
    shape_type = CYLINDER  # src/primitives.jl:28-38


class _ConeBase:
    shape_type = CONE

    def __init__(self, H, beta):  # src/primitives.jl:47-49 (field is `β`)
        self.H, self.beta = float(H), float(beta)
        self._init_pose()

    @property
    def β(self):  # noqa: PLC2401 - keep the reference's field name available
        return self.beta

    def shape_params(self):
        return np.array([self.H, self.beta])

    def _geom_key(self):
        return (float(self.H), float(self.beta))


class _SphereBase:
    shape_type = SPHERE

    def __init__(self, R):  # src/primitives.jl:58-60
        self.R = float(R)
        self._init_pose()

    def shape_params(self):
        return np.array([self.R])

    def _geom_key(self):
        return (float(self.R),)


class _PolygonBase:
    shape_type = POLYGON

    def __init__(self, A, b, R):  # src/primitives.jl:71-73
        A = np.asarray(A, dtype=np.float64)
        b = np.asarray(b, dtype=np.float64).reshape(-1)
        if A.ndim != 2 or A.shape[1] != 2 or A.shape[0] != b.shape[0]:
            raise ValueError("Polygon needs A (nh x 2) and b (nh)")
        self.A, self.b, self.R = A.copy(), b.copy(), float(R)
        self._init_pose()

    def shape_params(self):
        return np.concatenate([[float(self.A.shape[0])], self.A.reshape(-1), self.b, [self.R]])

    def _geom_key(self):
        return (_bytes(self.A), _bytes(self.b), float(self.R))


class _EllipsoidBase:
    shape_type = ELLIPSOID

    def __init__(self, P):  # src/primitives.jl:85-95: U = cholesky(P).U, F = eigen(P)
        P = np.asarray(P, dtype=np.float64)
        if P.shape != (3, 3):
            raise ValueError("Ellipsoid needs a 3x3 matrix P")
        self.P = P.copy()
        self.U = np.linalg.cholesky(P).T.copy()
        self.F = np.linalg.eigh(P)
        self._init_pose()

    def shape_params(self):
        return self.U.reshape(-1)


class Polytope(_PolytopeBase, AbstractPrimitive): pass
class Capsule(_CapsuleBase, AbstractPrimitive): pass
class Cylinder(_CylinderBase, AbstractPrimitive): pass
class Cone(_ConeBase, AbstractPrimitive): pass
class Sphere(_SphereBase, AbstractPrimitive): pass
class Polygon(_PolygonBase, AbstractPrimitive): pass
class Ellipsoid(_EllipsoidBase, AbstractPrimitive): pass


class PolytopeMRP(_PolytopeBase, AbstractPrimitiveMRP): pass
class CapsuleMRP(_CapsuleBase, AbstractPrimitiveMRP): pass
class CylinderMRP(_CylinderBase, AbstractPrimitiveMRP): pass
class ConeMRP(_ConeBase, AbstractPrimitiveMRP): pass
class SphereMRP(_SphereBase, AbstractPrimitiveMRP): pass
class PolygonMRP(_PolygonBase, AbstractPrimitiveMRP): pass
class EllipsoidMRP(_EllipsoidBase, AbstractPrimitiveMRP): pass


# --- misc constructors: src/misc_primitive_constructors.jl ---------------------------------
def create_rect_prism(len=20.0, wid=20.0, hei=2.0, attitude="MRP"):  # :1-27 (default :MRP, returns 3-tuple)
    ns = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [0, 0, -1]], dtype=np.float64)
    cs = np.array([[len / 2, 0, 0], [0, wid / 2, 0], [0, 0, hei / 2],
                   [-len / 2, 0, 0], [0, -wid / 2, 0], [0, 0, -hei / 2]], dtype=np.float64)
    A = ns.copy()
    b = np.einsum("ij,ij->i", ns, cs)
    mass = len * wid * hei
    inertia = (mass / 12) * np.diag([wid**2 + hei**2, len**2 + hei**2, len**2 + wid**2])
    if attitude in ("MRP", ":MRP"):
        return PolytopeMRP(A, b), mass, inertia
    if attitude in ("quat", ":quat"):
        return Polytope(A, b), mass, inertia
    raise ValueError("attitude must be :MRP or :quat")


def create_n_sided(N, d):  # :28-33
    th = np.array([2 * math.pi * k / N for k in range(N)])
    A = np.stack([np.cos(th), np.sin(th)], axis=1)
    b = d * np.ones(N)
    return A, b


# --- mass properties: src/mass_properties.jl:2-48 (cone, capsule, cylinder only; host scalars, not on the query path) ---
def mass_properties(prim, rho: float = 1.0):
    """(mass, 3x3 diagonal inertia about the body axes) for a uniform-density Cone, Capsule or Cylinder, with the
    reference's conventions: x is the symmetry axis; the cone's base radius is H*tan(beta); the capsule is a cylinder
    of length L plus two hemispherical caps.  The cylinder's transverse inertia keeps the reference's coefficient
    m/12*(2R^2 + L^2) (src/mass_properties.jl:44-45; the textbook value has 3R^2) so that simulations built on it
    reproduce the reference's trajectories."""
    t = prim.shape_type
    if t == CONE:
        rad = math.tan(prim.beta) * prim.H
        m = rho * math.pi * rad**2 * prim.H / 3.0
        ixx = 0.3 * m * rad**2
        iyy = m * (3.0 / 20.0 * rad**2 + 3.0 / 80.0 * prim.H**2)
        return m, np.diag([ixx, iyy, iyy])
    if t == CAPSULE:
        R, L = prim.R, prim.L
        m_cyl = rho * math.pi * R**2 * L
        m_cap = rho * 2.0 * math.pi * R**3 / 3.0          # one hemisphere
        ixx = 0.5 * m_cyl * R**2 + 2.0 * m_cap * (0.4 * R**2)
        iyy = m_cyl * (L**2 / 12.0 + R**2 / 4.0) + 2.0 * m_cap * (0.4 * R**2 + L**2 / 2.0 + 3.0 * L * R / 8.0)
        return m_cyl + 2.0 * m_cap, np.diag([ixx, iyy, iyy])
    if t == CYLINDER:
        R, L = prim.R, prim.L
        m = rho * math.pi * R**2 * L
        iyy = m / 12.0 * (2.0 * R**2 + L**2)
        return m, np.diag([0.5 * m * R**2, iyy, iyy])
    raise TypeError("mass_properties is defined for Cone, Capsule and Cylinder only (as in the reference)")


# --- attitude helpers: src/problem_matrices.jl:1-27 (host convenience; not on the device path) ----
def mrp_from_q(q):
    q = _vec(q, 4)
    return q[1:4] / (1 + q[0])


def q_from_mrp(p):
    p = _vec(p, 3)
    pp = float(p @ p)
    return (1 / (1 + pp)) * np.concatenate([[1 - pp], 2 * p])


def dcm_from_q(q):
    q4, q1, q2, q3 = _vec(q, 4) / np.linalg.norm(q)
    return np.array([
        [2 * q1**2 + 2 * q4**2 - 1, 2 * (q1 * q2 - q3 * q4), 2 * (q1 * q3 + q2 * q4)],
        [2 * (q1 * q2 + q3 * q4), 2 * q2**2 + 2 * q4**2 - 1, 2 * (q2 * q3 - q1 * q4)],
        [2 * (q1 * q3 - q2 * q4), 2 * (q2 * q3 + q1 * q4), 2 * q3**2 + 2 * q4**2 - 1]])


def dcm_from_mrp(p):
    return dcm_from_q(q_from_mrp(p))
