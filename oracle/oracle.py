"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes driver for oracle/libdcol_oracle.so, the C++ CPU restatement of the
reference's proximity hot path (PARITY UNPINNED: no Julia in this image, see
oracle/dcol_oracle.hpp).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product
package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


def build(force=False):
    so = os.path.join(_HERE, "libdcol_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("dcol_oracle_capi.cpp", "dcol_oracle.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.dcol_oracle_query.restype = C.c_int
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def max_threads():
    return int(lib().dcol_oracle_max_threads())


def set_max_iters(k=50):
    """Iteration cap of solve_socp (50 in the reference); see dcol_oracle.hpp max_iters_setting."""
    lib().dcol_oracle_set_max_iters(C.c_int(int(k)))


def query(kind, table_arrays, shape1, shape2, r1, att1, r2, att2, att_kind, pdip_tol=1e-6, nthreads=0):
    """kind: 0 proximity, 1 gradient, 2 jacobian, 3 jacobian_slow.  Returns dict of arrays.
    J is returned as (n, 4, ntheta) (converted from the column-major ABI layout)."""
    types, offs, params = table_arrays
    shape1 = np.ascontiguousarray(shape1, dtype=np.int32)
    shape2 = np.ascontiguousarray(shape2, dtype=np.int32)
    n = shape1.shape[0]
    na = 4 if att_kind == 0 else 3
    nt = 6 + 2 * na
    r1, att1, r2, att2 = (_f64(a) for a in (r1, att1, r2, att2))
    assert r1.shape == (n, 3) and r2.shape == (n, 3) and att1.shape == (n, na) and att2.shape == (n, na)
    alpha = np.empty(n)
    x = np.empty((n, 3))
    status = np.empty(n, dtype=np.int32)
    iters = np.empty(n, dtype=np.int32)
    deriv = None
    if kind == 1:
        deriv = np.empty((n, nt))
    elif kind in (2, 3):
        deriv = np.empty((n, nt, 4))
    rc = lib().dcol_oracle_query(
        C.c_int(kind), C.c_int(len(types)), _p(types, _ip), _p(offs, _lp), _p(params, _dp),
        C.c_int64(n), _p(shape1, _ip), _p(shape2, _ip), _p(r1, _dp), _p(att1, _dp), _p(r2, _dp), _p(att2, _dp),
        C.c_int(att_kind), C.c_double(pdip_tol), C.c_int(nthreads),
        _p(alpha, _dp), _p(x, _dp), _p(deriv, _dp), _p(status, _ip), _p(iters, _ip))
    if rc != 0:
        raise RuntimeError(f"dcol_oracle_query failed rc={rc}")
    out = {"alpha": alpha, "status": status, "iters": iters}
    if kind != 1:
        out["x"] = x
    if kind == 1:
        out["grad"] = deriv
    elif kind in (2, 3):
        out["J"] = np.ascontiguousarray(deriv.transpose(0, 2, 1))
    return out


def solve_socp(c, G, h, n_ort, d1, d2, pdip_tol=1e-4):
    G = _f64(G); h = _f64(h); c = _f64(c)
    ns, nx = G.shape
    x = np.empty(nx); s = np.empty(ns); z = np.empty(ns)
    it = C.c_int32(0)
    st = lib().dcol_oracle_solve_socp(C.c_int(nx), C.c_int(n_ort), C.c_int(d1), C.c_int(d2), _p(c, _dp), _p(G, _dp), _p(h, _dp),
                                      C.c_double(pdip_tol), _p(x, _dp), _p(s, _dp), _p(z, _dp), C.byref(it))
    return x, s, z, int(it.value), int(st)


def pair_problem(prim1, prim2):
    """(c, G, h, n_ort, d1, d2) for two host primitives at their current poses."""
    p1 = _f64(prim1.packed()); p2 = _f64(prim2.packed())
    dims = np.zeros(5, dtype=np.int32)
    G = np.zeros(264 * 8); h = np.zeros(264); c = np.zeros(8)
    rc = lib().dcol_oracle_pair_problem(C.c_int(prim1.shape_type), _p(p1, _dp), C.c_int(prim2.shape_type), _p(p2, _dp),
                                        _p(_f64(prim1.r), _dp), _p(_f64(prim1.att), _dp), _p(_f64(prim2.r), _dp), _p(_f64(prim2.att), _dp),
                                        C.c_int(prim1.att_kind), _p(dims, _ip), _p(G, _dp), _p(h, _dp), _p(c, _dp))
    if rc != 0:
        raise RuntimeError("pair_problem failed")
    nx, ns, n_ort, d1, d2 = (int(v) for v in dims)
    return c[:nx].copy(), G[: ns * nx].reshape(ns, nx).copy(), h[:ns].copy(), n_ort, d1, d2


def combine(Go1, ho1, Gs1, hs1, Go2, ho2, Gs2, hs2):
    Go1, Gs1, Go2, Gs2 = (_f64(a) for a in (Go1, Gs1, Go2, Gs2))
    ho1, hs1, ho2, hs2 = (_f64(a) for a in (ho1, hs1, ho2, hs2))
    L1, v1 = Go1.shape; d1 = Gs1.shape[0]
    L2, v2 = Go2.shape; d2 = Gs2.shape[0]
    dims = np.zeros(5, dtype=np.int32)
    G = np.zeros(264 * 8); h = np.zeros(264); c = np.zeros(8)
    rc = lib().dcol_oracle_combine(C.c_int(L1), C.c_int(d1), C.c_int(v1), _p(Go1, _dp), _p(ho1, _dp), _p(Gs1, _dp), _p(hs1, _dp),
                                   C.c_int(L2), C.c_int(d2), C.c_int(v2), _p(Go2, _dp), _p(ho2, _dp), _p(Gs2, _dp), _p(hs2, _dp),
                                   _p(dims, _ip), _p(G, _dp), _p(h, _dp), _p(c, _dp))
    if rc != 0:
        raise RuntimeError(f"combine failed rc={rc}")
    nx, ns, n_ort, dd1, dd2 = (int(v) for v in dims)
    return c[:nx].copy(), G[: ns * nx].reshape(ns, nx).copy(), h[:ns].copy(), n_ort, dd1, dd2


def nt(n_ort, d1, d2, s, z, op, g):
    s = _f64(s); z = _f64(z); g = _f64(g)
    ns = n_ort + d1 + d2
    out = np.empty(ns); Wd = np.empty((ns, ns))
    rc = lib().dcol_oracle_nt(C.c_int(n_ort), C.c_int(d1), C.c_int(d2), _p(s, _dp), _p(z, _dp), C.c_int(op), _p(g, _dp), _p(out, _dp), _p(Wd, _dp))
    if rc != 0:
        raise RuntimeError("nt failed")
    return out, Wd


def cone_op(n_ort, d1, d2, op, u, v=None):
    u = _f64(u); v = _f64(v) if v is not None else u
    out = np.empty(n_ort + d1 + d2)
    rc = lib().dcol_oracle_cone_op(C.c_int(n_ort), C.c_int(d1), C.c_int(d2), C.c_int(op), _p(u, _dp), _p(v, _dp), _p(out, _dp))
    if rc != 0:
        raise RuntimeError("cone_op failed")
    return out[0] if op == 3 else out


def att(op, a):
    a = _f64(a)
    out = np.empty(9 if op < 2 else (3 if op == 2 else 4))
    lib().dcol_oracle_att(C.c_int(op), _p(a, _dp), _p(out, _dp))
    return out.reshape(3, 3) if op < 2 else out
