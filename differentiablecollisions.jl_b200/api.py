"""Host-side mirror of the reference's query API over the C ABI.

Single-pair calls keep the reference's names, argument meaning and error behaviour:
  proximity(P1, P2; pdip_tol=1e-6)            -> (alpha, x)        src/proximity.jl:3-26
  proximity_gradient(P1, P2; pdip_tol=1e-6)   -> (alpha, grad)     src/proximity_gradient.jl:80-111
  proximity_jacobian(P1, P2; pdip_tol=1e-6)   -> (alpha, x, J)     src/proximity.jl:168-195
`RuntimeError("pdip failed")` mirrors `error("pdip failed")` (static_solver2.jl:188); mixed
quaternion/MRP pairs raise TypeError (the reference has no such method, src/proximity.jl:3,15).

Batched variants (`Context.proximity_batch` ...) take arrays of poses and return arrays plus the
per-pair status, never raising for solver failures.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .primitives import ATT_MRP, ATT_QUAT, AbstractPrimitive
from .problem_matrices import format_trace, pair_problem
from .shapes import ShapeTable

MEM_HOST, MEM_DEVICE = 0, 1


def _scene_arrays(shape, r, att, att_kind):
    shape = np.ascontiguousarray(shape, dtype=np.int32).reshape(-1)
    nb = shape.shape[0]
    na = 4 if att_kind == ATT_QUAT else 3
    return (shape, np.ascontiguousarray(r, dtype=np.float64).reshape(nb, 3), np.ascontiguousarray(att, dtype=np.float64).reshape(nb, na),
            nb, 6 + 2 * na)


def _pair_range(lib, nb, pair_range):
    n = int(lib.dcol_scene_pair_count(nb))
    k0, k1 = (0, n) if pair_range is None else (int(pair_range[0]), int(pair_range[1]))
    if not 0 <= k0 <= k1 <= n:
        raise ValueError(f"pair_range must satisfy 0 <= k0 <= k1 <= {n}")
    return k0, k1


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    # torch tensor (device or pinned host memory)
    return C.c_void_p(a.data_ptr())


class Context:
    """One CUDA device + one shape table (`dcol_ctx`).  Not thread-safe."""

    def __init__(self, device: int = 0):
        self._lib = _capi.load()
        h = C.c_void_p()
        rc = self._lib.dcol_create(C.byref(h), int(device))
        if rc != 0:
            raise RuntimeError(self._lib.dcol_last_error(None).decode())
        self._h = h
        self.device = int(device)
        self.table = ShapeTable()
        self._table_version = -1
        self._one_bufs = {}          # preallocated arrays and pointers of the single-pair calls, per attitude kind

    def close(self):
        if getattr(self, "_h", None):
            self._lib.dcol_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"dcol error {rc}: {self._lib.dcol_last_error(self._h).decode()}")

    # ---- shape table ---------------------------------------------------------------------
    def add_shape(self, prim) -> int:
        # ids returned here index self.table; after set_shape_arrays (raw mode) the device holds a different
        # table, so mark it stale: the next query re-uploads self.table instead of silently using the raw one
        if self._table_version is None:
            self._table_version = -1
        return self.table.add(prim)

    def set_shape_arrays(self, types, offs, params):
        types = np.ascontiguousarray(types, dtype=np.int32)
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        params = np.ascontiguousarray(params, dtype=np.float64)
        self._check(self._lib.dcol_set_shapes(self._h, len(types), _ptr(types), _ptr(offs), _ptr(params), params.shape[0]))
        self._table_version = None

    def sync_shapes(self):
        if self._table_version != self.table.version:
            t, o, p = self.table.arrays()
            self._check(self._lib.dcol_set_shapes(self._h, len(t), _ptr(t), _ptr(o), _ptr(p), p.shape[0]))
            self._table_version = self.table.version

    def set_stream(self, cuda_stream_ptr):
        self._check(self._lib.dcol_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    # ---- instrumentation -----------------------------------------------------------------
    def launch_count(self) -> int:
        return int(self._lib.dcol_launch_count(self._h))

    def last_kernel_ms(self) -> float:
        return float(self._lib.dcol_last_kernel_ms(self._h))

    def measure_fp64_tflops(self, seconds: float = 0.6, three_source: bool = False) -> float:
        """DFMA throughput of the device (`dcol_measure_fp64_tflops_mode`): the FP64 pipe's peak (x = fma(x, a, x)), or with
        `three_source` what a stream of DFMAs with three distinct register operands reaches (92 % of it on B200)."""
        return float(self._lib.dcol_measure_fp64_tflops_mode(self._h, float(seconds), 1 if three_source else 0))

    def set_copy_only(self, on: bool):
        """Measurement knob (dcol_set_copy_only): host-pointer calls do their copies but launch no kernels."""
        self._check(self._lib.dcol_set_copy_only(self._h, int(bool(on))))

    # ---- raw call (host numpy arrays or device/pinned torch tensors) --------------------------
    def query_raw(self, kind, n, shape1, shape2, r1, att1, r2, att2, att_kind, pdip_tol, mem,
                  alpha, x, deriv, status, iters):
        if self._table_version is not None:
            self.sync_shapes()
        args = [self._h, int(n), _ptr(shape1), _ptr(shape2), _ptr(r1), _ptr(att1), _ptr(r2), _ptr(att2),
                int(att_kind), float(pdip_tol), int(mem)]
        if kind == 0:
            rc = self._lib.dcol_proximity(*args, _ptr(alpha), _ptr(x), _ptr(status), _ptr(iters))
        elif kind == 1:
            rc = self._lib.dcol_proximity_gradient(*args, _ptr(alpha), _ptr(deriv), _ptr(status), _ptr(iters))
        else:
            rc = self._lib.dcol_proximity_jacobian(*args, _ptr(alpha), _ptr(x), _ptr(deriv), _ptr(status), _ptr(iters))
        self._check(rc)

    # ---- asynchronous submissions (host buffers; keep them alive until wait) ---------------------------
    def submit_raw(self, kind, n, shape1, shape2, r1, att1, r2, att2, att_kind, pdip_tol, alpha, x, deriv, status, iters) -> int:
        """`dcol_submit`: enqueue one host-buffer query and return its ticket; up to two may be in flight."""
        if self._table_version is not None:
            self.sync_shapes()
        t = self._lib.dcol_submit(self._h, int(kind), int(n), _ptr(shape1), _ptr(shape2), _ptr(r1), _ptr(att1), _ptr(r2), _ptr(att2),
                                  int(att_kind), float(pdip_tol), _ptr(alpha), _ptr(x), _ptr(deriv), _ptr(status), _ptr(iters))
        if t < 0:
            self._check(t)
        return t

    def wait(self, ticket: int):
        self._check(self._lib.dcol_wait(self._h, int(ticket)))

    # ---- pinned host memory (`dcol_host_register`) ---------------------------------------------
    def pin(self, arr: np.ndarray) -> np.ndarray:
        """Page-lock a C-contiguous numpy array in place: host-pointer calls then move it by DMA (55 GB/s measured on
        B200) instead of through the driver's staging copies (4 GB/s device-to-host).  Pin once, reuse across calls,
        `unpin` before the array is freed."""
        if not arr.flags["C_CONTIGUOUS"]:
            raise ValueError("pin: array must be C-contiguous")
        self._check(self._lib.dcol_host_register(self._h, arr.ctypes.data, arr.nbytes))
        return arr

    def unpin(self, arr: np.ndarray):
        self._check(self._lib.dcol_host_unregister(self._h, arr.ctypes.data))

    def alloc_outputs(self, kind, n, att_kind=ATT_QUAT, pinned=True):
        """Output buffers of one batched query (`out=` of the *_batch methods), page-locked by default."""
        nt = 6 + 2 * (4 if att_kind == ATT_QUAT else 3)
        out = {"alpha": np.empty(n), "status": np.empty(n, dtype=np.int32), "iters": np.empty(n, dtype=np.int32)}
        if kind != 1:
            out["x"] = np.empty((n, 3))
        if kind == 1:
            out["deriv"] = np.empty((n, nt))
        elif kind == 2:
            out["deriv"] = np.empty((n, nt, 4))
        if pinned:
            for a in out.values():
                if a.nbytes:
                    self.pin(a)
        return out

    # ---- batched host API ----------------------------------------------------------------------
    def solve_socp(self, c, G, h, n_ort, d1, d2, pdip_tol=1e-4, verbose=False, trace=False):
        """`dcol_solve_socp`: the reference's solve_socp (static_solver2.jl:106-191; default pdip_tol 1e-4 as there, :112) on
        one problem (G: ns x nx) or a batch (G: n x ns x nx).  Returns (x, s, z, status, iters); with `trace=True` also the
        per-iteration record of `dcol_solve_socp_trace` (n x 50 x 5, NaN past a problem's last iteration).  `verbose=True`
        prints the reference's iteration table (:117-120, :179-183) for every problem of the call."""
        G = np.ascontiguousarray(G, dtype=np.float64)
        single = G.ndim == 2
        if single:
            G = G[None]
        n, ns, nx = G.shape
        c = np.ascontiguousarray(np.broadcast_to(np.asarray(c, dtype=np.float64).reshape(-1, nx), (n, nx)))
        h = np.ascontiguousarray(np.broadcast_to(np.asarray(h, dtype=np.float64).reshape(-1, ns), (n, ns)))
        x = np.empty((n, nx)); s = np.empty((n, ns)); z = np.empty((n, ns))
        status = np.empty(n, dtype=np.int32); iters = np.empty(n, dtype=np.int32)
        if verbose or trace:
            tr = np.full((n, _capi.TRACE_ROWS, _capi.TRACE_COLS), np.nan)
            self._check(self._lib.dcol_solve_socp_trace(self._h, n, nx, int(n_ort), int(d1), int(d2), _ptr(c), _ptr(G), _ptr(h),
                                                        float(pdip_tol), MEM_HOST, _ptr(x), _ptr(s), _ptr(z), _ptr(status), _ptr(iters), _ptr(tr)))
            if verbose:
                for k in range(n):
                    print(format_trace(tr[k], iters[k]))
        else:
            self._check(self._lib.dcol_solve_socp(self._h, n, nx, int(n_ort), int(d1), int(d2), _ptr(c), _ptr(G), _ptr(h), float(pdip_tol),
                                                  MEM_HOST, _ptr(x), _ptr(s), _ptr(z), _ptr(status), _ptr(iters)))
        if single:
            return (x[0], s[0], z[0], int(status[0]), int(iters[0])) + ((tr[0],) if trace else ())
        return (x, s, z, status, iters) + ((tr,) if trace else ())

    def warm_stride(self) -> int:
        if self._table_version is not None:
            self.sync_shapes()
        return int(self._lib.dcol_warm_stride(self._h))

    def set_warm_start(self, state=None, capacity_pairs: int = 0, mu_save: float = 1e-2):
        """Opt-in warm start (`dcol_set_warm_start`; not in the reference, excluded from parity).  `state`: a zero-initialised
        CUDA tensor of capacity_pairs * warm_stride() doubles that the caller keeps alive; None switches it off."""
        self._warm_keepalive = state
        self._check(self._lib.dcol_set_warm_start(self._h, _ptr(state), int(capacity_pairs), float(mu_save)))

    def set_max_iters(self, k: int = 50):
        """Testing knob (dcol_set_max_iters): iteration cap of the solver, 50 in the reference and by default."""
        self._check(self._lib.dcol_set_max_iters(self._h, int(k)))

    def _one(self, kind, s1, s2, prim1, prim2, pdip_tol):
        """One pair through preallocated buffers (the single-pair drop-in calls sit in the inner loops of the reference's
        examples: no array allocation, conversion or pointer construction per call).  Same entry points as `_batch`."""
        att_kind = prim1.att_kind
        b = self._one_bufs.get(att_kind)
        if b is None:
            na = 4 if att_kind == ATT_QUAT else 3
            nt = 6 + 2 * na
            a = {"s1": np.empty(1, np.int32), "s2": np.empty(1, np.int32), "r1": np.empty((1, 3)), "a1": np.empty((1, na)),
                 "r2": np.empty((1, 3)), "a2": np.empty((1, na)), "alpha": np.empty(1), "x": np.empty((1, 3)),
                 "grad": np.empty((1, nt)), "J": np.empty((1, nt, 4)), "status": np.empty(1, np.int32), "iters": np.empty(1, np.int32)}
            b = self._one_bufs[att_kind] = (a, {k: _ptr(v) for k, v in a.items()})
        a, q = b
        a["s1"][0] = s1; a["s2"][0] = s2
        a["r1"][0] = prim1.r; a["a1"][0] = prim1.att; a["r2"][0] = prim2.r; a["a2"][0] = prim2.att
        if self._table_version is not None:
            self.sync_shapes()
        args = (self._h, 1, q["s1"], q["s2"], q["r1"], q["a1"], q["r2"], q["a2"], att_kind, float(pdip_tol), MEM_HOST)
        if kind == 0:
            rc = self._lib.dcol_proximity(*args, q["alpha"], q["x"], q["status"], q["iters"])
        elif kind == 1:
            rc = self._lib.dcol_proximity_gradient(*args, q["alpha"], q["grad"], q["status"], q["iters"])
        else:
            rc = self._lib.dcol_proximity_jacobian(*args, q["alpha"], q["x"], q["J"], q["status"], q["iters"])
        self._check(rc)
        return a

    def _batch(self, kind, shape1, shape2, r1, att1, r2, att2, att_kind, pdip_tol, out=None, want_mu=False):
        shape1 = np.ascontiguousarray(shape1, dtype=np.int32).reshape(-1)
        shape2 = np.ascontiguousarray(shape2, dtype=np.int32).reshape(-1)
        n = shape1.shape[0]
        na = 4 if att_kind == ATT_QUAT else 3
        nt = 6 + 2 * na
        r1 = np.ascontiguousarray(r1, dtype=np.float64).reshape(n, 3)
        r2 = np.ascontiguousarray(r2, dtype=np.float64).reshape(n, 3)
        att1 = np.ascontiguousarray(att1, dtype=np.float64).reshape(n, na)
        att2 = np.ascontiguousarray(att2, dtype=np.float64).reshape(n, na)
        if shape2.shape[0] != n:
            raise ValueError("shape1 and shape2 must have the same length")
        if out is not None:        # caller-owned (e.g. pinned) buffers from alloc_outputs
            alpha, status, iters = out["alpha"], out["status"], out["iters"]
            x, deriv = out.get("x"), out.get("deriv")
            want = (n, nt) if kind == 1 else (n, nt, 4)
            if alpha.shape != (n,) or (kind != 1 and x.shape != (n, 3)) or (kind != 0 and deriv.shape != want):
                raise ValueError("out: buffers do not match this query (use alloc_outputs with the same kind, n, att_kind)")
        else:
            alpha = np.empty(n)
            x = np.empty((n, 3)) if kind != 1 else None
            deriv = np.empty((n, nt)) if kind == 1 else (np.empty((n, nt, 4)) if kind == 2 else None)
            status = np.empty(n, dtype=np.int32)
            iters = np.empty(n, dtype=np.int32)
        mu = np.empty(n) if want_mu else None
        if want_mu:
            self._check(self._lib.dcol_set_mu_output(self._h, _ptr(mu)))
        try:
            self.query_raw(kind, n, shape1, shape2, r1, att1, r2, att2, att_kind, pdip_tol, MEM_HOST,
                           alpha, x, deriv, status, iters)
        finally:
            if want_mu:
                self._lib.dcol_set_mu_output(self._h, None)
        out = {"alpha": alpha, "status": status, "iters": iters}
        if want_mu:
            out["mu"] = mu
        if x is not None:
            out["x"] = x
        if kind == 1:
            out["grad"] = deriv
        elif kind == 2:
            out["J"] = deriv.transpose(0, 2, 1)  # (n, 4, nt) view of the column-major C layout, no copy
        return out

    def _alloc_out(self, kind, n, nt):
        out = {"alpha": np.empty(n), "status": np.empty(n, dtype=np.int32), "iters": np.empty(n, dtype=np.int32)}
        x = np.empty((n, 3)) if kind != 1 else None
        deriv = np.empty((n, nt)) if kind == 1 else (np.empty((n, nt, 4)) if kind == 2 else None)
        return out, x, deriv

    @staticmethod
    def _finish(out, kind, x, deriv):
        if x is not None:
            out["x"] = x
        if kind == 1:
            out["grad"] = deriv
        elif kind == 2:
            out["J"] = deriv.transpose(0, 2, 1)  # (n, 4, nt) view of the column-major C layout, no copy
        return out

    def query_indexed(self, kind, shape1, shape2, pose1, pose2, r1, att1, r2, att2, att_kind=ATT_QUAT, pdip_tol=1e-6):
        """Pose-table variant (`dcol_query_indexed`): pair k uses rows pose1[k] of (r1, att1) and pose2[k] of (r2, att2)."""
        if self._table_version is not None:
            self.sync_shapes()
        shape1 = np.ascontiguousarray(shape1, dtype=np.int32).reshape(-1)
        shape2 = np.ascontiguousarray(shape2, dtype=np.int32).reshape(-1)
        pose1 = np.ascontiguousarray(pose1, dtype=np.int32).reshape(-1)
        pose2 = np.ascontiguousarray(pose2, dtype=np.int32).reshape(-1)
        n = shape1.shape[0]
        na = 4 if att_kind == ATT_QUAT else 3
        nt = 6 + 2 * na
        r1 = np.ascontiguousarray(r1, dtype=np.float64).reshape(-1, 3); att1 = np.ascontiguousarray(att1, dtype=np.float64).reshape(-1, na)
        r2 = np.ascontiguousarray(r2, dtype=np.float64).reshape(-1, 3); att2 = np.ascontiguousarray(att2, dtype=np.float64).reshape(-1, na)
        if n and (pose1.max() >= r1.shape[0] or pose2.max() >= r2.shape[0] or pose1.min() < 0 or pose2.min() < 0):
            raise IndexError("pose index out of range")
        out, x, deriv = self._alloc_out(kind, n, nt)
        self._check(self._lib.dcol_query_indexed(self._h, int(kind), n, _ptr(shape1), _ptr(shape2), _ptr(pose1), _ptr(pose2),
                                                 r1.shape[0], _ptr(r1), _ptr(att1), r2.shape[0], _ptr(r2), _ptr(att2),
                                                 int(att_kind), float(pdip_tol), MEM_HOST,
                                                 _ptr(out["alpha"]), _ptr(x), _ptr(deriv), _ptr(out["status"]), _ptr(out["iters"])))
        return self._finish(out, kind, x, deriv)

    def scene_set_cull(self, alpha_cut: float = 0.0):
        """Opt-in broad phase of scene_allpairs / scene_wrench (`dcol_scene_set_cull`): pairs whose bounding balls prove
        alpha* > alpha_cut are answered without the conic solve (alpha = a lower bound > alpha_cut, status 3); 0 = off."""
        self._check(self._lib.dcol_scene_set_cull(self._h, float(alpha_cut)))

    def scene_allpairs(self, kind, shape, r, att, att_kind=ATT_QUAT, pdip_tol=1e-6, pair_range=None):
        """All i<j pairs of a scene in `combinations` order (`dcol_scene_allpairs`); pair list generated on the device.
        `pair_range=(k0, k1)`: only that slice of the enumeration (`dcol_scene_allpairs_range`, one shard of SURVEY.md 8e)."""
        if self._table_version is not None:
            self.sync_shapes()
        shape, r, att, nb, nt = _scene_arrays(shape, r, att, att_kind)
        k0, k1 = _pair_range(self._lib, nb, pair_range)
        out, x, deriv = self._alloc_out(kind, k1 - k0, nt)
        self._check(self._lib.dcol_scene_allpairs_range(self._h, int(kind), nb, k0, k1, _ptr(shape), _ptr(r), _ptr(att), int(att_kind), float(pdip_tol),
                                                        MEM_HOST, _ptr(out["alpha"]), _ptr(x), _ptr(deriv), _ptr(out["status"]), _ptr(out["iters"])))
        return self._finish(out, kind, x, deriv)

    def scene_wrench(self, shape, r, att, lam, att_kind=ATT_QUAT, pdip_tol=1e-6, per_pair=True, pair_range=None):
        """`dcol_scene_wrench`: alpha of all i<j pairs and, per body, sum_k lam[k] * [d alpha_k/d r_b ; G(q_b)' d alpha_k/d q_b]
        (the accumulation loop of examples/contact_physics/ncp_contact_simulation.jl:133-139); the gradient rows stay on
        the device.  per_pair=False skips the alpha / status / iters arrays.  `pair_range=(k0, k1)`: the partial sums of
        that slice of the pairs (`dcol_scene_wrench_range`; `lam` holds the slice's k1 - k0 weights)."""
        if self._table_version is not None:
            self.sync_shapes()
        shape, r, att, nb, _ = _scene_arrays(shape, r, att, att_kind)
        k0, k1 = _pair_range(self._lib, nb, pair_range)
        n = k1 - k0
        lam = np.ascontiguousarray(lam, dtype=np.float64).reshape(-1)
        if lam.shape[0] != n:
            raise ValueError(f"scene_wrench: lam must have one entry per pair ({n})")
        out = {"wrench": np.empty((nb, 6))}
        if per_pair:
            out.update(alpha=np.empty(n), status=np.empty(n, dtype=np.int32), iters=np.empty(n, dtype=np.int32))
        self._check(self._lib.dcol_scene_wrench_range(self._h, nb, k0, k1, _ptr(shape), _ptr(r), _ptr(att), _ptr(lam), int(att_kind), float(pdip_tol),
                                                      MEM_HOST, _ptr(out.get("alpha")), _ptr(out["wrench"]), _ptr(out.get("status")), _ptr(out.get("iters"))))
        return out

    def se2_constraints(self, xytheta, vehicle_shape, obs_shape, r_obs, p_obs, pdip_tol=1e-6):
        """`dcol_se2_constraints`: `ineq_con_x` / `ineq_con_x_jac` of examples/trajectory_optimization/piano_mover.jl:39-53 for a
        batch of planar states (x, y, theta) against K obstacles: c = 1 - alpha (n, K) and dc/d(x, y, theta) (n, K, 3)."""
        if self._table_version is not None:
            self.sync_shapes()
        xyt = np.ascontiguousarray(xytheta, dtype=np.float64).reshape(-1, 3)
        obs_shape = np.ascontiguousarray(obs_shape, dtype=np.int32).reshape(-1)
        K = obs_shape.shape[0]; n = xyt.shape[0]
        r_obs = np.ascontiguousarray(r_obs, dtype=np.float64).reshape(K, 3); p_obs = np.ascontiguousarray(p_obs, dtype=np.float64).reshape(K, 3)
        c = np.empty((n, K)); Jc = np.empty((n, K, 3)); status = np.empty((n, K), dtype=np.int32)
        self._check(self._lib.dcol_se2_constraints(self._h, n, _ptr(xyt), int(vehicle_shape), K, _ptr(obs_shape), _ptr(r_obs), _ptr(p_obs),
                                                   float(pdip_tol), MEM_HOST, _ptr(c), _ptr(Jc), _ptr(status)))
        return {"c": c, "Jc": Jc, "status": status}

    def se3_constraints(self, states, vehicle_shape, obs_shape, r_obs, p_obs, r_col=0, p_col=6, pdip_tol=1e-6):
        """`dcol_se3_constraints`: `ineq_con_x` / `ineq_con_x_jac` of examples/trajectory_optimization/cone_through_wall.jl:56-68
        and cluttered_hallway_quadrotor.jl:83-104 for a batch of states (n, stride) whose columns r_col.. / p_col.. hold the
        vehicle's position / MRP, against K obstacles: c = 1 - alpha (n, K) and the non-zero blocks of the constraint
        Jacobian rows, Jc[..., :3] = -d alpha/d r, Jc[..., 3:] = -d alpha/d p (n, K, 6)."""
        if self._table_version is not None:
            self.sync_shapes()
        X = np.ascontiguousarray(states, dtype=np.float64)
        if X.ndim != 2:
            raise ValueError("states must be (n, stride)")
        n, stride = X.shape
        obs_shape = np.ascontiguousarray(obs_shape, dtype=np.int32).reshape(-1)
        K = obs_shape.shape[0]
        r_obs = np.ascontiguousarray(r_obs, dtype=np.float64).reshape(K, 3); p_obs = np.ascontiguousarray(p_obs, dtype=np.float64).reshape(K, 3)
        c = np.empty((n, K)); Jc = np.empty((n, K, 6)); status = np.empty((n, K), dtype=np.int32)
        self._check(self._lib.dcol_se3_constraints(self._h, n, _ptr(X), int(stride), int(r_col), int(p_col), int(vehicle_shape), K, _ptr(obs_shape),
                                                   _ptr(r_obs), _ptr(p_obs), float(pdip_tol), MEM_HOST, _ptr(c), _ptr(Jc), _ptr(status)))
        return {"c": c, "Jc": Jc, "status": status}

    def proximity_batch(self, shape1, shape2, r1, att1, r2, att2, att_kind=ATT_QUAT, pdip_tol=1e-6, out=None):
        return self._batch(0, shape1, shape2, r1, att1, r2, att2, att_kind, pdip_tol, out)

    def proximity_gradient_batch(self, shape1, shape2, r1, att1, r2, att2, att_kind=ATT_QUAT, pdip_tol=1e-6, out=None):
        return self._batch(1, shape1, shape2, r1, att1, r2, att2, att_kind, pdip_tol, out)

    def proximity_jacobian_batch(self, shape1, shape2, r1, att1, r2, att2, att_kind=ATT_QUAT, pdip_tol=1e-6, out=None):
        return self._batch(2, shape1, shape2, r1, att1, r2, att2, att_kind, pdip_tol, out)


class Group:
    """Several GPUs from one process (`dcol_group`): one context per listed device, every batched query cut into contiguous
    index slices that run concurrently, one per member.  Same batch methods and result layout as `Context`; results are
    bit-identical to the single-context call.  Host arrays only."""

    def __init__(self, devices):
        self._lib = _capi.load()
        devs = np.ascontiguousarray(list(devices), dtype=np.int32)
        h = C.c_void_p()
        rc = self._lib.dcol_group_create(C.byref(h), _ptr(devs), len(devs))
        if rc != 0:
            raise RuntimeError(self._lib.dcol_group_last_error(None).decode())
        self._h = h
        self.devices = [int(d) for d in devs]
        self.table = ShapeTable()
        self._table_version = -1

    def close(self):
        if getattr(self, "_h", None):
            self._lib.dcol_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(self._lib.dcol_group_size(self._h))

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"dcol error {rc}: {self._lib.dcol_group_last_error(self._h).decode()}")

    def add_shape(self, prim) -> int:
        if self._table_version is None:
            self._table_version = -1
        return self.table.add(prim)

    def set_shape_arrays(self, types, offs, params):
        types = np.ascontiguousarray(types, dtype=np.int32)
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        params = np.ascontiguousarray(params, dtype=np.float64)
        self._check(self._lib.dcol_group_set_shapes(self._h, len(types), _ptr(types), _ptr(offs), _ptr(params), params.shape[0]))
        self._table_version = None

    def sync_shapes(self):
        if self._table_version != self.table.version:
            t, o, p = self.table.arrays()
            self._check(self._lib.dcol_group_set_shapes(self._h, len(t), _ptr(t), _ptr(o), _ptr(p), p.shape[0]))
            self._table_version = self.table.version

    def last_kernel_ms(self) -> float:
        return float(self._lib.dcol_group_last_kernel_ms(self._h))

    def pin(self, arr: np.ndarray) -> np.ndarray:
        """Page-lock a numpy array in place for every device of the group (`dcol_host_register` on member 0; the
        registration is portable across devices).  `unpin` before the array is freed."""
        if not arr.flags["C_CONTIGUOUS"]:
            raise ValueError("pin: array must be C-contiguous")
        c0 = self._lib.dcol_group_ctx(self._h, 0)
        if self._lib.dcol_host_register(c0, arr.ctypes.data, arr.nbytes) != 0:
            raise RuntimeError(self._lib.dcol_last_error(c0).decode())
        return arr

    def unpin(self, arr: np.ndarray):
        c0 = self._lib.dcol_group_ctx(self._h, 0)
        if self._lib.dcol_host_unregister(c0, arr.ctypes.data) != 0:
            raise RuntimeError(self._lib.dcol_last_error(c0).decode())

    def _batch(self, kind, shape1, shape2, r1, att1, r2, att2, att_kind=ATT_QUAT, pdip_tol=1e-6, out=None):
        if self._table_version is not None:
            self.sync_shapes()
        shape1 = np.ascontiguousarray(shape1, dtype=np.int32).reshape(-1)
        shape2 = np.ascontiguousarray(shape2, dtype=np.int32).reshape(-1)
        n = shape1.shape[0]
        if shape2.shape[0] != n:
            raise ValueError("shape1 and shape2 must have the same length")
        na = 4 if att_kind == ATT_QUAT else 3
        nt = 6 + 2 * na
        r1 = np.ascontiguousarray(r1, dtype=np.float64).reshape(n, 3)
        r2 = np.ascontiguousarray(r2, dtype=np.float64).reshape(n, 3)
        att1 = np.ascontiguousarray(att1, dtype=np.float64).reshape(n, na)
        att2 = np.ascontiguousarray(att2, dtype=np.float64).reshape(n, na)
        if out is not None:        # caller-owned (e.g. pinned) buffers, laid out as Context.alloc_outputs does
            alpha, status, iters = out["alpha"], out["status"], out["iters"]
            x, deriv = out.get("x"), out.get("deriv")
            want = (n, nt) if kind == 1 else (n, nt, 4)
            if alpha.shape != (n,) or (kind != 1 and x.shape != (n, 3)) or (kind != 0 and deriv.shape != want):
                raise ValueError("out: buffers do not match this query")
        else:
            alpha = np.empty(n)
            x = np.empty((n, 3)) if kind != 1 else None
            deriv = np.empty((n, nt)) if kind == 1 else (np.empty((n, nt, 4)) if kind == 2 else None)
            status = np.empty(n, dtype=np.int32)
            iters = np.empty(n, dtype=np.int32)
        self._check(self._lib.dcol_group_query(self._h, int(kind), n, _ptr(shape1), _ptr(shape2), _ptr(r1), _ptr(att1), _ptr(r2), _ptr(att2),
                                               int(att_kind), float(pdip_tol), _ptr(alpha), _ptr(x), _ptr(deriv), _ptr(status), _ptr(iters)))
        out = {"alpha": alpha, "status": status, "iters": iters}
        if x is not None:
            out["x"] = x
        if kind == 1:
            out["grad"] = deriv
        elif kind == 2:
            out["J"] = deriv.transpose(0, 2, 1)
        return out

    def scene_set_cull(self, alpha_cut: float = 0.0):
        """Broad phase of the group's scene operators: set on every member (`dcol_scene_set_cull`)."""
        for i in range(len(self)):
            if self._lib.dcol_scene_set_cull(self._lib.dcol_group_ctx(self._h, i), float(alpha_cut)) != 0:
                raise RuntimeError("dcol_scene_set_cull failed")

    def scene_allpairs(self, kind, shape, r, att, att_kind=ATT_QUAT, pdip_tol=1e-6):
        """`dcol_group_scene_allpairs`: member i solves the i-th slice of the i<j enumeration; rows as on one device."""
        if self._table_version is not None:
            self.sync_shapes()
        shape, r, att, nb, nt = _scene_arrays(shape, r, att, att_kind)
        n = int(self._lib.dcol_scene_pair_count(nb))
        out, x, deriv = Context._alloc_out(None, kind, n, nt)
        self._check(self._lib.dcol_group_scene_allpairs(self._h, int(kind), nb, _ptr(shape), _ptr(r), _ptr(att), int(att_kind), float(pdip_tol),
                                                        _ptr(out["alpha"]), _ptr(x), _ptr(deriv), _ptr(out["status"]), _ptr(out["iters"])))
        return Context._finish(out, kind, x, deriv)

    def scene_wrench(self, shape, r, att, lam, att_kind=ATT_QUAT, pdip_tol=1e-6, per_pair=True):
        """`dcol_group_scene_wrench`: the members' per-body partial sums added in member order."""
        if self._table_version is not None:
            self.sync_shapes()
        shape, r, att, nb, _ = _scene_arrays(shape, r, att, att_kind)
        n = int(self._lib.dcol_scene_pair_count(nb))
        lam = np.ascontiguousarray(lam, dtype=np.float64).reshape(-1)
        if lam.shape[0] != n:
            raise ValueError(f"scene_wrench: lam must have one entry per pair ({n})")
        out = {"wrench": np.empty((nb, 6))}
        if per_pair:
            out.update(alpha=np.empty(n), status=np.empty(n, dtype=np.int32), iters=np.empty(n, dtype=np.int32))
        self._check(self._lib.dcol_group_scene_wrench(self._h, nb, _ptr(shape), _ptr(r), _ptr(att), _ptr(lam), int(att_kind), float(pdip_tol),
                                                      _ptr(out.get("alpha")), _ptr(out["wrench"]), _ptr(out.get("status")), _ptr(out.get("iters"))))
        return out

    def proximity_batch(self, shape1, shape2, r1, att1, r2, att2, att_kind=ATT_QUAT, pdip_tol=1e-6):
        return self._batch(0, shape1, shape2, r1, att1, r2, att2, att_kind, pdip_tol)

    def proximity_gradient_batch(self, shape1, shape2, r1, att1, r2, att2, att_kind=ATT_QUAT, pdip_tol=1e-6):
        return self._batch(1, shape1, shape2, r1, att1, r2, att2, att_kind, pdip_tol)

    def proximity_jacobian_batch(self, shape1, shape2, r1, att1, r2, att2, att_kind=ATT_QUAT, pdip_tol=1e-6):
        return self._batch(2, shape1, shape2, r1, att1, r2, att2, att_kind, pdip_tol)


_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def _single(kind, prim1, prim2, pdip_tol, ctx, verbose=False):
    if not isinstance(prim1, AbstractPrimitive) or not isinstance(prim2, AbstractPrimitive):
        raise TypeError("proximity expects two primitives")
    if prim1.att_kind != prim2.att_kind:
        raise TypeError("no method: both primitives must be quaternion (AbstractPrimitive) or both MRP (AbstractPrimitiveMRP)")
    ctx = ctx or default_context()
    if verbose:
        # the reference prints one row per interior-point iteration (static_solver2.jl:117-120, 179-183).  The fast
        # kernels keep no per-iteration record, so the table comes from the pair's dense (c, G, h) solved by
        # dcol_solve_socp_trace on the GPU: the same iterates up to round-off; the values returned below are the fast path's.
        ctx.solve_socp(*pair_problem(prim1, prim2), pdip_tol=pdip_tol, verbose=True)
    s1, s2 = ctx.add_shape(prim1), ctx.add_shape(prim2)
    out = ctx._one(kind, s1, s2, prim1, prim2, pdip_tol)
    st = int(out["status"][0])
    if st == 1:
        raise RuntimeError("pdip failed")
    if st != 0:
        raise ArithmeticError("numerical breakdown in the conic solver")
    return out


def proximity(prim1, prim2, pdip_tol: float = 1e-6, verbose: bool = False, ctx: Context | None = None):
    out = _single(0, prim1, prim2, pdip_tol, ctx, verbose)
    return float(out["alpha"][0]), out["x"][0].copy()


def proximity_gradient(prim1, prim2, pdip_tol: float = 1e-6, verbose: bool = False, ctx: Context | None = None):
    out = _single(1, prim1, prim2, pdip_tol, ctx, verbose)
    return float(out["alpha"][0]), out["grad"][0].copy()


def proximity_jacobian(prim1, prim2, pdip_tol: float = 1e-6, verbose: bool = False, ctx: Context | None = None):
    out = _single(2, prim1, prim2, pdip_tol, ctx, verbose)
    return float(out["alpha"][0]), out["x"][0].copy(), out["J"][0].T.copy()      # (4, nt) from the column-major C layout
