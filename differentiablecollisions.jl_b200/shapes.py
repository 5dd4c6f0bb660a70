"""Immutable shape table in the packing `include/dcol_b200.h` documents.

type: 0 polytope, 1 capsule, 2 cylinder, 3 cone, 4 sphere, 5 ellipsoid, 6 polygon.
params (f64, packed at param_off[i]): polytope {n, A row-major n x 3, b[n]}; capsule {R,L};
cylinder {R,L}; cone {H,beta}; sphere {R}; ellipsoid {U row-major 3x3 (= chol(P).U)};
polygon {nh, A nh x 2, b[nh], R}; then r_offset[3], Q_offset[9] row-major for every shape.
"""
from __future__ import annotations

import numpy as np


class ShapeTable:
    """De-duplicating registry: primitives with identical geometry share one shape id."""

    def __init__(self):
        self._keys = {}
        self.types = []
        self._chunks = []
        self._offs = []
        self._n_params = 0
        self.version = 0

    def add(self, prim) -> int:
        key = prim.shape_key()
        sid = self._keys.get(key)
        if sid is not None:
            return sid
        packed = prim.packed()
        sid = len(self.types)
        self._keys[key] = sid
        self.types.append(int(prim.shape_type))
        self._offs.append(self._n_params)
        self._chunks.append(packed)
        self._n_params += packed.shape[0]
        self.version += 1
        return sid

    def __len__(self):
        return len(self.types)

    def arrays(self):
        """(type int32[n], param_off int64[n], params f64[n_params])"""
        t = np.asarray(self.types, dtype=np.int32)
        o = np.asarray(self._offs, dtype=np.int64)
        p = np.concatenate(self._chunks) if self._chunks else np.zeros(0)
        return t, o, np.ascontiguousarray(p, dtype=np.float64)
