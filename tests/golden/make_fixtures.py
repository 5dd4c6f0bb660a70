"""Decode the reference's JLD2 input fixtures into JSON (run in the build container only;
/root/reference does not exist on the GPU box).

  test/example_socps/polytopes.jld2      -> polytopes.json      (A1 14x3, b1, A2 8x3, b2; test/polytope_derivs_test.jl:3-8)
  test/example_socps/example_socp_2.jld2 -> example_socp_2.json (G 10x5, h, n_ort, n_soc1, n_soc2; test/solver_tests.jl:38-57)

JLD2 is HDF5-based; the datasets are contiguous little-endian f64 in Julia column-major order.
There is no h5py here, so arrays are located by scanning for runs of finite, plausibly-sized
doubles of the expected length and matched to names by their (unique) lengths.
"""
import json
import os
import struct
import sys

import numpy as np

REF = "/root/reference/test/example_socps"
HERE = os.path.dirname(os.path.abspath(__file__))


def runs_of_doubles(buf, min_len):
    """All maximal 8-byte-aligned runs of 'plausible' doubles (finite, |v| in [1e-20, 1e6] or 0)."""
    out = []
    for align in range(8):
        n = (len(buf) - align) // 8
        v = np.frombuffer(buf, dtype="<f8", count=n, offset=align)
        ok = np.isfinite(v) & ((np.abs(v) < 1e6) & ((np.abs(v) > 1e-20) | (v == 0)))
        i = 0
        while i < n:
            if ok[i]:
                j = i
                while j < n and ok[j]:
                    j += 1
                if j - i >= min_len:
                    out.append((align + 8 * i, v[i:j].copy()))
                i = j
            else:
                i += 1
    return out


def find_array(buf, length, pred):
    """First window of `length` doubles inside any plausible run satisfying pred(window)."""
    for off, run in runs_of_doubles(buf, length):
        for k in range(0, len(run) - length + 1):
            w = run[k:k + length]
            if pred(w):
                return w
    raise RuntimeError("array not found")


def main():
    if not os.path.isdir(REF):
        print("reference not mounted; fixtures are already committed", file=sys.stderr)
        return
    # --- polytopes.jld2: A x <= b for two polytopes.  Column-major A (n x 3).
    buf = open(os.path.join(REF, "polytopes.jld2"), "rb").read()

    def is_A(n):
        def pred(w):
            A = w.reshape(3, n).T                     # column-major -> (n,3)
            nz = np.abs(A).sum(axis=1)
            return bool((nz > 1e-3).all() and np.abs(w).max() < 10 and np.count_nonzero(np.abs(w) > 1e-12) > n)
        return pred
    A1 = find_array(buf, 42, lambda w: is_A(14)(w) and abs(w[1] - 1.732050807568876) < 1e-12).reshape(3, 14).T
    A2 = find_array(buf, 24, lambda w: abs(w[0] - 2.4494897427831774) < 1e-12).reshape(3, 8).T
    b1 = find_array(buf, 14, lambda w: abs(w[0] - 0.5629165124598854) < 1e-12)
    b2 = find_array(buf, 8, lambda w: abs(w[0] - 4.330127018922189) < 1e-12)
    json.dump({"A1": A1.tolist(), "b1": b1.tolist(), "A2": A2.tolist(), "b2": b2.tolist(),
               "source": "test/example_socps/polytopes.jld2"},
              open(os.path.join(HERE, "polytopes.json"), "w"), indent=1)
    # --- example_socp_2.jld2: G (10x5 column-major), h (10), n_ort=3, n_soc1=4, n_soc2=3
    buf = open(os.path.join(REF, "example_socp_2.jld2"), "rb").read()
    G = find_array(buf, 50, lambda w: abs(w[2] + 0.5978702047461094) < 1e-12).reshape(5, 10).T
    h = find_array(buf, 10, lambda w: abs(w[2] + 0.20271801759598262) < 1e-12 and w[0] == 0 and w[1] == 0)
    json.dump({"G": G.tolist(), "h": h.tolist(), "n_ort": 3, "n_soc1": 4, "n_soc2": 3, "c": [0, 0, 0, 1.0, 0],
               "source": "test/example_socps/example_socp_2.jld2 (dims from test/solver_tests.jl:38-57)"},
              open(os.path.join(HERE, "example_socp_2.json"), "w"), indent=1)
    print("A1", A1.shape, "A2", A2.shape, "G", G.shape)
    print(G)
    print(h)


if __name__ == "__main__":
    main()
