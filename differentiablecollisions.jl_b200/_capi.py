"""ctypes binding of libdcol_b200.so (the C ABI of include/dcol_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is
present, loading / context creation raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DCOL_B200_LIB") or os.path.join(_HERE, "libdcol_b200.so")

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)
lp = C.POINTER(C.c_int64)

_Q_ARGS = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
           C.c_int, C.c_double, C.c_int]

TRACE_ROWS, TRACE_COLS = 50, 5     # DCOL_TRACE_ROWS / DCOL_TRACE_COLS of include/dcol_b200.h

# name -> (restype, argtypes); this is the complete list of symbols include/dcol_b200.h declares
SYMBOLS = {
    "dcol_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "dcol_destroy": (None, [C.c_void_p]),
    "dcol_last_error": (C.c_char_p, [C.c_void_p]),
    "dcol_version": (C.c_char_p, []),
    "dcol_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dcol_set_shapes": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "dcol_proximity": (C.c_int, _Q_ARGS + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcol_proximity_gradient": (C.c_int, _Q_ARGS + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcol_proximity_jacobian": (C.c_int, _Q_ARGS + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcol_query_indexed": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                     C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcol_scene_pair_count": (C.c_int64, [C.c_int]),
    "dcol_scene_allpairs": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcol_scene_allpairs_range": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                            C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcol_scene_wrench_range": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                          C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcol_scene_set_cull": (C.c_int, [C.c_void_p, C.c_double]),
    "dcol_submit": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcol_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "dcol_scene_wrench": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcol_se2_constraints": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcol_se3_constraints": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int32, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcol_host_register": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "dcol_host_unregister": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dcol_solve_socp": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcol_solve_socp_trace": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcol_warm_stride": (C.c_int64, [C.c_void_p]),
    "dcol_set_warm_start": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_double]),
    "dcol_set_mu_output": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dcol_set_max_iters": (C.c_int, [C.c_void_p, C.c_int]),
    "dcol_launch_count": (C.c_int64, [C.c_void_p]),
    "dcol_last_kernel_ms": (C.c_double, [C.c_void_p]),
    "dcol_measure_fp64_tflops": (C.c_double, [C.c_void_p]),
    "dcol_measure_fp64_tflops_for": (C.c_double, [C.c_void_p, C.c_double]),
    "dcol_measure_fp64_tflops_mode": (C.c_double, [C.c_void_p, C.c_double, C.c_int]),
    "dcol_set_copy_only": (C.c_int, [C.c_void_p, C.c_int]),
    "dcol_host_copy_selftest": (C.c_int, [C.c_int, C.c_int64, C.c_int]),
    "dcol_group_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_int]),
    "dcol_group_destroy": (None, [C.c_void_p]),
    "dcol_group_last_error": (C.c_char_p, [C.c_void_p]),
    "dcol_group_size": (C.c_int, [C.c_void_p]),
    "dcol_group_ctx": (C.c_void_p, [C.c_void_p, C.c_int]),
    "dcol_group_set_shapes": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "dcol_group_query": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcol_group_scene_allpairs": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcol_group_scene_wrench": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcol_group_last_kernel_ms": (C.c_double, [C.c_void_p]),
}

_lib = None


def load():
    """Load the CUDA library; raise if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). dcol_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        if os.environ.get("DCOL_B200_LIB") and not hasattr(lib, name):
            continue             # A/B runs against an older build selected through DCOL_B200_LIB
        fn = getattr(lib, name)  # AttributeError if a declared symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
