"""CPU: the C-ABI library loads and exports every symbol include/dcol_b200.h declares; host-side logic."""
import ctypes
import os
import re

import numpy as np
import pytest

import dcol_b200 as dc
from dcol_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "dcol_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(dcol_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported():
    names = declared_symbols()
    assert len(names) >= 12
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dcol_b200.h but not exported"
    assert set(names) == set(_capi.SYMBOLS), "ctypes binding table out of sync with the header"


def test_version_and_no_cpu_fallback():
    lib = _capi.load()
    assert b"dcol_b200" in lib.dcol_version()
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
            dc.Context(0)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "differentiablecollisions.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), f"{f} references the oracle"


def test_shape_table_packing():
    t = dc.ShapeTable()
    A, b = dc.create_n_sided(5, 0.6)
    prims = [dc.create_rect_prism(1, 2, 3, attitude="quat")[0], dc.Capsule(.3, 1.2), dc.Cylinder(.6, 3.), dc.Cone(2., .4),
             dc.Sphere(.8), dc.Ellipsoid(np.diag([1., 2., 3.])), dc.Polygon(A, b, .2)]
    ids = [t.add(p) for p in prims]
    assert ids == list(range(7)) and t.add(dc.Capsule(.3, 1.2)) == 1          # de-duplicated
    types, offs, params = t.arrays()
    assert list(types) == [0, 1, 2, 3, 4, 5, 6]
    assert params[offs[0]] == 6 and params[offs[6]] == 5
    assert offs[1] - offs[0] == 1 + 24 + 12 and len(params) - offs[6] == 1 + 15 + 1 + 12
    np.testing.assert_allclose(params[offs[5]:offs[5] + 9].reshape(3, 3), np.diag(np.sqrt([1., 2., 3.])))
    # MRP twin has its own type object but the same geometry packing
    assert dc.CapsuleMRP(.3, 1.2).att_kind == dc.ATT_MRP and issubclass(dc.CapsuleMRP, dc.AbstractPrimitiveMRP)
    prism, mass, inertia = dc.create_rect_prism(1, 2, 3)
    assert isinstance(prism, dc.PolytopeMRP) and mass == 6 and inertia.shape == (3, 3)   # misc_primitive_constructors.jl:1,21


def test_flops_model_matches_survey_examples():
    fm = dc.flops_model
    assert fm.f_iter(4, 12, []) == pytest.approx(1585, abs=1)       # SURVEY §8d cube-cube
    assert fm.f_fixed(4, 12) == 1224


def test_mass_properties_host_helper():
    """src/mass_properties.jl:2-48 conventions (x = symmetry axis); checked against independently computed numbers."""
    import math
    m, I = dc.mass_properties(dc.Cylinder(0.6, 3.0), rho=2.0)
    assert m == pytest.approx(2.0 * math.pi * 0.36 * 3.0) and I[0, 0] == pytest.approx(0.5 * m * 0.36)
    assert I[1, 1] == pytest.approx(m / 12 * (2 * 0.36 + 9.0)) and I[2, 2] == I[1, 1]
    m, I = dc.mass_properties(dc.CapsuleMRP(0.3, 1.2))
    vol = math.pi * 0.09 * 1.2 + 4.0 / 3.0 * math.pi * 0.027
    assert m == pytest.approx(vol) and I[0, 0] < I[1, 1]
    m, I = dc.mass_properties(dc.Cone(2.0, math.radians(22)))
    rad = 2.0 * math.tan(math.radians(22))
    assert m == pytest.approx(math.pi * rad**2 * 2.0 / 3) and I[0, 0] == pytest.approx(0.3 * m * rad**2)
    with pytest.raises(TypeError):
        dc.mass_properties(dc.Sphere(1.0))


@pytest.mark.parametrize("threads", [1, 2, 7])
def test_host_copy_threads(threads):
    """The copy engine behind host-pointer calls on pageable memory (HostCopier, dcol_capi.cu) needs no GPU: ragged slices,
    many batches (workers go from spinning to sleeping and back in between), every byte checked."""
    lib = _capi.load()
    assert lib.dcol_host_copy_selftest(threads, 9 * 1024 * 1024 + 13, 12) == 0
    assert lib.dcol_host_copy_selftest(threads, 1, 3) == 0
    import time
    time.sleep(0.01)                                   # longer than the workers' spin window
    assert lib.dcol_host_copy_selftest(threads, 300_001, 6) == 0
    assert lib.dcol_host_copy_selftest(0, 10, 1) == -1
