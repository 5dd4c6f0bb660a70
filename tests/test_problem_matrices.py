"""Host mirror of the dense conic form (dcol_b200.problem_matrices, behind verbose=True and Context.solve_socp) against
the oracle's restatement of src/problem_matrices.jl + src/combine_problem_matrices.jl, for all 49 type pairs, both
attitude kinds, with shape offsets; and the trace table's format (static_solver2.jl:117-120, 179-183)."""
import numpy as np
import pytest

import dcol_b200 as dc
import util
from dcol_b200.problem_matrices import TRACE_HEADER, format_trace
from oracle import oracle


@pytest.mark.parametrize("att_kind", [dc.ATT_QUAT, dc.ATT_MRP])
def test_pair_problem_matches_oracle(att_kind):
    g = util.rng(21 + att_kind)
    for variant in ("cube", "scene"):
        shapes1, shapes2 = util.standard_shapes(att_kind, variant), util.standard_shapes(att_kind, variant)
        for p in shapes1 + shapes2:                     # offsets as the examples set them (r_offset, Q_offset)
            p.r_offset = 0.3 * g.standard_normal(3)
            p.Q_offset = dc.dcm_from_q(g.standard_normal(4))
        r1, a1, r2, a2 = util.random_poses(49, att_kind, seed=5)
        k = 0
        for p1 in shapes1:
            for p2 in shapes2:
                p1.r, p2.r = r1[k], r2[k]
                if att_kind == dc.ATT_QUAT:
                    p1.q, p2.q = 1.7 * a1[k], a2[k]    # not unit length: dcm_from_q normalises
                else:
                    p1.p, p2.p = a1[k], a2[k]
                k += 1
                c, G, h, n_ort, d1, d2 = dc.pair_problem(p1, p2)
                oc, oG, oh, on, o1, o2 = oracle.pair_problem(p1, p2)
                assert (n_ort, d1, d2) == (on, o1, o2) and G.shape == oG.shape
                np.testing.assert_array_equal(c, oc)
                np.testing.assert_allclose(G, oG, rtol=0, atol=1e-13)
                np.testing.assert_allclose(h, oh, rtol=0, atol=1e-12)


def test_trace_table_format():
    tr = np.full((50, 5), np.nan)
    tr[0] = [1.2345, 3.2e-1, 4.5e-16, 2.2e-3, 0.98765]
    tr[1] = [0.5, 1e-7, 1e-15, 1e-9, 1.0]
    txt = format_trace(tr, 2).split("\n")
    assert "\n".join(txt[:2]) == TRACE_HEADER and len(txt) == 4
    assert txt[2] == "  1    1.234e+00   3.20e-01   4.50e-16   2.20e-03   0.9877"   # "%3d   %10.3e  %9.2e  %9.2e  %9.2e  % 6.4f"
    assert txt[3].endswith(" 1.0000")
