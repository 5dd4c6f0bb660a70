"""Extract the known-answer tables of SURVEY.md Appendix C into appendix_c.json.

Provenance: these numbers come from an independent NumPy restatement of the reference algorithm
(per-block Cholesky `W\\`, complex-step dR/dtheta, dense solves) made during the survey — NOT from
Julia (Julia is not installed in this image).  They pin the C++ oracle to a second, independently
written implementation; they do not prove parity with Julia output.
"""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
SURVEY = os.path.join(HERE, "..", "..", "SURVEY.md")


def main():
    txt = open(SURVEY).read().split("## Appendix C")[1].split("## Appendix D")[0]
    rows = []
    for line in txt.splitlines():
        m = re.match(r"\|\s*(\w+)\s*\|\s*(\w+)\s*\|\s*(\d+)\s*\|\s*(\d+)\s*\|\s*(\d+)\s*\|\s*(\d+)\s*\|\s*(\d+)\s*\|\s*([-\d.]+)\s*\|\s*([-\d.]+)\s*\|\s*([-\d.]+)\s*\|\s*([-\d.]+)\s*\|", line)
        if m and m.group(1) != "P1":
            g = m.groups()
            rows.append({"p1": g[0], "p2": g[1], "nx": int(g[2]), "n_ort": int(g[3]), "d1": int(g[4]), "d2": int(g[5]),
                         "iters": int(g[6]), "alpha": float(g[7]), "x": [float(g[8]), float(g[9]), float(g[10])]})
    assert len(rows) == 49, len(rows)
    jac = {}
    block = txt.split("```")[1]
    cur = None
    for line in block.splitlines():
        m = re.match(r"^(\S+) \[(quat|MRP)\]", line)
        if m:
            cur = m.group(1) + "/" + m.group(2)
            jac[cur] = {"J": [], "grad": None}
            continue
        nums = [float(v) for v in re.findall(r"-?\d+\.\d+", line)]
        if cur is None or not nums:
            continue
        if "∇α" in line:
            jac[cur]["grad"] = nums
        else:
            jac[cur]["J"].append(nums)
    for k, v in jac.items():
        nt = 14 if k.endswith("quat") else 12
        assert len(v["J"]) == 4 and all(len(r) == nt for r in v["J"]) and len(v["grad"]) == nt, k
    out = {
        "provenance": "SURVEY.md Appendix C: independent NumPy restatement of the reference algorithm, NOT Julia output",
        "pdip_tol": 1e-6,
        "pose1": {"r": [0.1, -0.2, 0.3], "q_unnormalised": [1, 0.1, 0.2, 0.3]},
        "pose2": {"r": [1.5, 0.3, -0.2], "q_unnormalised": [1, 0.2, -0.3, 0.4]},
        "pairs": rows,
        "jacobians": jac,
        "fixture_socp": {
            "tol1e-6": {"iters": 5, "x": [-0.128792427397, -0.091923544485, -0.213354575493, 1.626482022973, -0.168678665753]},
            "tol1e-12": {"iters": 9, "x": [-0.129236812242, -0.092306939614, -0.213119833927, 1.626481006844, -0.168162432719]}},
        "fixture_polytopes": {"tol1e-6": {"iters": 9, "alpha": 0.685853085869, "x": [0.486896874298, 0.092664820593, -0.017712796626]},
                              "tol1e-8": {"iters": 10, "alpha": 0.685852496383}},
    }
    json.dump(out, open(os.path.join(HERE, "appendix_c.json"), "w"), indent=1)
    print(len(rows), "pairs;", list(jac))


if __name__ == "__main__":
    main()
