"""Summarise an `ncu --page source --csv` export by function and by opcode (stall samples vs executed instructions).
usage: ncu_source_summary.py gpurun_out/mixed_source.csv > profiles/r01e_source_hotspots.txt"""
import collections, csv, re, sys
csv.field_size_limit(1 << 30)
rows = list(csv.reader(open(sys.argv[1])))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
num = lambda x: int(float(x)) if x not in ("", None) else 0
secs = []
for n, hi in enumerate(heads):
    end = heads[n + 1] - 1 if n + 1 < len(heads) else len(rows)
    name = rows[hi - 1][1] if hi > 0 and len(rows[hi - 1]) > 1 else "?"
    ix = {h: i for i, h in enumerate(rows[hi])}
    data = [r for r in rows[hi + 1:end] if len(r) > ix["# Samples"]]
    samples = sum(num(r[ix["# Samples"]]) for r in data)
    execd = sum(num(r[ix["Instructions Executed"]]) for r in data)
    secs.append((name, samples, execd, data, ix))
tot_s = sum(s[1] for s in secs); tot_e = sum(s[2] for s in secs)
print(f"# warp-stall samples and executed warp instructions by function ({tot_s} samples, {tot_e} instructions)")
agg = collections.defaultdict(lambda: [0, 0, 0])
for name, s, e, data, ix in secs:
    key = re.sub(r"<\(int\)(\d), \(int\)(\d), \(bool\)(\d)>", r"<\1,\2,\3>", name)
    key = re.sub(r"solve_entry<(\d),(\d),(\d)>.*", r"solve_entry<\1,\2,\3>", key)
    agg[key][0] += s; agg[key][1] += e; agg[key][2] += len(data)
for k, (s, e, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:24]:
    print(f"  {100*s/tot_s:5.1f}% samples  {100*e/tot_e:5.1f}% executed  {n:6d} SASS instrs  {k[:100]}")
ops = collections.Counter(); opx = collections.Counter()
for name, s, e, data, ix in secs:
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        op = m.group(2).split(".")[0] if m else "?"
        ops[op] += num(r[ix["# Samples"]]); opx[op] += num(r[ix["Instructions Executed"]])
print("# by opcode")
for k, c in ops.most_common(16):
    print(f"  {k:8s} {100*c/tot_s:5.1f}% samples  {100*opx[k]/tot_e:5.1f}% executed")
