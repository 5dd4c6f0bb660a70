"""Join an `ncu --page source --print-source sass --csv` export with `nvdisasm -g -c` of the same cubin and aggregate
warp-stall samples / executed instructions by CUDA source line (and by stall reason).

usage: ncu_lines.py <sass.csv> <nvdisasm.dis> [top_n]  > profiles/rNN_source_lines.txt
"""
import collections, csv, re, sys
csv.field_size_limit(1 << 30)
num = lambda x: int(float(x)) if x not in ("", None, "-") else 0

def load_dis(path):
    """function (mangled) -> list of (file, line) per instruction, in order"""
    fun = None; cur = ("?", 0); out = collections.OrderedDict()
    for l in open(path, errors="replace"):
        m = re.match(r"\.text\.(\S+):", l)
        if m:
            fun = m.group(1); out[fun] = []; cur = ("?", 0); continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
        if fun and re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            out[fun].append(cur)
    return out

def demangle_key(mangled):
    m = re.match(r"_ZN4dcol11solve_entryILi(\d)ELi(\d)ELi(\d)E", mangled)
    if m: return "solve_entry<%s,%s,%s>" % m.groups()
    m = re.match(r"_ZN4dcol20proximity_persistentILi(\d)E", mangled)
    if m: return "proximity_persistent<%s>" % m.group(1)
    m = re.match(r"_ZN4dcol\d+(\w+?)E?PK", mangled)
    for name in ("build_frame", "grad_side"):
        if name in mangled: return name
    return mangled

def name_key(name):
    m = re.search(r"solve_entry<\(int\)(\d), \(int\)(\d), \(int\)(\d)>", name)
    if m: return "solve_entry<%s,%s,%s>" % m.groups()
    m = re.search(r"proximity_persistent<\(int\)(\d)>", name)
    if m: return "proximity_persistent<%s>" % m.group(1)
    for n in ("build_frame", "grad_side"):
        if n in name: return n
    return name

def main():
    rows = list(csv.reader(open(sys.argv[1])))
    dis = {demangle_key(k): v for k, v in load_dis(sys.argv[2]).items()}
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    by_line = collections.defaultdict(lambda: collections.Counter())
    by_fun = collections.defaultdict(lambda: collections.Counter())
    tot = collections.Counter(); miss = 0
    for n, hi in enumerate(heads):
        name = name_key(rows[hi - 1][1])
        end = heads[n + 1] - 1 if n + 1 < len(heads) else len(rows)
        ix = {h: i for i, h in enumerate(rows[hi])}
        data = [r for r in rows[hi + 1:end] if len(r) > ix["# Samples"]]
        lines = dis.get(name)
        if lines is None or len(lines) != len(data):
            miss += 1; lines = [("?" + name, 0)] * len(data)
        stall_cols = [c for c in ix if c.startswith("stall_") and "Not Issued" not in c]
        for r, ln in zip(data, lines):
            s = num(r[ix["# Samples"]]); e = num(r[ix["Instructions Executed"]])
            c = by_line[ln]; c["samples"] += s; c["exec"] += e
            f = by_fun[name]; f["samples"] += s; f["exec"] += e
            tot["samples"] += s; tot["exec"] += e
            for k in stall_cols:
                v = num(r[ix[k]])
                if v: c[k] += v; tot[k] += v
    print(f"# {tot['samples']} warp-stall samples, {tot['exec']} executed warp instructions; {miss} functions without a line map")
    print("# stall reasons overall: " + " ".join(f"{k[6:]}={100*v/tot['samples']:.1f}%" for k, v in tot.most_common() if k.startswith("stall_"))[:400])
    print("# by function")
    for k, c in sorted(by_fun.items(), key=lambda kv: -kv[1]["samples"])[:12]:
        print(f"  {100*c['samples']/tot['samples']:5.1f}% samples {100*c['exec']/tot['exec']:5.1f}% executed  {k}")
    print("# by source line (samples%, executed%, cycles-per-instruction relative to the average, top stall reasons)")
    avg = tot["samples"] / max(tot["exec"], 1)
    for ln, c in sorted(by_line.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = sorted(((k[6:], v) for k, v in c.items() if k.startswith("stall_")), key=lambda kv: -kv[1])[:3]
        rel = (c["samples"] / max(c["exec"], 1)) / avg
        print(f"  {100*c['samples']/tot['samples']:5.2f}% {100*c['exec']/tot['exec']:5.2f}%  x{rel:4.1f}  {ln[0]}:{ln[1]:<5d} " + " ".join(f"{k}:{100*v/max(c['samples'],1):.0f}" for k, v in st))

main()
