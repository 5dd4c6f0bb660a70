"""Development aid: attribute SASS instruction counts of one solver specialisation to source functions.
usage: sass_attr.py <object.o> <mangled-substring e.g. ILi2ELi2ELb0>"""
import collections, os, re, subprocess, sys, tempfile
obj, pat = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
out = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cubin)], capture_output=True, text=True).stdout.split("\n")
src_dir = os.path.dirname(os.path.abspath(obj))
fn_at = {}
for f in ("dcol_solver.cuh", "dcol_deriv.cuh"):
    cur = "?"
    for i, l in enumerate(open(os.path.join(src_dir, f)), 1):
        m = re.search(r"__device__ (?:static )?(?:__forceinline__|__noinline__) [\w:<>,\s&\*]*?\b(\w+)\(", l)
        if m:
            cur = m.group(1)
        fn_at[(f, i)] = cur
start = end = None
for i, l in enumerate(out):
    if l.startswith(".text.") and pat in l:
        start = i
    elif start is not None and l.startswith("\t.section"):
        end = i
        break
cur = None
cnt = collections.Counter()
tot = 0
for l in out[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l):
        key = fn_at.get(cur, cur[0] if cur else "<none>")
        cnt[key] += 1
        tot += 1
print("total instrs", tot, f"({tot*16/1024:.0f} KB)")
for k, c in cnt.most_common(30):
    print(f"{c:6d} {100*c/tot:5.1f}%  {k}")
