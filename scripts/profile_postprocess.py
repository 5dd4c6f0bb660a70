"""Turn the outputs of scripts/r02_profile.sh (gpurun_out/r02p_*) into the summaries committed under profiles/:
per-launch ncu tables of one timed step, profiles/ncu_traffic.json (read by bench.py for roofline.traffic), the launch
list, and copies of the bench lines / latency / configs / warm-start logs.  Run in the build container."""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from ncu_summary import KEYS
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02p"

SCALE = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
traffic = {}
for f, name, pairs in (("prox", "proximity", 1000000), ("jac", "jacobian", 2000000)):
    rows = list(csv.reader(open(os.path.join(G, f"{TAG}_{f}_step_raw.csv"))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    labels = ["<=8-row tier, first pass", "<=8-row tier, resume pass", "<=16-row tier, first pass", "<=16-row tier, resume pass"]
    with open(os.path.join(P, f"{TAG}_{name}_persistent_step.txt"), "w") as o:
        o.write("# ncu --set full --clock-control none: the four solver launches of ONE timed step of `python bench.py --steps 2 --warmup 3 --profile`\n"
                f"# ({pairs} pairs, round-2 final kernels); cold-cache, serialised replays: compare shares, not absolutes\n")
        for r, lab in zip(data, labels):
            o.write(f"# kernel: {r[hdr.index('Kernel Name')]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}  -- {lab}\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k); o.write(f"  {k:95s} {r[i]:>20s} {units[i]}\n")
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    num = lambda v: float(v.replace(",", ""))
    L = [{"ms": num(r[it]) * (1e-3 if units[it] == "us" else 1.0), "dram_read": num(r[ir]) * SCALE[units[ir]], "dram_write": num(r[iw]) * SCALE[units[iw]]} for r in data]
    traffic[name] = {"pairs": pairs, "launches": L, "dram_bytes_per_step": sum(l["dram_read"] + l["dram_write"] for l in L),
                     "source": "ncu --set full --clock-control none on `python bench.py --steps 2 --warmup 3 --profile` (round-2 final kernels), sum over the 4 "
                               "proximity_persistent launches of one timed step (2 shared-memory tiers x {first pass, resume pass}); cold-cache, serialised "
                               "replays; the binning / pack / un-permute kernels around them add ~250 MB per 1M pairs (profiles/r02r_step_all_kernels.txt)"}
    print(name, f"{traffic[name]['dram_bytes_per_step'] / 1e6:.0f} MB per step", [round(l["ms"], 3) for l in L])
json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
with open(os.path.join(P, f"{TAG}_launches.txt"), "w") as o:
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), "launches", os.path.join(G, f"{TAG}_launches.csv")], stdout=o, check=False)
for src, dst in ((f"{TAG}_bench.json", f"{TAG}_bench.json"), (f"{TAG}_bench_ref.json", f"{TAG}_bench_reference.json"), (f"{TAG}_configs.txt", "r02_configs.txt"),
                 (f"{TAG}_warm.txt", "r02_warm_start.txt"), (f"{TAG}_hostpipe.txt", "r02_host_pipeline.txt")):
    shutil.copy(os.path.join(G, src), os.path.join(P, dst))
with open(os.path.join(P, "r02_latency.txt"), "w") as o:
    o.write("# C ABI wall-clock latency (scripts/latency_capi.cu), B200, round-2 final build\n" + open(os.path.join(G, f"{TAG}_latency_capi.log")).read()
            + "\n# Python wrapper: the `latency` object of profiles/{TAG}_bench.json\n")
d = json.loads(open(os.path.join(G, f"{TAG}_bench.json")).read().strip().splitlines()[-1])
r = d["roofline"]; j = d["jacobian"]
print(f"bench: {d['value']:.3e} pairs/s, {d['ms_per_step']:.3f} ms/step, solver {r['kernel_ms_per_step']:.3f} ms, {r['achieved']:.2f} of {r['peak']:.2f} TFLOP/s = {r['frac']:.3f} ({r['frac_of_nominal_37.2']:.3f} of nominal)")
print(f"jacobian: {j['value']:.3e}/s solver {j['roofline']['kernel_ms_per_step']:.3f} ms frac {j['roofline']['frac']:.3f}; scene {d['scene']['value']:.3e} frac {d['scene']['roofline_frac']:.3f}")
print("e2e", d["e2e"]["value"], d["e2e"]["frac_of_copy_bound"], "jac e2e", j["e2e"]["value"], j["e2e"]["frac_of_copy_bound"])
print("latency", d["latency"]); print("cpu", d["cpu_baseline"])
