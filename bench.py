#!/usr/bin/env python
"""bench.py — throughput of the DCOL proximity hot path (BASELINE.json metric: pair-queries/sec).

One "step" = one pass of the hot path over one batch of synthetic random-pose pairs:
  headline  : BASELINE config 2 — proximity, all 7x7 primitive-type pairs, quaternion, 1M pairs per GPU
  secondary : BASELINE config 3 — proximity_jacobian, mixed primitives, MRP, 2M pairs per GPU
              (16M sharded over 8 GPUs = 2M per GPU; reported under "jacobian" in the same JSON line)
Pairs are independent, so N GPUs run N index shards with no data-path collective ("scaling": "weak");
torch.distributed (NCCL) is used only for the barrier and the max-over-ranks of the device time.

  value : device-resident whole-call throughput (inputs already in HBM, results left in HBM),
          CUDA events on the stream the library launches on, L2 flushed between timed steps
  e2e   : same metric through the C ABI with HOST buffers (pinned), H2D + D2H inside the timed region
  roofline     : FP64 CUDA-core roofline of the solver kernels (algorithmic FLOPs from the frozen
                 flops_model with each pair's actual iteration count / kernel time from CUDA events
                 inside the library); peak = DFMA micro-benchmark (pipe peak, two-source form) run in the same process, because
                 MEASURED_PEAKS.json carries no FP64 figure.  "hbm" sub-object: pose-in/result-out bytes.
  cpu_baseline : the C++ oracle (restatement of the Julia reference) on all host threads, bounded sample; rebuilt on the box
                 it runs on with -O3 -march=native, flags and thread count in `sample`
  e2e.copy_bound : the same end-to-end loop with the kernels switched off (dcol_set_copy_only): what the host can feed
  scene        : the same proximity workload on the reference's contact-example shapes (8-face fixture polytope,
                 Capsule(0.2, 5.0)): kernel time, roofline fraction, mean iterations
  large_batch  : the same workload at 8 M pairs per GPU in one call (fixed part of a step amortised): kernel ms per 1 M pairs
  latency      : wall time of single-pair and config-1 (10k cube-cube) calls through the Python wrapper

`--impl reference` times the reference's CPU implementation of the path; Julia is not in the image, so
this is the oracle port (kind "port") on all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "proximity pair-queries/sec"
UNIT = "pairs/s"
PDIP_TOL = 1e-6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1_000_000, help="proximity pairs per GPU per step (config 2)")
    ap.add_argument("--jac-pairs", type=int, default=2_000_000, help="proximity_jacobian pairs per GPU per step (config 3)")
    ap.add_argument("--no-jacobian", action="store_true")
    ap.add_argument("--no-large", action="store_true", help="skip the 8x larger batch (large_batch object)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", default="cube", choices=["cube", "scene"])
    ap.add_argument("--profile", action="store_true",
                    help="for runs under ncu: the same device-resident timed region, without the end-to-end, scene, latency and CPU legs")
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {"workload": "BASELINE config 2: proximity, all 7x7 primitive-type pairs (polytope/capsule/cylinder/cone/sphere/"
                        "ellipsoid/polygon), quaternion attitude, random poses r1=2N(0,I) r2=N(0,I) q=normalize(N(0,I4))",
            "pairs_per_gpu": args.pairs, "global_pairs": args.pairs * n_gpus, "pdip_tol": PDIP_TOL,
            "shapes": args.variant, "parallelism": f"index-sharded x{n_gpus}, no data-path collective",
            "l2": "flushed between timed steps (256 MiB memset); inputs+outputs 160 MB per step"}


class ClockSampler:
    """nvidia-smi sampled every 20 ms from before the warm-up; only samples whose timestamp falls inside the timed
    region [t0, t1] are reported (the recipe's clocks line, /opt/skills/guides/B200_PROFILING.md)."""
    FIELDS = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self, t0, t1):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, power, reasons, sm_all = [], [], [], set(), []
        for r in rows:
            if len(r) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                c, m = float(r[2]), float(r[3])
            except ValueError:
                continue
            sm_all.append(c)
            if ts < t0 - 0.02 or ts > t1 + 0.02:
                continue
            sm.append(c); mx.append(m)
            try:
                power.append(float(r[4]))
            except ValueError:
                pass
            for k, nm in enumerate(self.NAMES):
                if r[6 + k].strip().lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power) if power else None)
        elif sm_all:
            out.update(sm_mhz=float(np.median(sm_all)), samples=0, note="no sample fell inside the timed region; median over the whole run")
        return out


ORACLE_FLAGS = "-O3 -march=native"


def native_oracle():
    """The oracle rebuilt for THIS host (g++ -O3 -march=native) into a temporary directory; falls
    back to the shipped -march=x86-64-v3 build if the compiler is not there.  Returns (module, flags description)."""
    from oracle import oracle as orc
    try:
        import ctypes
        out = os.path.join(tempfile.mkdtemp(prefix="dcol_oracle_"), "libdcol_oracle_native.so")
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "-B", "ARCHFLAGS=-march=native", f"OUT={out}"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        orc._LIB = ctypes.CDLL(out)
        orc._LIB.dcol_oracle_query.restype = ctypes.c_int
        return orc, "g++ " + ORACLE_FLAGS + " (built on this host)"
    except Exception:
        return orc, "g++ -O3 -march=x86-64-v3 (shipped build; native rebuild failed)"


def reference_arm(args, rank, world):
    """The reference's CPU path: C++ oracle port on all host threads, bounded sample per step."""
    if rank != 0:
        return
    import dcol_b200 as dc
    import util
    orc, flags = native_oracle()
    threads = orc.max_threads()
    sample = 100_000
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(sample, dc.ATT_QUAT, seed=1234, variant=args.variant)
    ta = t.arrays()
    for _ in range(max(1, args.warmup)):
        orc.query(0, ta, s1[:20000], s2[:20000], r1[:20000], a1[:20000], r2[:20000], a2[:20000], 0, PDIP_TOL)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.query(0, ta, s1, s2, r1, a1, r2, a2, 0, PDIP_TOL)
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    cfg = workload_config(args, args.gpus)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{sample} pairs of the config-2 workload per step, C++ restatement of the Julia reference "
                                       f"(Julia is not installed in this image), {flags}, std::thread x{threads} (dynamic chunks of 64 pairs)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import dcol_b200 as dc
    import util

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; dcol_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = dc.bind_to_gpu_numa_node(local)          # pinned buffers local to the GPU's PCIe root (end-to-end path at N > 1)
    if not numa.get("bound") and world > 1:
        numa = dict(numa, fallback=dc.bind_disjoint_cpus(local, int(os.environ.get("LOCAL_WORLD_SIZE", world))))
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = dc.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def make(n, att_kind, seed, variant):
        t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(n, att_kind, seed=seed + 7919 * rank, variant=variant)
        host = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (s1, s2, r1, a1, r2, a2)]
        devt = [h.to(dev) for h in host]
        return t, host, devt

    def run_config(kind, att_kind, n, steps, warmup, seed, variant=None, e2e=True):
        variant = variant or args.variant
        na = 4 if att_kind == 0 else 3
        nt = 6 + 2 * na
        table, host, devt = make(n, att_kind, seed, variant)
        ctx.set_shape_arrays(*table.arrays())
        d_alpha = torch.empty(n, dtype=torch.float64, device=dev)
        d_x = torch.empty((n, 3), dtype=torch.float64, device=dev)
        d_der = torch.empty((n, nt, 4), dtype=torch.float64, device=dev) if kind == 2 else None
        d_st = torch.empty(n, dtype=torch.int32, device=dev)
        d_it = torch.empty(n, dtype=torch.int32, device=dev)
        h_alpha = torch.empty(n, dtype=torch.float64).pin_memory()
        h_x = torch.empty((n, 3), dtype=torch.float64).pin_memory()
        h_der = torch.empty((n, nt, 4), dtype=torch.float64).pin_memory() if kind == 2 else None
        h_st = torch.empty(n, dtype=torch.int32).pin_memory()
        h_it = torch.empty(n, dtype=torch.int32).pin_memory()

        def dev_step():
            ctx.query_raw(kind, n, *devt, att_kind, PDIP_TOL, dc.MEM_DEVICE, d_alpha, d_x, d_der, d_st, d_it)

        def host_step():
            ctx.query_raw(kind, n, *host, att_kind, PDIP_TOL, dc.MEM_HOST, h_alpha, h_x, h_der, h_st, h_it)

        # three sets of pinned output buffers: the end-to-end loop keeps three submissions in flight (dcol_submit/dcol_wait)
        DEPTH = 3
        outs = [(h_alpha, h_x, h_der, h_st, h_it)]
        for _ in range(DEPTH - 1):
            outs.append(tuple(torch.empty_like(tn).pin_memory() if tn is not None else None for tn in (h_alpha, h_x, h_der, h_st, h_it)))

        def host_submit(k):
            o = outs[k % DEPTH]
            return ctx.submit_raw(kind, n, *host, att_kind, PDIP_TOL, o[0], o[1] if kind != 1 else None, o[2], o[3], o[4])

        sampler = ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
        for _ in range(warmup):
            dev_step()
        barrier()
        t_begin = time.time()
        l0 = ctx.launch_count()
        evs = []
        kms = []
        for _ in range(steps):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            dev_step()
            e1.record(stream)
            evs.append((e0, e1))
            kms.append(ctx.last_kernel_ms())
        barrier()
        t_end = time.time()
        launches = ctx.launch_count() - l0
        clocks = sampler.stop(t_begin, t_end) if sampler else None
        dev_ms = sum(a.elapsed_time(b) for a, b in evs)
        dev_ms = max_over_ranks(dev_ms)
        e2e_s = e2e_sync_s = copy_s = None
        e2e_steps = max(4, min(12, steps // 2))
        if e2e:
            # end to end: host (pinned) buffers through the C ABI, H2D and D2H inside the timed region
            host_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                host_step()
            torch.cuda.synchronize()
            e2e_sync_s = max_over_ranks(time.perf_counter() - t0)
            barrier()

            # pipelined: every step still copies its inputs host->device and its results device->host inside the timed
            # region; submission k+1 is issued before submission k is waited for
            def pipelined():
                for k in range(DEPTH):             # warm the submission slots (their staging buffers are allocated on first use)
                    ctx.wait(host_submit(k))
                barrier()
                t0 = time.perf_counter()
                tickets = []
                for k in range(e2e_steps):
                    if len(tickets) >= DEPTH:
                        ctx.wait(tickets[-DEPTH])
                    tickets.append(host_submit(k))
                for tk in tickets[-DEPTH:]:
                    ctx.wait(tk)
                torch.cuda.synchronize()
                dt = max_over_ranks(time.perf_counter() - t0)
                barrier()
                return dt
            e2e_s = pipelined()
            # the same loop with the kernels switched off: the copy-only bound of this host / PCIe path at this N
            ctx.set_copy_only(True)
            try:
                copy_s = pipelined()
            finally:
                ctx.set_copy_only(False)
        # the default caller path: blocking calls on PAGEABLE arrays (plain numpy / Julia Arrays, outputs reused) — the library's
        # own staging pipeline (copy threads + pinned staging, dcol_capi.cu HostCopier) instead of page-locked caller buffers
        pageable_s = None
        if e2e and world == 1:
            np_in = [h.numpy().copy() for h in host]
            np_out = [np.empty_like(tn.numpy()) if tn is not None else None for tn in (h_alpha, h_x, h_der, h_st, h_it)]

            def pageable_step():
                ctx.query_raw(kind, n, *np_in, att_kind, PDIP_TOL, dc.MEM_HOST, *np_out)
            pageable_step(); pageable_step()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                pageable_step()
            pageable_s = time.perf_counter() - t0
            del np_in, np_out
        # optional result gather over NVLink: all-gather of the alpha rows, outside the timed step
        gather = None
        if world > 1:
            parts = [torch.empty_like(d_alpha) for _ in range(world)]
            dist.all_gather(parts, d_alpha)
            barrier()
            g0 = torch.cuda.Event(enable_timing=True); g1 = torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(5):
                dist.all_gather(parts, d_alpha)
            g1.record(); torch.cuda.synchronize()
            gather = {"op": "nccl all_gather of alpha", "bytes_per_rank": int(d_alpha.numel() * 8), "ms": max_over_ranks(g0.elapsed_time(g1) / 5)}
        iters = d_it.cpu().numpy()
        status = d_st.cpu().numpy()
        types, offs, params = table.arrays()
        n_ort = np.array([int(params[o]) if t in (0, 6) else {1: 2, 2: 4, 3: 1, 4: 0, 5: 0}[int(t)] for t, o in zip(types, offs)])
        flops = dc.flops_model.batch_flops(kind, types, n_ort, host[0].numpy(), host[1].numpy(), iters, att_kind)
        in_b = n * (8 + 8 * (6 + 2 * na))
        out_b = n * (8 + 8 * (1 + 3 + (4 * nt if kind == 2 else 0)))
        return {"n": n, "dev_ms": dev_ms, "steps": steps, "kernel_ms": float(np.mean(kms)), "launches": launches, "clocks": clocks,
                "e2e_s": e2e_s, "e2e_sync_s": e2e_sync_s, "copy_s": copy_s, "pageable_s": pageable_s, "e2e_steps": e2e_steps, "variant": variant, "flops": flops, "in_bytes": in_b, "out_bytes": out_b,
                "mean_iters": float(iters.mean()), "max_iters": int(iters.max()), "failed": int((status != 0).sum()), "gather": gather}

    # FP64 roofline denominator: 0.6 s of back-to-back DFMA-only launches (median launch), SM clock sampled meanwhile
    peak_sampler = ClockSampler(local) if rank == 0 else None
    if peak_sampler:
        peak_sampler.start()
        time.sleep(0.1)
    tp0 = time.time()
    fp64_peak = ctx.measure_fp64_tflops(0.02 if args.profile else 0.6)
    peak_clocks = peak_sampler.stop(tp0, time.time()) if peak_sampler else None
    fp64_peak3 = ctx.measure_fp64_tflops(0.02 if args.profile else 0.2, three_source=True)
    prox = run_config(0, dc.ATT_QUAT, args.pairs, args.steps, args.warmup, 1234, e2e=not args.profile)
    other = "scene" if args.variant == "cube" else "cube"
    prox2 = None if args.profile else run_config(0, dc.ATT_QUAT, args.pairs, max(3, args.steps // 4), args.warmup, 1234, variant=other, e2e=False)
    # the same workload at 8 M pairs per GPU: what the solver does once the fixed part of a step (first tiles, class-pair
    # transitions of the tile queue, the tail of the last tiles and the resume launches: ~0.14 ms) is amortised
    big = None if (args.profile or args.no_large) else run_config(0, dc.ATT_QUAT, 8 * args.pairs, 3, 3, 1234, e2e=False)
    jac = None
    if not args.no_jacobian:
        jac = run_config(2, dc.ATT_MRP, args.jac_pairs, max(2, args.steps // 2), args.warmup, 4321, e2e=not args.profile)

    # latency of the small calls every reference example makes in its inner loop (src/proximity.jl:3-14): wall clock
    # through the Python wrapper over the C ABI, pageable numpy arrays and device tensors
    latency = None
    if rank == 0 and world == 1 and not args.profile:
        def wall(f, reps):
            for _ in range(10):
                f()
            t0 = time.perf_counter()
            for _ in range(reps):
                f()
            return 1e6 * (time.perf_counter() - t0) / reps
        latency = {"unit": "us per call (wall clock, Python wrapper over the C ABI)"}
        cone = dc.Cone(2.0, np.deg2rad(22)); caps = dc.Capsule(0.3, 1.2)
        cone.r = np.array([0.3, 0.2, 1.0]); caps.r = np.array([-1.0, 0.4, 0.1])
        lctx = dc.Context(local)
        latency["single_pair_proximity"] = wall(lambda: dc.proximity(cone, caps, ctx=lctx), 300)
        latency["single_pair_jacobian"] = wall(lambda: dc.proximity_jacobian(cone, caps, ctx=lctx), 300)
        for nn in (1000, 10000):
            tb, c1, c2, q1, b1, q2, b2 = util.cube_batch(nn, 0, 5)
            lctx.set_shape_arrays(*tb.arrays())
            latency[f"config1_cube_cube_{nn}_host"] = wall(lambda: lctx.proximity_batch(c1, c2, q1, b1, q2, b2, att_kind=0, pdip_tol=PDIP_TOL), 100)
            latency[f"config1_cube_cube_{nn}_kernels"] = 1e3 * lctx.last_kernel_ms()
            dv = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (c1, c2, q1, b1, q2, b2)]
            oa = torch.empty(nn, dtype=torch.float64, device=dev); ox = torch.empty((nn, 3), dtype=torch.float64, device=dev)
            os_ = torch.empty(nn, dtype=torch.int32, device=dev); oi = torch.empty(nn, dtype=torch.int32, device=dev)
            latency[f"config1_cube_cube_{nn}_device"] = wall(lambda: lctx.query_raw(0, nn, *dv, 0, PDIP_TOL, dc.MEM_DEVICE, oa, ox, None, os_, oi), 100)
        lctx.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.profile:
        orc, oflags = native_oracle()
        threads = orc.max_threads()
        sample = min(250_000 * threads, 8_000_000)     # ~10-30 s of CPU work in total across the threads, bounded
        t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(sample, dc.ATT_QUAT, seed=1234, variant=args.variant)
        orc.query(0, t.arrays(), s1[:20000], s2[:20000], r1[:20000], a1[:20000], r2[:20000], a2[:20000], 0, PDIP_TOL)
        t0 = time.perf_counter()
        orc.query(0, t.arrays(), s1, s2, r1, a1, r2, a2, 0, PDIP_TOL)
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        orc.query(0, t.arrays(), s1[:200000], s2[:200000], r1[:200000], a1[:200000], r2[:200000], a2[:200000], 0, PDIP_TOL, nthreads=1)
        dt1 = time.perf_counter() - t0
        cpu = {"value": sample / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{sample} pairs of the config-2 workload, C++ restatement of the Julia reference, {oflags}, std::thread x{threads}"}
        cpu["single_thread_us_per_pair"] = 1e6 * dt1 / 200000
        sj = min(100_000 * threads, 3_000_000)
        t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(sj, dc.ATT_MRP, seed=4321, variant=args.variant)
        t0 = time.perf_counter()
        orc.query(2, t.arrays(), s1, s2, r1, a1, r2, a2, 1, PDIP_TOL)
        cpu["jacobian_value"] = sj / (time.perf_counter() - t0)
        cpu["jacobian_sample"] = f"{sj} pairs of the config-3 workload (mixed, MRP)"

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback"

        traffic = {}
        try:   # DRAM bytes per launch of the solver kernel from one `ncu --set full` capture of this workload (profiles/)
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except Exception:
            pass

        def roof(r, key):
            ach = r["flops"] / (r["kernel_ms"] * 1e-3) / 1e12
            gbs = (r["in_bytes"] + r["out_bytes"]) / (r["kernel_ms"] * 1e-3) / 1e9
            return {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                    "peak_source": "DFMA-only kernel x = fma(x, a, x) launched back to back for 0.6 s in this process, median launch: the FP64 "
                                   "pipe ~100 % busy under ncu (MEASURED_PEAKS.json has no FP64 figure; nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz "
                                   "= 37.2 TFLOP/s)",
                    "peak_clocks": peak_clocks, "frac_of_nominal_37.2": ach / 37.2,
                    "peak_three_source_dfma": fp64_peak3, "frac_of_three_source_peak": ach / fp64_peak3,
                    "peak_three_source_note": "x = fma(x, a, b) with three distinct 64-bit register sources keeps the pipe 92 % busy whatever the "
                                              "unrolling or CTA shape (profiles/r02z_k_dfma_ncu.txt); earlier bench lines of this repo used this figure",
                    "traffic": traffic.get(key, {}).get("dram_bytes_per_step") if traffic.get(key, {}).get("pairs") == r["n"] else None,
                    "traffic_source": traffic.get(key, {}).get("source"),
                    "kernel_ms_per_step": r["kernel_ms"], "algorithmic_flops_per_pair": r["flops"] / r["n"],
                    "hbm": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                            "peak_source": hbm_src, "algorithmic_bytes_per_pair": (r["in_bytes"] + r["out_bytes"]) / r["n"]}}

        max_over_ranks_done = lambda r: r["dev_ms"]          # dev_ms already is the max over ranks
        value = world * prox["n"] * prox["steps"] / (prox["dev_ms"] * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": prox["steps"], "warmup": args.warmup,
                "ms_per_step": prox["dev_ms"] / prox["steps"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
                "e2e": None if args.profile else {"value": world * prox["n"] * prox["e2e_steps"] / prox["e2e_s"], "unit": UNIT,
                        "h2d_bytes_per_step": prox["in_bytes"], "d2h_bytes_per_step": prox["out_bytes"],
                        "api": "dcol_submit/dcol_wait, three submissions in flight, pinned host buffers",
                        "one_call_at_a_time": world * prox["n"] * prox["e2e_steps"] / prox["e2e_sync_s"],
                        "copy_bound": world * prox["n"] * prox["e2e_steps"] / prox["copy_s"],
                        "frac_of_copy_bound": prox["copy_s"] / prox["e2e_s"],
                        "copy_bound_how": "same loop, buffers, streams and submission depth with the kernels switched off (dcol_set_copy_only)",
                        "pageable_arrays_one_call_at_a_time": (prox["n"] * prox["e2e_steps"] / prox["pageable_s"]) if prox["pageable_s"] else None,
                        "pageable_how": "blocking dcol_proximity on plain (pageable) numpy arrays, outputs reused: the library stages through its own pinned buffers with copy threads"},
                "gpu_launches": prox["launches"], "clocks": prox["clocks"], "roofline": roof(prox, "proximity"),
                "solver": {"mean_iters": prox["mean_iters"], "max_iters": prox["max_iters"], "failed_pairs": prox["failed"]}}
        r2 = roof(prox2, "proximity_" + prox2["variant"]) if prox2 else None
        if prox2:
            line[prox2["variant"]] = {"what": "same workload on the '%s' shape set" % prox2["variant"],
                                      "value": world * prox2["n"] * prox2["steps"] / (max_over_ranks_done(prox2) * 1e-3), "unit": UNIT,
                                      "ms_per_step": prox2["dev_ms"] / prox2["steps"], "kernel_ms_per_step": prox2["kernel_ms"],
                                      "roofline_frac": r2["frac"], "achieved_tflops": r2["achieved"],
                                      "algorithmic_flops_per_pair": r2["algorithmic_flops_per_pair"],
                                      "mean_iters": prox2["mean_iters"], "max_iters": prox2["max_iters"], "failed_pairs": prox2["failed"]}
        if big:
            rb = roof(big, "proximity_large")
            line["large_batch"] = {"what": "same workload, %d pairs per GPU in one call (the fixed part of a step amortised)" % big["n"],
                                   "value": world * big["n"] * big["steps"] / (big["dev_ms"] * 1e-3), "unit": UNIT,
                                   "ms_per_step": big["dev_ms"] / big["steps"], "kernel_ms_per_step": big["kernel_ms"],
                                   "kernel_ms_per_1M_pairs": big["kernel_ms"] / (big["n"] / 1e6),
                                   "roofline_frac": rb["frac"], "achieved_tflops": rb["achieved"], "mean_iters": big["mean_iters"],
                                   "failed_pairs": big["failed"]}
        if latency:
            line["latency"] = latency
        if prox.get("gather"):
            line["gather"] = prox["gather"]
        line["host_affinity"] = numa
        if cpu:
            line["cpu_baseline"] = cpu
        if jac:
            line["jacobian"] = {
                "metric": "proximity_jacobian queries/sec", "unit": UNIT,
                "value": world * jac["n"] * jac["steps"] / (jac["dev_ms"] * 1e-3), "steps": jac["steps"],
                "ms_per_step": jac["dev_ms"] / jac["steps"],
                "config": {"workload": "BASELINE config 3: proximity_jacobian, mixed 7x7 primitive types, MRP attitude",
                           "pairs_per_gpu": jac["n"], "global_pairs": jac["n"] * world},
                "e2e": None if args.profile else {"value": world * jac["n"] * jac["e2e_steps"] / jac["e2e_s"], "unit": UNIT,
                        "h2d_bytes_per_step": jac["in_bytes"], "d2h_bytes_per_step": jac["out_bytes"],
                        "api": "dcol_submit/dcol_wait, three submissions in flight, pinned host buffers",
                        "one_call_at_a_time": world * jac["n"] * jac["e2e_steps"] / jac["e2e_sync_s"],
                        "copy_bound": world * jac["n"] * jac["e2e_steps"] / jac["copy_s"],
                        "frac_of_copy_bound": jac["copy_s"] / jac["e2e_s"],
                        "pageable_arrays_one_call_at_a_time": (jac["n"] * jac["e2e_steps"] / jac["pageable_s"]) if jac["pageable_s"] else None},
                "gpu_launches": jac["launches"], "clocks": jac["clocks"], "roofline": roof(jac, "jacobian"),
                "solver": {"mean_iters": jac["mean_iters"], "max_iters": jac["max_iters"], "failed_pairs": jac["failed"]}}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
