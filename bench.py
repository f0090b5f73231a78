"""bench.py -- headline benchmark of the B200 scoring path (see the contract in DESIGN.md).

A "step" is one complete exact-mode wGCL scoring run (distance tiles -> alpha grid with the
fixed point, local 1-AUC score and global JS score).  Metric (BASELINE.json): pair-alphas per
second = n(n+1)/2 * (#alpha values evaluated) / time, over all GPUs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N = 1 : BASELINE.json configs[1], the reference's 10k example with --force-exact --seed 42
        (tests/golden/example10k.npz, produced from the reference's example files).
N > 1 : weak scaling -- a synthetic planted-partition graph with round(10000*sqrt(N)) vertices
        (d = 32, 64 communities), i.e. the same number of pairs per GPU, its pair-matrix tiles
        sharded over the ranks with one NCCL all-reduce of the n-length degree sums per pass.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
METRIC = "exact global+local score throughput (node-pairs x alphas / s)"
UNIT = "pair-alphas/s"
EMPTY = (np.zeros(0), np.zeros(0, dtype=np.int64), np.zeros((0, 0), dtype=np.int64), np.zeros(0),
         np.zeros((0, 0)))


def workload(n_gpus):
    from cge_jl_b200.divergence import draw_samples

    if n_gpus == 1:
        z = np.load(os.path.join(ROOT, "tests", "golden", "example10k.npz"))
        edges, ew, vw, comm, emb = (z[k] for k in ("edges", "eweights", "vweights", "comm",
                                                   "embedding"))
        name = "CGE.jl example/10k (n=10000, m=41536, d=32, k=64) --force-exact --seed 42"
        data = "reference example 10k graph (fixture tests/golden/example10k.npz)"
    else:
        from cge_jl_b200.synth import planted_partition

        n = int(round(10000 * np.sqrt(n_gpus)))
        edges, ew, vw, comm, emb = planted_partition(n, k=64, d=32, seed=1000 + n_gpus)
        name = f"synthetic planted partition n={n} d=32 k=64, exact, --seed 42"
        data = "synthetic"
    n = vw.shape[0]
    samples = draw_samples(edges, ew, n, 10000, 42, directed=False, exact=True)
    return dict(edges=edges, ew=ew, vw=vw, comm=comm, emb=emb, n=n, samples=samples, name=name,
                data=data)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                    "-i", str(self.index)], capture_output=True, text=True, timeout=5)
                if o.returncode == 0 and o.stdout.strip():
                    self.rows.append([c.strip() for c in o.stdout.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({nm for r in self.rows for nm, v in zip(names, r[3:7]) if v == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def oracle_sample(w, max_alphas):
    """The CPU port (oracle/) on a bounded prefix of the alpha grid of the same workload.

    The first alphas need more fixed-point passes than the average one (50 and 28 against a mean
    of 25.65 for the 10k example), so pairs*alphas/time of the prefix understates the full-run
    throughput.  Where the full pass profile is known (the frozen oracle run of the 10k example,
    tests/golden/oracle_example10k_exact.npz) the sample time is extrapolated linearly in the
    number of O(n^2) sweeps (per alpha: kernel, passes, P, B; plus the distance build), as SURVEY.md
    section 8(d) prescribes, and the value is the full-run equivalent.
    """
    import oracle

    t0 = time.perf_counter()
    _, tr = oracle.wgcl(w["edges"], w["ew"], w["comm"], w["emb"], np.zeros(w["n"]), w["vw"],
                        samples=w["samples"], max_alphas=max_alphas)
    dt = time.perf_counter() - t0
    pairs = w["n"] * (w["n"] + 1) // 2
    a_run, passes = int(tr.n_alpha_run), int(sum(tr.iters))
    value, note = pairs * a_run / dt, "not extrapolated"
    gold = os.path.join(ROOT, "tests", "golden", "oracle_example10k_exact.npz")
    if w["n"] == 10000 and w["data"].startswith("reference example") and os.path.exists(gold):
        g = np.load(gold)
        it = g["iters"].astype(int)
        if list(it[:a_run]) == list(tr.iters)[:a_run]:
            sweeps_full = int(it.sum()) + 3 * int((it > 0).sum()) + 1
            sweeps_sample = passes + 3 * a_run + 1
            t_full = dt * sweeps_full / sweeps_sample
            value = pairs * int((it > 0).sum()) / t_full
            note = (f"extrapolated to the full run linearly in O(n^2) sweeps "
                    f"({sweeps_sample} of {sweeps_full}; full run measured once: {float(g['seconds']):.0f} s)")
    return value, dt, a_run, passes, note


def oracle_parallel_sample(w, max_alphas):
    """the same algorithm on ALL host cores (oracle/cge_oracle_mt.c) on the
    same bounded prefix -- what a parallel CPU implementation would do on this box.  The reference
    has no threading, so this is reported beside the faithful single-thread port, not instead."""
    import oracle

    t0 = time.perf_counter()
    _, tr = oracle.wgcl_mt(w["edges"], w["ew"], w["comm"], w["emb"], w["vw"], samples=w["samples"],
                           max_alphas=max_alphas)
    dt = time.perf_counter() - t0
    pairs = w["n"] * (w["n"] + 1) // 2
    a_run, passes = int(tr.n_alpha_run), int(sum(tr.iters))
    value, note = pairs * a_run / dt, "not extrapolated"
    gold = os.path.join(ROOT, "tests", "golden", "oracle_example10k_exact.npz")
    if w["n"] == 10000 and w["data"].startswith("reference example") and os.path.exists(gold):
        it = np.load(gold)["iters"].astype(int)
        if list(it[:a_run]) == list(tr.iters)[:a_run]:
            # this port sweeps the pair array (passes + kernel + B) times per alpha, plus the build
            sweeps_full = int(it.sum()) + 2 * int((it > 0).sum()) + 1
            sweeps_sample = passes + 2 * a_run + 1
            value = pairs * int((it > 0).sum()) / (dt * sweeps_full / sweeps_sample)
            note = f"extrapolated linearly in O(n^2) sweeps ({sweeps_sample} of {sweeps_full})"
    return {"value": value, "unit": UNIT, "cores": int(tr.threads), "kind": "port, multi-threaded",
            "sample": f"oracle/cge_oracle_mt.c (same algorithm, rows dealt to {int(tr.threads)} "
                      f"threads), first {a_run} of 40 alpha values ({passes} passes), {dt:.1f} s; {note}"}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (the line-by-line C port in oracle/; the
    Julia original cannot run in this image) on the host cores.  The reference is single-threaded
    (no @threads / Distributed anywhere in src/), so cores = 1."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(1 if args.gpus == 1 else args.gpus)
    if args.gpus > 1:
        args.ref_alphas = 1  # the weak-scaling graphs are N times larger: keep a step under a minute
    vals, times = [], []
    t_start = time.perf_counter()
    budget_s = 240.0  # the whole reference arm must end within a few minutes
    done_warm = 0
    while done_warm < args.warmup:
        t0 = time.perf_counter()
        v, dt, a_run, sweeps, note = oracle_sample(w, args.ref_alphas)
        done_warm += 1
        if (time.perf_counter() - t_start) + (args.steps + args.warmup - done_warm) * dt > budget_s:
            vals.append(v)  # too slow to afford untimed runs: this one counts as the first step
            times.append(dt)
            break
    while len(vals) < args.steps:
        if vals and (time.perf_counter() - t_start) + times[-1] > budget_s:
            break
        v, dt, a_run, sweeps, note = oracle_sample(w, args.ref_alphas)
        vals.append(v)
        times.append(dt)
    args.steps = len(vals)
    val = float(np.mean(vals))
    sample = (f"first {args.ref_alphas} of 40 alpha values ({sweeps} fixed-point passes) of the "
              f"same workload per step; {note}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": w["data"], "config": {"workload": w["name"], "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "host_cores": os.cpu_count(),
                         "kind": "port", "sample": sample,
                         "parallel_port": oracle_parallel_sample(w, args.ref_alphas)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_b200(args):
    import torch
    import torch.distributed as dist

    from cge_jl_b200 import _lib
    from cge_jl_b200 import divergence as dv

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the scoring path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
        args.gpus = world
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    w = workload(world)
    n = w["n"]
    pairs = n * (n + 1) // 2
    sc = dv.Scorer(local_rank)
    if world > 1:
        ids = [dv.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        sc.comm_init(ids[0], rank, world)
        if not args.no_p2p:  # per-pass exchange over NVLink peer memory inside the kernel
            handles = [None] * world
            dist.all_gather_object(handles, sc.p2p_export(w["n"]))
            sc.p2p_import(handles)
    problem, keep = dv.make_problem(w["edges"], w["ew"], w["comm"], w["emb"], np.zeros(n), w["vw"],
                                    None, None, None, False, False, w["samples"], 0, args.driver,
                                    args.regime)
    h2d = sum(a.nbytes for a in keep)
    d2h = 7 * 8

    def timed(fn, count):
        """`count` calls of fn bracketed by barrier+sync; returns max-over-ranks seconds."""
        barrier()
        t0 = time.perf_counter()
        for _ in range(count):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        barrier()
        return dt

    last = {}

    def step_resident():
        last["out"], last["stats"] = sc.run()

    def step_e2e():
        sc.upload(problem, keep)
        last["out"], last["stats"] = sc.run()

    # device-resident: inputs already in HBM
    sc.upload(problem, keep)
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev_ms, sweep_ms, sweeps, launches = [], 0.0, 0, 0
    dt_res = 0.0
    for _ in range(args.steps):  # one timed() per step so per-step device times can be collected
        dt_res += timed(step_resident, 1)
        st = last["stats"]
        ev_ms.append(st.ms_build + st.ms_solve)
        sweep_ms += st.ms_sweeps
        sweeps += st.fp_sweeps
        launches += st.launches
    stats = last["stats"]
    out = last["out"]
    a_run = int(stats.n_alpha_run)
    # end to end through the public call: host buffers, H2D + D2H inside the timed region
    if args.no_e2e:
        dt_e2e = float("nan")
    else:
        for _ in range(max(1, args.warmup // 3)):
            step_e2e()
        dt_e2e = timed(step_e2e, args.steps)
    sampler.stop_flag = True
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    sampler.join(timeout=2)
    value = pairs * a_run * args.steps / dt_res
    e2e = pairs * a_run * args.steps / dt_e2e
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # dominant kernel: the fixed point.  Persistent drivers: one launch per alpha runs all its
    # passes; host loop: one launch per pass.  Algorithmic bytes = 8 B per unordered pair and pass.
    persistent = int(stats.driver) in (2, 3)
    fp_launches = (a_run if persistent else int(stats.fp_sweeps)) * args.steps
    passes_per_launch = sweeps / max(fp_launches, 1)
    bytes_per_launch = 8.0 * pairs / world * passes_per_launch
    avg_launch_s = 1e-3 * sweep_ms / max(fp_launches, 1)
    achieved = bytes_per_launch / avg_launch_s / 1e9 if avg_launch_s > 0 else 0.0
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum from the committed ncu capture
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if world == 1 and tr["workload_pairs"] == pairs and int(stats.regime) == 1:
            traffic = (tr["dram_bytes_per_pass"] * passes_per_launch if persistent
                       else tr["hostloop_k_sweep_bytes_per_launch"])
    except Exception:
        pass
    kname = {1: "k_sweep<M,false> (one fixed-point pass per launch)",
             2: "k_fixed_point<M,false> (all passes of one alpha per cooperative launch)",
             3: "k_fixed_point_ring<M,false> (all passes of one alpha, cp.async.bulk ring)"}
    if int(stats.regime) in (2, 3):
        kname = {1: "k_sweep_rc<false>", 2: "k_fixed_point_rc<false>", 3: "k_fixed_point_rc<false>"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt_res / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": w["data"],
        "config": {"workload": w["name"], "alphas_evaluated": a_run,
                   "fixed_point_passes": int(stats.fp_sweeps), "b_passes": int(stats.b_sweeps),
                   "pairs": pairs, "samples_local": 10000, "tiles": int(stats.n_tiles),
                   "l2": "inputs larger than L2 (q matrix %.0f MB per GPU vs 126 MB L2)"
                         % (stats.matrix_bytes / 1e6),
                   "driver": {1: "hostloop", 2: "persistent", 3: "ring"}.get(int(stats.driver), "?"),
                   "regime": {1: "stored", 2: "recompute", 3: "recompute, row-norm/dot form"}.get(int(stats.regime), "?"),
                   "result": [float(x) for x in out],
                   "device_ms_per_step": float(np.mean(ev_ms)),
                   "ms_breakdown_last_step": {
                       "upload": float(stats.ms_upload), "build": float(stats.ms_build),
                       "solve": float(stats.ms_solve), "fp_kernels": float(stats.ms_sweeps),
                       "b_kernels": float(stats.ms_bsweeps), "total": float(stats.ms_total)}},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * dt_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic,
                     "kernel": kname.get(int(stats.driver), "?"),
                     "bytes_per_launch": bytes_per_launch, "avg_launch_us": 1e6 * avg_launch_s,
                     "passes_per_launch": passes_per_launch,
                     "avg_pass_us": 1e6 * avg_launch_s / max(passes_per_launch, 1e-9),
                     "launches_timed": int(fp_launches),
                     "note": "achieved = 8 B x unordered pairs x passes in the launch / CUDA-event "
                             "duration of the launch, events recorded by the library on its stream",
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)"
                                    if "hbm_gbs" in peaks else "fallback 6650 GB/s"},
    }
    if int(stats.regime) in (2, 3):
        # recompute regime: FP64-pipe roofline, algorithmic work (2d + 6) flop per pair and pass
        # (SURVEY.md 8(d)); peak = FP64 FMA throughput measured on this device by the library
        d_emb = w["emb"].shape[1]
        flops = (2.0 * d_emb + 6.0) * pairs / world * passes_per_launch
        fp64_peak = sc.fp64_peak_tflops()
        ach = flops / avg_launch_s / 1e12 if avg_launch_s > 0 else 0.0
        line["roofline"].update({
            "bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": ach / fp64_peak if fp64_peak else None, "traffic": None,
            "flops_per_launch": flops,
            "note": "achieved = (2d+6) flop x unordered pairs x passes in the launch / CUDA-event "
                    "duration; sqrt, divide and the power are counted as 0 flop, so this is a lower "
                    "bound on executed work",
            "peak_source": "cge_b200_measure_fp64_peak (DFMA microbenchmark on this device)"})
    if world == 1 and not args.no_cpu_baseline:
        v, dt, a, sw, note = oracle_sample(w, args.ref_alphas)
        line["cpu_baseline"] = {
            "value": v, "unit": UNIT, "cores": 1, "host_cores": os.cpu_count(), "kind": "port",
            "threads_note": "the reference has no threading (no @threads/@spawn/Distributed in "
                            "src/): JULIA_NUM_THREADS does not change it, so one core is used",
            "sample": f"oracle/ C port of wGCL, first {a} of 40 alpha values ({sw} fixed-point "
                      f"passes) of the same workload, {dt:.1f} s; {note}",
            "parallel_port": oracle_parallel_sample(w, args.ref_alphas)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--driver", type=int, default=0, help="0 auto, 1 host loop, 2 persistent")
    ap.add_argument("--regime", type=int, default=0, help="0 auto, 1 stored, 2 recompute, 3 recompute with the row-norm/dot form")
    ap.add_argument("--ref-alphas", type=int, default=2,
                    help="alpha values per CPU sample (bounds the CPU baseline's run time)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only")
    ap.add_argument("--no-p2p", action="store_true",
                    help="multi-GPU: NCCL all-reduce per pass from the host instead of the "
                         "in-kernel NVLink exchange")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
