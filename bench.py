"""bench.py -- headline benchmark of the B200 scoring path (contract: DESIGN.md section 7).

A "step" is one complete exact-mode wGCL scoring run (distance tiles -> alpha grid with the
fixed point, local 1-AUC score and global JS score) of the workload below.  Metric (BASELINE.json):
pair-alphas per second = n(n+1)/2 * (#alpha values evaluated) / time, over all GPUs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload 4|2]

Workload (every N): BASELINE.json configs[3] -- synthetic ABCD-style undirected graph, 200 000
vertices, 64 communities, d = 128, exact mode, --seed 42, the configuration the metric's
"1/2/4/8 B200" is quoted on.  Its 160 GB pair matrix fits one B200 (stored regime); at N > 1 the
tile sequence is sharded over the ranks (strong scaling, the same graph at every N) with the per-pass
exchange of the degree sums over NVLink peer memory inside the persistent kernel.
--workload 2 selects BASELINE.json configs[1] (the reference's 10k example, --force-exact), which is
also reported as a secondary record of the N = 1 line.

Every timed step uploads the problem from host buffers (H2D) and runs it: `e2e` is the whole step,
`value` the device-resident part (the run() call alone, inputs already in HBM), both as the sum
over the K steps of the max over ranks, each part bracketed by barrier + synchronize.
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
GOLDEN = os.path.join(ROOT, "tests", "golden")


def workload(cfg, with_samples=True):
    from cge_jl_b200.divergence import draw_samples

    t0 = time.perf_counter()
    if cfg == 2:
        z = np.load(os.path.join(GOLDEN, "example10k.npz"))
        edges, ew, vw, comm, emb = (z[k] for k in ("edges", "eweights", "vweights", "comm",
                                                   "embedding"))
        name = "CGE.jl example/10k (n=10000, m=41536, d=32, k=64) --force-exact --seed 42"
        data = "reference example 10k graph (fixture tests/golden/example10k.npz)"
    elif cfg == 4:
        from cge_jl_b200.synth import abcd_like

        edges, ew, vw, comm, emb = abcd_like(200000, k=64, d=128, seed=1004)
        name = ("BASELINE config 4: synthetic ABCD-style undirected graph, n=200000, 64 communities, "
                "d=128, exact, --seed 42")
        data = "synthetic"
    else:
        raise SystemExit("unknown workload")
    t1 = time.perf_counter()
    n = vw.shape[0]
    samples = draw_samples(edges, ew, n, 10000, 42, directed=False, exact=True) if with_samples else None
    return dict(cfg=cfg, edges=edges, ew=ew, vw=vw, comm=comm, emb=emb, n=n, samples=samples,
                name=name, data=data, t_generate=t1 - t0, t_sampling=time.perf_counter() - t1)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period=0.1):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.period = index, [], False, period

    def run(self):
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                    "-i", str(self.index)], capture_output=True, text=True, timeout=5)
                if o.returncode == 0 and o.stdout.strip():
                    self.rows.append([c.strip() for c in o.stdout.strip().split(",")])
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({nm for r in self.rows for nm, v in zip(names, r[3:7]) if v == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------
# the reference's CPU path (oracle/: the line-by-line C port; Julia is not in this image)
# ---------------------------------------------------------------------------------------------
def pass_profile(cfg):
    """(alphas, fixed-point passes, B sweeps) of the FULL run of the workload.  Pass counts are a
    property of the input (GPU and oracle agree on them, tests/), frozen once from a run:
    tests/golden/config4_n1.json (make: scripts/freeze_config4.py) and the frozen oracle run of the
    10k example."""
    if cfg == 2:
        it = np.load(os.path.join(GOLDEN, "oracle_example10k_exact.npz"))["iters"].astype(int)
        return int((it > 0).sum()), int(it.sum()), int((it > 0).sum()), "frozen oracle run (260 s)"
    try:
        g = json.load(open(os.path.join(GOLDEN, "config4_n1.json")))
        return int(g["alphas"]), int(sum(g["iters"])), int(g["b_passes"]), "tests/golden/config4_n1.json"
    except Exception:
        return 40, 1041, 40, "round-1 run (profiles/r01_config_runs.jsonl)"


def reference_sample(cfg, sample_n=4000, max_alphas=2):
    """One bounded sample of the reference's algorithm on this box's host cores, extrapolated to
    the full workload phase by phase.

    The reference (single-threaded Julia, restated line by line in oracle/cge_oracle.c) keeps D, GD
    and P as packed n(n+1)/2 arrays: 3 x 160 GB at 200 000 vertices, which no host holds, so the
    sample is the same generator at `sample_n` vertices (same d, same 64 communities, first
    `max_alphas` alpha values).  Its seconds per pair in each O(n^2) phase -- D build (d-dim
    distances), GD = (1-D)^alpha (one pow per pair and alpha), fixed-point passes, P, B -- are
    scaled to the full run's pairs and pass profile:
        T_full = pairs_full * (c_build + A*(c_pow + c_P) + W*c_pass + b*c_B),   value = pairs_full*A / T_full.
    For workload 2 the sample is the 10k example itself (first 2 alpha values), same formula."""
    import oracle

    if cfg == 2:
        w = workload(2)
        what = "the workload itself"
    else:
        from cge_jl_b200.divergence import draw_samples
        from cge_jl_b200.synth import abcd_like

        edges, ew, vw, comm, emb = abcd_like(sample_n, k=64, d=128, seed=1004)
        w = dict(edges=edges, ew=ew, vw=vw, comm=comm, emb=emb, n=sample_n,
                 samples=draw_samples(edges, ew, sample_n, 10000, 42, directed=False, exact=True))
        what = f"the same generator (ABCD-style, 64 communities, d=128) at n={sample_n}"
    t0 = time.perf_counter()
    _, tr = oracle.wgcl(w["edges"], w["ew"], w["comm"], w["emb"], np.zeros(w["n"]), w["vw"],
                        samples=w["samples"], max_alphas=max_alphas)
    dt = time.perf_counter() - t0
    ps = w["n"] * (w["n"] + 1) // 2
    a_s, w_s = int(tr.n_alpha_run), int(sum(tr.iters))
    b_s = int(np.sum(~np.isnan(np.array(tr.div))))
    t = list(tr.t_phase)
    c_build, c_pow, c_pass, c_p, c_b = (t[0] / ps, t[1] / (ps * a_s), t[2] / (ps * w_s),
                                        t[3] / (ps * a_s), t[4] / (ps * max(b_s, 1)))
    n_full = 200000 if cfg == 4 else 10000
    pf = n_full * (n_full + 1) // 2
    A, W, B, src = pass_profile(cfg)
    t_full = pf * (c_build + A * (c_pow + c_p) + W * c_pass + B * c_b)
    t_other = dt - sum(t)  # O(m), O(K): negligible, kept as measured
    value = pf * A / (t_full + t_other)
    note = (f"oracle/cge_oracle.c (C port of wGCL, 1 thread) on {what}, first {a_s} of 40 alpha values "
            f"({w_s} passes), {dt:.1f} s; ns per pair: D build {1e9 * c_build:.1f}, pow {1e9 * c_pow:.1f} "
            f"per alpha, pass {1e9 * c_pass:.2f}, P {1e9 * c_p:.2f}, B {1e9 * c_b:.2f}; extrapolated phase "
            f"by phase to the full run ({pf} pairs, {A} alphas, {W} passes, {B} B sweeps: {src}) = "
            f"{t_full:.0f} s of CPU")
    return value, dt, note


def run_reference(args):
    """--impl reference: the reference's CPU algorithm on the host cores.  The reference is
    single-threaded (no @threads / Distributed anywhere in src/), so cores = 1; the same sample at
    every N (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, times, note = [], [], ""
    t_start = time.perf_counter()
    budget_s = 240.0  # the whole reference arm must end within a few minutes
    warm = 0
    while warm < args.warmup:
        v, dt, note = reference_sample(args.workload)
        warm += 1
        if (time.perf_counter() - t_start) + (args.steps + args.warmup - warm) * dt > budget_s:
            vals.append(v)  # too slow to afford untimed runs: this one counts as the first step
            times.append(dt)
            break
    while len(vals) < args.steps:
        if vals and (time.perf_counter() - t_start) + times[-1] > budget_s:
            break
        v, dt, note = reference_sample(args.workload)
        vals.append(v)
        times.append(dt)
    val = float(np.mean(vals))
    w = workload(args.workload, with_samples=False) if args.workload == 2 else None
    name = w["name"] if w else ("BASELINE config 4: synthetic ABCD-style undirected graph, n=200000, 64 "
                                "communities, d=128, exact, --seed 42")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(vals), "steps_requested": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic" if args.workload == 4 else w["data"],
        "config": {"workload": name, "sample": note,
                   "julia": "not installed in this image: the C port stands in for CGE.jl (DESIGN.md 6)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "host_cores": os.cpu_count(),
                         "kind": "port", "sample": note},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------
# the B200 arm
# ---------------------------------------------------------------------------------------------
def parity_checks(cfg, out, stats, world):
    """Rank 0: the result of this run against the frozen single-GPU run of the same workload and
    against the frozen CPU oracle (streaming oracle for config 4: first alpha; line-by-line oracle
    for the 10k example: the full run)."""
    res = {}
    iters = [int(x) for x in list(stats.iters)[: int(stats.n_alpha_run)]]
    if cfg == 4:
        try:
            g = json.load(open(os.path.join(GOLDEN, "config4_n1.json")))
            rel = float(np.max(np.abs(np.array(out) - np.array(g["result"])) /
                               np.maximum(np.abs(np.array(g["result"])), 1e-300)))
            res["parity_vs_n1"] = {"iters_equal": iters == g["iters"],
                                   "best_alphas_equal": bool(out[0] == g["result"][0] and out[4] == g["result"][4]),
                                   "max_rel_diff": rel, "ok": bool(iters == g["iters"] and rel <= 1e-9),
                                   "frozen": "tests/golden/config4_n1.json (1 GPU, stored regime)"}
        except Exception as e:  # noqa: BLE001
            res["parity_vs_n1"] = {"ok": None, "error": str(e)}
        try:
            o = json.load(open(os.path.join(GOLDEN, "config4_oracle_alpha1.json")))
            na = int(o["alphas"])
            dd = [abs(stats.div[a] - o["div"][a]) / abs(o["div"][a]) for a in range(na)]
            da = [abs(stats.auc[a] - o["auc"][a]) / max(abs(o["auc"][a]), 1e-300) for a in range(na)]
            res["parity_vs_oracle"] = {"alphas_compared": na, "iters_equal": iters[:na] == o["iters"][:na],
                                       "div_rel_diff": float(max(dd)), "auc_rel_diff": float(max(da)),
                                       "ok": bool(iters[:na] == o["iters"][:na] and max(dd) <= 1e-9 and max(da) <= 1e-9),
                                       "frozen": "tests/golden/config4_oracle_alpha1.json (oracle/cge_oracle_stream.c)"}
        except Exception as e:  # noqa: BLE001
            res["parity_vs_oracle"] = {"ok": None, "error": str(e)}
    else:
        g = np.load(os.path.join(GOLDEN, "oracle_example10k_exact.npz"))
        ref = g["out"] if "out" in g.files else None
        it = [int(x) for x in g["iters"] if x > 0]
        rel = float(np.max(np.abs(out - ref) / np.maximum(np.abs(ref), 1e-300))) if ref is not None else None
        res["parity_vs_oracle"] = {"iters_equal": iters == it, "max_rel_diff": rel,
                                   "ok": bool(iters == it and (rel is None or rel <= 1e-9)),
                                   "frozen": "tests/golden/oracle_example10k_exact.npz (oracle/cge_oracle.c)"}
    return res


def secondary_config2(dv, local_rank):
    """BASELINE configs[1] (10k example, --force-exact) beside the headline workload: a cold
    one-shot call (fresh handle, first use of the kernels) and the steady state."""
    w = workload(2)
    n, pairs = w["n"], w["n"] * (w["n"] + 1) // 2
    problem, keep = dv.make_problem(w["edges"], w["ew"], w["comm"], w["emb"], np.zeros(n), w["vw"],
                                    None, None, None, False, False, w["samples"])
    t0 = time.perf_counter()
    sc = dv.Scorer(local_rank)
    sc.upload(problem, keep)
    out, st = sc.run()
    cold = time.perf_counter() - t0
    for _ in range(3):
        sc.upload(problem, keep)
        sc.run()
    ts, te, sweeps, sweep_ms = [], [], 0, 0.0  # (k_fixed_point launches only: without the fused first passes)
    for _ in range(5):
        t0 = time.perf_counter()
        sc.upload(problem, keep)
        t1 = time.perf_counter()
        out, st = sc.run()
        t2 = time.perf_counter()
        ts.append(t2 - t1)
        te.append(t2 - t0)
        sweeps += st.fp_sweeps - st.b_fused
        sweep_ms += st.ms_sweeps - st.ms_fused
    sc.close()
    a = int(st.n_alpha_run)
    # the first call of a FRESH process (kernel modules not loaded yet, CUDA's default lazy loading):
    # what a one-shot cge_b200_score from the Julia host pays
    fresh = None
    try:
        o = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "run_config.py"), "--config", "2"],
                           capture_output=True, text=True, timeout=300).stdout.strip().splitlines()
        d = json.loads(o[-1])
        fresh = {"s_upload": d["s_upload"], "s_run": d["s_run"], "s_sampling": d["s_sampling"]}
    except Exception as e:  # noqa: BLE001
        fresh = {"error": str(e)[:200]}
    return {"workload": w["name"], "value": pairs * a / float(np.mean(ts)), "unit": UNIT,
            "ms_per_step": 1e3 * float(np.mean(ts)), "e2e_value": pairs * a / float(np.mean(te)),
            "e2e_ms_per_step": 1e3 * float(np.mean(te)), "steps": 5, "warmup": 3,
            "first_call_in_this_process_s": cold, "first_call_in_a_fresh_process": fresh,
            "t_host_sampling_s": w["t_sampling"],
            "fixed_point_gbs": 8.0 * pairs * sweeps / (sweep_ms * 1e-3) / 1e9,
            "result": [float(x) for x in out],
            "parity": parity_checks(2, out, st, 1)}


def run_b200(args):
    import torch
    import torch.distributed as dist

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

    def max_over_ranks(vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(vals, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    t_wall0 = time.perf_counter()
    w = workload(args.workload)
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

    def step():
        """upload from host buffers + run; returns (seconds of the whole step, seconds of run())."""
        barrier()
        t0 = time.perf_counter()
        sc.upload(problem, keep)
        torch.cuda.synchronize()
        barrier()
        t1 = time.perf_counter()
        out, st = sc.run()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        whole, resident = max_over_ranks([t2 - t0, t2 - t1])
        barrier()
        return whole, resident, out, st

    warm_s = []
    for _ in range(args.warmup):
        t0 = time.perf_counter()
        step()
        warm_s.append(time.perf_counter() - t0)
    # time budget: a 200 000-vertex run takes ~30 s on one GPU; never let K steps run past the budget
    steps = args.steps
    if warm_s:
        left = args.budget_s - (time.perf_counter() - t_wall0)
        fit = max(1, int(left / max(warm_s[-1], 1e-3)))
        if world > 1:
            fit = int(min(max_over_ranks([-fit]))) * -1  # the same (smallest) count on every rank
        steps = max(1, min(steps, fit))
    sampler = ClockSampler(local_rank, period=0.1 if args.workload == 2 else 1.0)
    if rank == 0:
        sampler.start()
    dt_e2e = dt_res = 0.0
    ev_ms, sweep_ms, sweeps, launches, bs_ms, b_alone, fused_ms, fused = [], 0.0, 0, 0, 0.0, 0, 0.0, 0
    out = st = None
    for _ in range(steps):
        whole, resident, out, st = step()
        dt_e2e += whole
        dt_res += resident
        ev_ms.append(st.ms_build + st.ms_solve)
        sweep_ms += st.ms_sweeps
        bs_ms += st.ms_bsweeps
        b_alone += int(st.b_sweeps) - int(st.b_fused)
        fused_ms += float(st.ms_fused)
        fused += int(st.b_fused)
        sweeps += st.fp_sweeps
        launches += st.launches
    sampler.stop_flag = True
    stats = st
    a_run = int(stats.n_alpha_run)
    if rank != 0:
        sc.close()
        if world > 1:
            dist.destroy_process_group()
        return
    sampler.join(timeout=3)
    value = pairs * a_run * steps / dt_res
    e2e = pairs * a_run * steps / dt_e2e
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # dominant kernel: the fixed point.  Persistent drivers: one launch per alpha runs all its
    # passes; host loop: one launch per pass.  Algorithmic bytes = 8 B per unordered pair and pass.
    # The first pass of an alpha whose predecessor's B sweep was deferred runs in k_bfp<M> (B sweep + pass in
    # one matrix read); those launches are timed separately (stats.ms_fused) and reported as `fused_pass`,
    # so the roofline below is the fixed-point kernel's own launches and the passes they executed.
    persistent = int(stats.driver) in (2, 3)
    fp_launches = (a_run if persistent else int(stats.fp_sweeps)) * steps
    fp_passes = sweeps - fused
    passes_per_launch = fp_passes / max(fp_launches, 1)
    bytes_per_launch = 8.0 * pairs / world * passes_per_launch
    avg_launch_s = 1e-3 * (sweep_ms - fused_ms) / max(fp_launches, 1)
    achieved = bytes_per_launch / avg_launch_s / 1e9 if avg_launch_s > 0 else 0.0
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum from the committed ncu capture of this workload
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))[str(args.workload)]
        if world == 1 and tr["workload_pairs"] == pairs and int(stats.regime) == 1:
            traffic = tr["dram_bytes_per_pass"] * passes_per_launch
    except Exception:
        pass
    stored = int(stats.regime) == 1
    kname = {1: "k_sweep<M,false> (one fixed-point pass per launch)",
             2: "k_fixed_point<M,false> (all passes of one alpha per cooperative launch)",
             3: "k_fixed_point_ring<M,false> (all passes of one alpha, cp.async.bulk ring)"}
    if not stored:
        kname = {1: "k_sweep_rc<false>", 2: "k_fixed_point_rc<false>", 3: "k_fixed_point_rc<false>"}
    elif int(stats.b_fused) > 0:
        kname[2] = ("k_fixed_point<M,false> (all passes of one alpha per cooperative launch but the first, which "
                    "k_bfp<M> runs together with the previous alpha's B sweep: see fused_pass)")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
        "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt_res / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": w["data"],
        "config": {"workload": w["name"], "alphas_evaluated": a_run,
                   "fixed_point_passes": int(stats.fp_sweeps), "b_passes": int(stats.b_sweeps),
                   "b_passes_fused_with_a_fixed_point_pass": int(stats.b_fused),
                   "pairs": pairs, "samples_local": 10000, "tiles": int(stats.n_tiles),
                   "l2": "inputs larger than L2 (q matrix %.1f GB per GPU vs 126 MB L2)"
                         % (stats.matrix_bytes / 1e9),
                   "driver": {1: "hostloop", 2: "persistent", 3: "ring"}.get(int(stats.driver), "?"),
                   "regime": {1: "stored", 2: "recompute (row-norm/dot form)", 3: "recompute (row-norm/dot form)",
                              4: "recompute (difference form)"}.get(int(stats.regime), "?"),
                   "exchange": None if world == 1 else ("NCCL all-reduce per pass (host loop)" if args.no_p2p
                                                        else "NVLink peer memory inside the kernel"),
                   "timed_region": "each step = upload from host buffers (e2e only) + run(); value = run() alone",
                   "result": [float(x) for x in out],
                   "iters": [int(x) for x in list(stats.iters)[:a_run]],
                   "device_ms_per_step": float(np.mean(ev_ms)),
                   "t_host_s": {"generate_graph": w["t_generate"], "sampling_draw_samples": w["t_sampling"]},
                   "ms_breakdown_last_step": {
                       "upload": float(stats.ms_upload), "build": float(stats.ms_build),
                       "solve": float(stats.ms_solve), "fp_kernels": float(stats.ms_sweeps),
                       "b_kernels": float(stats.ms_bsweeps), "total": float(stats.ms_total)},
                   "budget_s": args.budget_s},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * dt_e2e / steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic,
                     "kernel": kname.get(int(stats.driver), "?"),
                     "bytes_per_launch": bytes_per_launch, "avg_launch_us": 1e6 * avg_launch_s,
                     "passes_per_launch": passes_per_launch,
                     "avg_pass_us": 1e6 * avg_launch_s / max(passes_per_launch, 1e-9),
                     "launches_timed": int(fp_launches),
                     "b_sweep_gbs": 8.0 * pairs / world * b_alone / (bs_ms * 1e-3) / 1e9
                                    if bs_ms > 0 and b_alone > 0 else None,
                     "fused_pass": None if fused == 0 else {
                         "kernel": "k_bfp<M> (B sweep of alpha-1/4 + first fixed-point pass of alpha, one matrix read)",
                         "launches_timed": int(fused), "avg_launch_us": 1e3 * fused_ms / fused,
                         "gbs": 8.0 * pairs / world * fused / (fused_ms * 1e-3) / 1e9,
                         "frac": 8.0 * pairs / world * fused / (fused_ms * 1e-3) / 1e9 / peak if peak else None,
                         "note": "8 B x unordered pairs credited ONCE per launch although it does the work of two sweeps"},
                     "note": "achieved = 8 B x unordered pairs x passes in the launch / CUDA-event "
                             "duration of the launch, events recorded by the library on its stream; per GPU. "
                             "The B sweep of an alpha rides on the first fixed-point pass of the next alpha "
                             "(k_bfp: one matrix read for both): those launches and the pass they carry are "
                             "reported under fused_pass, not here; b_sweep_gbs covers the stand-alone B sweeps",
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)"
                                    if "hbm_gbs" in peaks else "fallback 6650 GB/s"},
    }
    if not stored:
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
    line.update(parity_checks(args.workload, out, stats, world))
    sc.close()
    if world == 1 and not args.no_cpu_baseline:
        v, dt, note = reference_sample(args.workload, sample_n=10000)  # ~12 s of CPU
        line["cpu_baseline"] = {
            "value": v, "unit": UNIT, "cores": 1, "host_cores": os.cpu_count(), "kind": "port",
            "threads_note": "the reference has no threading (no @threads/@spawn/Distributed in "
                            "src/): JULIA_NUM_THREADS does not change it, so one core is used",
            "sample": note}
    if world == 1 and args.workload == 4 and not args.no_secondary:
        line["secondary"] = secondary_config2(dv, local_rank)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", type=int, default=4, choices=[2, 4],
                    help="BASELINE.json config: 4 = 200k ABCD graph d=128 (default), 2 = 10k example")
    ap.add_argument("--driver", type=int, default=0, help="0 auto, 1 host loop, 2 persistent")
    ap.add_argument("--regime", type=int, default=0,
                    help="0 auto, 1 stored, 2 recompute (row-norm/dot form), 4 recompute (difference form)")
    ap.add_argument("--budget-s", type=float, default=float(os.environ.get("CGE_BENCH_BUDGET_S", "1500")),
                    help="wall-clock budget: fewer than --steps timed steps run when K steps would not fit "
                         "(the line then carries the real `steps` next to `steps_requested`)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
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
