"""Runs one of BASELINE.json's configurations end to end on the B200 path and checks the
size-independent properties that stand in for the oracle where the reference cannot allocate its
arrays (SURVEY.md section 7 "Reference infeasibility"):

  * the fixed point converged for every alpha: max |w - S| <= 0.001 for the final S (debug probe);
  * scores are finite, best alphas lie on the grid, the local-score error matches its formula;
  * (multi-GPU) every rank returns the same vector.

  python scripts/run_config.py --config 3                       # 1 GPU
  torchrun --nproc-per-node 8 scripts/run_config.py --config 4  # tiles sharded over 8 GPUs

  torchrun --nproc-per-node 8 scripts/run_config.py --config 5 --exact --max-alphas 2   # 1M vertices

Appends one JSON line per run to gpurun_out/config_runs.jsonl (rank 0); the committed copies are
profiles/r01_config_runs.jsonl and profiles/r02_config_runs.jsonl.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cge_jl_b200 import divergence as dv  # noqa: E402
from cge_jl_b200.landmarks import landmarks, split_cluster_rss  # noqa: E402
from cge_jl_b200.synth import abcd_like, planted_partition  # noqa: E402

LANDMARKS = 0
EXACT = False
EMPTY = (np.zeros(0), np.zeros(0, dtype=np.int64), np.zeros((0, 0), dtype=np.int64), np.zeros(0),
         np.zeros((0, 0)))


def build(cfg, dev=None):
    """dev: a Scorer -- landmark selection and aggregation then run on the GPU."""
    t0 = time.perf_counter()
    directed, lm = False, None
    t_lm = None
    if isinstance(cfg, str):  # "n,d,k,directed" -- ad-hoc synthetic problem (profiling)
        n, d, k, dr = (int(x) for x in cfg.split(","))
        directed = bool(dr)
        edges, ew, vw, comm, emb = planted_partition(n, k=k, d=d, seed=7, directed=directed,
                                                     weighted=directed)
        name = f"synthetic planted partition n={n} d={d} k={k} directed={directed}"
        if LANDMARKS > 0:
            by = {}
            for v, c in enumerate(comm[:, 0], start=1):
                by.setdefault(int(c), []).append(v)
            t1 = time.perf_counter()
            lm = landmarks(edges, ew, vw, [np.asarray(v) for v in by.values()], comm, emb, False,
                           LANDMARKS, 4, split_cluster_rss, directed, device=dev)
            t_lm = time.perf_counter() - t1
            name += f" landmarks -l {LANDMARKS}"
    elif cfg in (1, 2):
        z = np.load(os.path.join(ROOT, "tests", "golden", "example10k.npz"))
        edges, ew, vw, comm, emb = (z[k] for k in ("edges", "eweights", "vweights", "comm", "embedding"))
        name = "10k example, -l 200 --seed 42 (landmarks)" if cfg == 1 else "10k example --force-exact"
        if cfg == 1:
            by = {}
            for v, c in enumerate(comm[:, 0], start=1):
                by.setdefault(int(c), []).append(v)
            t1 = time.perf_counter()
            lm = landmarks(edges, ew, vw, [np.asarray(v) for v in by.values()], comm, emb, False,
                           200, 4, split_cluster_rss, False, device=dev)
            t_lm = time.perf_counter() - t1
    elif cfg == 3:
        edges, ew, vw, comm, emb = planted_partition(50000, k=32, d=64, seed=1003, directed=True,
                                                     weighted=True)
        directed = True
        name = "synthetic 50k-node weighted planted-partition digraph, d=64, k=32, exact, directed"
    elif cfg == 4:
        edges, ew, vw, comm, emb = abcd_like(200000, k=64, d=128, seed=1004)
        name = "synthetic 200k-node ABCD-style graph, 64 communities, d=128, exact"
    elif cfg == 5 and EXACT:
        # the exact half of BASELINE.json configs[4]: 1M vertices, d = 128, 5e11 pairs = 4 TB of q:
        # recompute regime (+ what fits of the matrix kept in HBM), super-tiles sharded over the ranks
        edges, ew, vw, comm, emb = planted_partition(1000000, k=64, d=128, seed=1005)
        name = "synthetic 1M-node planted-partition graph, d=128, k=64, exact (--force-exact)"
    elif cfg == 5:
        # the landmark half of BASELINE.json configs[4]: 1M vertices, d = 128, rss landmarks -l 4000
        # on one GPU
        edges, ew, vw, comm, emb = planted_partition(1000000, k=64, d=128, seed=1005)
        name = "synthetic 1M-node planted-partition graph, d=128, k=64, rss landmarks -l 4000 (1 GPU)"
        by = {}
        for v, c in enumerate(comm[:, 0], start=1):
            by.setdefault(int(c), []).append(v)
        t1 = time.perf_counter()
        lm = landmarks(edges, ew, vw, [np.asarray(v) for v in by.values()], comm, emb, False,
                       4000, 4, split_cluster_rss, False, device=dev)
        t_lm = time.perf_counter() - t1
        print(f"landmarks(): {t_lm:.1f} s, N = {lm[1].shape[0]}", flush=True)
    else:
        raise SystemExit("unknown config")
    return dict(edges=edges, ew=ew, vw=vw, comm=comm, emb=emb, directed=directed, lm=lm, name=name,
                t_gen=time.perf_counter() - t0, t_lm=t_lm)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=0)
    ap.add_argument("--synthetic", default="", help="n,d,k,directed instead of --config")
    ap.add_argument("--landmarks", type=int, default=0, help="with --synthetic: landmark mode, -l L")
    ap.add_argument("--samples", type=int, default=10000)
    ap.add_argument("--no-p2p", action="store_true")
    ap.add_argument("--regime", type=int, default=0)
    ap.add_argument("--driver", type=int, default=0, help="0 auto, 1 host loop, 2 persistent, 3 TMA ring")
    ap.add_argument("--max-alphas", type=int, default=0)
    ap.add_argument("--exact", action="store_true", help="config 5: the exact half (--force-exact)")
    ap.add_argument("--spot", type=int, default=0, help="vertices of the NumPy degree-sum spot check")
    ap.add_argument("--host-landmarks", action="store_true",
                    help="landmark mode: runsplit + aggregation in NumPy instead of on the GPU")
    args = ap.parse_args()
    global EXACT
    EXACT = args.exact
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    global LANDMARKS
    LANDMARKS = args.landmarks
    sc = dv.Scorer(local_rank)
    c = build(args.synthetic or args.config, None if args.host_landmarks else sc)
    n = c["vw"].shape[0]
    t0 = time.perf_counter()
    if c["lm"] is None:
        samples = dv.draw_samples(c["edges"], c["ew"], n, args.samples, 42, c["directed"], True)
        prob = (c["edges"], c["ew"], c["comm"], c["emb"], np.zeros(n), c["vw"], None, None, None)
        n_scored, target = n, c["vw"]
    else:
        dii, lemb, lcomm, ledges, lw, lweight, v2l = c["lm"]
        samples = dv.draw_samples(c["edges"], c["ew"], n, args.samples, 42, c["directed"], False)
        prob = (ledges, lw, lcomm, lemb, dii, lweight, c["vw"], v2l, c["emb"])
        n_scored, target = lemb.shape[0], lweight
    t_sample = time.perf_counter() - t0
    if world > 1:
        ids = [dv.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        sc.comm_init(ids[0], rank, world)
        if not args.no_p2p:
            handles = [None] * world
            dist.all_gather_object(handles, sc.p2p_export(n_scored))
            sc.p2p_import(handles)
    p, keep = dv.make_problem(*prob, False, c["directed"], samples, args.max_alphas, args.driver, args.regime)
    t0 = time.perf_counter()
    sc.upload(p, keep)
    t_upload = time.perf_counter() - t0
    t0 = time.perf_counter()
    out, st = sc.run()
    t_run = time.perf_counter() - t0
    # properties
    checks = {}
    if c["directed"]:
        e = c["edges"]
        din, dout = np.zeros(n), np.zeros(n)
        np.add.at(dout, e[:, 0] - 1, c["ew"])
        np.add.at(din, e[:, 1] - 1, c["ew"])
        sin, sout = sc.debug_read(3, n_scored), sc.debug_read(4, n_scored)
        res = max(np.abs(din - sin)[din > 0].max(), np.abs(dout - sout)[dout > 0].max())
    else:
        res = float(np.abs(target - sc.debug_read(3, n_scored)).max())
    checks["final_residual"] = float(res)
    checks["converged"] = bool(res <= 0.001)
    checks["finite"] = bool(np.all(np.isfinite(out)))
    checks["alphas_on_grid"] = bool(out[0] * 4 == round(out[0] * 4) and out[4] * 4 == round(out[4] * 4))
    checks["auc_err_formula"] = bool(np.isclose(out[6], 1.96 * np.sqrt(out[5] * (1 - out[5]) / args.samples)))
    if world > 1:
        outs = [None] * world
        dist.all_gather_object(outs, out.tolist())
        checks["ranks_agree"] = bool(all(o == outs[0] for o in outs))
    if args.spot > 0 and rank == 0 and not c["directed"] and c["lm"] is None:
        # sub-block spot check: S of a few vertices from scratch in NumPy (scripts/spotcheck.py)
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        from spotcheck import degree_sum_spot_check
        t0 = time.perf_counter()
        rows = np.random.default_rng(5).integers(0, n, size=args.spot)
        worst = degree_sum_spot_check(c["emb"], sc.debug_read(1, n), sc.debug_read(3, n), c["vw"],
                                      float(st.hi), 0.25 * int(st.n_alpha_run), rows)
        checks["spot_check_rows"] = int(args.spot)
        checks["spot_check_max_rel_diff"] = float(worst)
        checks["spot_check_ok"] = bool(worst <= 1e-9)
        checks["spot_check_s"] = time.perf_counter() - t0
    pairs = n_scored * (n_scored + 1) // 2 if not c["directed"] else n_scored * n_scored
    line = {
        "config": args.config, "workload": c["name"], "n_gpus": world, "n_scored": int(n_scored),
        "result": [float(x) for x in out], "alphas": int(st.n_alpha_run),
        "fp_passes": int(st.fp_sweeps), "b_passes": int(st.b_sweeps), "b_fused": int(st.b_fused),
        "driver": int(st.driver), "regime": int(st.regime),
        "iters": [int(x) for x in list(st.iters)[: int(st.n_alpha_run)]],
        "div": [float(x) for x in list(st.div)[: int(st.n_alpha_run)]],
        "auc": [float(x) for x in list(st.auc)[: int(st.n_alpha_run)]],
        "matrix_gb_per_gpu": st.matrix_bytes / 1e9,
        "s_run": t_run, "s_upload": t_upload, "s_sampling": t_sample, "s_generate": c["t_gen"],
        "s_landmarks": c["t_lm"], "landmarks_on": None if c["lm"] is None else ("host" if args.host_landmarks else "gpu"),
        "ms_fp_kernels": float(st.ms_sweeps), "ms_b_kernels": float(st.ms_bsweeps), "ms_fused_kernels": float(st.ms_fused),
        "pair_alphas_per_s": pairs * int(st.n_alpha_run) / t_run,
        # the fixed-point kernel's own launches (the first pass of an alpha that carries the previous alpha's
        # B sweep runs in k_bfp and is timed as ms_fused)
        "avg_pass_ms": float(st.ms_sweeps - st.ms_fused) / max(int(st.fp_sweeps) - int(st.b_fused), 1),
        "gbps_per_gpu_fp": 8.0 * (n_scored * (n_scored + 1) // 2) / world * (int(st.fp_sweeps) - int(st.b_fused))
                           / (float(st.ms_sweeps - st.ms_fused) * 1e-3) / 1e9 if st.ms_sweeps > st.ms_fused else None,
        "checks": checks,
    }
    if int(st.regime) >= 2 and st.ms_sweeps > 0:
        # recompute regime: FP64 roofline, (2d + 6) flop per recomputed pair and pass;
        # pairs whose q tile is kept in HBM ("store what fits") are not counted as flops
        d_emb = c["emb"].shape[1] if c["lm"] is None else c["lm"][1].shape[1]
        upairs = n_scored * (n_scored + 1) // 2
        rc_pairs = max(upairs / world - st.matrix_bytes / 8.0, 0.0)
        tf = (2.0 * d_emb + 6.0) * rc_pairs * int(st.fp_sweeps) / (float(st.ms_sweeps) * 1e-3) / 1e12
        peak = sc.fp64_peak_tflops()
        line.update({"recomputed_pair_share": rc_pairs / (upairs / world), "tflops_per_gpu_fp": tf,
                     "fp64_peak_tflops": peak, "fp64_frac": tf / peak,
                     "fp64_frac_note": "upper bound when part of the matrix is read from HBM: the pass "
                                       "time also covers the stored tiles"})
    if rank == 0:
        print(json.dumps(line))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "config_runs.jsonl"), "a") as f:
            f.write(json.dumps(line) + "\n")
    sc.close()
    if world > 1:
        dist.destroy_process_group()
    if not all(v for k, v in checks.items() if isinstance(v, bool)):
        raise SystemExit("property check failed: " + json.dumps(checks))


if __name__ == "__main__":
    main()
