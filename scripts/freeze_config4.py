"""Freezes the single-GPU result of BASELINE config 4 (what bench.py's `parity_vs_n1` and
`pass_profile` compare against) from a bench line:

  python bench.py --steps 1 --warmup 3 > gpurun_out/bench_cfg4_n1.json      # on a B200
  python scripts/freeze_config4.py gpurun_out/bench_cfg4_n1.json           # here

Writes tests/golden/config4_n1.json: result vector, per-alpha pass counts, B sweeps.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    line = None
    for ln in open(sys.argv[1]):
        ln = ln.strip()
        if ln.startswith("{") and '"metric"' in ln:
            line = json.loads(ln)
    if line is None or line["n_gpus"] != 1 or "config 4" not in line["config"]["workload"]:
        raise SystemExit("not a single-GPU config-4 bench line")
    c = line["config"]
    out = {"workload": c["workload"], "result": c["result"], "iters": c["iters"],
           "alphas": c["alphas_evaluated"], "fp_passes": c["fixed_point_passes"],
           "b_passes": c["b_passes"], "regime": c["regime"], "driver": c["driver"],
           "made_by": "scripts/freeze_config4.py from a bench.py line (1 B200, stored regime)"}
    with open(os.path.join(ROOT, "tests", "golden", "config4_n1.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out)[:400])


if __name__ == "__main__":
    main()
