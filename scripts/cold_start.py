"""Cold-start cost of one exact run in a fresh process under lazy / eager CUDA module loading."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for mode in ("LAZY", "EAGER", "LAZY"):
    env = dict(os.environ, CUDA_MODULE_LOADING=mode)
    t0 = time.perf_counter()
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "run_config.py"), "--config", "2"],
                         env=env, capture_output=True, text=True).stdout.strip().splitlines()
    wall = time.perf_counter() - t0
    d = json.loads(out[-1])
    print(mode, "process wall %.2f s" % wall, "s_run %.3f" % d["s_run"], "s_upload %.3f" % d["s_upload"],
          "fp_kernels %.1f ms" % d["ms_fp_kernels"], flush=True)
