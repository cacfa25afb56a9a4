"""A/B of one environment switch at the step level, same box, alternating runs:
    python scripts/ab_bench.py RLCF_GEMM_MULTICAST 0 1 [--rounds 2] [-- extra bench.py args]
Prints value / e2e / median SM clock per run (bench.py without the baseline and other-mode legs)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
extra = []
if "--" in args:
    i = args.index("--")
    args, extra = args[:i], args[i + 1:]
rounds = 2
if "--rounds" in args:
    i = args.index("--rounds")
    rounds = int(args[i + 1])
    del args[i:i + 2]
var, values = args[0], args[1:]
for _ in range(rounds):
    for v in values:
        env = dict(os.environ, **{var: v})
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "5",
                              "--no-other-modes", "--no-torch-gpu-baseline"] + extra, env=env, capture_output=True,
                             text=True, timeout=900).stdout.strip().splitlines()
        b = json.loads(out[-1])
        print(f"{var}={v}: value {b['value']:.2f} e2e {b['e2e']['value']:.2f} ms/step {b['ms_per_step']:.2f} "
              f"sm_mhz {b['clocks']['sm_mhz']} power_w_max {b['clocks'].get('power_w_max')}", flush=True)
