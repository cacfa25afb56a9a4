"""Runs bench.py in the other modes at a few batch sizes and prints one line each (smoke test of the standalone
--mode / --config paths + a look at how the throughput moves with the images per step)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
runs = [("prompt", 2, 8), ("prompt", 2, 16), ("full", 2, 8), ("full", 2, 16), ("ln", 5, 8)]
if len(sys.argv) > 1:
    runs = [(m, int(c), int(b)) for m, c, b in (a.split(":") for a in sys.argv[1:])]
for mode, config, b in runs:
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--mode", mode, "--config", str(config),
                        "--images-per-step", str(b), "--steps", "6", "--warmup", "3", "--no-other-modes",
                        "--no-torch-gpu-baseline"], capture_output=True, text=True, timeout=900)
    lines = r.stdout.strip().splitlines()
    if r.returncode != 0 or not lines:
        print(f"{mode} config {config} B={b}: FAILED rc={r.returncode}: {r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ''}")
        continue
    j = json.loads(lines[-1])
    ws = j["roofline"].get("whole_step", {})
    print(f"{mode} config {config} B={b}: value {j['value']:.1f} e2e {j['e2e']['value']:.1f} ms/step {j['ms_per_step']:.1f} "
          f"sm_mhz {j['clocks']['sm_mhz']} needed GF/image {ws.get('algorithmic_gflop_per_image', 0):.0f} "
          f"(reference {ws.get('reference_gflop_per_image') or 0:.0f}) frac {ws.get('frac', 0):.3f}", flush=True)
