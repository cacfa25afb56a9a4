"""Aggregates an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list by
kernel name; with --traffic-json OUT also writes the average DRAM bytes per tcgen05-GEMM launch (bench.py's
roofline.traffic).  usage: summarize_launches.py LIST.csv [--images N] [--traffic-json OUT.json]"""
import collections
import csv
import json
import re
import sys

path = sys.argv[1]
images = sys.argv[sys.argv.index("--images") + 1] if "--images" in sys.argv else "?"
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.OrderedDict()
for r in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    a = agg.setdefault(name, {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
    m, u = r["Metric Name"], r["Metric Unit"]
    if m == "gpu__time_duration.sum":
        a["n"] += 1
        a["us"] += v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
    elif m.startswith("dram__bytes"):
        a["rd" if "read" in m else "wr"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
tot = sum(a["us"] for a in agg.values())
print("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one eager step,")
print(f"# config 2, {images} images per step (python scripts/profile_step.py {images}). Serialised, cold-cache per-launch times: compare SHARES.")
print(f"{'kernel':58s} {'launches':>8s} {'total_us':>10s} {'share':>6s} {'avg_us':>8s} {'dram_rd_MB':>10s} {'dram_wr_MB':>10s} {'GB/s':>7s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    print(f"{k[:58]:58s} {a['n']:8d} {a['us']:10.1f} {100 * a['us'] / tot:5.1f}% {a['us'] / a['n']:8.1f} {a['rd'] / 1e6:10.1f} "
          f"{a['wr'] / 1e6:10.1f} {(a['rd'] + a['wr']) / a['us'] / 1e3:7.0f}")
print(f"{'TOTAL':58s} {sum(a['n'] for a in agg.values()):8d} {tot:10.1f}")
if "--traffic-json" in sys.argv:
    out = sys.argv[sys.argv.index("--traffic-json") + 1]
    g = [a for k, a in agg.items() if "gemm_f16_kernel" in k]
    n, b = sum(a["n"] for a in g), sum(a["rd"] + a["wr"] for a in g)
    with open(out, "w") as f:
        json.dump({"dram_bytes_per_launch_avg": b / n, "launches": n,
                   "source": f"{path} (ncu dram__bytes_read.sum + dram__bytes_write.sum, scripts/profile_step.py {images})"}, f)
