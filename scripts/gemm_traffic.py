"""From an ncu launch list with gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum (CSV): the
per-kernel summary with DRAM traffic and the GEMM traffic record bench.py reports as roofline.traffic.
usage: python scripts/gemm_traffic.py launches.csv B summary.txt traffic.json"""
import collections, csv, json, re, sys
path, B, out_txt, out_json = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
per = collections.OrderedDict()
for r in csv.DictReader(lines):
    d = per.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"])})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["us"] = v / 1e3 if unit.startswith("n") else (v if unit.startswith("u") else v * 1e3)
    else:
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        d["rd" if "read" in r["Metric Name"] else "wr"] = v * mult
agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get("us", 0); a[2] += d.get("rd", 0); a[3] += d.get("wr", 0)
tot = sum(a[1] for a in agg.values())
with open(out_txt, "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one eager step,\n"
            f"# config 2, {B} images per step (python scripts/profile_step.py {B}). Serialised, cold-cache per-launch times: compare SHARES.\n")
    f.write(f"{'kernel':58s} {'launches':>8s} {'total_us':>10s} {'share':>6s} {'avg_us':>8s} {'dram_rd_MB':>10s} {'dram_wr_MB':>10s} {'GB/s':>7s}\n")
    for k, (n, t, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k[:58]:58s} {n:8d} {t:10.1f} {100*t/tot:5.1f}% {t/n:8.1f} {rd/1e6:10.1f} {wr/1e6:10.1f} {(rd+wr)/t/1e3:7.0f}\n")
    f.write(f"{'TOTAL':58s} {sum(a[0] for a in agg.values()):8d} {tot:10.1f}\n")
g = [(n, t, rd, wr) for k, (n, t, rd, wr) in agg.items() if "gemm_f16_kernel" in k]
n = sum(x[0] for x in g); t = sum(x[1] for x in g); b = sum(x[2] + x[3] for x in g)
json.dump({"kernel": f"gemm_f16_kernel (all launches of one {B}-image step)", "launches": n,
           "dram_bytes_per_launch_avg": b / n, "us_per_launch_avg_under_ncu": t / n,
           "gemm_share_of_step_under_ncu": t / tot, "images_per_step": B,
           "source": f"{out_txt.replace('_summary.txt', '.csv')} (ncu dram__bytes_read.sum + dram__bytes_write.sum, scripts/profile_step.py {B})"},
          open(out_json, "w"), indent=1)
print(open(out_txt).read())
