"""Per-kernel evidence table from `ncu --set full` captures of one step (run here, no GPU needed):

    ncu --set full --clock-control none --profile-from-start off -o gpurun_out/step_full python scripts/profile_step.py 8
    ncu -i gpurun_out/step_full.ncu-rep --page raw --csv > step_full_raw.csv
    python scripts/kernel_evidence.py step_full_raw.csv [more.csv[.gz] ...]     (several captures of the same step merge)

For every kernel of the library: number of launches captured, and -- for its LONGEST launch, i.e. the hot-path shape --
duration, DRAM bytes and achieved HBM bandwidth (absolute and against the measured copy peak of MEASURED_PEAKS.json),
tensor-pipe activity, SM throughput, registers, achieved occupancy, IPC.  north_star asks for exactly these two columns
(achieved HBM GB/s, tensor-pipe %) for each kernel."""
import csv
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
        PEAK = float(json.load(f)["hbm_gbs"])
except Exception:
    PEAK = 6546.9
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
        "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}


def strip_args(name):
    """'f<(int)2>(A, B)' -> 'f<(int)2>': drops the trailing parenthesised argument list only"""
    name = name.strip()
    if not name.endswith(")"):
        return name
    depth = 0
    for i in range(len(name) - 1, -1, -1):
        depth += name[i] == ")"
        depth -= name[i] == "("
        if depth == 0:
            return name[:i]
    return name


class Row:
    def __init__(self, col, units, cells):
        self.col, self.units, self.cells = col, units, cells

    def num(self, name, default=0.0):
        i = self.col.get(name)
        if i is None or self.cells[i] == "":
            return default
        return float(self.cells[i].replace(",", ""))

    def scaled(self, name):
        """a byte or time metric in bytes / microseconds (ncu picks a unit per column)"""
        i = self.col.get(name)
        if i is None or self.cells[i] == "":
            return 0.0
        if self.units[i] not in UNIT:
            raise SystemExit(f"unit {self.units[i]!r} of {name} not handled")
        return float(self.cells[i].replace(",", "")) * UNIT[self.units[i]]


def main():
    kern = {}
    for path in sys.argv[1:]:
        with (gzip.open if path.endswith(".gz") else open)(path, "rt", newline="") as f:
            rows = list(csv.reader(f))
        hdr, units = rows[0], rows[1]
        col = {n: i for i, n in enumerate(hdr)}
        for cells in rows[2:]:
            if len(cells) < len(hdr):
                continue
            name = strip_args(cells[col["Kernel Name"]])
            for junk in ("void ", "rlcf::", "(int)", "(bool)"):
                name = name.replace(junk, "")
            if name.startswith("at::") or "elementwise_kernel" in name:   # torch fills between the library's launches
                continue
            r = Row(col, units, cells)
            dur = r.scaled("gpu__time_duration.sum")
            k = kern.setdefault(name, dict(n=0, total=0.0, best=None, best_dur=-1.0))
            k["n"] += 1
            k["total"] += dur
            if dur > k["best_dur"]:
                k["best_dur"], k["best"] = dur, r
    total = sum(k["total"] for k in kern.values())
    print(f"# HBM peak for the %HBM column: {PEAK:.1f} GB/s (MEASURED_PEAKS.json, copy kernel)")
    print(f"# {'kernel':44s} {'n':>4s} {'share':>6s} {'us':>8s} {'rd MB':>8s} {'wr MB':>8s} {'GB/s':>7s} {'%HBM':>5s} "
          f"{'tensor%':>7s} {'SM%':>5s} {'regs':>4s} {'occ%':>5s} {'IPC':>5s}")
    for name, k in sorted(kern.items(), key=lambda kv: -kv[1]["total"]):
        r, dur = k["best"], k["best_dur"]
        rd, wr = r.scaled("dram__bytes_read.sum"), r.scaled("dram__bytes_write.sum")
        gbs = (rd + wr) / dur / 1e3 if dur > 0 else 0.0
        print(f"  {name[:44]:44s} {k['n']:4d} {100 * k['total'] / total:5.1f}% {dur:8.1f} {rd / 1e6:8.1f} {wr / 1e6:8.1f} "
              f"{gbs:7.0f} {100 * gbs / PEAK:5.1f} "
              f"{r.num('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):7.1f} "
              f"{r.num('sm__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} "
              f"{int(r.num('launch__registers_per_thread')):4d} "
              f"{r.num('sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} "
              f"{r.num('sm__inst_executed.avg.per_cycle_elapsed'):5.2f}")
    print(f"# total of the captured launches: {total / 1e3:.2f} ms (serialised, cold-cache replay; us / MB columns are "
          "the longest launch of each kernel)")


if __name__ == "__main__":
    main()
