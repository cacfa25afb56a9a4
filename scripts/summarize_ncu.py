"""Key metrics of an `ncu --set full` report (run here, no GPU needed): ncu -i X.ncu-rep --page raw --csv | this."""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active", "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "lts__t_bytes.sum", "sm__pipe_xu_cycles_active", "smsp__inst_executed.sum", "gpc__cycles_elapsed.max"]
for h, u, v in zip(hdr, units, vals):
    if any(h == w or h.startswith(w) for w in want) and "per_second" not in h and ".pct_of_peak_sustained_elapsed" not in h.replace("avg.pct_of_peak_sustained_elapsed", ""):
        print(f"{h} = {v} {u}")
