"""Pick the metrics that matter out of `ncu -i X.ncu-rep --page raw --csv` (one line per metric, per kernel)."""
import csv, sys
KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit", "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor",
        "sm__pipe_fma_cycles_active.avg.pct", "sm__pipe_fmaheavy", "sm__inst_executed_pipe_xu.avg.pct", "sm__inst_executed_pipe_alu.avg.pct",
        "sm__inst_executed_pipe_lsu.avg.pct", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__issue_active.avg.pct", "lts__t_sector_hit_rate.pct", "smsp__average_warp", "launch__shared_mem_per_block",
        "sm__cycles_elapsed.max", "smsp__warp_issue_stalled", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct")
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print(f"## {d.get('Kernel Name', '?')[:100]}  grid={d.get('Grid Size')} block={d.get('Block Size')}")
    for h, u, v in zip(hdr, units, vals):
        if any(k in h for k in KEYS) and v not in ("", "n/a"):
            print(f"  {h} [{u}] = {v}")
