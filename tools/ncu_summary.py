"""Print the key counters of every kernel in an `ncu --page raw --csv` export.  Usage: python tools/ncu_summary.py raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max",
        "smsp__warps_eligible.avg.per_cycle_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    print("----", r[hdr.index("Kernel Name")][:60])
    for w in want:
        if w in hdr:
            print("   %-70s %s" % (w, r[hdr.index(w)]))
    items = [(float(r[i] or 0), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    print("   stalls:", ", ".join("%s=%.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v)
                                  for v, h in sorted(items, reverse=True)[:7]))
