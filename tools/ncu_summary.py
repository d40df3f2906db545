"""Key metrics of one kernel from an `ncu --set full` report, as metric,unit,value rows.
usage: python tools/ncu_summary.py <report.ncu-rep> <out.csv>   (needs ncu on PATH)"""
import csv
import subprocess
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
names, units, vals = rows[0], rows[1], rows[2]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value"])
    w.writerow(["kernel", "", vals[names.index("Kernel Name")][:80]])
    for n, u, v in zip(names, units, vals):
        key = n.split(".", 1)[1] if n.split(".", 1)[0].isupper() and "." in n else n
        if n in KEEP or key in KEEP:
            w.writerow([key if key in KEEP else n, u, v])
print("wrote", sys.argv[2])
