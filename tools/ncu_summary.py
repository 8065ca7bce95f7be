"""Key metrics of an ncu --set full report (.ncu-rep) as text: python tools/ncu_summary.py <rep> [> profiles/...]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "sm__cycles_elapsed.max", "gpc__cycles_elapsed.avg.per_second",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_uniform.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]
for r in rows[2:]:
    print("----")
    for k in want:
        for i, h in enumerate(hdr):
            if h == k:
                print("%-95s %s %s" % (k, r[i], units[i]))
