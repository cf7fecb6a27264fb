"""Print the headline metrics of an .ncu-rep (raw page) per captured kernel."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rd = list(csv.reader(raw.splitlines()))
hdr, units, rows = rd[0], rd[1], rd[2:]
want = ["Kernel Name", "launch__grid_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tc", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_uniform", "smsp__inst_executed.sum", "sm__pipe_tc_cycles_active", "sm__inst_executed_pipe_tmem"]
for i, hname in enumerate(hdr):
    if any(hname == w or hname.startswith(w + ".") or hname == w for w in want):
        print(f"{hname} [{units[i]}]: " + " | ".join(r[i] for r in rows))
if len(sys.argv) > 2:
    for i, hname in enumerate(hdr):
        if sys.argv[2] in hname:
            print(f"{hname} [{units[i]}]: " + " | ".join(r[i] for r in rows))
