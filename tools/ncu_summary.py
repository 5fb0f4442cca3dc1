"""Turn an Nsight Compute report (.ncu-rep, read here without a GPU) into the short per-launch summary kept under
profiles/: duration, DRAM bytes, DRAM %, tensor-pipe %, SM %, registers, achieved occupancy.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep "header line" > profiles/x_summary.txt"""
import csv, subprocess, sys

rep, header = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
idx = [hdr.index(w) for w in want if w in hdr]
print(header)
print("columns: " + " | ".join(f"{hdr[i]} [{units[i]}]" for i in idx))
for r in rows[2:]:
    print(" | ".join(r[i] for i in idx))
