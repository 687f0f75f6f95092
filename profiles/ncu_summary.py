"""Key metrics of an `ncu --set full` report as text (the .ncu-rep itself stays in gpurun_out/ scratch).
Usage: python profiles/ncu_summary.py report.ncu-rep > profiles/rNN_ncu_<kernel>.txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("kernel:", d["Kernel Name"])
    print("grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for k in KEYS:
        if k in d:
            print("  %-72s %s %s" % (k, d[k], units[hdr.index(k)]))
# stall reasons summed over the source page
src = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
if len(rows) > 2:
    hdr = rows[1]
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = {hdr[i]: sum(int(r[i] or 0) for r in rows[2:] if len(r) > i) for i in stall}
    tot = sum(agg.values()) or 1
    print("warp stall samples (all):", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
