"""Print selected metrics of every kernel in an ncu report: python tools/ncu_raw.py report.ncu-rep [substring ...]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
default = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
           "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "lts__t_bytes.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "sm__inst_executed_pipe",
           "sm__pipe_fma_cycles_active", "sm__pipe_tensor", "issue_stalled", "bank_conflicts", "l1tex__data_pipe_lsu_wavefronts_mem_shared",
           "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
subs = sys.argv[2:] or default
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    print("====", r[ki][:100])
    for i, h in enumerate(hdr):
        if any(s in h for s in subs) and r[i] not in ("", "0", "n/a"):
            name = h.split(".", 2)[-1] if h.count(".") > 2 and h.split(".")[1].startswith("Triage") else h
            print(f"  {name[:95]:95s} {r[i]:>16s} {units[i]}")
