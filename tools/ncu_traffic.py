"""Per-kernel table from an `ncu --set full` report: duration, DRAM bytes read + written, achieved DRAM throughput,
L2 bytes, SM busy -- and (optionally) the summed DRAM traffic of one kernel family as JSON for bench.py's
roofline.traffic.   python tools/ncu_traffic.py report.ncu-rep [name-regex]"""
import csv, io, json, re, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def get(r, name, default=0.0):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return default
    return float(r[i].replace(",", ""))
def scaled(r, name, want):
    """value converted to `want` units (byte / Kbyte / Mbyte / Gbyte; nsecond / usecond / msecond)"""
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return 0.0
    v, u = float(r[i].replace(",", "")), units[i]
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1,
         "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(u, 1)
    g = {"byte": 1, "MB": 1e6, "us": 1e-6}[want]
    return v * f / g
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
print(f"{'us':>8} {'dram rd MB':>10} {'dram wr MB':>10} {'GB/s':>7} {'%dram':>6} {'L2 MB':>8} {'%sm':>5} {'regs':>4} {'grid':>7}  kernel")
tot = {"us": 0.0, "rd": 0.0, "wr": 0.0, "n": 0}
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
    name = re.sub(r"b200ocr::|\(anonymous namespace\)::|<unnamed>::|void ", "", name)
    us = scaled(r, "gpu__time_duration.sum", "us")
    rd, wr = scaled(r, "dram__bytes_read.sum", "MB"), scaled(r, "dram__bytes_write.sum", "MB")
    l2 = scaled(r, "lts__t_bytes.sum", "MB")
    print(f"{us:8.1f} {rd:10.2f} {wr:10.2f} {(rd + wr) / us * 1e3 if us else 0:7.0f} "
          f"{get(r, 'dram__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} {l2:8.1f} "
          f"{get(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} {int(get(r, 'launch__registers_per_thread')):4d} "
          f"{int(get(r, 'launch__grid_size')):7d}  {name[:60]}")
    if pat is None or pat.search(name):
        tot["us"] += us; tot["rd"] += rd; tot["wr"] += wr; tot["n"] += 1
if tot["n"]:
    print(f"# family {sys.argv[2] if pat else '(all)'}: {tot['n']} launches, {tot['us']:.1f} us, dram read {tot['rd']:.1f} MB + write {tot['wr']:.1f} MB"
          f" = {(tot['rd'] + tot['wr']) / tot['n']:.2f} MB per launch")
    print("# json", json.dumps({"launches": tot["n"], "traffic_bytes_per_launch": (tot["rd"] + tot["wr"]) * 1e6 / tot["n"],
                                "us_per_launch_under_ncu": tot["us"] / tot["n"]}))
