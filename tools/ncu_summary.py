"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/ncu_summary.py file.csv"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = []
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
        u = row.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1e-3)
        rows.append((row["Kernel Name"], v))
    except Exception:
        pass
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for k, v in rows:
    k = re.sub(r"\(.*", "", k)
    k = re.sub(r"b200ocr::|\(anonymous namespace\)::|<unnamed>::", "", k)[:70]
    agg[k][0] += 1; agg[k][1] += v; agg[k][2] = max(agg[k][2], v)
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot / 1e3:.2f} ms of kernel time")
print(f"{'total us':>10} {'share':>6} {'n':>6} {'avg us':>8} {'max us':>8}  kernel")
for k, (c, t, m) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} {100 * t / tot:5.1f}% {c:6d} {t / c:8.1f} {m:8.1f}  {k}")
