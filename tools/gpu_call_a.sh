#!/bin/bash
# one-GPU validation call: DB post-process (path halving, 512-thread boxes), filter multicast, A/B bench, kernel times
o=gpurun_out
timeout 400 python -m pytest tests/test_dbpost_gpu.py tests/test_stages_gpu.py tests/test_configs_gpu.py -x -q -m gpu > $o/t_a1.log 2>&1; tail -2 $o/t_a1.log
B200OCR_CONV_MULTICAST=1 B200OCR_CONV_MULTICAST_MIN=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "conv" > $o/t_a2.log 2>&1; tail -2 $o/t_a2.log
timeout 300 python -m pytest tests/test_env_paths_gpu.py -x -q -m gpu -k "MULTICAST or fused" > $o/t_a3.log 2>&1; tail -2 $o/t_a3.log
B200OCR_CONV_MULTICAST=1 timeout 300 python -m pytest tests/test_configs_gpu.py tests/test_net_parity_gpu.py -x -q -m gpu > $o/t_a4.log 2>&1; tail -2 $o/t_a4.log
for v in 0 1 0 1; do
  B200OCR_CONV_MULTICAST=$v timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-latency > $o/bench_mc$v.json 2> $o/bench_mc$v.err
  python -c "
import json;d=json.load(open('$o/bench_mc$v.json'));print('multicast', $v, round(d['value']), round(d['e2e']['value']), d['roofline']['net_ms_at_timed_shape'])"
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $o/launches_pipeline.csv python tools/ncu_workload.py pipeline > $o/ncu_pipeline.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/launches_pipeline.csv", errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
k, v = rows[hdr].index("Kernel Name"), rows[hdr].index("Metric Value")
acc = collections.OrderedDict()
for r in rows[hdr + 2:]:
    if len(r) <= v: continue
    name = r[k].split("(")[0].split("::")[-1]
    if any(s in name for s in ("ccl", "mark", "boxes", "slot", "list", "sort_cand", "fused_stem", "dbhead", "nest")):
        acc.setdefault(name, []).append(float(r[v].replace(",", "")) / 1e3)
for n, t in acc.items(): print(f"{n:40s} n={len(t):3d} avg {sum(t)/len(t):8.1f} us")
PY
