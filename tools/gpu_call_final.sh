#!/bin/bash
# round-2 closing call (one GPU): full GPU test suite, the driver's bench line + ncu traffic of the dominant family,
# the other configurations, the reference arm, a launch list, REC_FILL check
o=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > $o/r02_tests_gpu.log 2>&1; tail -3 $o/r02_tests_gpu.log
timeout 400 bash tools/capture_traffic.sh c4 $o/r02; cp $o/r02_traffic_c4.json profiles/r02_traffic.json 2>/dev/null
for c in c2 c3 c5; do timeout 200 python bench.py --config $c --steps 10 --warmup 3 > $o/r02_bench_$c.json 2> $o/r02_bench_$c.err; tail -c 300 $o/r02_bench_$c.json | head -c 0; python -c "
import json;d=json.load(open('$o/r02_bench_$c.json'));print('$c', round(d['value'],1), round(d['e2e']['value'],1), d['unit'], d['roofline']['kernel'], round(d['roofline']['frac'],3))"; done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $o/r02_bench_reference_c4.json 2> $o/r02_bench_reference_c4.err; head -c 400 $o/r02_bench_reference_c4.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $o/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-roofline --no-latency --prime-seconds 0.1 > $o/ncu_bench_r02.log 2>&1
python tools/ncu_summary.py $o/r02_launches.csv > $o/r02_launches_summary.txt; head -14 $o/r02_launches_summary.txt
for v in 0.85 0.92 0.85 0.92; do
  B200OCR_REC_FILL=$v timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-latency --no-roofline > $o/bench_fill$v.json 2> $o/bench_fill$v.err
  python -c "
import json;d=json.load(open('$o/bench_fill$v.json'));print('fill', $v, round(d['value']), round(d['e2e']['value']))"
done
timeout 200 python bench.py --steps 10 --warmup 3 > $o/r02_bench_c4_final.json 2> $o/r02_bench_c4_final.err; python -c "
import json;d=json.load(open('$o/r02_bench_c4_final.json'));print('c4 final', round(d['value']), round(d['e2e']['value']), d['roofline']['kernel'], round(d['roofline']['frac'],3), d['roofline']['traffic'], d['cpu_baseline']['value'])"
