#!/bin/bash
# one-GPU call: ncu --set full of the 480-wide convolutions and of the fused stem, REC_FILL A/B
o=gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv_tc_kernel" -c 3 -f -o $o/r02_conv480 python tools/ncu_workload.py rec "[205,28,704]" > $o/ncu_conv480.log 2>&1; tail -2 $o/ncu_conv480.log
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"fused_stem|boxes_kernel|ccl_runs_merge" -f -o $o/r02_stem_dbpost python tools/ncu_workload.py pipeline > $o/ncu_stem.log 2>&1; tail -2 $o/ncu_stem.log
for v in 0.75 0.85 0.75 0.85; do
  B200OCR_REC_FILL=$v timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-latency --no-roofline > $o/bench_fill$v.json 2> $o/bench_fill$v.err
  python -c "
import json;d=json.load(open('$o/bench_fill$v.json'));print('fill', $v, round(d['value']), round(d['e2e']['value']))"
done
ls -la $o/*.ncu-rep
timeout 120 python -m pytest tests/test_service.py tests/test_stages_gpu.py -x -q -m gpu -k "pool or service" > $o/pool_tests.log 2>&1; tail -2 $o/pool_tests.log
timeout 200 python tools/pool_sweep.py 1 f16:3:64:16 > $o/pool_sweep_n1e.txt 2>&1; cat $o/pool_sweep_n1e.txt
