#!/bin/bash
# GPU box: throughput of the C4 bench under different launch-granularity knobs (one JSON value per line)
run() { echo -n "$1 => "; env $1 timeout 300 python bench.py --steps 8 --warmup 3 --workers ${W:-2} --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), d['ms_per_step'])"; }
run "X=1"
run "B200OCR_REC_MAX_COLS=200000"
run "B200OCR_REC_MAX_COLS=100000"
run "B200OCR_REC_MAX_COLS=50000"
run "B200OCR_REC_MAX_COLS=25000"
run "B200OCR_REC_FILL=0.9"
run "B200OCR_REC_FILL=0.9 B200OCR_REC_MAX_COLS=50000"
run "B200OCR_DET_MAX_BATCH=16"
run "B200OCR_DET_MAX_BATCH=8"
run "B200OCR_CLS_MAX_BATCH=128"
run "B200OCR_CLS_MAX_BATCH=64"
W=3 run "B200OCR_REC_FILL=0.9 B200OCR_REC_MAX_COLS=50000"
W=1 run "B200OCR_REC_FILL=0.9 B200OCR_REC_MAX_COLS=50000"
