#!/bin/bash
run() { echo -n "$1 => "; env $1 timeout 300 python bench.py --steps 8 --warmup 3 --workers ${W:-2} --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), d['ms_per_step'])"; }
run "X=1"
run "B200OCR_REC_MAX_COLS=250000"
run "B200OCR_REC_MAX_COLS=150000"
run "B200OCR_REC_MAX_COLS=800000"
run "B200OCR_DET_MAX_BATCH=64"
run "B200OCR_CLS_MAX_BATCH=1024"
run "B200OCR_REC_FILL=0.6"
run "B200OCR_REC_FILL=0.85"
