#!/bin/bash
# knob sweep at the round-2 default batch (run on the GPU box): python bench.py lines, resident / e2e images/s
run() { env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline --no-latency 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$*', '=>', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2))"; }
run X=1
run B200OCR_DET_MAX_BATCH=64
run B200OCR_REC_MAX_COLS=800000
run B200OCR_REC_MAX_COLS=800000 B200OCR_REC_MAX_ROWS=2048
run B200OCR_DET_MAX_BATCH=64 B200OCR_REC_MAX_COLS=800000 B200OCR_REC_MAX_ROWS=2048
run B200OCR_CLS_MAX_BATCH=1024
run B200OCR_REC_FILL=0.85
run X=2
