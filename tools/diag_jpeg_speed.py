"""Diagnostic (GPU box): time of the device JPEG decode inside Worker.process_encoded (B200OCR_TRACE=1 prints it)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cpp-paddle-ocr_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
os.environ["B200OCR_TRACE"] = "1"
import cv2, numpy as np
import b200ocr, make_synth_weights, synth_data
models = make_synth_weights.ensure_models()
w = b200ocr.Worker(0, models, enable_cls=True)
for q, rst in ((90, 0), (90, 8), (75, 0)):
    files = [cv2.imencode(".jpg", synth_data.card(100 + i), [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_RST_INTERVAL, rst])[1].tobytes()
             for i in range(21)]
    print(f"quality {q} restart {rst}: {sum(map(len, files)) // len(files)} bytes per file", file=sys.stderr, flush=True)
    for rep in range(3):
        t0 = time.perf_counter()
        w.process_encoded(list(range(21)), files)
        print(f"   process_encoded of 21 files: {(time.perf_counter() - t0) * 1e3:.1f} ms", file=sys.stderr, flush=True)
