"""Throughput of BASELINE.json configs 2, 3 and 5 on one GPU through the reference-facing stage calls (host images in,
host results out; wall clock over `reps` repetitions after two warm-up calls).  Run on the GPU box:
    python tools/bench_configs.py > profiles/rNN_configs.txt"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cpp-paddle-ocr_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import numpy as np
import b200ocr, make_synth_weights, synth_data

models = make_synth_weights.ensure_models()
label = f"{models}/rec/ppocr_keys_v1.txt"


def timeit(fn, reps=5):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


# C2: recognition-only, 4096 crops of 48x320
crops = list(synth_data.rec_crops(4096, 48, 320, seed=0))
rec = b200ocr.Recognizer(f"{models}/rec", label, rec_batch_num=6, rec_img_h=48, rec_img_w=320)
dt = timeit(lambda: rec.run(crops))
print(f"C2 recognition-only: 4096 crops 48x320 (CRNN/SVTR forward + CTC greedy decode): {dt * 1e3:.1f} ms per call, "
      f"{4096 / dt:.0f} crops/s (1.405 GFLOP per crop -> {4096 * 1.405 / dt / 1e3:.1f} TFLOP/s)")

# C3: detection-only, batch 64 at 960 max side
imgs = [synth_data.card(500 + i) for i in range(64)]
det = b200ocr.Detector(f"{models}/det", limit_type="max", limit_side_len=960, det_db_thresh=0.3, det_db_box_thresh=0.5,
                       det_db_unclip_ratio=2.0, det_db_score_mode="fast")
dt = timeit(lambda: det.run_batch(imgs))
print(f"C3 detection-only: 64 cards 1024x640 at limit 960 ([64,3,608,960] DB forward + DBPostProcess): {dt * 1e3:.1f} ms per "
      f"call, {64 / dt:.0f} images/s")

# C5: dense 2048x2048 pages, det at 960 + all crops through one recognizer call
pages = [synth_data.page(3 + i) for i in range(4)]
from b200ocr import Detector, Recognizer
det5 = Detector(f"{models}/det", limit_type="max", limit_side_len=960, det_db_thresh=0.2, det_db_box_thresh=0.4,
                det_db_unclip_ratio=1.8, det_db_score_mode="fast")
rec5 = Recognizer(f"{models}/rec", label, rec_batch_num=16, rec_img_h=28, rec_img_w=192)


def c5():
    n = 0
    all_boxes = det5.run_batch(pages)
    for page, boxes in zip(pages, all_boxes):
        cs = []
        for b in boxes:
            xs, ys = b[:, 0], b[:, 1]
            x0, y0 = max(int(xs.min()), 0), max(int(ys.min()), 0)
            x1, y1 = min(int(xs.max()) + 1, page.shape[1]), min(int(ys.max()) + 1, page.shape[0])
            if x1 > x0 and y1 > y0:
                cs.append(page[y0:y1, x0:x1])
        rec5.run(cs)
        n += len(cs)
    return n


lines = c5()
dt = timeit(c5, reps=3)
print(f"C5 dense pages: 4 pages 2048x2048, {lines} text lines in total (det at 960 + ragged rec of every line): {dt * 1e3:.1f} ms "
      f"per call, {4 / dt:.1f} pages/s, {lines / dt:.0f} lines/s")
