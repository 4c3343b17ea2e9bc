"""smoke(): one small det -> cls -> rec pass on cuda:0 through the C ABI, checked against the CPU oracle."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "cpp-paddle-ocr_b200"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)


def run():
    import cv2
    import numpy as np
    import b200ocr
    import make_synth_weights
    from oracle.pipeline import OracleWorker

    if b200ocr.device_count() < 1:
        raise RuntimeError("smoke() needs a CUDA device: b200ocr has no CPU fallback")
    models = make_synth_weights.ensure_models()
    img = cv2.imread(os.path.join(ROOT, "tests", "golden", "card-jd.jpg"))
    worker = b200ocr.Worker(0, models, gpu_id=0, enable_cls=True)
    line = worker.process(1, img)
    res = json.loads(line)
    assert res["success"] and res["width"] == img.shape[1] and res["height"] == img.shape[0], line
    oracle = OracleWorker(0, models, enable_cls=True)
    ref_boxes = oracle.det.run(img)
    got_boxes = [w["box"] for w in res["words"]]
    assert len(got_boxes) == len(ref_boxes), (len(got_boxes), len(ref_boxes))
    for g, r in zip(got_boxes, ref_boxes):
        assert np.abs(np.asarray(g) - np.asarray(r)).max() <= 1, (g, r)
    ref_words, raw = oracle.process_words(img, det_boxes=got_boxes, want_raw=True)
    for w, (t, s, _b), (_idx, mx, second) in zip(res["words"], ref_words, raw):
        assert w["text"] == t, (w["text"], t)  # every recognised string identical to the oracle's: N / N or fail
        # confidence within the 1e-2 contract, except in a line where the oracle itself has a near-tie step
        assert abs(w["confidence"] - s) < 1e-2 or (mx - second).min() < 1e-2, (w, s)
    print(f"smoke ok: {len(got_boxes)} boxes, {len(ref_words)}/{len(ref_words)} lines identical to the oracle, "
          f"{worker.launches} kernel launches, {b200ocr.version()}")
