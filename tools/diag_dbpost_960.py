"""Diagnostic (GPU box): which boxes of the oracle's DBPostProcess are missing from the GPU's, on the oracle's own
probability map, for an S-page at limit 960 and an S-card at limit 960."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cpp-paddle-ocr_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import numpy as np
import b200ocr, make_synth_weights, synth_data
from oracle import ocr_ops
from oracle.pipeline import OracleDetector

models = make_synth_weights.ensure_models()
for name, im, th, bt, ur in (("page3", synth_data.page(3), 0.2, 0.4, 1.8), ("card500", synth_data.card(500), 0.3, 0.5, 2.0)):
    odet = OracleDetector(f"{models}/det", "max", 960, th, bt, ur, "fast", False)
    det = b200ocr.Detector(f"{models}/det", limit_type="max", limit_side_len=960, det_db_thresh=th, det_db_box_thresh=bt,
                           det_db_unclip_ratio=ur, det_db_score_mode="fast")
    pred, rh, rw = odet.forward(im)
    trace = []
    ref, bitmap = odet.post(pred, rh, rw, im.shape[0], im.shape[1], trace=trace)
    mine, bm = det.postprocess(pred, im.shape[0], im.shape[1], want_bitmap=True)
    print(name, "pred", pred.shape, "oracle boxes", len(ref), "gpu boxes", len(mine), "bitmap equal", np.array_equal(bm, bitmap),
          "contours", len(trace))
    from collections import Counter
    print("  oracle stages", Counter(t["stage"] for t in trace))
    mine_l = [m.tolist() for m in mine]
    for r in ref:
        if not any(np.abs(np.asarray(m) - np.asarray(r)).max() <= 3 for m in mine_l):
            print("  missing on GPU:", r)
    for m in mine_l:
        if not any(np.abs(np.asarray(m) - np.asarray(r)).max() <= 3 for r in ref):
            print("  extra on GPU:", m)
    # the oracle's candidates near a decision threshold
    for t in trace:
        if "score" in t and abs(t["score"] - bt) < 0.02:
            print("  near box_thresh:", t["start"], t["npts"], round(t["score"], 4), t["stage"])
        if t["stage"] in ("ssid", "ssid2", "unclip"):
            print("  dropped by size rule:", t["start"], t["npts"], t["stage"], t.get("ssid"), t.get("ssid2"))
