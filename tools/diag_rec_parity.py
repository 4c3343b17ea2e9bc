"""Diagnostic (GPU box): recognizer parity statistics of the CUDA path against the oracle on S-card lines.
    python tools/diag_rec_parity.py [--rec-params file.pdiparams] [--cards 30]
Prints: lines compared, identical strings, max |confidence difference| over identical strings, and -- through the
network-level entry point on the oracle's own input tensors -- the per-step |max-prob difference| distribution and the
number of arg-max flips with the oracle's top-2 margin at each flip."""
import argparse, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cpp-paddle-ocr_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import cv2, numpy as np
import b200ocr, make_synth_weights, synth_data
from oracle import ocr_ops
from oracle.pipeline import OracleWorker, OracleRecognizer

ap = argparse.ArgumentParser()
ap.add_argument("--rec-params", default="")
ap.add_argument("--cards", type=int, default=30)
ap.add_argument("--h", type=int, default=28)
ap.add_argument("--w", type=int, default=192)
a = ap.parse_args()
models = make_synth_weights.ensure_models()
if a.rec_params:
    shutil.copyfile(a.rec_params, os.path.join(models, "rec", "inference.pdiparams"))
H, W, B = a.h, a.w, 16
label = f"{models}/rec/ppocr_keys_v1.txt"
det = b200ocr.Detector(f"{models}/det", limit_type="max", limit_side_len=512, det_db_thresh=0.2, det_db_box_thresh=0.4,
                       det_db_unclip_ratio=1.8, det_db_score_mode="fast")
rec = b200ocr.Recognizer(f"{models}/rec", label, rec_batch_num=B, rec_img_h=H, rec_img_w=W)
orec = OracleRecognizer(f"{models}/rec", label, B, H, W)
net = b200ocr.Net(f"{models}/rec", 0, b200ocr.NET_NO_GRAPH)
imgs = [cv2.imread(os.path.join(ROOT, "tests/golden/card-jd.jpg")), synth_data.reference_test_image()] + \
       [synth_data.card(1000 + i) for i in range(a.cards)]
n_lines = same = flips = steps = string_flips = 0
dscore, dstep, flip_margin = [], [], []
for im in imgs:
    crops = []
    for b in det.run(im):
        r = ocr_ops.bounding_rect_crop(b.tolist(), im.shape[0], im.shape[1])
        if r:
            x, y, w, h = r
            crops.append(im[y:y + h, x:x + w])
    if not crops:
        continue
    texts, scores = rec.run(crops)
    rt, rs, raw = orec.run(crops, want_raw=True)
    for i in range(len(crops)):
        n_lines += 1
        if texts[i] == rt[i]:
            same += 1
            dscore.append(abs(float(scores[i]) - float(rs[i])))
        else:
            string_flips += 1
            print("  differs:", repr(texts[i]), "vs oracle", repr(rt[i]), "min margin", float((raw[i][1] - raw[i][2]).min()))
    for idx, x in ocr_ops.rec_batches(crops, B, H, W):
        prob, am = net.forward(x)
        for m, i in enumerate(idx):
            ridx, rmx, rsec = raw[i]
            d = np.abs(prob[m] - rmx)
            ok = am[m] == ridx
            dstep.extend(d[ok].tolist())
            steps += len(d)
            flips += int((~ok).sum())
            flip_margin.extend((rmx - rsec)[~ok].tolist())
dstep = np.array(dstep)
print(f"lines {n_lines}  identical strings {same}  differing {string_flips}")
print(f"confidence |diff| over identical strings: max {max(dscore):.4f}  p99 {np.quantile(dscore, 0.99):.4f}  >1e-2: {sum(d > 1e-2 for d in dscore)}")
print(f"steps {steps}  arg-max flips {flips}  oracle top-2 margin at the flips: {sorted(round(f, 4) for f in flip_margin)[:20]}")
print(f"per-step |max-prob diff| (same arg-max): max {dstep.max():.4f}  p999 {np.quantile(dstep, 0.999):.4f}  p99 {np.quantile(dstep, 0.99):.4f}  >1e-2: {int((dstep > 1e-2).sum())}")
