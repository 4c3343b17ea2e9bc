"""Diagnostic: where does the Recognizer stage differ from the oracle?  (run on the GPU box)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cpp-paddle-ocr_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import cv2, numpy as np
import b200ocr, make_synth_weights, synth_data
from oracle import ocr_ops
from oracle.pipeline import OracleWorker, OracleRecognizer

models = make_synth_weights.ensure_models()
ow = OracleWorker(0, models, enable_cls=False)
imgs = [cv2.imread(os.path.join(ROOT, "tests/golden/card-jd.jpg")), synth_data.reference_test_image(), synth_data.card(0), synth_data.card(1)]
crops = []
for im in imgs:
    for b in ow.det.run(im):
        x, y, w, h = ocr_ops.bounding_rect_crop(b, im.shape[0], im.shape[1])
        crops.append(im[y:y + h, x:x + w])
crops = crops[:40]
H, W, B = 48, 320, 6
label = f"{models}/rec/ppocr_keys_v1.txt"
rec = b200ocr.Recognizer(f"{models}/rec", label, rec_batch_num=B, rec_img_h=H, rec_img_w=W)
orec = OracleRecognizer(f"{models}/rec", label, B, H, W)
texts, scores = rec.run(crops)
rt, rs, raw = orec.run(crops, want_raw=True)
net = b200ocr.Net(f"{models}/rec", 0, b200ocr.NET_NO_GRAPH)
for idx, x in ocr_ops.rec_batches(crops, B, H, W):
    mine = b200ocr.crop_preprocess([crops[i] for i in idx], "rec", H, x.shape[3])
    # crop_preprocess uses wh ratio = img_w/img_h for resize_w cap; here the batch width is x.shape[3]
    dpre = np.abs(mine - x).max()
    prob, am = net.forward(x)
    from oracle.interp import run_program
    ref = run_program(orec.net.prog, orec.net.params, x)[0]
    dnet = np.abs(prob - ref.max(-1)).max()
    print("batch", idx, "W", x.shape[3], "pre diff", f"{dpre:.4f}", "net max|dp|", f"{dnet:.4f}",
          "stage-vs-oracle score diffs", [f"{abs(scores[i] - rs[i]):.3f}{'' if texts[i] == rt[i] else '*'}" for i in idx],
          "crop sizes", [crops[i].shape[:2] for i in idx])
